// lz_nn.cu -- fused elementwise epilogues around the network's convolutions (bf16, channels-last).
//
// The dense contraction itself (3x3 convolutions as implicit GEMMs) stays with PyTorch/cuDNN on the tensor
// cores (north_star); measured on B200 those run at ~90 % of the sustained bf16 peak, but PyTorch's eval-mode
// BatchNorm + ReLU + residual add around them are separate HBM-bound passes that cost more than the convs.
// A pre-activation residual block  x' = x + conv2(relu(bn2(conv1(relu(bn1(x))))))  needs, per conv, exactly
// one such pass if the passes are fused:
//     bn_relu      : a  = relu(scale * u + shift)
//     add_bn_relu  : x' = u + v ;  a = relu(scale * x' + shift)          (scale/shift of the NEXT block's bn1)
// Rows are (n, h, w) positions, C channels contiguous (NHWC); each thread handles 8 channels (16 B vectors).
// HBM-bound: 2 x 2 B/element for bn_relu, 4 x 2 B/element for add_bn_relu.
#include <cuda_bf16.h>

#include "lz_common.cuh"

namespace lzb {
namespace {

struct alignas(16) bf16x8 { __nv_bfloat162 v[4]; };

template <bool kAdd, bool kWriteSum>
__global__ void __launch_bounds__(256)
bn_relu_kernel(const bf16x8* __restrict__ u, const bf16x8* __restrict__ v, const float* __restrict__ scale,
               const float* __restrict__ shift, int64_t total_vec, int vec_per_row, bf16x8* __restrict__ out_sum,
               bf16x8* __restrict__ out_act) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % vec_per_row) * 8;
        bf16x8 a = u[i];
        bf16x8 s, r;
        bf16x8 b;
        if (kAdd) b = v[i];
        const float4 sc0 = *reinterpret_cast<const float4*>(scale + c0), sc1 = *reinterpret_cast<const float4*>(scale + c0 + 4);
        const float4 sh0 = *reinterpret_cast<const float4*>(shift + c0), sh1 = *reinterpret_cast<const float4*>(shift + c0 + 4);
        const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
        const float shv[8] = {sh0.x, sh0.y, sh0.z, sh0.w, sh1.x, sh1.y, sh1.z, sh1.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float2 x = __bfloat1622float2(a.v[k]);
            if (kAdd) {
                const float2 y = __bfloat1622float2(b.v[k]);
                // the residual stream is stored in bf16 (as PyTorch's bf16 add does) and BN reads the stored value
                s.v[k] = __floats2bfloat162_rn(x.x + y.x, x.y + y.y);
                x = __bfloat1622float2(s.v[k]);
            }
            const float r0 = fmaxf(fmaf(x.x, scv[2 * k], shv[2 * k]), 0.0f);
            const float r1 = fmaxf(fmaf(x.y, scv[2 * k + 1], shv[2 * k + 1]), 0.0f);
            r.v[k] = __floats2bfloat162_rn(r0, r1);
        }
        if (kAdd && kWriteSum) out_sum[i] = s;
        out_act[i] = r;
    }
}

}  // namespace
}  // namespace lzb

using namespace lzb;

extern "C" int lzb_bn_relu_bf16(const void* u, const void* v, const float* scale, const float* shift, int64_t rows,
                                int32_t channels, void* out_sum, void* out_act, void* stream) {
    LZB_REQUIRE(rows >= 0 && channels > 0 && channels % 8 == 0, "channels must be a positive multiple of 8");
    if (rows == 0) return LZB_OK;
    LZB_REQUIRE(u && scale && shift && out_act, "null pointer");
    LZB_REQUIRE(((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out_sum) |
                  reinterpret_cast<uintptr_t>(out_act) | reinterpret_cast<uintptr_t>(scale) |
                  reinterpret_cast<uintptr_t>(shift)) & 15) == 0, "pointers must be 16-byte aligned");
    const int vec_per_row = channels / 8;
    const int64_t total = rows * vec_per_row;
    const int grid = thread_grid(total, 256);
    cudaStream_t s = (cudaStream_t)stream;
    const bf16x8* pu = reinterpret_cast<const bf16x8*>(u);
    const bf16x8* pv = reinterpret_cast<const bf16x8*>(v);
    bf16x8* ps = reinterpret_cast<bf16x8*>(out_sum);
    bf16x8* pa = reinterpret_cast<bf16x8*>(out_act);
    if (!v) bn_relu_kernel<false, false><<<grid, 256, 0, s>>>(pu, pv, scale, shift, total, vec_per_row, ps, pa);
    else if (out_sum) bn_relu_kernel<true, true><<<grid, 256, 0, s>>>(pu, pv, scale, shift, total, vec_per_row, ps, pa);
    else bn_relu_kernel<true, false><<<grid, 256, 0, s>>>(pu, pv, scale, shift, total, vec_per_row, ps, pa);
    return check_launch("bn_relu_kernel");
}

// ------------------------------------------------------------------------------------------------------
// Fused network heads (src/neural_network.py:98-151 + project_policy_logits_fast + bucket_logits_to_scalar).
// Input: pv = relu(bn(conv1x1([policy.conv1 ; value.conv1])(trunk)))  bf16 [n, 36, pc + vc] channels-last, i.e.
// the two heads' 1x1 convolutions run as ONE cuDNN conv; everything after it -- global pooling (mean / max /
// std), gpool_linear, bn2 + relu, the three 1-channel output convs, log-softmax, the value MLP, the bucket
// expectation and the masked softmax over the legal actions of the packed state -- is this one kernel
// (~30 PyTorch launches in the unfused path).
//
// A 128-thread block processes a TILE of kTile = 16 states per iteration so that every weight element
// (pre-transposed: consecutive threads read consecutive addresses) is loaded once per tile and used for 16
// states from registers; pooled features / hidden activations live in shared memory (fp32 math throughout).
// ------------------------------------------------------------------------------------------------------
namespace lzb {
namespace {

constexpr int kHeadsThreads = 128;
constexpr int kTile = 16;
constexpr int kMaxHeadCh = 64;     // policy_channels, value_channels <= 64
constexpr int kMaxMlp = 128;       // value_mlp_channels <= 128
constexpr int kMaxBins = 128;      // value buckets <= 128

struct HeadsParams {
    const __nv_bfloat16* pv; int64_t n; int pc, vc, mlp, bins;
    const float *wgl_t, *bn2_scale, *bn2_shift, *wout, *wfc1_t, *bfc1, *wfc2_t, *bfc2;
    const uint64_t* states; float *priors, *values, *log_heads, *value_logits;
};

__global__ void __launch_bounds__(kHeadsThreads)
heads_tail_kernel(HeadsParams P) {
    __shared__ float pooled_p[kTile][3 * kMaxHeadCh];
    __shared__ float pooled_v[kTile][3 * kMaxHeadCh];
    __shared__ float g[kTile][kMaxHeadCh];
    __shared__ float hid[kTile][kMaxMlp];
    // value logits reuse pooled_v's storage (dead after the fc1 phase; a __syncthreads separates the two uses)
    float (*vlog)[kMaxBins + 1] = reinterpret_cast<float (*)[kMaxBins + 1]>(&pooled_v[0][0]);
    static_assert(sizeof(float) * kTile * (kMaxBins + 1) <= sizeof(float) * kTile * 3 * kMaxHeadCh, "vlog alias too big");
    __shared__ float lp[kTile][3][36];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pc = P.pc, vc = P.vc, c2 = pc + vc, mlp = P.mlp, bins = P.bins;
    const int64_t num_tiles = (P.n + kTile - 1) / kTile;
    for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int64_t base = tile * kTile;
        const int ns = (int)min((int64_t)kTile, P.n - base);
        __syncthreads();
        // 1. global pooling per (state, channel): mean / max / std (biased variance + 1e-6)  neural_network.py:68-81
        for (int idx = t; idx < ns * c2; idx += kHeadsThreads) {
            const int s = idx / c2, c = idx - s * c2;
            const __nv_bfloat16* src = P.pv + (base + s) * 36 * c2 + c;
            float x[36], sum = 0.0f, mx = -INFINITY;
#pragma unroll
            for (int cell = 0; cell < 36; ++cell) { x[cell] = __bfloat162float(src[cell * c2]); sum += x[cell]; mx = fmaxf(mx, x[cell]); }
            const float mean = sum * (1.0f / 36.0f);
            float var = 0.0f;
#pragma unroll
            for (int cell = 0; cell < 36; ++cell) { const float d = x[cell] - mean; var = fmaf(d, d, var); }
            const float sd = sqrtf(var * (1.0f / 36.0f) + 1e-6f);
            if (c < pc) { pooled_p[s][c] = mean; pooled_p[s][pc + c] = mx; pooled_p[s][2 * pc + c] = sd; }
            else { const int cv = c - pc; pooled_v[s][cv] = mean; pooled_v[s][vc + cv] = mx; pooled_v[s][2 * vc + cv] = sd; }
        }
        __syncthreads();
        // 2a. g = gpool_linear(pooled_p) (no bias): thread -> (output j, half of the tile)
        {
            const int halves = kHeadsThreads / kMaxHeadCh;            // 2
            const int j = t % kMaxHeadCh, hsel = t / kMaxHeadCh;
            if (j < pc) {
                float acc[kTile / 2];
#pragma unroll
                for (int q = 0; q < kTile / 2; ++q) acc[q] = 0.0f;
                for (int k = 0; k < 3 * pc; ++k) {
                    const float w = __ldg(P.wgl_t + k * pc + j);
#pragma unroll
                    for (int q = 0; q < kTile / 2; ++q) acc[q] = fmaf(w, pooled_p[hsel * (kTile / halves) + q][k], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < kTile / 2; ++q) g[hsel * (kTile / halves) + q][j] = acc[q];
            }
        }
        // 2b. hid = relu(fc1(pooled_v)): thread -> output m
        for (int m = t; m < mlp; m += kHeadsThreads) {
            float acc[kTile];
            const float b = __ldg(P.bfc1 + m);
#pragma unroll
            for (int q = 0; q < kTile; ++q) acc[q] = b;
            for (int k = 0; k < 3 * vc; ++k) {
                const float w = __ldg(P.wfc1_t + k * mlp + m);
#pragma unroll
                for (int q = 0; q < kTile; ++q) acc[q] = fmaf(w, pooled_v[q][k], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < kTile; ++q) hid[q][m] = fmaxf(acc[q], 0.0f);
        }
        __syncthreads();
        // 3a. value logits = fc2(hid): thread -> bucket k
        for (int k = t; k < bins; k += kHeadsThreads) {
            float acc[kTile];
            const float b = __ldg(P.bfc2 + k);
#pragma unroll
            for (int q = 0; q < kTile; ++q) acc[q] = b;
            for (int m = 0; m < mlp; ++m) {
                const float w = __ldg(P.wfc2_t + m * bins + k);
#pragma unroll
                for (int q = 0; q < kTile; ++q) acc[q] = fmaf(w, hid[q][m], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < kTile; ++q) vlog[q][k] = acc[q];
        }
        // 3b. policy: p2 = relu(bn2(p + g)); three 1-channel output convs -> raw logits (stored in lp)
        for (int idx = t; idx < ns * 36; idx += kHeadsThreads) {
            const int s = idx / 36, cell = idx - s * 36;
            const __nv_bfloat16* src = P.pv + ((base + s) * 36 + cell) * c2;
            float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
            for (int ch = 0; ch < pc; ++ch) {
                const float p2 = fmaxf(fmaf(__ldg(P.bn2_scale + ch), __bfloat162float(src[ch]) + g[s][ch],
                                            __ldg(P.bn2_shift + ch)), 0.0f);
                a0 = fmaf(__ldg(P.wout + ch), p2, a0);
                a1 = fmaf(__ldg(P.wout + pc + ch), p2, a1);
                a2 = fmaf(__ldg(P.wout + 2 * pc + ch), p2, a2);
            }
            lp[s][0][cell] = a0; lp[s][1][cell] = a1; lp[s][2][cell] = a2;
        }
        __syncthreads();
        // 4. log-softmax of each (state, head) row over the 36 cells; value expectation per state
        for (int row = warp; row < ns * 3; row += kHeadsThreads / 32) {
            const int s = row / 3, h = row - s * 3;
            const float x0 = lp[s][h][lane], x1 = lane < 4 ? lp[s][h][32 + lane] : -INFINITY;
            float mx = fmaxf(x0, x1);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float se = expf(x0 - mx) + (lane < 4 ? expf(x1 - mx) : 0.0f);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
            const float lse = mx + logf(se);
            lp[s][h][lane] = x0 - lse;
            if (lane < 4) lp[s][h][32 + lane] = x1 - lse;
            if (P.log_heads) {
                float* dst = P.log_heads + ((base + s) * 3 + h) * 36;
                dst[lane] = x0 - lse;
                if (lane < 4) dst[32 + lane] = x1 - lse;
            }
        }
        for (int s = warp; s < ns; s += kHeadsThreads / 32) {
            float mx = -INFINITY;
            for (int k = lane; k < bins; k += 32) mx = fmaxf(mx, vlog[s][k]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float se = 0.0f, sc = 0.0f;
            const float step = bins > 1 ? 2.0f / (float)(bins - 1) : 0.0f;
            for (int k = lane; k < bins; k += 32) {
                const float e = expf(vlog[s][k] - mx);
                se += e; sc += e * (-1.0f + step * (float)k);
                if (P.value_logits) P.value_logits[(base + s) * bins + k] = vlog[s][k];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                se += __shfl_xor_sync(0xffffffffu, se, off);
                sc += __shfl_xor_sync(0xffffffffu, sc, off);
            }
            if (lane == 0 && P.values) P.values[base + s] = sc / se;
        }
        __syncthreads();
        // 5. priors = softmax of the combined logits over the legal actions of the packed state (warp per state)
        if (P.priors && P.states) {
            for (int s = warp; s < ns; s += kHeadsThreads / 32) {
                const int64_t i = base + s;
                lz::State<int> st;
                lz::Packed pk;
                const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(P.states + 4 * i);
                const ulonglong2 a = sp[0], b = sp[1];
                pk.w[0] = a.x; pk.w[1] = a.y; pk.w[2] = b.x; pk.w[3] = b.y;
                lz::unpack(pk, st);
                lz::Legal L;
                lz::legal_actions<int, true>(st, L, true);
                float logit[7], mx = -INFINITY;
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int ac = lane + 32 * k;
                    float v = -INFINITY;
                    if (ac < lz::kActionDim && lz::legal_test(L, ac)) {
                        if (ac < 36) v = lp[s][0][ac];
                        else if (ac < 180) {
                            const int from = (ac - 36) >> 2, d = (ac - 36) & 3;
                            v = lp[s][1][from] + lp[s][0][from + (d == 0 ? -6 : d == 1 ? 6 : d == 2 ? -1 : 1)];
                        } else if (ac < 216) v = lp[s][2][ac - 180];
                        else v = 0.0f;
                    }
                    logit[k] = v; mx = fmaxf(mx, v);
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
                const bool ok = mx > -INFINITY && mx < INFINITY;
                float e[7], sum = 0.0f;
#pragma unroll
                for (int k = 0; k < 7; ++k) { e[k] = (ok && logit[k] > -INFINITY) ? expf(logit[k] - mx) : 0.0f; sum += e[k]; }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const int ac = lane + 32 * k;
                    if (ac < lz::kActionDim) P.priors[i * lz::kActionDim + ac] = ok ? e[k] / sum : 0.0f;
                }
            }
        }
    }
}

}  // namespace
}  // namespace lzb

extern "C" int lzb_heads_tail(const void* pv, int64_t n, int32_t pc, int32_t vc, int32_t mlp, int32_t bins,
                              const float* wgl_t, const float* bn2_scale, const float* bn2_shift, const float* wout,
                              const float* wfc1_t, const float* bfc1, const float* wfc2_t, const float* bfc2,
                              const uint64_t* states, float* priors, float* values, float* log_heads,
                              float* value_logits, void* stream) {
    LZB_REQUIRE(n >= 0, "bad batch");
    LZB_REQUIRE(pc >= 1 && pc <= lzb::kMaxHeadCh && vc >= 1 && vc <= lzb::kMaxHeadCh, "head channels must be in [1, 64]");
    LZB_REQUIRE(mlp >= 1 && mlp <= lzb::kMaxMlp && bins >= 2 && bins <= lzb::kMaxBins, "value MLP / bins too large");
    if (n == 0) return LZB_OK;
    LZB_REQUIRE(pv && wgl_t && bn2_scale && bn2_shift && wout && wfc1_t && bfc1 && wfc2_t && bfc2, "null weights");
    LZB_REQUIRE(!priors || states, "priors need the packed states");
    lzb::HeadsParams P;
    P.pv = reinterpret_cast<const __nv_bfloat16*>(pv); P.n = n; P.pc = pc; P.vc = vc; P.mlp = mlp; P.bins = bins;
    P.wgl_t = wgl_t; P.bn2_scale = bn2_scale; P.bn2_shift = bn2_shift; P.wout = wout; P.wfc1_t = wfc1_t; P.bfc1 = bfc1;
    P.wfc2_t = wfc2_t; P.bfc2 = bfc2; P.states = states; P.priors = priors; P.values = values;
    P.log_heads = log_heads; P.value_logits = value_logits;
    const int64_t tiles = (n + lzb::kTile - 1) / lzb::kTile;
    const int64_t blocks = tiles < (int64_t)lzb::kNumSMs * 4 ? tiles : (int64_t)lzb::kNumSMs * 4;
    lzb::heads_tail_kernel<<<(int)blocks, lzb::kHeadsThreads, 0, (cudaStream_t)stream>>>(P);
    return lzb::check_launch("heads_tail_kernel");
}
