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
// (~30 PyTorch launches in the unfused path).  One 128-thread block per state (grid-strided), fp32 math,
// weights pre-transposed so that consecutive threads read consecutive addresses.
// ------------------------------------------------------------------------------------------------------
namespace lzb {
namespace {

constexpr int kHeadsThreads = 128;
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
    __shared__ float act[36][2 * kMaxHeadCh + 1];
    __shared__ float pooled_p[3 * kMaxHeadCh], pooled_v[3 * kMaxHeadCh];
    __shared__ float g[kMaxHeadCh], hid[kMaxMlp], vlog[kMaxBins];
    __shared__ float logit3[3][36], lp[3][36];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int pc = P.pc, vc = P.vc, c2 = pc + vc;
    for (int64_t i = blockIdx.x; i < P.n; i += gridDim.x) {
        __syncthreads();
        // 1. load + global pooling (mean / max / std, biased variance + 1e-6)   neural_network.py:68-81
        if (t < c2) {
            const __nv_bfloat16* src = P.pv + i * 36 * c2 + t;
            float s = 0.0f, mx = -INFINITY;
#pragma unroll 4
            for (int cell = 0; cell < 36; ++cell) {
                const float x = __bfloat162float(src[cell * c2]);
                act[cell][t] = x; s += x; mx = fmaxf(mx, x);
            }
            const float mean = s * (1.0f / 36.0f);
            float var = 0.0f;
#pragma unroll 4
            for (int cell = 0; cell < 36; ++cell) { const float d = act[cell][t] - mean; var += d * d; }
            const float sd = sqrtf(var * (1.0f / 36.0f) + 1e-6f);
            if (t < pc) { pooled_p[t] = mean; pooled_p[pc + t] = mx; pooled_p[2 * pc + t] = sd; }
            else { const int c = t - pc; pooled_v[c] = mean; pooled_v[vc + c] = mx; pooled_v[2 * vc + c] = sd; }
        }
        __syncthreads();
        // 2. policy gpool_linear (no bias) and value fc1 + relu
        if (t < pc) {
            float a = 0.0f;
            for (int k = 0; k < 3 * pc; ++k) a = fmaf(P.wgl_t[k * pc + t], pooled_p[k], a);
            g[t] = a;
        }
        for (int m = t; m < P.mlp; m += kHeadsThreads) {
            float a = P.bfc1[m];
            for (int k = 0; k < 3 * vc; ++k) a = fmaf(P.wfc1_t[k * P.mlp + m], pooled_v[k], a);
            hid[m] = fmaxf(a, 0.0f);
        }
        __syncthreads();
        // 3. p2 = relu(bn2(p + g)); three 1-channel output convs; value fc2
        if (t < 108) {
            const int cell = t % 36, h = t / 36;
            float a = 0.0f;
            for (int ch = 0; ch < pc; ++ch) {
                const float p2 = fmaxf(fmaf(P.bn2_scale[ch], act[cell][ch] + g[ch], P.bn2_shift[ch]), 0.0f);
                a = fmaf(P.wout[h * pc + ch], p2, a);
            }
            logit3[h][cell] = a;
        }
        for (int k = t; k < P.bins; k += kHeadsThreads) {
            float a = P.bfc2[k];
            for (int m = 0; m < P.mlp; ++m) a = fmaf(P.wfc2_t[m * P.bins + k], hid[m], a);
            vlog[k] = a;
        }
        __syncthreads();
        // 4. log-softmax of each policy head over the 36 cells (warps 0..2); value expectation (warp 3)
        if (warp < 3) {
            const float x0 = logit3[warp][lane], x1 = lane < 4 ? logit3[warp][32 + lane] : -INFINITY;
            float mx = fmaxf(x0, x1);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float se = expf(x0 - mx) + (lane < 4 ? expf(x1 - mx) : 0.0f);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
            const float lse = mx + logf(se);
            lp[warp][lane] = x0 - lse;
            if (lane < 4) lp[warp][32 + lane] = x1 - lse;
            if (P.log_heads) {
                P.log_heads[(i * 3 + warp) * 36 + lane] = x0 - lse;
                if (lane < 4) P.log_heads[(i * 3 + warp) * 36 + 32 + lane] = x1 - lse;
            }
        } else {
            float mx = -INFINITY;
            for (int k = lane; k < P.bins; k += 32) mx = fmaxf(mx, vlog[k]);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float se = 0.0f, sc = 0.0f;
            const float step = P.bins > 1 ? 2.0f / (float)(P.bins - 1) : 0.0f;
            for (int k = lane; k < P.bins; k += 32) {
                const float e = expf(vlog[k] - mx);
                se += e; sc += e * (-1.0f + step * (float)k);
                if (P.value_logits) P.value_logits[i * P.bins + k] = vlog[k];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                se += __shfl_xor_sync(0xffffffffu, se, off);
                sc += __shfl_xor_sync(0xffffffffu, sc, off);
            }
            if (lane == 0 && P.values) P.values[i] = sc / se;
        }
        __syncthreads();
        // 5. priors = softmax of the combined logits over the legal actions of the packed state (warp 0)
        if (warp == 0 && P.priors && P.states) {
            lz::State<int> s;
            lz::Packed pk;
            const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(P.states + 4 * i);
            const ulonglong2 a = sp[0], b = sp[1];
            pk.w[0] = a.x; pk.w[1] = a.y; pk.w[2] = b.x; pk.w[3] = b.y;
            lz::unpack(pk, s);
            lz::Legal L;
            lz::legal_actions<int, true>(s, L, true);
            float logit[7], mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const int ac = lane + 32 * k;
                float v = -INFINITY;
                if (ac < lz::kActionDim && lz::legal_test(L, ac)) {
                    if (ac < 36) v = lp[0][ac];
                    else if (ac < 180) {
                        const int from = (ac - 36) >> 2, d = (ac - 36) & 3;
                        v = lp[1][from] + lp[0][from + (d == 0 ? -6 : d == 1 ? 6 : d == 2 ? -1 : 1)];
                    } else if (ac < 216) v = lp[2][ac - 180];
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
    int64_t blocks = n < (int64_t)lzb::kNumSMs * 8 ? n : (int64_t)lzb::kNumSMs * 8;
    lzb::heads_tail_kernel<<<(int)blocks, lzb::kHeadsThreads, 0, (cudaStream_t)stream>>>(P);
    return lzb::check_launch("heads_tail_kernel");
}
