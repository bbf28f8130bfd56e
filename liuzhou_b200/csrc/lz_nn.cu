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
