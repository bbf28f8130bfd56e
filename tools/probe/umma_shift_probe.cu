// Probe: does a tcgen05 shared-memory matrix descriptor (K-major, SWIZZLE_128B) accept a start address that is NOT
// 1024-byte aligned (a row shift inside the 8-row swizzle atom), and what does the base_offset field (bits 49-51) do?
// One CTA, one M=128 x N=16 x K=16 bf16 MMA per variant: B = identity, so D[i][k] = A[row read for i][k].
// A is a 192-row x 64-col bf16 array written by hand in the SWIZZLE_128B layout (16 B chunk j of row r at r*128 + ((j ^ (r&7))*16)),
// A[r][k] = (r + 3k) % 251 for k < 16.  For a shift of `o` rows the wanted result is D[i][k] = A[o + i][k].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_shift_probe umma_shift_probe.cu && ./umma_shift_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int kRows = 192;
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(128, 1) probe(const int* shifts, const int* modes, int nvar, float* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    uint8_t* gen = raw + (base - smem_u32(raw));
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(gen);                       // 192 x 128 B = 24 KB
    __nv_bfloat16* B = reinterpret_cast<__nv_bfloat16*>(gen + kRows * 128);         // 16 rows x 128 B (2 KB), 1024-aligned
    const uint32_t a_s = base, b_s = base + kRows * 128;
    const uint32_t bar = b_s + 2048, slot = bar + 16;
    volatile uint32_t* slot_p = reinterpret_cast<volatile uint32_t*>(gen + kRows * 128 + 2048 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kRows * 64; i += 128) {
        const int r = i / 64, k = i % 64;
        const float v = k < 16 ? (float)((r + 3 * k) % 251) : 0.0f;
        const int chunk = k / 8, phys = chunk ^ (r & 7);
        A[r * 64 + phys * 8 + (k % 8)] = __float2bfloat16(v);
    }
    for (int i = tid; i < 16 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        const int chunk = k / 8, phys = chunk ^ (n & 7);
        B[n * 64 + phys * 8 + (k % 8)] = __float2bfloat16(n == k ? 1.0f : 0.0f);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *slot_p;
    uint32_t parity = 0;
    for (int v = 0; v < nvar; ++v) {
        const int o = shifts[v], mode = modes[v];
        if (tid == 0) {
            const uint32_t a_addr = a_s + (uint32_t)o * 128u;
            uint64_t adesc = (uint64_t)((a_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            if (mode == 1) adesc |= (uint64_t)((a_addr >> 7) & 7u) << 49;            // documented formula
            if (mode == 2) adesc |= (uint64_t)((8u - ((a_addr >> 7) & 7u)) & 7u) << 49;   // the opposite sign, in case
            const uint64_t bdesc = (uint64_t)((b_s >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
        }
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        parity ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int row = warp * 32 + lane;
        for (int k = 0; k < 16; ++k) out[(v * 128 + row) * 16 + k] = __uint_as_float(r[k]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    const int shifts_h[] = {0, 8, 1, 1, 1, 6, 6, 6, 7, 7, 13, 13, 42, 42};
    const int modes_h[]  = {0, 0, 0, 1, 2, 0, 1, 2, 0, 1, 0,  1,  0,  1};
    const int nvar = sizeof(shifts_h) / sizeof(int);
    int *shifts, *modes; float* out;
    cudaMalloc(&shifts, sizeof(shifts_h)); cudaMalloc(&modes, sizeof(modes_h)); cudaMalloc(&out, nvar * 128 * 16 * 4);
    cudaMemcpy(shifts, shifts_h, sizeof(shifts_h), cudaMemcpyHostToDevice);
    cudaMemcpy(modes, modes_h, sizeof(modes_h), cudaMemcpyHostToDevice);
    cudaMemset(out, 0xFF, nvar * 128 * 16 * 4);
    const int smem = 1024 + kRows * 128 + 2048 + 64;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<<<1, 128, smem>>>(shifts, modes, nvar, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h[64 * 128 * 16];
    cudaMemcpy(h, out, nvar * 128 * 16 * 4, cudaMemcpyDeviceToHost);
    for (int v = 0; v < nvar; ++v) {
        const int o = shifts_h[v];
        int bad = 0, first_bad = -1;
        for (int i = 0; i < 128; ++i)
            for (int k = 0; k < 16; ++k)
                if (h[(v * 128 + i) * 16 + k] != (float)((o + i + 3 * k) % 251)) { ++bad; if (first_bad < 0) first_bad = i; }
        printf("shift %2d rows, base_offset mode %d: %s (%d wrong of 2048, first bad row %d); rows read for i=0..9 (col 0):", o,
               modes_h[v], bad ? "MISMATCH" : "OK", bad, first_bad);
        for (int i = 0; i < 10; ++i) printf(" %g", h[(v * 128 + i) * 16]);
        printf(" | col1:");
        for (int i = 0; i < 4; ++i) printf(" %g", h[(v * 128 + i) * 16 + 1]);
        printf("\n");
    }
    return 0;
}
