// lz_trunk.cu -- the WHOLE convolutional trunk of ChessNet (stem conv + 2 x blocks 3x3 convs + the heads' 1x1 conv,
// src/neural_network.py:83-96,213-259) as ONE persistent tcgen05 kernel: activations never leave the SM.
//
// Why this is possible: a convolution only mixes pixels of the same board, and a CTA tile here is 3 WHOLE boards (6 per
// CTA pair), so layer l + 1 of a tile depends on nothing but layer l of the same tile -- no grid-wide synchronisation
// between layers.  The per-layer kernels (lz_conv.cu) spend more time moving activations (L2 -> SM operand loads, the
// epilogue's global stores / residual loads, DRAM spill of the 113 MB working set) than in the tensor core
// (profiles/r02_conv_decompose_*.txt, profiles/r02_conv_pad_ncu_full.csv); here a tile's activations stay in shared
// memory / TMEM through all 22 layers and only the weights stream (L2-resident, 5.9 MB for the whole net).
//
// Data flow of one tile (CTA pair, cta_group::2: M = 256 = 2 x 128 accumulator rows, N = 128, K = 16 per MMA):
//   * A operand = the previous layer's activations as bf16 "padded boards" in shared memory: row(b, y, x) = b * 42 +
//     (y + 1) * 6 + x (a zero row above every board), three copies pre-shifted in x (dx = -1, 0, +1, zero column where
//     the neighbour is off the board); a tap (dy, dx) is an MMA whose A descriptor starts 6 (dy + 1) rows into copy dx
//     (the tensor core swizzles on absolute address bits, so a descriptor may start on any 128-byte row:
//     tools/probe/umma_shift_probe.cu).  The stem's copies come from global memory by TMA (zeros out of bounds); every
//     other layer's copies are written by the epilogue of the layer before it.
//   * accumulators in TMEM: acc_h (stem, conv1, heads conv) and acc_x = the RESIDUAL STREAM in fp32: conv2's MMAs
//     accumulate straight onto it (x' = x + conv2(h) is done by the tensor core, never rounded to bf16).
//   * epilogue warps (8): TMEM -> registers -> bias / BatchNorm / ReLU -> bf16 -> the three shifted copies in shared
//     memory (swizzled 16-byte stores), 64 channels at a time: the next layer's MMAs on channels 0-63 start while the
//     epilogue still produces channels 64-127 (K-chunk pipelining).
//   * weights: one (tap, 64-channel chunk) slice (8 KB per CTA) per pipeline stage, streamed by TMA through a 12-stage
//     ring across layer and tile boundaries.
// Per layer and tile: 72 MMAs (36 for the stem, 8 for the 1x1 heads conv); junk accumulator rows (the pad row between
// boards, 20 of 128) are the price of serving all taps from one resident tile.
#include <cuda.h>
#include <stdlib.h>

#include "lz_common.cuh"
#include "lz_tc.cuh"

namespace lzb {
namespace {
using namespace tc;

constexpr int kTrunkThreads = 320;                     // warp 0 producer, warp 1 MMA, warps 2-9 epilogue
constexpr int kBoards = 3;                             // boards per CTA tile
constexpr int kBoardRows = 42;                         // padded rows per board
constexpr int kBoxRows = kBoards * kBoardRows;         // 126
constexpr int kBufRows = 136;                          // rows per chunk buffer (126 + 6 zero + 4 spare)
constexpr uint32_t kChunkBytes = kBufRows * 128;       // 17,408: [rows][64 ch] bf16, SWIZZLE_128B
constexpr uint32_t kBoxBytes = kBoxRows * 128;         // 16,128 per TMA box
constexpr uint32_t kCopyBytes = 2 * kChunkBytes;       // one dx copy, 128 channels
constexpr uint32_t kWSliceBytes = 64 * 64 * 2;         // 8 KB: 64 couts (this CTA's half) x 64 cin of one tap
constexpr int kWStages = 12;
constexpr uint32_t kAccH = 0, kAccX = 128;             // TMEM column offsets of the two accumulators
constexpr uint32_t kTrunkTmemCols = 256;

struct TrunkParams {
    const float* params;       // [layers][3][128] : bias | scale | shift
    __nv_bfloat16* out;        // [images * 36][128] : relu(heads conv + bias), the input of heads_tail_kernel
    int images;
    int blocks;                // residual blocks; layers = 2 * blocks + 2
    int debug;
};

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // the 8 epilogue warps

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTrunkThreads, 1)
trunk_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW0,
             const __grid_constant__ CUtensorMap tmW1, const TrunkParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base;                                    // 3 copies x 2 chunks
    const uint32_t w_smem = a_smem + 3 * kCopyBytes;                 // weight ring
    const uint32_t bar0 = w_smem + kWStages * kWSliceBytes;
    const uint32_t in_full = bar0;                                   // leader: stem input copies landed
    const uint32_t in_empty = in_full + 8;                           // both  : the tile's last MMAs are done
    const uint32_t ready_bar = in_empty + 8;                         // [2] leader: chunk kc of the copies written (16 arrivals)
    const uint32_t wfull_bar = ready_bar + 16;                       // [kWStages] leader
    const uint32_t wempty_bar = wfull_bar + 8 * kWStages;            // [kWStages] both
    const uint32_t accfull_bar = wempty_bar + 8 * kWStages;          // both  : the layer's accumulator is complete
    const uint32_t epidone_bar = accfull_bar + 8;                    // leader: last layer's accumulator drained (16 arrivals)
    const uint32_t tmem_slot = epidone_bar + 8;
    const uint32_t vec_smem = (tmem_slot + 16 + 15u) & ~15u;         // [2][3][128] f32 layer parameters, double buffered
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    float* vec = reinterpret_cast<float*>(gen + (vec_smem - base));
    volatile uint32_t* tmem_slot_p = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int64_t pair_tiles = (P.images + 2 * kBoards - 1) / (2 * kBoards);
    const int layers = 2 * P.blocks + 2;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmIn); prefetch_tmap(&tmW0); prefetch_tmap(&tmW1);
        mbar_init(in_full, 1); mbar_init(in_empty, 1);
        mbar_init(ready_bar, 16); mbar_init(ready_bar + 8, 16);
        for (int i = 0; i < kWStages; ++i) { mbar_init(wfull_bar + 8 * i, 1); mbar_init(wempty_bar + 8 * i, 1); }
        mbar_init(accfull_bar, 1); mbar_init(epidone_bar, 16);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTrunkTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // the copies start out all zero: pad rows / columns are never written afterwards (or are rewritten with zeros)
    for (uint32_t i = threadIdx.x; i < 3 * kCopyBytes / 16; i += kTrunkThreads)
        *reinterpret_cast<uint4*>(gen + (a_smem - base) + i * 16) = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_p;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == 0) {
        // ===================== TMA producer (both CTAs): stem input copies + every weight slice, in MMA order ===========
        const uint32_t in_full_leader = mapa_rank(in_full, 0), wfull_leader = mapa_rank(wfull_bar, 0);
        uint32_t wst = 0, wph = 0, tile_ph = 0;
        for (int64_t pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
            const int board0 = (int)(2 * pt + rank) * kBoards;
            mbar_wait(in_empty, tile_ph ^ 1);                       // the previous tile no longer reads the copies
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(in_full, 2 * 3 * kBoxBytes);
#pragma unroll
                for (int dxi = 0; dxi < 3; ++dxi)                   // planes (64-channel padded), chunk 0 of copy dx
                    tma_tile4d_2sm(a_smem + dxi * kCopyBytes, &tmIn, in_full_leader, 0, dxi - 1, -1, board0);
            }
            __syncwarp();
            tile_ph ^= 1;
            for (int l = 0; l < layers; ++l) {
                const int taps = l == layers - 1 ? 1 : 9, kch = l == 0 ? 1 : 2;
                const int tap_base = l == 0 ? 0 : (l - 1) * 9;       // index of the layer's first tap in its weight tensor
                for (int kc = 0; kc < kch; ++kc) {
                    for (int t = 0; t < taps; ++t) {
                        // MMA order within a chunk: dx-major, dy-minor; tap index in the weight tensor = (dy+1)*3 + (dx+1)
                        const int tap = taps == 9 ? (t % 3) * 3 + (t / 3) : 0;
                        mbar_wait(wempty_bar + 8 * wst, wph ^ 1);
                        if (elect_one()) {
                            if (leader) mbar_arrive_expect_tx(wfull_bar + 8 * wst, 2 * kWSliceBytes);
                            if (l == 0)
                                tma_tile2d_2sm(w_smem + wst * kWSliceBytes, &tmW0, wfull_leader + 8 * wst, 0,
                                               tap * 128 + (int)rank * 64);
                            else
                                tma_tile2d_2sm(w_smem + wst * kWSliceBytes, &tmW1, wfull_leader + 8 * wst, kc * 64,
                                               (tap_base + tap) * 128 + (int)rank * 64);
                        }
                        __syncwarp();
                        if (++wst == kWStages) { wst = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA) =====================
        if (leader) {
            uint32_t wst = 0, wph = 0, tile_ph = 0, ready_ph = 0;
            for (int64_t pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
                mbar_wait_cluster(epidone_bar, tile_ph ^ 1);         // acc_h of the previous tile has been drained
                mbar_wait(in_full, tile_ph);                         // stem input copies are in shared memory
                tc_fence_after();
                for (int l = 0; l < layers; ++l) {
                    const int taps = l == layers - 1 ? 1 : 9, kch = l == 0 ? 1 : 2;
                    const bool conv2 = l >= 2 && l < layers - 1 && (l & 1) == 0;     // accumulates onto the residual stream
                    const uint32_t d_tmem = tmem_base + (conv2 ? kAccX : kAccH);
                    for (int kc = 0; kc < kch; ++kc) {
                        if (l > 0) {                                 // chunk kc of the copies written by layer l - 1's epilogue
                            mbar_wait_cluster(ready_bar + 8 * kc, ready_ph);
                            tc_fence_after();
                        }
                        for (int t = 0; t < taps; ++t) {
                            const int dxi = taps == 9 ? t / 3 : 1, dyi = taps == 9 ? t % 3 : 1;
                            mbar_wait(wfull_bar + 8 * wst, wph);
                            tc_fence_after();
                            if (elect_one()) {
                                const uint64_t adesc = umma_desc_sw128(a_smem + dxi * kCopyBytes + kc * kChunkBytes +
                                                                       (uint32_t)(dyi * 6) * 128u);
                                const uint64_t bdesc = umma_desc_sw128(w_smem + wst * kWSliceBytes);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    if (!(P.debug & 4))
                                        umma_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, (uint32_t)(conv2 || (kc | t | k) != 0));
                                umma_commit_2sm(wempty_bar + 8 * wst, 3);
                                if (kc == kch - 1 && t == taps - 1) {
                                    umma_commit_2sm(accfull_bar, 3);                 // the layer's accumulator is complete
                                    if (l == layers - 1) umma_commit_2sm(in_empty, 3);   // ... and the copies are free
                                }
                            }
                            __syncwarp();
                            if (++wst == kWStages) { wst = 0; wph ^= 1; }
                        }
                    }
                    if (l > 0) ready_ph ^= 1;
                }
                tile_ph ^= 1;
            }
        }
    } else {
        // ===================== epilogue: warps 2..9 =====================
        // thread <-> accumulator row i = q * 32 + lane (TMEM lane quarter q = warp % 4), 32 of the 64 columns of the
        // current chunk (col_half = (warp - 2) / 4).  Row i = b * 42 + y * 6 + x; rows with i % 42 >= 36 or b = 3 are junk.
        const int q = warp & 3, col_half = (warp - 2) >> 2;
        const int epi_tid = threadIdx.x - 64;
        const int i_row = q * 32 + lane;
        const int b = i_row / kBoardRows, rem = i_row - b * kBoardRows;
        const bool row_ok = b < kBoards && rem < 36;
        const int y = rem / 6, x = rem - y * 6;
        // shared-memory targets of this pixel in the three copies: copy dx holds pixel (y, x' + dx) at (y, x'); the column
        // x' that has no source (its neighbour is off the board) is written with zeros by the thread that wraps onto it
        uint32_t dst[3], swz[3];
        bool zero[3];
#pragma unroll
        for (int dxi = 0; dxi < 3; ++dxi) {
            const int xs = x - (dxi - 1);
            zero[dxi] = xs < 0 || xs > 5;
            const int xp = (xs + 6) % 6;
            const int rho = b * kBoardRows + (y + 1) * 6 + xp;
            dst[dxi] = a_smem + dxi * kCopyBytes + (uint32_t)rho * 128u;
            swz[dxi] = (uint32_t)(rho & 7);
        }
        const uint32_t ready_leader = mapa_rank(ready_bar, 0), epidone_leader = mapa_rank(epidone_bar, 0);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t acc_ph = 0;
        // layer parameters: bias | scale | shift, 3 x 128 f32 per layer, double buffered in shared memory
        for (int i = epi_tid; i < 384; i += 256) vec[i] = P.params[i];
        for (int64_t pt = cluster_id; pt < pair_tiles; pt += num_clusters) {
            const int board0 = (int)(2 * pt + rank) * kBoards;
            for (int l = 0; l < layers; ++l) {
                epi_bar_sync();                                        // everyone is done with layer l - 1: its buffer is free
                {
                    const int ln = l + 1 < layers ? l + 1 : 0;         // prefetch the next layer's parameters
                    float* vn = vec + ((l + 1) & 1) * 384;
                    for (int i = epi_tid; i < 384; i += 256) vn[i] = P.params[ln * 384 + i];
                }
                const float* vp = vec + (l & 1) * 384;
                const bool last = l == layers - 1;
                const bool conv2 = l >= 2 && !last && (l & 1) == 0;
                const bool stem = l == 0;
                const uint32_t acc_col = conv2 ? kAccX : kAccH;
                const int ncopies_lo = l + 1 == layers - 1 ? 1 : 0, ncopies_hi = l + 1 == layers - 1 ? 2 : 3;   // 1x1 next: copy 0 only
                mbar_wait(accfull_bar, acc_ph);
                acc_ph ^= 1;
                tc_fence_after();
#pragma unroll 1
                for (int kc = 0; kc < 2; ++kc) {
                    const int c0 = kc * 64 + col_half * 32;            // first of this thread's 32 output channels
                    uint32_t v[32];
                    tmem_ld32(lane_addr + acc_col + (uint32_t)c0, v);
                    tmem_ld_wait();
                    uint32_t pk[16];
                    const float4* vb = reinterpret_cast<const float4*>(vp + c0);            // bias  (broadcast reads)
                    const float4* vs = reinterpret_cast<const float4*>(vp + 128 + c0);      // scale
                    const float4* vt = reinterpret_cast<const float4*>(vp + 256 + c0);      // shift
                    if (stem) {
                        // x0 = relu(conv + bias) -> residual stream (fp32, TMEM); a0 = relu(scale * x0 + shift) -> copies
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = vb[j];
                            v[4 * j] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.0f));
                            v[4 * j + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.0f));
                            v[4 * j + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.0f));
                            v[4 * j + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.0f));
                        }
                        tmem_st32(lane_addr + kAccX + (uint32_t)c0, v);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 ss = vs[j], tt = vt[j];
                            pk[2 * j] = pack_bf16(fmaxf(fmaf(ss.x, __uint_as_float(v[4 * j]), tt.x), 0.0f),
                                                  fmaxf(fmaf(ss.y, __uint_as_float(v[4 * j + 1]), tt.y), 0.0f));
                            pk[2 * j + 1] = pack_bf16(fmaxf(fmaf(ss.z, __uint_as_float(v[4 * j + 2]), tt.z), 0.0f),
                                                      fmaxf(fmaf(ss.w, __uint_as_float(v[4 * j + 3]), tt.w), 0.0f));
                        }
                        tmem_st_wait();
                    } else if (conv2) {
                        // the accumulator IS x' = x + conv2(h); a' = relu(scale * x' + shift)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 ss = vs[j], tt = vt[j];
                            pk[2 * j] = pack_bf16(fmaxf(fmaf(ss.x, __uint_as_float(v[4 * j]), tt.x), 0.0f),
                                                  fmaxf(fmaf(ss.y, __uint_as_float(v[4 * j + 1]), tt.y), 0.0f));
                            pk[2 * j + 1] = pack_bf16(fmaxf(fmaf(ss.z, __uint_as_float(v[4 * j + 2]), tt.z), 0.0f),
                                                      fmaxf(fmaf(ss.w, __uint_as_float(v[4 * j + 3]), tt.w), 0.0f));
                        }
                    } else {
                        // conv1 / heads conv: relu(acc + bias)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 bb = vb[j];
                            pk[2 * j] = pack_bf16(fmaxf(__uint_as_float(v[4 * j]) + bb.x, 0.0f),
                                                  fmaxf(__uint_as_float(v[4 * j + 1]) + bb.y, 0.0f));
                            pk[2 * j + 1] = pack_bf16(fmaxf(__uint_as_float(v[4 * j + 2]) + bb.z, 0.0f),
                                                      fmaxf(__uint_as_float(v[4 * j + 3]) + bb.w, 0.0f));
                        }
                    }
                    if (last) {
                        if (row_ok && board0 + b < P.images && !(P.debug & 1)) {
                            uint4* o = reinterpret_cast<uint4*>(P.out + ((int64_t)(board0 + b) * 36 + rem) * 128 + c0);
#pragma unroll
                            for (int t = 0; t < 4; ++t) o[t] = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
                        }
                    } else {
                        if (row_ok) {
#pragma unroll
                            for (int dxi = 0; dxi < 3; ++dxi) {
                                if (dxi < ncopies_lo || dxi >= ncopies_hi) continue;
                                const uint32_t rowaddr = dst[dxi] + (uint32_t)kc * kChunkBytes;
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const uint32_t c16 = (uint32_t)(col_half * 4 + t) ^ swz[dxi];
                                    if (zero[dxi]) st_shared_v4(rowaddr + c16 * 16, 0u, 0u, 0u, 0u);
                                    else st_shared_v4(rowaddr + c16 * 16, pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
                                }
                            }
                        }
                        fence_proxy_async();          // generic-proxy stores -> visible to the tensor core (async proxy)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_release_cluster(ready_leader + 8 * kc);
                    }
                }
                if (last) {                           // acc_h has been read: the next tile's stem may overwrite it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_release_cluster(epidone_leader);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTrunkTmemCols) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace
}  // namespace lzb

// planes bf16 [n,6,6,64] (channel-padded input), w_stem bf16 [9][128][64], w_trunk bf16 [2*blocks*9 + 1][128][128]
// (conv1_0, conv2_0, ..., conv2_{blocks-1}, heads 1x1; BatchNorm folded where it follows a conv), params f32
// [2*blocks+2][3][128] (bias | scale | shift per layer), out bf16 [n,6,6,128] = relu(heads conv + bias).
extern "C" int lzb_trunk_bf16(const void* planes, int64_t n, const void* w_stem, const void* w_trunk, const float* params,
                              int32_t blocks, void* out, void* stream) {
    using namespace lzb;
    LZB_REQUIRE(n > 0 && n < (1ll << 30), "bad batch size");
    LZB_REQUIRE(blocks >= 1 && blocks <= 64, "blocks must be in [1, 64]");
    LZB_REQUIRE(planes && w_stem && w_trunk && params && out, "null pointer");
    LZB_REQUIRE(((reinterpret_cast<uintptr_t>(planes) | reinterpret_cast<uintptr_t>(w_stem) | reinterpret_cast<uintptr_t>(w_trunk) |
                  reinterpret_cast<uintptr_t>(out)) & 15) == 0, "pointers must be 16-byte aligned");
    static EncodeTiledFn encode_tiled = nullptr;
    if (!encode_tiled) {
        cudaDriverEntryPointQueryResult st;
        void* fn = nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st) != cudaSuccess ||
            st != cudaDriverEntryPointSuccess || !fn) {
            set_error("lzb_trunk: cuTensorMapEncodeTiled unavailable");
            return LZB_ERR_CUDA;
        }
        encode_tiled = reinterpret_cast<EncodeTiledFn>(fn);
    }
    alignas(64) CUtensorMap tmIn, tmW0, tmW1;
    {
        const cuuint64_t dim[4] = {64, 6, 6, (cuuint64_t)n};
        const cuuint64_t stride[3] = {128, 128 * 6, 128 * 36};
        const cuuint32_t box[4] = {64, 6, 7, (cuuint32_t)kBoards};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult rc = encode_tiled(&tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(planes), dim, stride, box,
                                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("lzb_trunk: tensor map (planes) failed (%d)", (int)rc); return LZB_ERR_CUDA; }
    }
    for (int which = 0; which < 2; ++which) {
        const cuuint64_t cin = which == 0 ? 64 : 128;
        const cuuint64_t rows = which == 0 ? 9 * 128 : ((cuuint64_t)blocks * 18 + 1) * 128;
        const cuuint64_t dim[2] = {cin, rows};
        const cuuint64_t stride[1] = {cin * 2};
        const cuuint32_t box[2] = {64, 64};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult rc = encode_tiled(which == 0 ? &tmW0 : &tmW1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                                         const_cast<void*>(which == 0 ? w_stem : w_trunk), dim, stride, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) { set_error("lzb_trunk: tensor map (weights %d) failed (%d)", which, (int)rc); return LZB_ERR_CUDA; }
    }
    constexpr size_t smem = 1024 + 3 * (size_t)kCopyBytes + (size_t)kWStages * kWSliceBytes + 8 * (6 + 2 * kWStages) + 32 +
                            2 * 384 * sizeof(float) + 256;
    static int sm_count[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { set_error("lzb_trunk: bad current device"); return LZB_ERR_CUDA; }
    if (sm_count[dev] == 0) {
        if (cudaFuncSetAttribute(trunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("lzb_trunk: cannot raise dynamic shared memory to %zu", smem);
            return LZB_ERR_CUDA;
        }
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 2) sms = kNumSMs;
        sm_count[dev] = sms;
    }
    TrunkParams P;
    P.params = params; P.out = reinterpret_cast<__nv_bfloat16*>(out); P.images = (int)n; P.blocks = blocks;
    static const int debug = getenv("LZB_TRUNK_DEBUG") ? atoi(getenv("LZB_TRUNK_DEBUG")) : 0;
    P.debug = debug;
    const int64_t pair_tiles = (n + 2 * kBoards - 1) / (2 * kBoards);
    const int clusters = (int)(pair_tiles < sm_count[dev] / 2 ? pair_tiles : sm_count[dev] / 2);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kTrunkThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, trunk_kernel, tmIn, tmW0, tmW1, P) != cudaSuccess) return check_launch("trunk_kernel");
    return check_launch("trunk_kernel");
}
