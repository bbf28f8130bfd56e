// lz_native.cu -- kernels on the native HBM layout: one game state = 4 x uint64 packed bitboards (32 B),
// loaded / stored as two 16-byte vectors per thread (fully coalesced: 32 lanes x 32 B = 8 sectors x 128 B).
// One THREAD per game here (the whole state fits in 5 registers pairs), so a warp advances 32 games.
#include "lz_common.cuh"

using namespace lz;

namespace lzb {
namespace {

__device__ __forceinline__ Packed load_packed(const uint64_t* p, int64_t i) {
    const ulonglong2* v = reinterpret_cast<const ulonglong2*>(p + 4 * i);
    const ulonglong2 a = v[0], b = v[1];
    Packed r;
    r.w[0] = a.x; r.w[1] = a.y; r.w[2] = b.x; r.w[3] = b.y;
    return r;
}
__device__ __forceinline__ void store_packed(uint64_t* p, int64_t i, const Packed& s) {
    ulonglong2* v = reinterpret_cast<ulonglong2*>(p + 4 * i);
    v[0] = make_ulonglong2(s.w[0], s.w[1]);
    v[1] = make_ulonglong2(s.w[2], s.w[3]);
}

// byte layout -> packed: one warp per state (ballot build), lane 0 stores
__global__ void __launch_bounds__(kThreads)
pack_kernel(lzb_states_in st, int64_t B, uint64_t* __restrict__ packed) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t b = warp; b < B; b += nwarps) {
        State<int> s;
        int8_t blo, bhi;
        warp_load_state<int>(st, b, lane, s, blo, bhi);
        if (lane == 0) store_packed(packed, b, pack(s));
    }
}

// packed -> byte layout: one warp per state, coalesced byte stores
__global__ void __launch_bounds__(kThreads)
unpack_kernel(const uint64_t* __restrict__ packed, int64_t B, lzb_states_out out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t b = warp; b < B; b += nwarps) {
        State<int> s;
        unpack(load_packed(packed, b), s);
        for (int cell = lane; cell < 36; cell += 32) {
            const uint64_t m = 1ULL << cell;
            out.board[b * 36 + cell] = (s.black & m) ? (int8_t)1 : (s.white & m) ? (int8_t)-1 : (int8_t)0;
            out.marks_black[b * 36 + cell] = (uint8_t)((s.mb >> cell) & 1);
            out.marks_white[b * 36 + cell] = (uint8_t)((s.mw >> cell) & 1);
        }
        if (lane == 0) {
            out.phase[b] = s.phase; out.current_player[b] = s.player;
            out.pending_marks_required[b] = s.pm_req; out.pending_marks_remaining[b] = s.pm_rem;
            out.pending_captures_required[b] = s.pc_req; out.pending_captures_remaining[b] = s.pc_rem;
            out.forced_removals_done[b] = s.forced;
            if (out.move_count) out.move_count[b] = s.move_count;
            if (out.moves_since_capture) out.moves_since_capture[b] = s.msc;
        }
    }
}

__global__ void __launch_bounds__(kThreads)
init_kernel(uint64_t* __restrict__ packed, int64_t B) {
    State<int> s;
    set_initial(s);
    const Packed p = pack(s);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
        store_packed(packed, i, p);
}

template <bool kScalar>
__global__ void __launch_bounds__(kThreads)
legal_masks_kernel(const uint64_t* __restrict__ packed, int64_t B, uint64_t* __restrict__ mask_words,
                   int32_t* __restrict__ counts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        State<int> s;
        unpack(load_packed(packed, i), s);
        Legal L;
        legal_actions<int, kScalar>(s, L, true);
        uint64_t w[4];
        legal_to_words(L, w);
        Packed o; o.w[0] = w[0]; o.w[1] = w[1]; o.w[2] = w[2]; o.w[3] = w[3];
        store_packed(mask_words, i, o);
        if (counts) counts[i] = legal_count(L);
    }
}

__global__ void __launch_bounds__(kThreads)
apply_actions_kernel(const uint64_t* __restrict__ parents, const int64_t* __restrict__ parent_indices,
                     const int32_t* __restrict__ actions, int64_t N, uint64_t* __restrict__ children) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = parent_indices ? parent_indices[i] : i;
        State<int> s;
        unpack(load_packed(parents, p), s);
        const int a = actions[i];
        if (a >= 0 && a < kActionDim) apply_index(s, a);
        store_packed(children, i, pack(s));
    }
}

// Config-2 workload: every thread owns one game and plays up to max_steps plies with the state in
// registers: legal set (scalar-engine semantics) -> counter-based uniform pick -> apply -> terminal check.
template <bool kHash>
__global__ void __launch_bounds__(kThreads)
playout_kernel(uint64_t* __restrict__ packed, int32_t* __restrict__ plies, int8_t* __restrict__ result,
               uint64_t* __restrict__ hash, int64_t B, uint64_t seed, uint64_t game_offset, int max_steps,
               int max_game_plies) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        int res = result[i];
        if (res != 2) continue;
        State<int> s;
        unpack(load_packed(packed, i), s);
        int ply = plies[i];
        uint64_t h = kHash ? hash[i] : 0;
        playout_advance<kHash>(s, ply, res, h, seed, game_offset + (uint64_t)i, max_steps, max_game_plies);
        store_packed(packed, i, pack(s));
        plies[i] = ply;
        result[i] = (int8_t)res;
        if (kHash) hash[i] = h;
    }
}

}  // namespace
}  // namespace lzb

using namespace lzb;

extern "C" int lzb_pack_states(const lzb_states_in* st, int64_t B, uint64_t* packed, void* stream) {
    LZB_REQUIRE(st && B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(packed && (reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed must be 16-byte aligned");
    pack_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(*st, B, packed);
    return check_launch("pack_kernel");
}

extern "C" int lzb_unpack_states(const uint64_t* packed, int64_t B, const lzb_states_out* st, void* stream) {
    LZB_REQUIRE(st && B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(packed && (reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed must be 16-byte aligned");
    unpack_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(packed, B, *st);
    return check_launch("unpack_kernel");
}

extern "C" int lzb_init_states(uint64_t* packed, int64_t B, void* stream) {
    LZB_REQUIRE(B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(packed && (reinterpret_cast<uintptr_t>(packed) & 15) == 0, "packed must be 16-byte aligned");
    init_kernel<<<thread_grid(B), kThreads, 0, (cudaStream_t)stream>>>(packed, B);
    return check_launch("init_kernel");
}

extern "C" int lzb_legal_masks_packed(const uint64_t* packed, int64_t B, int scalar_semantics, uint64_t* mask_words,
                                      int32_t* counts, void* stream) {
    LZB_REQUIRE(B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(packed && mask_words, "null pointer");
    if (scalar_semantics)
        legal_masks_kernel<true><<<thread_grid(B), kThreads, 0, (cudaStream_t)stream>>>(packed, B, mask_words, counts);
    else
        legal_masks_kernel<false><<<thread_grid(B), kThreads, 0, (cudaStream_t)stream>>>(packed, B, mask_words, counts);
    return check_launch("legal_masks_kernel");
}

extern "C" int lzb_apply_actions_packed(const uint64_t* parents, const int64_t* parent_indices, const int32_t* actions,
                                        int64_t N, uint64_t* children, void* stream) {
    LZB_REQUIRE(N >= 0, "bad arguments");
    if (N == 0) return LZB_OK;
    LZB_REQUIRE(parents && actions && children, "null pointer");
    apply_actions_kernel<<<thread_grid(N), kThreads, 0, (cudaStream_t)stream>>>(parents, parent_indices, actions, N,
                                                                               children);
    return check_launch("apply_actions_kernel");
}

extern "C" int lzb_playout_run(uint64_t* packed, int32_t* plies, int8_t* result, uint64_t* hash, int64_t B,
                               uint64_t seed, uint64_t game_offset, int32_t max_steps, int32_t max_game_plies,
                               void* stream) {
    LZB_REQUIRE(B >= 0 && max_steps >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(packed && plies && result, "null pointer");
    // 128 threads/block: one game per thread is latency-bound integer work; smaller blocks spread the
    // 65,536 games of config 2 over all 148 SMs (512 blocks = 3.46 per SM).
    const int threads = 128;
    const int grid = thread_grid(B, threads);
    if (hash)
        playout_kernel<true><<<grid, threads, 0, (cudaStream_t)stream>>>(packed, plies, result, hash, B, seed,
                                                                       game_offset, max_steps, max_game_plies);
    else
        playout_kernel<false><<<grid, threads, 0, (cudaStream_t)stream>>>(packed, plies, result, hash, B, seed,
                                                                        game_offset, max_steps, max_game_plies);
    return check_launch("playout_kernel");
}
