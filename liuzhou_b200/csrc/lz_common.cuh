// lz_common.cuh -- launch plumbing shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/liuzhou_b200.h"
#include "lz_rules.cuh"

namespace lzb {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

inline int check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(err));
        return LZB_ERR_CUDA;
    }
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return LZB_OK;
}

#define LZB_REQUIRE(cond, msg)                      \
    do {                                            \
        if (!(cond)) {                              \
            lzb::set_error("%s: %s", __func__, msg); \
            return LZB_ERR_INVALID_ARGUMENT;        \
        }                                           \
    } while (0)

constexpr int kWarpsPerBlock = 8;                 // 256 threads
constexpr int kThreads = kWarpsPerBlock * 32;
constexpr int kNumSMs = 148;                      // B200

// Grid for "one warp per item" kernels: enough blocks to cover the items, capped at a multiple of the SM
// count (8 resident 256-thread blocks per SM); kernels grid-stride over the remainder.
inline int warp_grid(int64_t items) {
    int64_t blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}
inline int thread_grid(int64_t items, int threads = kThreads) {
    int64_t blocks = (items + threads - 1) / threads;
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// ---- device helpers -----------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t ballot36(bool lo_pred, bool hi_pred) {
    // lo_pred: predicate for cell == lane (0..31); hi_pred: predicate for cell == 32 + lane (lanes 0..3)
    const uint32_t lo = __ballot_sync(0xffffffffu, lo_pred);
    const uint32_t hi = __ballot_sync(0xffffffffu, hi_pred) & 0xFu;
    return (uint64_t)lo | ((uint64_t)hi << 32);
}

using lz::legal_to_words;
using lz::legal_kth_words;

// The same load split in two, so that a kernel can have the NEXT state's loads in flight while it works on (and stores
// the wide outputs of) the current one: under saturating write traffic a dependent load costs several thousand cycles.
template <typename I>
struct RawState {
    int8_t blo, bhi;
    uint8_t mblo, mbhi, mwlo, mwhi;
    I phase, player, pm_req, pm_rem, pc_req, pc_rem, forced, move_count, msc;
};
template <typename I>
__device__ __forceinline__ void warp_fetch_state(const lzb_states_in& st, int64_t b, int lane, RawState<I>& r) {
    const int8_t* bp = st.board + b * 36;
    const uint8_t* mbp = st.marks_black + b * 36;
    const uint8_t* mwp = st.marks_white + b * 36;
    const bool hi = lane < 4;
    r.blo = bp[lane];
    r.bhi = hi ? bp[32 + lane] : (int8_t)0;
    r.mblo = mbp[lane]; r.mwlo = mwp[lane];
    r.mbhi = hi ? mbp[32 + lane] : (uint8_t)0; r.mwhi = hi ? mwp[32 + lane] : (uint8_t)0;
    r.phase = (I)st.phase[b];
    r.player = (I)st.current_player[b];
    r.pm_req = (I)st.pending_marks_required[b];
    r.pm_rem = (I)st.pending_marks_remaining[b];
    r.pc_req = (I)st.pending_captures_required[b];
    r.pc_rem = (I)st.pending_captures_remaining[b];
    r.forced = (I)st.forced_removals_done[b];
    r.move_count = st.move_count ? (I)st.move_count[b] : (I)0;
    r.msc = st.moves_since_capture ? (I)st.moves_since_capture[b] : (I)0;
}
template <typename I>
__device__ __forceinline__ void warp_build_state(const RawState<I>& r, int lane, lz::State<I>& s) {
    const bool hi = lane < 4;
    s.black = ballot36(r.blo == 1, hi && r.bhi == 1);
    s.white = ballot36(r.blo == -1, hi && r.bhi == -1);
    s.other = ballot36(r.blo != 0 && r.blo != 1 && r.blo != -1, hi && r.bhi != 0 && r.bhi != 1 && r.bhi != -1);
    s.mb = ballot36(r.mblo != 0, hi && r.mbhi != 0);
    s.mw = ballot36(r.mwlo != 0, hi && r.mwhi != 0);
    s.phase = r.phase; s.player = r.player; s.pm_req = r.pm_req; s.pm_rem = r.pm_rem; s.pc_req = r.pc_req;
    s.pc_rem = r.pc_rem; s.forced = r.forced; s.move_count = r.move_count; s.msc = r.msc;
}

// Load one state in the reference byte layout into bitboards, one warp per state.
template <typename I>
__device__ __forceinline__ void warp_load_state(const lzb_states_in& st, int64_t b, int lane, lz::State<I>& s,
                                                int8_t& byte_lo, int8_t& byte_hi) {
    const int8_t* bp = st.board + b * 36;
    const uint8_t* mbp = st.marks_black + b * 36;
    const uint8_t* mwp = st.marks_white + b * 36;
    const bool hi = lane < 4;
    byte_lo = bp[lane];
    byte_hi = hi ? bp[32 + lane] : (int8_t)0;
    const uint8_t mb_lo = mbp[lane], mw_lo = mwp[lane];
    const uint8_t mb_hi = hi ? mbp[32 + lane] : (uint8_t)0, mw_hi = hi ? mwp[32 + lane] : (uint8_t)0;
    s.black = ballot36(byte_lo == 1, hi && byte_hi == 1);
    s.white = ballot36(byte_lo == -1, hi && byte_hi == -1);
    s.other = ballot36(byte_lo != 0 && byte_lo != 1 && byte_lo != -1,
                       hi && byte_hi != 0 && byte_hi != 1 && byte_hi != -1);
    s.mb = ballot36(mb_lo != 0, hi && mb_hi != 0);
    s.mw = ballot36(mw_lo != 0, hi && mw_hi != 0);
    s.phase = (I)st.phase[b];
    s.player = (I)st.current_player[b];
    s.pm_req = (I)st.pending_marks_required[b];
    s.pm_rem = (I)st.pending_marks_remaining[b];
    s.pc_req = (I)st.pending_captures_required[b];
    s.pc_rem = (I)st.pending_captures_remaining[b];
    s.forced = (I)st.forced_removals_done[b];
    s.move_count = st.move_count ? (I)st.move_count[b] : (I)0;
    s.msc = st.moves_since_capture ? (I)st.moves_since_capture[b] : (I)0;
}
#endif

}  // namespace lzb
