// lz_dropin.cu -- sm_100a kernels behind the reference's v0_core tensor-op surface (reference byte layout).
//
// Design: one warp per state / action.  The 36 board bytes and 2x36 mark bytes are read by the 32 lanes
// (cells lane and 32+lane) and turned into bitboards with warp ballots, so every lane holds the whole
// state in registers; the rule logic (lz_rules.cuh) then runs warp-uniform without divergence, and the
// wide outputs (3,520 B of metadata per state, 180 B per child) are written with fully coalesced
// 16 B / 1 B per-lane stores.  These ops are HBM-bound byte work: 3,888 B/state for the legal mask,
// 384 B/action for apply.
#include "lz_common.cuh"

using namespace lz;

namespace lzb {
namespace {

// ------------------------------------------------------------------------------------------------------
// (a2) encode_actions_fast  -- fast_legal_mask_cuda.cu:282-404
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
encode_actions_kernel(lzb_states_in st, int64_t B, int pd, int md, int sd, int ad, uint8_t* __restrict__ mask,
                      int32_t* __restrict__ metadata) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int total = pd + md + sd + ad;
    const bool standard = pd == 36 && md == 144 && sd == 36 && ad == 4;
    RawState<int> nxt;
    if (warp < B) warp_fetch_state<int>(st, warp, lane, nxt);
    for (int64_t b = warp; b < B; b += nwarps) {
        const RawState<int> cur = nxt;
        if (b + nwarps < B) warp_fetch_state<int>(st, b + nwarps, lane, nxt);   // in flight during this state's 3.7 KB of stores
        State<int> s;
        warp_build_state<int>(cur, lane, s);
        Legal L;
        legal_actions<int, false>(s, L, ad > 0);
        uint8_t* m = mask + b * total;
        int4* meta = reinterpret_cast<int4*>(metadata) + b * total;
        if (standard) {
            // the reference's own dims (36 / 144 / 36 / 4): regions cannot overlap, 220 = 6 x 32 + 28 actions
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int a = lane + 32 * j;
                if (j == 6 && a >= 220) break;
                int4 code = make_int4(-1, -1, -1, -1);
                bool legal;
                if (a < 36) {
                    legal = (L.place >> a) & 1;
                    if (legal) code = make_int4(kActPlace, a, -1, -1);
                } else if (a < 180) {
                    const int mo = a - 36, from = mo >> 2, d = mo & 3;
                    const uint64_t mvd = d == 0 ? L.mv[0] : d == 1 ? L.mv[1] : d == 2 ? L.mv[2] : L.mv[3];
                    legal = (mvd >> from) & 1;
                    if (legal) code = make_int4(kActMove, from, d, from + (d == 0 ? -6 : d == 1 ? 6 : d == 2 ? -1 : 1));
                } else if (a < 216) {
                    legal = (L.sel >> (a - 180)) & 1;
                    if (legal) code = make_int4(L.sel_kind, a - 180, -1, -1);
                } else {
                    legal = a == 216 && L.process;
                    if (legal) code = make_int4(kActProcess, -1, -1, -1);
                }
                m[a] = legal ? 1 : 0;
                meta[a] = code;
            }
            continue;
        }
        const int sel_off = pd + md, rem_idx = pd + md + sd;
        for (int a = lane; a < total; a += 32) {
            // Region resolution in the reference's write order (placement, movement, selection, removal;
            // later writers win if a caller passes overlapping dims).
            int4 code = make_int4(-1, -1, -1, -1);
            bool legal = false;
            if (a < 36 && ((L.place >> a) & 1)) { legal = true; code = make_int4(kActPlace, a, -1, -1); }
            const int mo = a - pd;
            if (mo >= 0 && mo < 144 && mo < md) {
                const int from = mo >> 2, d = mo & 3;
                if ((L.mv[d] >> from) & 1) {
                    legal = true;
                    code = make_int4(kActMove, from, d, from + (d == 0 ? -6 : d == 1 ? 6 : d == 2 ? -1 : 1));
                }
            }
            const int so = a - sel_off;
            if (so >= 0 && so < 36 && so < sd && ((L.sel >> so) & 1)) {
                legal = true; code = make_int4(L.sel_kind, so, -1, -1);
            }
            if (a == rem_idx && L.process) { legal = true; code = make_int4(kActProcess, -1, -1, -1); }
            m[a] = legal ? 1 : 0;
            meta[a] = code;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a4) batch_apply_moves / _inplace -- fast_apply_moves_cuda.cu:548-917
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_store_state(const lzb_states_out& out, int64_t i, int lane,
                                                 const State<long long>& s, int8_t blo, int8_t bhi) {
    int8_t* bp = out.board + i * 36;
    uint8_t* mbp = out.marks_black + i * 36;
    uint8_t* mwp = out.marks_white + i * 36;
    auto cell_byte = [&](int cell, int8_t orig) -> int8_t {
        const uint64_t b = 1ULL << cell;
        return (s.black & b) ? (int8_t)1 : (s.white & b) ? (int8_t)-1 : (s.other & b) ? orig : (int8_t)0;
    };
    bp[lane] = cell_byte(lane, blo);
    mbp[lane] = (uint8_t)((s.mb >> lane) & 1);
    mwp[lane] = (uint8_t)((s.mw >> lane) & 1);
    if (lane < 4) {
        bp[32 + lane] = cell_byte(32 + lane, bhi);
        mbp[32 + lane] = (uint8_t)((s.mb >> (32 + lane)) & 1);
        mwp[32 + lane] = (uint8_t)((s.mw >> (32 + lane)) & 1);
    }
    // nine int64 scalars: lane l writes scalar l
    long long v = s.phase;
    int64_t* dst = out.phase;
    switch (lane) {
        case 1: v = s.player; dst = out.current_player; break;
        case 2: v = s.pm_req; dst = out.pending_marks_required; break;
        case 3: v = s.pm_rem; dst = out.pending_marks_remaining; break;
        case 4: v = s.pc_req; dst = out.pending_captures_required; break;
        case 5: v = s.pc_rem; dst = out.pending_captures_remaining; break;
        case 6: v = s.forced; dst = out.forced_removals_done; break;
        case 7: v = s.move_count; dst = out.move_count; break;
        case 8: v = s.msc; dst = out.moves_since_capture; break;
        default: break;
    }
    if (lane < 9) dst[i] = v;
}

template <bool kInplace>
__global__ void __launch_bounds__(kThreads)
apply_moves_kernel(lzb_states_in in, int64_t B, const int32_t* __restrict__ codes, const int64_t* __restrict__ parents,
                   int64_t N, lzb_states_out out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t i = warp; i < N; i += nwarps) {
        const int64_t p = parents[i];
        if (p < 0 || p >= B) continue;                         // :584-587 / :770-773
        State<long long> s;
        int8_t blo, bhi;
        warp_load_state<long long>(in, p, lane, s, blo, bhi);
        const int4 code = reinterpret_cast<const int4*>(codes)[i];
        apply_action(s, code.x, code.y, code.z);
        warp_store_state(out, kInplace ? p : i, lane, s, blo, bhi);
    }
}

// Byte <-> bit conversions, four cells at a time.  gather4: the low bit of each byte of x (x & 0x01010101 == x) -> a nibble;
// scatter4: a nibble -> 0x00/0x01 bytes.  Both are one multiplication by 2^0 + 2^7 + 2^14 + 2^21: the sixteen partial
// products land on distinct bit positions (no carries), the wanted ones on 21..24 resp. 0, 8, 16, 24.
__device__ __forceinline__ uint32_t gather4(uint32_t x01) { return ((x01 * 0x00204081u) >> 21) & 0xFu; }
__device__ __forceinline__ uint32_t scatter4(uint32_t nib) { return (nib * 0x00204081u) & 0x01010101u; }
__device__ __forceinline__ void words_to_boards(const uint32_t (&bw)[9], uint64_t& black, uint64_t& white, uint64_t& other) {
    black = white = other = 0;
#pragma unroll
    for (int w = 0; w < 9; ++w) {
        const uint32_t is_b = __vcmpeq4(bw[w], 0x01010101u), is_w = __vcmpeq4(bw[w], 0xFFFFFFFFu);   // 0xFF per equal byte
        const uint32_t is_o = __vcmpne4(bw[w], 0u) & ~(is_b | is_w);
        black |= (uint64_t)gather4(is_b & 0x01010101u) << (4 * w);
        white |= (uint64_t)gather4(is_w & 0x01010101u) << (4 * w);
        other |= (uint64_t)gather4(is_o & 0x01010101u) << (4 * w);
    }
}
__device__ __forceinline__ uint64_t words_to_marks(const uint32_t (&mw)[9]) {
    uint64_t m = 0;
#pragma unroll
    for (int w = 0; w < 9; ++w) m |= (uint64_t)gather4(__vcmpne4(mw[w], 0u) & 0x01010101u) << (4 * w);
    return m;
}
// the child's bytes of cells 4w..4w+3: +1 / -1 from the bitboards, any other original byte carried over where `other` is set
__device__ __forceinline__ uint32_t board_word(const State<long long>& s, uint32_t orig, int w) {
    const uint32_t b = scatter4((uint32_t)(s.black >> (4 * w)) & 0xFu), wh = scatter4((uint32_t)(s.white >> (4 * w)) & 0xFu),
                   o = scatter4((uint32_t)(s.other >> (4 * w)) & 0xFu);
    return b | (wh * 0xFFu) | (orig & (o * 0xFFu) & ~((b | wh) * 0xFFu));
}

// Thread-per-action variant for large batches.  A warp per action keeps only 64 actions in flight per SM and one
// action's dependent loads cost ~9 us, i.e. ~1 action/ns for the whole chip (6 % of the HBM roof at 384 B/action);
// with one THREAD per action 2,048 actions per SM are in flight.  Each thread gathers its parent's 3 x 36 bytes as
// 27 aligned 32-bit words, builds the bitboards four cells per multiplication, applies the action with the same
// apply_action() and hands the child's 27 words to its warp through shared memory, so that the 32 children of a warp
// (3 x 1,152 contiguous bytes when the output rows are the action indices) leave as fully coalesced 16-byte stores --
// written directly by the threads, a warp's store instruction touches 36 sectors for 128 bytes.
constexpr int kStageWords = 32 * 9;                     // one byte tensor's words of a warp's 32 children
template <bool kInplace>
__global__ void __launch_bounds__(kThreads, 3)
apply_moves_thread_kernel(lzb_states_in in, int64_t B, const int32_t* __restrict__ codes, const int64_t* __restrict__ parents,
                          int64_t N, lzb_states_out out, int vec_ok) {
    __shared__ __align__(16) uint32_t stage[kInplace ? 1 : kWarpsPerBlock][kInplace ? 1 : 3][kInplace ? 4 : kStageWords];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31); i0 < N; i0 += stride) {
        const int64_t i = i0 + lane;
        const int64_t p = i < N ? parents[i] : -1;
        const bool live = p >= 0 && p < B;                     // :584-587 / :770-773: other rows are left untouched
        uint32_t bw[9], mbw[9], mww[9];
        State<long long> s;
        if (live) {
            const uint32_t* bp = reinterpret_cast<const uint32_t*>(in.board + p * 36);
            const uint32_t* mbp = reinterpret_cast<const uint32_t*>(in.marks_black + p * 36);
            const uint32_t* mwp = reinterpret_cast<const uint32_t*>(in.marks_white + p * 36);
#pragma unroll
            for (int w = 0; w < 9; ++w) { bw[w] = bp[w]; mbw[w] = mbp[w]; mww[w] = mwp[w]; }
            s.mb = words_to_marks(mbw);
            s.mw = words_to_marks(mww);
            s.phase = in.phase[p]; s.player = in.current_player[p];
            s.pm_req = in.pending_marks_required[p]; s.pm_rem = in.pending_marks_remaining[p];
            s.pc_req = in.pending_captures_required[p]; s.pc_rem = in.pending_captures_remaining[p];
            s.forced = in.forced_removals_done[p];
            s.move_count = in.move_count ? in.move_count[p] : 0;
            s.msc = in.moves_since_capture ? in.moves_since_capture[p] : 0;
            const int4 code = reinterpret_cast<const int4*>(codes)[i];
            words_to_boards(bw, s.black, s.white, s.other);
            apply_action(s, code.x, code.y, code.z);
#pragma unroll
            for (int w = 0; w < 9; ++w) {
                bw[w] = board_word(s, bw[w], w);
                mbw[w] = scatter4((uint32_t)(s.mb >> (4 * w)) & 0xFu);
                mww[w] = scatter4((uint32_t)(s.mw >> (4 * w)) & 0xFu);
            }
            const int64_t o = kInplace ? p : i;
            out.phase[o] = s.phase; out.current_player[o] = s.player;
            out.pending_marks_required[o] = s.pm_req; out.pending_marks_remaining[o] = s.pm_rem;
            out.pending_captures_required[o] = s.pc_req; out.pending_captures_remaining[o] = s.pc_rem;
            out.forced_removals_done[o] = s.forced; out.move_count[o] = s.move_count; out.moves_since_capture[o] = s.msc;
        }
        // the three byte tensors: whole warp live and rows contiguous -> through shared memory, coalesced; otherwise directly
        const bool staged = !kInplace && __all_sync(0xFFFFFFFFu, live);
        if (staged) {
            if constexpr (!kInplace) {
#pragma unroll
                for (int w = 0; w < 9; ++w) {               // lane stride 9 words: conflict-free
                    stage[wib][0][lane * 9 + w] = bw[w];
                    stage[wib][1][lane * 9 + w] = mbw[w];
                    stage[wib][2][lane * 9 + w] = mww[w];
                }
                __syncwarp();
                uint8_t* const dst[3] = {reinterpret_cast<uint8_t*>(out.board), out.marks_black, out.marks_white};
#pragma unroll
                for (int arr = 0; arr < 3; ++arr) {
                    uint8_t* g = dst[arr] + i0 * 36;             // 1,152 contiguous bytes
                    if (vec_ok) {
                        const uint4* src = reinterpret_cast<const uint4*>(stage[wib][arr]);
                        for (int k = lane; k < kStageWords / 4; k += 32) reinterpret_cast<uint4*>(g)[k] = src[k];
                    } else {
                        for (int k = lane; k < kStageWords; k += 32) reinterpret_cast<uint32_t*>(g)[k] = stage[wib][arr][k];
                    }
                }
                __syncwarp();
            }
        } else if (live) {
            const int64_t o = kInplace ? p : i;
            uint32_t* ob = reinterpret_cast<uint32_t*>(out.board + o * 36);
            uint32_t* omb = reinterpret_cast<uint32_t*>(out.marks_black + o * 36);
            uint32_t* omw = reinterpret_cast<uint32_t*>(out.marks_white + o * 36);
#pragma unroll
            for (int w = 0; w < 9; ++w) { ob[w] = bw[w]; omb[w] = mbw[w]; omw[w] = mww[w]; }
        }
    }
}

// word access needs 4-byte aligned byte tensors (always true for whole torch tensors; a sliced view may not be)
inline bool words_ok(const void* a, const void* b, const void* c, const void* d, const void* e, const void* f) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
             reinterpret_cast<uintptr_t>(d) | reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f)) & 3) == 0;
}
constexpr int64_t kThreadApplyMin = 8192;      // below this the warp-per-action kernel's latency is as good

// ------------------------------------------------------------------------------------------------------
// (a6) states_to_model_input -- encoding.cpp:26-79 : f32[B,11,6,6]
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 6)
model_input_kernel(lzb_states_in st, int64_t B, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    // A state's 396 floats are 99 float4; float4 k holds cells 4 (k % 9) .. + 3 of plane k / 9 (36 = 9 x 4, so a float4
    // never straddles planes).  Lane l writes k = l, l + 32, l + 64, l + 96: its planes and cell offsets are fixed.
    int pl[4], c0[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int k = lane + 32 * j; pl[j] = k / 9; c0[j] = 4 * (k - pl[j] * 9); }
    const bool hi = lane < 4;
    int8_t n_blo = 0, n_bhi = 0;
    uint8_t n_mblo = 0, n_mwlo = 0, n_mbhi = 0, n_mwhi = 0;
    int64_t n_cur = 0, n_phase = 0;
    auto fetch = [&](int64_t b) {
        const int8_t* bp = st.board + b * 36;
        const uint8_t* mbp = st.marks_black + b * 36;
        const uint8_t* mwp = st.marks_white + b * 36;
        n_blo = bp[lane]; n_bhi = hi ? bp[32 + lane] : (int8_t)0;
        n_mblo = mbp[lane]; n_mwlo = mwp[lane];
        n_mbhi = hi ? mbp[32 + lane] : (uint8_t)0; n_mwhi = hi ? mwp[32 + lane] : (uint8_t)0;
        n_cur = st.current_player[b]; n_phase = st.phase[b];
    };
    if (warp < B) fetch(warp);
    for (int64_t b = warp; b < B; b += nwarps) {
        const int8_t blo = n_blo, bhi = n_bhi;
        const uint8_t mblo = n_mblo, mwlo = n_mwlo, mbhi = n_mbhi, mwhi = n_mwhi;
        // `current` is cast to the board dtype before the compare (encoding.cpp:51)
        const int64_t cur64 = n_cur, phase = n_phase;
        if (b + nwarps < B) fetch(b + nwarps);                  // next state's loads fly during this state's stores
        const int8_t cur = (int8_t)cur64, neg = (int8_t)(-cur);
        const uint64_t self_p = ballot36(blo == cur, hi && bhi == cur);
        const uint64_t opp_p = ballot36(blo == neg, hi && bhi == neg);
        const uint64_t mb = ballot36(mblo != 0, mbhi != 0);
        const uint64_t mw = ballot36(mwlo != 0, mwhi != 0);
        const bool is_black = cur64 == 1;
        const uint64_t own_marks = is_black ? mb : mw, opp_marks = is_black ? mw : mb;
        float4* o = reinterpret_cast<float4*>(out + b * 396);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j;
            if (k < 99) {
                const uint64_t pb = pl[j] == 0 ? self_p : pl[j] == 1 ? opp_p : pl[j] == 2 ? own_marks : opp_marks;
                const uint32_t nib = pl[j] < 4 ? (uint32_t)(pb >> c0[j]) & 0xFu : (phase == pl[j] - 3 ? 0xFu : 0u);
                o[k] = make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f,
                                   (nib & 8u) ? 1.0f : 0.0f);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a7) project_policy_logits_fast -- project_policy_logits_fast.cpp:16-164 (fp32, dims 36/144/36/4)
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float combined_logit(const float* p1, const float* p2, const float* pmc, int a) {
    if (a < 36) return p1[a];
    if (a < 180) {
        const int from = (a - 36) >> 2, d = (a - 36) & 3;
        const int r = from / 6, c = from - r * 6;
        const int rt = r + (d == 0 ? -1 : d == 1 ? 1 : 0), ct = c + (d == 2 ? -1 : d == 3 ? 1 : 0);
        if (rt < 0 || rt >= 6 || ct < 0 || ct >= 6) return -INFINITY;
        return p2[from] + p1[rt * 6 + ct];
    }
    if (a < 216) return pmc[a - 180];
    return 0.0f;
}

__global__ void __launch_bounds__(kThreads)
project_policy_kernel(const float* __restrict__ log_p1, const float* __restrict__ log_p2,
                      const float* __restrict__ log_pmc, const uint8_t* __restrict__ legal, int64_t B,
                      float* __restrict__ probs, float* __restrict__ masked_logits) {
    __shared__ float heads[kWarpsPerBlock][3][36];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + w;
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t b = warp; b < B; b += nwarps) {
        __syncwarp();
        for (int i = lane; i < 36; i += 32) {
            heads[w][0][i] = log_p1[b * 36 + i];
            heads[w][1][i] = log_p2[b * 36 + i];
            heads[w][2][i] = log_pmc[b * 36 + i];
        }
        __syncwarp();
        float logit[7];
        bool leg[7];
        float mx = -INFINITY;
        bool any_legal = false;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int a = lane + 32 * k;
            leg[k] = a < 220 && legal[b * 220 + a] != 0;
            logit[k] = leg[k] ? combined_logit(heads[w][0], heads[w][1], heads[w][2], a) : -INFINITY;
            mx = fmaxf(mx, logit[k]);
            any_legal |= leg[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
        any_legal = __any_sync(0xffffffffu, any_legal);
        const bool has_finite = mx > -INFINITY && mx < INFINITY;   // row has a finite legal logit
        float e[7], sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            e[k] = (has_finite && logit[k] > -INFINITY) ? expf(logit[k] - mx) : 0.0f;
            sum += e[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const int a = lane + 32 * k;
            if (a >= 220) continue;
            probs[b * 220 + a] = has_finite ? e[k] / sum : 0.0f;
            // rows with legal actions but no finite logit: legal entries of masked_logits become 0 (:150-159)
            masked_logits[b * 220 + a] = (any_legal && !has_finite && leg[k]) ? 0.0f : logit[k];
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a9) root_puct_allocate_visits -- root_puct_fused.cu:12-117
// One warp per root; lane l owns actions l, l+32, ... (kSlots per lane) with N / W / c*P in registers.
// Per simulation: recompute u for the owned slots (sqrt(total+1) changes every step), keep q cached
// (it only changes for the chosen action), warp argmax with lowest-index tie-break via shuffles.
// fp32 expression order is the reference's: u = ((c * p) * sqrt_total) / (1 + n), score = q + u.
// ------------------------------------------------------------------------------------------------------
template <int kSlots>
__global__ void __launch_bounds__(kThreads)
root_puct_kernel(const float* __restrict__ priors, const float* __restrict__ leaf, const uint8_t* __restrict__ valid,
                 int64_t R, int M, int64_t sims, float c_puct, float* __restrict__ visits,
                 float* __restrict__ value_sum, float* __restrict__ root_values) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t r = warp; r < R; r += nwarps) {
        float cp[kSlots], lv[kSlots], n[kSlots], w[kSlots], q[kSlots];
        bool ok[kSlots];
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int a = lane + 32 * k;
            ok[k] = a < M && valid[r * M + a] != 0;
            cp[k] = ok[k] ? __fmul_rn(c_puct, priors[r * M + a]) : 0.0f;
            lv[k] = ok[k] ? leaf[r * M + a] : 0.0f;
            n[k] = 0.0f; w[k] = 0.0f; q[k] = 0.0f;
        }
        float total = 0.0f;
        for (int64_t sim = 0; sim < sims; ++sim) {
            const float sqrt_total = __fsqrt_rn(__fadd_rn(total, 1.0f));
            float best = -INFINITY;
            int best_idx = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                const float u = __fdiv_rn(__fmul_rn(cp[k], sqrt_total), __fadd_rn(1.0f, n[k]));
                const float score = __fadd_rn(q[k], u);
                // ascending index within a lane, so strict > keeps the lowest index on ties; a NaN score is
                // never selected and a -inf score only as the first candidate (root_puct_fused.cu:58)
                if (ok[k] && (score > best || (score == best && best_idx == 0x7fffffff))) {
                    best = score; best_idx = lane + 32 * k;
                }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, off);
                const int oi = __shfl_xor_sync(0xffffffffu, best_idx, off);
                const bool take = (oi != 0x7fffffff) &&
                                  (best_idx == 0x7fffffff || ob > best || (ob == best && oi < best_idx));
                if (take) { best = ob; best_idx = oi; }
            }
            if (best_idx == 0x7fffffff) continue;             // no valid action: the reference does nothing
#pragma unroll
            for (int k = 0; k < kSlots; ++k) {
                if (best_idx == lane + 32 * k) {
                    n[k] = __fadd_rn(n[k], 1.0f);
                    w[k] = __fadd_rn(w[k], lv[k]);
                    q[k] = __fdiv_rn(w[k], fmaxf(n[k], 1e-8f));
                }
            }
            total = __fadd_rn(total, 1.0f);
        }
        float pv = 0.0f, pw = 0.0f;
#pragma unroll
        for (int k = 0; k < kSlots; ++k) {
            const int a = lane + 32 * k;
            if (a < M) { visits[r * M + a] = n[k]; value_sum[r * M + a] = w[k]; }
            pv += n[k]; pw += w[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            pv += __shfl_xor_sync(0xffffffffu, pv, off);
            pw += __shfl_xor_sync(0xffffffffu, pw, off);
        }
        if (lane == 0) root_values[r] = pw / fmaxf(pv, 1.0f);
    }
}

}  // namespace
}  // namespace lzb

using namespace lzb;

extern "C" int lzb_encode_actions_fast(const lzb_states_in* st, int64_t B, int64_t pd, int64_t md, int64_t sd,
                                       int64_t ad, uint8_t* mask, int32_t* metadata, void* stream) {
    LZB_REQUIRE(st && B >= 0 && pd >= 0 && md >= 0 && sd >= 0 && ad >= 0, "bad arguments");
    LZB_REQUIRE(pd + md + sd + ad < (1 << 20), "action dims too large");
    if (B == 0 || pd + md + sd + ad == 0) return LZB_OK;
    LZB_REQUIRE(mask && metadata, "null output");
    LZB_REQUIRE((reinterpret_cast<uintptr_t>(metadata) & 15) == 0, "metadata must be 16-byte aligned");
    encode_actions_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(*st, B, (int)pd, (int)md, (int)sd,
                                                                                (int)ad, mask, metadata);
    return check_launch("encode_actions_kernel");
}

static lzb_states_in as_in(const lzb_states_out* o) {
    lzb_states_in in;
    in.board = o->board; in.marks_black = o->marks_black; in.marks_white = o->marks_white;
    in.phase = o->phase; in.current_player = o->current_player;
    in.pending_marks_required = o->pending_marks_required; in.pending_marks_remaining = o->pending_marks_remaining;
    in.pending_captures_required = o->pending_captures_required;
    in.pending_captures_remaining = o->pending_captures_remaining;
    in.forced_removals_done = o->forced_removals_done; in.move_count = o->move_count;
    in.moves_since_capture = o->moves_since_capture;
    return in;
}

extern "C" int lzb_batch_apply_moves(const lzb_states_in* parents, int64_t B, const int32_t* codes,
                                     const int64_t* parent_indices, int64_t N, const lzb_states_out* children,
                                     void* stream) {
    LZB_REQUIRE(parents && children && B >= 0 && N >= 0, "bad arguments");
    if (N == 0) return LZB_OK;
    LZB_REQUIRE(codes && parent_indices, "null action arrays");
    LZB_REQUIRE(parents->move_count && parents->moves_since_capture, "move_count / moves_since_capture required");
    LZB_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0, "action_codes must be 16-byte aligned");
    if (N >= kThreadApplyMin && words_ok(parents->board, parents->marks_black, parents->marks_white, children->board,
                                         children->marks_black, children->marks_white)) {
        // 16-byte stores of a warp's 1,152-byte row group need 16-byte aligned tensors (whole torch tensors are)
        const int vec_ok = ((reinterpret_cast<uintptr_t>(children->board) | reinterpret_cast<uintptr_t>(children->marks_black) |
                             reinterpret_cast<uintptr_t>(children->marks_white)) & 15) == 0;
        apply_moves_thread_kernel<false><<<thread_grid(N), kThreads, 0, (cudaStream_t)stream>>>(*parents, B, codes,
                                                                                                 parent_indices, N, *children, vec_ok);
        return check_launch("apply_moves_thread_kernel");
    }
    apply_moves_kernel<false><<<warp_grid(N), kThreads, 0, (cudaStream_t)stream>>>(*parents, B, codes, parent_indices,
                                                                                  N, *children);
    return check_launch("apply_moves_kernel");
}

extern "C" int lzb_batch_apply_moves_inplace(const lzb_states_out* states, int64_t B, const int32_t* codes,
                                             const int64_t* slot_indices, int64_t N, void* stream) {
    LZB_REQUIRE(states && B >= 0 && N >= 0, "bad arguments");
    if (N == 0) return LZB_OK;
    LZB_REQUIRE(codes && slot_indices, "null action arrays");
    LZB_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0, "action_codes must be 16-byte aligned");
    if (N >= kThreadApplyMin && words_ok(states->board, states->marks_black, states->marks_white, states->board,
                                         states->marks_black, states->marks_white)) {
        apply_moves_thread_kernel<true><<<thread_grid(N), kThreads, 0, (cudaStream_t)stream>>>(as_in(states), B, codes,
                                                                                                slot_indices, N, *states, 0);
        return check_launch("apply_moves_inplace_thread_kernel");
    }
    apply_moves_kernel<true><<<warp_grid(N), kThreads, 0, (cudaStream_t)stream>>>(as_in(states), B, codes,
                                                                                 slot_indices, N, *states);
    return check_launch("apply_moves_inplace_kernel");
}

extern "C" int lzb_states_to_model_input(const lzb_states_in* st, int64_t B, float* out, void* stream) {
    LZB_REQUIRE(st && B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(out && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "out must be 16-byte aligned");
    model_input_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(*st, B, out);
    return check_launch("model_input_kernel");
}

extern "C" int lzb_project_policy_logits_fast(const float* log_p1, const float* log_p2, const float* log_pmc,
                                              const uint8_t* legal_mask, int64_t B, float* probs,
                                              float* masked_logits, void* stream) {
    LZB_REQUIRE(B >= 0, "bad arguments");
    if (B == 0) return LZB_OK;
    LZB_REQUIRE(log_p1 && log_p2 && log_pmc && legal_mask && probs && masked_logits, "null pointer");
    project_policy_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(log_p1, log_p2, log_pmc, legal_mask, B,
                                                                                probs, masked_logits);
    return check_launch("project_policy_kernel");
}

extern "C" int lzb_root_puct_allocate_visits(const float* priors, const float* leaf_values, const uint8_t* valid_mask,
                                             int64_t R, int64_t M, int64_t sims, float c_puct, float* visits,
                                             float* value_sum, float* root_values, void* stream) {
    LZB_REQUIRE(R >= 0 && M >= 0, "bad shape");
    LZB_REQUIRE(sims > 0, "num_simulations must be positive");        // module.cpp:191
    LZB_REQUIRE(M <= 1024, "at most 1024 actions per root");
    if (R == 0) return LZB_OK;
    if (M == 0) return LZB_OK;
    LZB_REQUIRE(priors && leaf_values && valid_mask && visits && value_sum && root_values, "null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int grid = warp_grid(R);
#define LZB_PUCT(K)                                                                                             \
    root_puct_kernel<K><<<grid, kThreads, 0, s>>>(priors, leaf_values, valid_mask, R, (int)M, sims, c_puct, visits, \
                                                  value_sum, root_values)
    if (M <= 32) LZB_PUCT(1);
    else if (M <= 64) LZB_PUCT(2);
    else if (M <= 96) LZB_PUCT(3);
    else if (M <= 128) LZB_PUCT(4);
    else if (M <= 256) LZB_PUCT(8);
    else if (M <= 512) LZB_PUCT(16);
    else LZB_PUCT(32);
#undef LZB_PUCT
    return check_launch("root_puct_kernel");
}
