// lz_rules.cuh -- Liuzhou Chess rule engine on packed 6x6 bitboards (sm_100a device code; the same
// header compiles for the host so tests/host_shim can check the bit logic against the oracle on CPU).
//
// Cell index = r*6 + c  <->  bit index of a uint64 (36 bits used).  All set operations below are
// branch-light bit ops: Fang (2x2) and Zhou (full row / column) detection are shifts + ANDs, legal-move
// generation is four shifted intersections, phase transitions are scalar updates.
//
// Semantics follow the reference exactly (file:line cited per function, all under /root/reference):
//   legal sets      v0/src/game/fast_legal_mask.cpp:110-249,326-414  (== fast_legal_mask_cuda.cu)
//   apply           v0/src/game/fast_apply_moves_cuda.cu:201-546,624-743 (silent no-op on illegal input)
//   scalar variants v0/src/moves/move_generator.cpp:143-170,242-297, v0/src/game/game_state.cpp:59-79
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LZ_HD __host__ __device__ __forceinline__
#else
#define LZ_HD inline
#endif

namespace lz {

constexpr uint64_t kFull = 0xFFFFFFFFFULL;          // 36 cells
constexpr uint64_t kCol0 = 0x041041041ULL;          // c == 0
constexpr uint64_t kCol5 = kCol0 << 5;              // c == 5
constexpr uint64_t kRow0 = 0x3FULL;                 // r == 0
constexpr uint64_t kAnchor = 0x1F7DF7DFULL;         // r <= 4 && c <= 4 (top-left corners of 2x2 blocks)

enum Phase : int { kPlacement = 1, kMark = 2, kRemoval = 3, kMovement = 4, kCapture = 5, kForced = 6, kCounter = 7 };
enum Action : int { kActPlace = 1, kActMove = 2, kActMark = 3, kActCapture = 4, kActForced = 5,
                    kActCounter = 6, kActNoMoves = 7, kActProcess = 8 };

constexpr int kMaxMoveCount = 144;      // game_state.hpp:14
constexpr int kLoseThreshold = 4;       // game_state.hpp:15
constexpr int kNoCaptureLimit = 36;     // game_state.hpp:16
constexpr int kActionDim = 220;         // 36 + 144 + 36 + 4

LZ_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
LZ_HD int ctz64(uint64_t x) {  // x != 0
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
LZ_HD uint64_t bit(int cell) { return 1ULL << cell; }
LZ_HD uint64_t row_mask(int r) { return kRow0 << (6 * r); }
LZ_HD uint64_t col_mask(int c) { return kCol0 << c; }

// ---- shape detection --------------------------------------------------------------------------------
// Cells covered by a complete 2x2 block of `e` (e = pieces of one colour that are not in the marked set).
// rule_engine.cpp:57-89: a block counts only if all four cells are own AND unmarked.
LZ_HD uint64_t in_square(uint64_t e) {
    uint64_t a = e & (e >> 1) & (e >> 6) & (e >> 7) & kAnchor;
    return a | (a << 1) | (a << 6) | (a << 7);
}
// Cells x of `own` for which CheckLines(x) holds (rule_engine.cpp:91-136): the other five cells of x's row
// (or column) are all in e.  x itself is NOT tested against the marked set -- reference quirk.
LZ_HD uint64_t in_line(uint64_t own, uint64_t e) {
    // All six rows and all six columns at once (SWAR on the 36-bit board; equal to the per-line loop
    //   miss = line & ~e;  res |= miss == 0 ? line : (popcount(miss) == 1 ? miss & own : 0)
    // on 30 M random boards).  m = missing cells; a line contributes its cells if m is empty on it, or its single
    // missing cell if that cell is own (an own piece that is marked: the reference quirk above).
    e &= kFull;
    const uint64_t m = kFull & ~e;
    constexpr uint64_t kC01 = kCol0 | (kCol0 << 1), kC0123 = kC01 | (kC01 << 2);
    // rows: neighbouring columns are adjacent bits; the masks keep every shift inside its row
    uint64_t t = e & (e >> 1);
    const uint64_t rf = t & (t >> 2) & (t >> 4) & kCol0;        // row complete, flagged at its column-0 bit
    uint64_t p = m | ((m << 1) & ~kCol0);
    p |= (p << 2) & ~kC01;
    p |= (p << 4) & ~kC0123;                                     // inclusive prefix OR of m along the row
    uint64_t d = m & ((p << 1) & ~kCol0);                        // missing cells with another one to their left
    uint64_t a = d | ((d >> 1) & ~kCol5);
    a |= (a >> 2) & kC0123;
    a |= (a >> 4) & kC01;                                        // column-0 bit: the row misses two or more cells
    uint64_t res = rf * 63u | (m & own & ~((a & kCol0) * 63u));
    // columns: neighbouring rows are 6 bits apart; shifts by multiples of 6 never leave the column
    t = e & (e >> 6);
    const uint64_t cf = t & (t >> 12) & (t >> 24) & kRow0;
    p = m | (m << 6);
    p |= p << 12;
    p |= p << 24;
    d = m & (p << 6) & kFull;
    a = d | (d >> 6);
    a |= a >> 12;
    a |= a >> 24;
    res |= cf * kCol0 | (m & own & ~((a & kRow0) * kCol0));
    return res;
}
// IsPieceInShape over a whole colour at once (rule_engine.cpp:194-208).
LZ_HD uint64_t in_shape(uint64_t own, uint64_t marked) {
    uint64_t e = own & ~marked;
    return in_square(e) | in_line(own, e);
}
// DetectShapeFormed at one cell (rule_engine.cpp:138-154): 2 = line (Zhou), 1 = square (Fang), 0 = none.
LZ_HD int detect_shape(uint64_t own, uint64_t marked, int cell) {
    uint64_t e = own & ~marked, b = bit(cell);
    uint64_t rm = row_mask(cell / 6), cm = col_mask(cell % 6);
    bool line = (((e | b) & rm) == rm) || (((e | b) & cm) == cm);
    if (line) return 2;
    return (in_square(e) & b) ? 1 : 0;
}
// fast_legal_mask.cpp:110-129: drop candidates that sit in a shape unless that leaves nothing.
LZ_HD uint64_t prefer_normal(uint64_t cands, uint64_t own, uint64_t marked) {
    uint64_t normal = cands & ~in_shape(own, marked);
    return normal ? normal : cands;
}

// ---- working state -----------------------------------------------------------------------------------
// I = int (native packed states) or long long (drop-in kernels: the reference keeps nine int64 scalars and
// its kernels do int64 arithmetic on them).
template <typename I>
struct State {
    uint64_t black, white;   // pieces (+1 / -1)
    uint64_t other;          // cells whose byte is neither -1, 0 nor +1 (drop-in layout only; 0 natively)
    uint64_t mb, mw;         // marks_black / marks_white
    I phase, player, pm_req, pm_rem, pc_req, pc_rem, forced, move_count, msc;
};

template <typename I> LZ_HD uint64_t occupied(const State<I>& s) { return s.black | s.white | s.other; }
template <typename I> LZ_HD uint64_t empty_cells(const State<I>& s) { return kFull & ~occupied(s); }
// cells whose board byte equals `v` (the reference compares bytes with an arbitrary int)
template <typename I> LZ_HD uint64_t pieces(const State<I>& s, long long v) {
    return v == 1 ? s.black : (v == -1 ? s.white : (v == 0 ? empty_cells(s) : 0ULL));
}
template <typename I> LZ_HD void set_initial(State<I>& s) {
    s.black = s.white = s.other = s.mb = s.mw = 0;
    s.phase = kPlacement; s.player = 1;
    s.pm_req = s.pm_rem = s.pc_req = s.pc_rem = s.forced = s.move_count = s.msc = 0;
}

// ---- legal actions -----------------------------------------------------------------------------------
struct Legal {
    uint64_t place;      // action a = cell                     (kind 1)
    uint64_t mv[4];      // action a = 36 + from*4 + dir, set of `from` cells per dir (up, down, left, right)
    uint64_t sel;        // action a = 180 + cell               (kind = sel_kind)
    int sel_kind;        // 3 mark, 4 capture, 5 forced, 6 counter, 7 no-moves removal
    bool process;        // action 216                          (kind 8)
};

LZ_HD int legal_count(const Legal& L) {
    return popc64(L.place) + popc64(L.mv[0]) + popc64(L.mv[1]) + popc64(L.mv[2]) + popc64(L.mv[3]) +
           popc64(L.sel) + (L.process ? 1 : 0);
}
LZ_HD bool legal_test(const Legal& L, int a) {
    if (a < 36) return (L.place >> a) & 1;
    if (a < 180) return (L.mv[(a - 36) & 3] >> ((a - 36) >> 2)) & 1;
    if (a < 216) return (L.sel >> (a - 180)) & 1;
    return a == 216 && L.process;
}
// number of legal actions with index < a
LZ_HD int legal_rank(const Legal& L, int a) {
    if (a <= 36) return popc64(L.place & (bit(a) - 1));
    int n = popc64(L.place);
    if (a <= 180) {
        int from = (a - 36) >> 2, d = (a - 36) & 3;
        uint64_t below = bit(from) - 1;
        n += popc64(L.mv[0] & below) + popc64(L.mv[1] & below) + popc64(L.mv[2] & below) + popc64(L.mv[3] & below);
#pragma unroll
        for (int k = 0; k < 3; ++k) n += (k < d) ? (int)((L.mv[k] >> from) & 1) : 0;
        return n;
    }
    n += popc64(L.mv[0]) + popc64(L.mv[1]) + popc64(L.mv[2]) + popc64(L.mv[3]);
    if (a <= 216) return n + popc64(L.sel & (bit(a - 180) - 1));
    return n + popc64(L.sel) + (L.process ? 1 : 0);
}
// k-th legal action in ascending index order, 0 <= k < legal_count
LZ_HD int select_bit(uint64_t x, int k) {
    for (int i = 0; i < k; ++i) x &= x - 1;
    return ctz64(x);
}
LZ_HD int legal_kth(const Legal& L, int k) {
    int n = popc64(L.place);
    if (k < n) return select_bit(L.place, k);
    k -= n;
    uint64_t u = L.mv[0] | L.mv[1] | L.mv[2] | L.mv[3];
    int nm = popc64(L.mv[0]) + popc64(L.mv[1]) + popc64(L.mv[2]) + popc64(L.mv[3]);
    if (k < nm) {
        while (true) {
            int from = ctz64(u);
            int here = (int)((L.mv[0] >> from) & 1) + (int)((L.mv[1] >> from) & 1) + (int)((L.mv[2] >> from) & 1) +
                       (int)((L.mv[3] >> from) & 1);
            if (k < here) {
#pragma unroll
                for (int d = 0; d < 4; ++d) {
                    if ((L.mv[d] >> from) & 1) {
                        if (k == 0) return 36 + from * 4 + d;
                        --k;
                    }
                }
            }
            k -= here;
            u &= u - 1;
        }
    }
    k -= nm;
    n = popc64(L.sel);
    if (k < n) return 180 + select_bit(L.sel, k);
    return 216;
}
// (kind, primary, secondary, extra) exactly as encode_actions_fast writes metadata (fast_legal_mask.cpp:326-414)
LZ_HD void action_code(const Legal& L, int a, int& kind, int& primary, int& secondary, int& extra) {
    if (a < 36) { kind = kActPlace; primary = a; secondary = -1; extra = -1; return; }
    if (a < 180) {
        int from = (a - 36) >> 2, d = (a - 36) & 3;
        kind = kActMove; primary = from; secondary = d;
        extra = from + (d == 0 ? -6 : d == 1 ? 6 : d == 2 ? -1 : 1);
        return;
    }
    if (a < 216) { kind = L.sel_kind; primary = a - 180; secondary = -1; extra = -1; return; }
    kind = kActProcess; primary = -1; secondary = -1; extra = -1;
}

// Tensor-op semantics (encode_actions_fast): no game-over check; forced removal falls back to all pieces
// when every piece is in a shape.  kScalar = true gives v0::GenerateAllLegalMoves semantics instead:
// empty on game over and no forced-removal fallback (move_generator.cpp:143-170,242-245).
template <typename I, bool kScalar>
LZ_HD void legal_actions(const State<I>& s, Legal& L, bool aux_enabled = true);

template <typename I> LZ_HD int winner(const State<I>& s) {   // game_state.cpp:59-75 (+1 black, -1 white, 0 none)
    if (s.phase != kMovement && s.phase != kCapture && s.phase != kCounter) return 0;
    if (popc64(s.black) < kLoseThreshold) return -1;
    if (popc64(s.white) < kLoseThreshold) return 1;
    return 0;
}
template <typename I> LZ_HD bool draw_limit(const State<I>& s) {   // game_state.hpp:128-131
    return s.move_count >= kMaxMoveCount || s.msc >= kNoCaptureLimit;
}
template <typename I> LZ_HD bool game_over(const State<I>& s) { return winner(s) != 0 || draw_limit(s); }

template <typename I, bool kScalar>
LZ_HD void legal_actions(const State<I>& s, Legal& L, bool aux_enabled) {
    L.place = 0; L.mv[0] = L.mv[1] = L.mv[2] = L.mv[3] = 0; L.sel = 0; L.sel_kind = 0; L.process = false;
    if (kScalar && game_over(s)) return;
    const int phase = (int)s.phase;
    const long long cur = (long long)(int)s.player;
    const uint64_t emp = empty_cells(s);
    if (phase == kPlacement) { L.place = emp; return; }
    if (phase == kRemoval) { L.process = aux_enabled; return; }
    const uint64_t opp_marked = (cur == 1) ? s.mw : s.mb;   // fast_legal_mask.cpp:381,391
    const uint64_t opp = pieces(s, -cur);
    // Every selection phase ends in the same "prefer pieces outside a shape" filter; the phases only choose its
    // operands.  One in_shape() evaluation after the phase switch instead of one per branch: a warp whose threads hold
    // states in different phases (thread-per-state kernels, playouts) runs the expensive part once, not once per phase.
    bool need = false, strict = false;                          // strict: scalar-engine forced removal has no fallback
    uint64_t cands = 0, sh_own = 0, sh_marked = 0;
    if (phase == kMovement) {
        const uint64_t own = pieces(s, cur);
        L.mv[0] = own & (emp << 6);                 // up:    cell-6 empty
        L.mv[1] = own & (emp >> 6);                 // down:  cell+6 empty
        L.mv[2] = own & ((emp << 1) & ~kCol0);      // left:  cell-1 empty, c > 0
        L.mv[3] = own & ((emp >> 1) & ~kCol5);      // right: cell+1 empty, c < 5
        if ((L.mv[0] | L.mv[1] | L.mv[2] | L.mv[3]) == 0) {   // no-moves removal :164-177,:405-408
            L.sel_kind = kActNoMoves;
            need = true; cands = opp; sh_own = opp; sh_marked = 0;
        }
    } else if (phase == kMark) {                                // :204-226
        L.sel_kind = kActMark;
        if (s.pm_rem > 0) { need = true; cands = opp & ~opp_marked; sh_own = opp; sh_marked = opp_marked; }
    } else if (phase == kCapture) {                             // :228-249
        L.sel_kind = kActCapture;
        if (s.pc_rem > 0) { need = true; cands = opp; sh_own = opp; sh_marked = opp_marked; }
    } else if (phase == kForced) {                              // :131-147
        L.sel_kind = kActForced;
        if (s.forced == 0 || s.forced == 1 || (!kScalar && s.forced < 2)) {
            const uint64_t tgt = (s.forced == 0) ? s.black : s.white;
            need = true; cands = tgt; sh_own = tgt; sh_marked = 0;
            strict = kScalar;                                   // move_generator.cpp:159-167: no fallback
        }
    } else if (phase == kCounter) {                             // :149-162
        L.sel_kind = kActCounter;
        need = true; cands = opp; sh_own = opp; sh_marked = 0;
    }
    if (need) {
        const uint64_t normal = cands & ~in_shape(sh_own, sh_marked);
        L.sel = (normal || strict) ? normal : cands;
    }
}

// ---- apply one atomic action (fast_apply_moves_cuda.cu semantics) --------------------------------------
template <typename I> LZ_HD void set_piece(State<I>& s, int cell, long long v) {
    const uint64_t b = bit(cell);
    s.black &= ~b; s.white &= ~b; s.other &= ~b;
    if (v == 1) s.black |= b; else if (v == -1) s.white |= b; else if (v != 0) s.other |= b;
}
// has_unmarked_normal_piece (fast_apply_moves_cuda.cu:290-306)
LZ_HD bool has_unmarked_normal(uint64_t own, uint64_t marked) {
    return (own & ~in_shape(own, marked) & ~marked) != 0;
}

// Returns true if applied, false for a silent no-op.  move_count / moves_since_capture bookkeeping of the
// kernel body (fast_apply_moves_cuda.cu:624-743) included.
// kTrusted = true: the action is known to come from legal_actions() on this very state (playouts, tree expansion), so the
// "is the piece inside a shape while a normal piece exists" re-validation -- two in_shape() evaluations per removal -- is
// compiled out; the legal set has already applied exactly that preference.  Everything else is unchanged.
template <typename I, bool kTrusted = false>
LZ_HD bool apply_action(State<I>& s, int kind, int primary, int secondary) {
    const I phase_before = s.phase;
    const int old_total = popc64(occupied(s));
    bool ok = false;
    const long long cur = (long long)s.player;
    switch (kind) {
    case kActPlace: {                                           // :239-288
        const int cell = primary;
        if (s.phase != kPlacement || cell < 0 || cell >= 36) break;
        if (occupied(s) & bit(cell)) break;
        const uint64_t opp_marked = (cur == 1) ? s.mw : s.mb;
        if (opp_marked & bit(cell)) break;
        set_piece(s, cell, (long long)(int8_t)cur);            // board byte = (int8) current_player
        const uint64_t own_marked = (cur == 1) ? s.mb : s.mw;
        ok = true;
        int shape = 0;
        if (!(own_marked & bit(cell))) shape = detect_shape(pieces(s, (long long)(int)cur), own_marked, cell);
        if (shape) {
            s.pm_req = s.pm_rem = (shape == 2) ? 2 : 1;
            s.phase = kMark;
        } else {
            s.pm_req = s.pm_rem = 0;
            if (empty_cells(s) == 0) s.phase = kRemoval;
            else { s.player = -s.player; s.phase = kPlacement; }
        }
        s.move_count += 1;                                      // only when applied (:273,:279,:287)
        break;
    }
    case kActMark: {                                            // :308-348
        s.move_count += 1;
        const int cell = primary;
        if (s.phase != kMark || s.pm_rem <= 0 || cell < 0 || cell >= 36) break;
        const long long oppv = (long long)(int)(-cur);
        uint64_t& opp_marked = (oppv == -1) ? s.mw : s.mb;
        const uint64_t opp = pieces(s, oppv);
        if (!(opp & bit(cell)) || (opp_marked & bit(cell))) break;
        if (!kTrusted && (in_shape(opp, opp_marked) & bit(cell)) && has_unmarked_normal(opp, opp_marked)) break;
        opp_marked |= bit(cell);
        s.pm_rem -= 1;
        ok = true;
        if (s.pm_rem > 0) break;
        s.pm_req = s.pm_rem = 0;
        if (empty_cells(s) == 0) s.phase = kRemoval;
        else { s.player = -s.player; s.phase = kPlacement; }
        break;
    }
    case kActProcess: {                                         // :201-237 (no phase check in the reference)
        s.move_count += 1;
        ok = true;
        const uint64_t marked = s.mb | s.mw;
        if (marked == 0) { s.phase = kForced; s.player = -1; s.forced = 0; break; }
        s.black &= ~marked; s.white &= ~marked; s.other &= ~marked;
        s.mb = s.mw = 0;
        s.phase = kMovement; s.player = -1;                     // removed > 0 always: marked cells are counted
        break;
    }
    case kActForced: {                                          // :350-386
        s.move_count += 1;
        const int cell = primary;
        if (s.phase != kForced || cell < 0 || cell >= 36) break;
        if (s.forced == 0) {
            if (s.player != -1 || !(s.black & bit(cell))) break;
            if (!kTrusted && (in_shape(s.black, 0) & bit(cell))) break;
            s.black &= ~bit(cell);
            s.forced = 1; s.player = 1; ok = true;
        } else if (s.forced == 1) {
            if (s.player != 1 || !(s.white & bit(cell))) break;
            if (!kTrusted && (in_shape(s.white, 0) & bit(cell))) break;
            s.white &= ~bit(cell);
            s.forced = 2; s.phase = kMovement; s.player = -1; ok = true;
        }
        break;
    }
    case kActMove: {                                            // :488-546
        s.move_count += 1;
        const int from = primary, d = secondary;
        if (s.phase != kMovement || d < 0 || d >= 4) break;
        if (from < 0 || from >= 36) break;                      // reference: unchecked (UB); here: no-op
        const int r = from / 6, c = from % 6;
        const int rt = r + (d == 0 ? -1 : d == 1 ? 1 : 0), ct = c + (d == 2 ? -1 : d == 3 ? 1 : 0);
        if (rt < 0 || rt >= 6 || ct < 0 || ct >= 6) break;
        const int to = rt * 6 + ct;
        if (!(pieces(s, cur) & bit(from)) || (occupied(s) & bit(to))) break;
        if (cur == 0) break;                                    // "piece" would be an empty cell: nothing moves
        set_piece(s, to, cur);
        set_piece(s, from, 0);
        ok = true;
        const int shape = detect_shape(pieces(s, (long long)(int)cur), 0, to);
        if (shape) { s.pc_req = s.pc_rem = (shape == 2) ? 2 : 1; s.phase = kCapture; }
        else { s.pc_req = s.pc_rem = 0; s.player = -s.player; }
        break;
    }
    case kActNoMoves: {                                         // :388-416
        s.move_count += 1;
        const int cell = primary;
        if (s.phase != kMovement || cell < 0 || cell >= 36) break;
        const long long oppv = (long long)(int)(-cur);
        const uint64_t opp = pieces(s, oppv);
        if (!(opp & bit(cell)) || oppv == 0) break;
        if (!kTrusted && (in_shape(opp, 0) & bit(cell)) && has_unmarked_normal(opp, 0)) break;
        set_piece(s, cell, 0);
        ok = true;
        if (popc64(pieces(s, oppv)) < kLoseThreshold) break;
        s.phase = kCounter; s.player = -s.player;
        break;
    }
    case kActCapture: {                                         // :418-456
        s.move_count += 1;
        const int cell = primary;
        if (s.phase != kCapture || s.pc_rem <= 0 || cell < 0 || cell >= 36) break;
        const long long oppv = (long long)(int)(-cur);
        const uint64_t opp_marked = (oppv == -1) ? s.mw : s.mb;
        const uint64_t opp = pieces(s, oppv);
        if (!(opp & bit(cell)) || oppv == 0) break;
        if (!kTrusted && (in_shape(opp, opp_marked) & bit(cell)) && has_unmarked_normal(opp, opp_marked)) break;
        set_piece(s, cell, 0);
        s.pc_rem -= 1;
        ok = true;
        if (popc64(pieces(s, oppv)) < kLoseThreshold || s.pc_rem > 0) break;
        s.pc_req = s.pc_rem = 0;
        s.player = -s.player; s.phase = kMovement;
        break;
    }
    case kActCounter: {                                         // :458-486
        s.move_count += 1;
        const int cell = primary;
        if (s.phase != kCounter || cell < 0 || cell >= 36) break;
        const long long stuckv = (long long)(int)(-cur);
        const uint64_t stuck = pieces(s, stuckv);
        if (!(stuck & bit(cell)) || stuckv == 0) break;
        if (!kTrusted && (in_shape(stuck, 0) & bit(cell)) && has_unmarked_normal(stuck, 0)) break;
        set_piece(s, cell, 0);
        ok = true;
        if (popc64(pieces(s, stuckv)) < kLoseThreshold) break;
        s.phase = kMovement; s.player = -s.player;
        break;
    }
    default: break;
    }
    if (phase_before == kPlacement || phase_before == kMark) {  // :728-743
        s.msc = 0;
    } else {
        const int new_total = popc64(occupied(s));
        s.msc = (new_total < old_total) ? (I)0 : (I)(s.msc + 1);
    }
    return ok;
}

// Apply by 220-d action index (kind resolved from the phase, as the mask encoder would emit it).
template <typename I, bool kTrusted = false>
LZ_HD bool apply_index(State<I>& s, int a) {
    int kind, primary = -1, secondary = -1;
    if (a < 36) { kind = kActPlace; primary = a; }
    else if (a < 180) { kind = kActMove; primary = (a - 36) >> 2; secondary = (a - 36) & 3; }
    else if (a < 216) {
        primary = a - 180;
        const int ph = (int)s.phase;
        kind = ph == kMark ? kActMark : ph == kCapture ? kActCapture : ph == kForced ? kActForced
             : ph == kCounter ? kActCounter : kActNoMoves;
    } else { kind = kActProcess; }
    return apply_action<I, kTrusted>(s, kind, primary, secondary);
}

// ---- packed native state: 4 x u64 = 32 B ---------------------------------------------------------------
// w0 = black | meta << 36, w1 = white, w2 = marks_black, w3 = marks_white (upper 28 bits of w1..w3 free).
// meta (28 bits): phase:3 | white_to_move:1 | forced:2 | pm_req:2 | pm_rem:2 | pc_req:2 | pc_rem:2 |
//                 move_count:8 | moves_since_capture:6
struct __attribute__((aligned(16))) Packed { uint64_t w[4]; };

LZ_HD Packed pack(const State<int>& s) {
    uint64_t m = (uint64_t)(s.phase & 7) | ((uint64_t)(s.player == -1 ? 1 : 0) << 3) | ((uint64_t)(s.forced & 3) << 4) |
                 ((uint64_t)(s.pm_req & 3) << 6) | ((uint64_t)(s.pm_rem & 3) << 8) | ((uint64_t)(s.pc_req & 3) << 10) |
                 ((uint64_t)(s.pc_rem & 3) << 12) | ((uint64_t)(s.move_count & 255) << 14) |
                 ((uint64_t)(s.msc & 63) << 22);
    Packed p;
    p.w[0] = (s.black & kFull) | (m << 36);
    p.w[1] = s.white & kFull; p.w[2] = s.mb & kFull; p.w[3] = s.mw & kFull;
    return p;
}
LZ_HD void unpack(const Packed& p, State<int>& s) {
    const uint64_t m = p.w[0] >> 36;
    s.black = p.w[0] & kFull; s.white = p.w[1] & kFull; s.mb = p.w[2] & kFull; s.mw = p.w[3] & kFull;
    s.other = 0;
    s.phase = (int)(m & 7); s.player = ((m >> 3) & 1) ? -1 : 1; s.forced = (int)((m >> 4) & 3);
    s.pm_req = (int)((m >> 6) & 3); s.pm_rem = (int)((m >> 8) & 3);
    s.pc_req = (int)((m >> 10) & 3); s.pc_rem = (int)((m >> 12) & 3);
    s.move_count = (int)((m >> 14) & 255); s.msc = (int)((m >> 22) & 63);
}

// ---- counter-based RNG for the playout workload (same definition as oracle/lz_oracle.c:or_playout_pick) --
LZ_HD uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
LZ_HD uint32_t playout_pick(uint64_t seed, uint64_t game, uint32_t ply, uint32_t n) {
    uint64_t h = mix64(mix64(seed ^ (game * 0xD1342543DE82EF95ULL)) + (uint64_t)ply);
    return (uint32_t)(((h >> 32) * (uint64_t)n) >> 32);
}

// FNV-1a over the reference byte layout, identical to oracle/lz_oracle.c:or_state_hash
LZ_HD uint64_t state_hash(const State<int>& s) {
    uint64_t h = 0xCBF29CE484222325ULL;
#define LZ_MIXB(v) do { h ^= (uint64_t)((uint32_t)(v) & 0xFFu); h *= 0x100000001B3ULL; } while (0)
    for (int c = 0; c < 36; ++c) LZ_MIXB(((s.black >> c) & 1) ? 1u : ((s.white >> c) & 1) ? 0xFFu : 0u);
    for (int c = 0; c < 36; ++c) LZ_MIXB((uint32_t)((s.mb >> c) & 1));
    for (int c = 0; c < 36; ++c) LZ_MIXB((uint32_t)((s.mw >> c) & 1));
    LZ_MIXB(s.phase); LZ_MIXB(s.player); LZ_MIXB(s.pm_req); LZ_MIXB(s.pm_rem);
    LZ_MIXB(s.pc_req); LZ_MIXB(s.pc_rem); LZ_MIXB(s.forced); LZ_MIXB(s.move_count); LZ_MIXB(s.msc);
#undef LZ_MIXB
    return h;
}

#if defined(__CUDACC__)
// ---- device-only helpers: the legal set as a 220-bit mask, k-th legal action by popcount + fns ---------------------
// 16 cells (bits 0..15) -> bits 0, 4, 8, ..., 60
__device__ __forceinline__ uint64_t spread16x4(uint64_t x) {
    x = (x | (x << 24)) & 0x000000FF000000FFULL;
    x = (x | (x << 12)) & 0x000F000F000F000FULL;
    x = (x | (x << 6)) & 0x0303030303030303ULL;
    x = (x | (x << 3)) & 0x1111111111111111ULL;
    return x;
}
__device__ __forceinline__ void legal_to_words(const Legal& L, uint64_t w[4]) {
    // bit a of word a/64, a in [0,220): place 0..35 | movement 36..179 (from*4+dir) | select 180..215 | 216
    // movement bits m = from * 4 + dir as a 144-bit vector (m0: cells 0..15, m1: 16..31, m2: 32..35), branch-free
    uint64_t m0 = 0, m1 = 0, m2 = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        m0 |= spread16x4(L.mv[d] & 0xFFFFULL) << d;
        m1 |= spread16x4((L.mv[d] >> 16) & 0xFFFFULL) << d;
        m2 |= spread16x4((L.mv[d] >> 32) & 0xFULL) << d;
    }
    w[0] = L.place | (m0 << 36);
    w[1] = (m0 >> 28) | (m1 << 36);
    w[2] = (m1 >> 28) | (m2 << 36) | ((L.sel & 0xFFFULL) << 52);   // m2: 16 bits -> 164..179; select cells 0..11 -> 180..191
    w[3] = (L.sel >> 12) | (L.process ? (1ULL << (216 - 192)) : 0ULL);   // cells 12..35 -> 192..215; process 216
}

// k-th legal action in ascending index order (0 <= k < legal_count) from the 220-bit mask words: prefix popcounts over the
// seven 32-bit pieces, then find-n-th-set (fns) inside the piece -- ~40 instructions whatever the phase, where
// lz::legal_kth peels bits one at a time.
__device__ __forceinline__ int legal_kth_words(const uint64_t (&w)[4], int k) {
    uint32_t p[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) p[j] = (uint32_t)(w[j >> 1] >> (32 * (j & 1)));
    int base = 0, piece = 0;
    uint32_t word = p[0];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int c = __popc(p[j]);
        if (piece == j && k >= base + c) { base += c; piece = j + 1; word = p[j + 1]; }
    }
    return 32 * piece + (int)__fns(word, 0, k - base + 1);
}

#endif

// Config-2 workload step loop: legal set (scalar-engine semantics) -> counter-based uniform pick -> apply ->
// terminal check, for up to max_steps plies.  res: 2 = still running, else result_from_black.
template <bool kHash>
LZ_HD void playout_advance(State<int>& s, int& ply, int& res, uint64_t& h, uint64_t seed, uint64_t game,
                           int max_steps, int max_game_plies) {
    for (int step = 0; step < max_steps; ++step) {
        if (game_over(s)) { res = winner(s); return; }
        if (ply >= max_game_plies) { res = 0; return; }
        Legal L;
        legal_actions<int, true>(s, L, true);
        const int n = legal_count(L);
        if (n == 0) { res = -s.player; return; }                  // module.cpp:733-735: side to move loses
        const int k = (int)playout_pick(seed, game, (uint32_t)ply, (uint32_t)n);
#if defined(__CUDA_ARCH__)
        uint64_t lw[4];                       // branch-free on the device: threads of a warp pick different ranks
        legal_to_words(L, lw);
        const int a = legal_kth_words(lw, k);
#else
        const int a = legal_kth(L, k);
#endif
        apply_index<int, true>(s, a);         // the action comes from this state's legal set
        ++ply;
        if (kHash) h = mix64(h ^ state_hash(s));
    }
}

}  // namespace lz
