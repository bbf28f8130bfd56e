// TEST-ONLY host build of liuzhou_b200/csrc/lz_rules.cuh (the bitboard rule engine the CUDA kernels use).
// Lets the CPU test-suite check the bit logic against the oracle without a GPU.  Not part of the product:
// liuzhou_b200 never loads this library and has no CPU execution path.
#include <stdint.h>
#include <string.h>

#include "../../liuzhou_b200/csrc/lz_rules.cuh"

using namespace lz;

namespace {
template <typename I>
void load(State<I>& s, const int8_t* board, const uint8_t* mb, const uint8_t* mw, const int64_t* sc) {
    s.black = s.white = s.other = s.mb = s.mw = 0;
    for (int c = 0; c < 36; ++c) {
        if (board[c] == 1) s.black |= bit(c);
        else if (board[c] == -1) s.white |= bit(c);
        else if (board[c] != 0) s.other |= bit(c);
        if (mb[c]) s.mb |= bit(c);
        if (mw[c]) s.mw |= bit(c);
    }
    s.phase = (I)sc[0]; s.player = (I)sc[1]; s.pm_req = (I)sc[2]; s.pm_rem = (I)sc[3]; s.pc_req = (I)sc[4];
    s.pc_rem = (I)sc[5]; s.forced = (I)sc[6]; s.move_count = (I)sc[7]; s.msc = (I)sc[8];
}
template <typename I>
void store(const State<I>& s, int8_t* board, uint8_t* mb, uint8_t* mw, int64_t* sc) {
    for (int c = 0; c < 36; ++c) {
        if (s.black & bit(c)) board[c] = 1;
        else if (s.white & bit(c)) board[c] = -1;
        else if (!(s.other & bit(c))) board[c] = 0;
        mb[c] = (s.mb >> c) & 1;
        mw[c] = (s.mw >> c) & 1;
    }
    sc[0] = s.phase; sc[1] = s.player; sc[2] = s.pm_req; sc[3] = s.pm_rem; sc[4] = s.pc_req; sc[5] = s.pc_rem;
    sc[6] = s.forced; sc[7] = s.move_count; sc[8] = s.msc;
}
}  // namespace

extern "C" {

// scalars: int64[n,9] = phase, player, pm_req, pm_rem, pc_req, pc_rem, forced, move_count, msc
void hs_legal(int64_t n, const int8_t* board, const uint8_t* mb, const uint8_t* mw, const int64_t* sc, int scalar,
              uint8_t* mask /*[n,220]*/, int32_t* meta /*[n,220,4]*/, int32_t* counts, int32_t* kth /*[n,220]*/,
              int32_t* rank /*[n,220]*/) {
    for (int64_t i = 0; i < n; ++i) {
        State<int> s;
        load(s, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
        Legal L;
        if (scalar) legal_actions<int, true>(s, L, true); else legal_actions<int, false>(s, L, true);
        const int cnt = legal_count(L);
        counts[i] = cnt;
        for (int a = 0; a < 220; ++a) {
            const bool ok = legal_test(L, a);
            mask[i * 220 + a] = ok;
            int k = -1, p = -1, se = -1, e = -1;
            if (ok) action_code(L, a, k, p, se, e);
            int32_t* m = meta + (i * 220 + a) * 4;
            m[0] = k; m[1] = p; m[2] = se; m[3] = e;
            rank[i * 220 + a] = legal_rank(L, a);
            kth[i * 220 + a] = a < cnt ? legal_kth(L, a) : -1;
        }
    }
}

void hs_apply(int64_t n, int8_t* board, uint8_t* mb, uint8_t* mw, int64_t* sc, const int32_t* codes, uint8_t* applied) {
    for (int64_t i = 0; i < n; ++i) {
        State<long long> s;
        load(s, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
        applied[i] = apply_action(s, codes[i * 4], codes[i * 4 + 1], codes[i * 4 + 2]);
        store(s, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
    }
}

void hs_apply_index(int64_t n, int8_t* board, uint8_t* mb, uint8_t* mw, int64_t* sc, const int32_t* actions) {
    for (int64_t i = 0; i < n; ++i) {
        State<int> s;
        load(s, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
        Packed p = pack(s);              // exercise the packed round trip too
        State<int> t;
        unpack(p, t);
        apply_index(t, actions[i]);
        store(t, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
    }
}

void hs_status(int64_t n, const int8_t* board, const uint8_t* mb, const uint8_t* mw, const int64_t* sc, int32_t* win,
               uint8_t* over, uint64_t* hash) {
    for (int64_t i = 0; i < n; ++i) {
        State<int> s;
        load(s, board + i * 36, mb + i * 36, mw + i * 36, sc + i * 9);
        win[i] = winner(s);
        over[i] = game_over(s);
        hash[i] = state_hash(s);
    }
}

int hs_playout(uint64_t seed, uint64_t game, int max_plies, int chunk, int* result, uint64_t* hash, int8_t* board,
               uint8_t* mb, uint8_t* mw, int64_t* sc) {
    State<int> s;
    set_initial(s);
    int ply = 0, res = 2;
    uint64_t h = 0;
    while (res == 2) {   // advance in chunks like repeated kernel launches do, via the packed form
        Packed p = pack(s);
        unpack(p, s);
        playout_advance<true>(s, ply, res, h, seed, game, chunk, max_plies);
    }
    *result = res; *hash = h;
    store(s, board, mb, mw, sc);
    return ply;
}

}  // extern "C"
