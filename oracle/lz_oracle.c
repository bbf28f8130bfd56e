/*
 * lz_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See lz_oracle.h.
 *
 * Plain byte-per-cell C restatement of the reference's rule engine, legal-mask encoder, atomic
 * move application, model-input encoding and root-PUCT loop.  Each function cites the reference
 * file:line it follows.  Loops are kept in the reference's order on purpose.
 */
#include "lz_oracle.h"

#include <math.h>
#include <string.h>

/* directions up, down, left, right: fast_legal_mask_common.hpp:24-29 */
static const int kDr[4] = {-1, 1, 0, 0};
static const int kDc[4] = {0, 0, -1, 1};

#define IDX(r, c) ((r) * OR_SIZE + (c))

void or_load(const or_batch *b, int64_t i, or_state *s) {
    memcpy(s->board, b->board + i * OR_CELLS, OR_CELLS);
    memcpy(s->marks_black, b->marks_black + i * OR_CELLS, OR_CELLS);
    memcpy(s->marks_white, b->marks_white + i * OR_CELLS, OR_CELLS);
    s->phase = b->phase[i];
    s->current_player = b->current_player[i];
    s->pending_marks_required = b->pending_marks_required[i];
    s->pending_marks_remaining = b->pending_marks_remaining[i];
    s->pending_captures_required = b->pending_captures_required[i];
    s->pending_captures_remaining = b->pending_captures_remaining[i];
    s->forced_removals_done = b->forced_removals_done[i];
    s->move_count = b->move_count ? b->move_count[i] : 0;
    s->moves_since_capture = b->moves_since_capture ? b->moves_since_capture[i] : 0;
}

void or_store(const or_state *s, or_batch *b, int64_t i) {
    memcpy(b->board + i * OR_CELLS, s->board, OR_CELLS);
    memcpy(b->marks_black + i * OR_CELLS, s->marks_black, OR_CELLS);
    memcpy(b->marks_white + i * OR_CELLS, s->marks_white, OR_CELLS);
    b->phase[i] = s->phase;
    b->current_player[i] = s->current_player;
    b->pending_marks_required[i] = s->pending_marks_required;
    b->pending_marks_remaining[i] = s->pending_marks_remaining;
    b->pending_captures_required[i] = s->pending_captures_required;
    b->pending_captures_remaining[i] = s->pending_captures_remaining;
    b->forced_removals_done[i] = s->forced_removals_done;
    if (b->move_count) b->move_count[i] = s->move_count;
    if (b->moves_since_capture) b->moves_since_capture[i] = s->moves_since_capture;
}

/* GpuStateBatch.initial: v1/python/mcts_gpu.py:123-145 */
void or_initial(or_state *s) {
    memset(s, 0, sizeof(*s));
    s->phase = OR_PHASE_PLACEMENT;
    s->current_player = 1;
}

static int is_marked(const uint8_t *marked, int idx) { return marked != 0 && marked[idx] != 0; }

/* rule_engine.cpp:57-89 / fast_legal_mask.cpp:17-46 */
int or_check_squares(const int8_t *board, const uint8_t *marked, int r, int c, int player_value) {
    static const int offsets[2] = {0, -1};
    for (int a = 0; a < 2; ++a) {
        for (int b = 0; b < 2; ++b) {
            int rr = r + offsets[a];
            int cc = c + offsets[b];
            if (rr >= 0 && rr < OR_SIZE - 1 && cc >= 0 && cc < OR_SIZE - 1) {
                int ok = 1;
                const int cells[4] = {IDX(rr, cc), IDX(rr, cc + 1), IDX(rr + 1, cc), IDX(rr + 1, cc + 1)};
                for (int k = 0; k < 4; ++k) {
                    if (board[cells[k]] != player_value || is_marked(marked, cells[k])) {
                        ok = 0;
                        break;
                    }
                }
                if (ok) return 1;
            }
        }
    }
    return 0;
}

/* rule_engine.cpp:91-136 / fast_legal_mask.cpp:48-95.  NB: (r,c) itself is counted without looking
 * at its own mark or even its colour. */
int or_check_lines(const int8_t *board, const uint8_t *marked, int r, int c, int player_value) {
    int count = 1;
    for (int dc = c - 1; dc >= 0; --dc) {
        int idx = IDX(r, dc);
        if (board[idx] == player_value && !is_marked(marked, idx)) ++count; else break;
    }
    for (int dc = c + 1; dc < OR_SIZE; ++dc) {
        int idx = IDX(r, dc);
        if (board[idx] == player_value && !is_marked(marked, idx)) ++count; else break;
    }
    if (count >= 6) return 1;
    count = 1;
    for (int dr = r - 1; dr >= 0; --dr) {
        int idx = IDX(dr, c);
        if (board[idx] == player_value && !is_marked(marked, idx)) ++count; else break;
    }
    for (int dr = r + 1; dr < OR_SIZE; ++dr) {
        int idx = IDX(dr, c);
        if (board[idx] == player_value && !is_marked(marked, idx)) ++count; else break;
    }
    return count >= 6;
}

/* rule_engine.cpp:194-208 / fast_legal_mask.cpp:97-108 */
int or_is_piece_in_shape(const int8_t *board, const uint8_t *marked, int r, int c, int player_value) {
    if (board[IDX(r, c)] != player_value) return 0;
    return or_check_squares(board, marked, r, c, player_value) ||
           or_check_lines(board, marked, r, c, player_value);
}

/* fast_legal_mask.cpp:110-129: keep candidates that are not in a shape; if none, keep all. */
static int prefer_normal(const int8_t *board, const uint8_t *marked, const int *cand, int n,
                         int player_value, int *out) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
        int idx = cand[i];
        if (!or_is_piece_in_shape(board, marked, idx / OR_SIZE, idx % OR_SIZE, player_value)) out[m++] = idx;
    }
    if (m > 0) return m;
    for (int i = 0; i < n; ++i) out[i] = cand[i];
    return n;
}

static int collect_value(const int8_t *board, int value, const uint8_t *skip_marked, int *out) {
    int n = 0;
    for (int idx = 0; idx < OR_CELLS; ++idx) {
        if (board[idx] == value && !(skip_marked && skip_marked[idx])) out[n++] = idx;
    }
    return n;
}

static void set_meta(int32_t *meta, int64_t a, int32_t kind, int32_t p, int32_t s, int32_t e) {
    meta[a * 4 + 0] = kind; meta[a * 4 + 1] = p; meta[a * 4 + 2] = s; meta[a * 4 + 3] = e;
}

/* fast_legal_mask.cpp:253-418 */
void or_encode_actions(int64_t B, const or_batch *in, int64_t placement_dim, int64_t movement_dim,
                       int64_t selection_dim, int64_t auxiliary_dim, uint8_t *mask, int32_t *metadata) {
    const int64_t total = placement_dim + movement_dim + selection_dim + auxiliary_dim;
    memset(mask, 0, (size_t)(B * total));
    for (int64_t i = 0; i < B * total * 4; ++i) metadata[i] = -1;
    for (int64_t b = 0; b < B; ++b) {
        const int8_t *board = in->board + b * OR_CELLS;
        const uint8_t *mb = in->marks_black + b * OR_CELLS;
        const uint8_t *mw = in->marks_white + b * OR_CELLS;
        const int phase = (int)in->phase[b];
        const int cur = (int)in->current_player[b];
        const int pm_rem = (int)in->pending_marks_remaining[b];
        const int pc_rem = (int)in->pending_captures_remaining[b];
        const int forced = (int)in->forced_removals_done[b];
        uint8_t *m = mask + b * total;
        int32_t *meta = metadata + b * total * 4;

        if (phase == OR_PHASE_PLACEMENT) {                       /* :326-337 */
            for (int idx = 0; idx < OR_CELLS; ++idx) {
                if (board[idx] == 0) { m[idx] = 1; set_meta(meta, idx, OR_ACT_PLACE, idx, -1, -1); }
            }
        }
        int has_movement = 0;
        if (phase == OR_PHASE_MOVEMENT) {                        /* :339-368 (CUDA adds the bound :337) */
            for (int r = 0; r < OR_SIZE; ++r) for (int c = 0; c < OR_SIZE; ++c) {
                int base = IDX(r, c);
                if (board[base] != cur) continue;
                for (int d = 0; d < 4; ++d) {
                    int nr = r + kDr[d], nc = c + kDc[d];
                    if (nr >= 0 && nr < OR_SIZE && nc >= 0 && nc < OR_SIZE && board[IDX(nr, nc)] == 0) {
                        int64_t a = placement_dim + base * 4 + d;
                        if (a < placement_dim + movement_dim) {
                            m[a] = 1; set_meta(meta, a, OR_ACT_MOVE, base, d, IDX(nr, nc));
                            has_movement = 1;
                        }
                    }
                }
            }
        }
        int cand[OR_CELLS], sel[OR_CELLS], n = 0, kind = 0;
        const uint8_t *opp_marked = (cur == 1) ? mw : mb;        /* :381,:391 */
        if (phase == OR_PHASE_MARK) {                            /* :204-226 */
            int nc_ = collect_value(board, -cur, opp_marked, cand);
            n = prefer_normal(board, opp_marked, cand, nc_, -cur, sel);
            if (pm_rem <= 0) n = 0;
            kind = OR_ACT_MARK;
        } else if (phase == OR_PHASE_CAPTURE) {                  /* :228-249 */
            int nc_ = collect_value(board, -cur, 0, cand);
            n = (pc_rem > 0) ? prefer_normal(board, opp_marked, cand, nc_, -cur, sel) : 0;
            kind = OR_ACT_CAPTURE;
        } else if (phase == OR_PHASE_FORCED) {                   /* :131-147 */
            if (forced < 2) {
                int value = forced == 0 ? 1 : -1;
                int nc_ = collect_value(board, value, 0, cand);
                n = prefer_normal(board, 0, cand, nc_, value, sel);
            }
            kind = OR_ACT_FORCED;
        } else if (phase == OR_PHASE_COUNTER) {                  /* :149-162 */
            int nc_ = collect_value(board, -cur, 0, cand);
            n = prefer_normal(board, 0, cand, nc_, -cur, sel);
            kind = OR_ACT_COUNTER;
        } else if (phase == OR_PHASE_MOVEMENT && !has_movement) { /* :164-177,:405-408 */
            int nc_ = collect_value(board, -cur, 0, cand);
            n = prefer_normal(board, 0, cand, nc_, -cur, sel);
            kind = OR_ACT_NOMOVES;
        }
        for (int i = 0; i < n; ++i) {                            /* emit_selection :370-379 */
            int idx = sel[i];
            if (idx >= 0 && idx < selection_dim) {
                int64_t a = placement_dim + movement_dim + idx;
                m[a] = 1; set_meta(meta, a, kind, idx, -1, -1);
            }
        }
        if (phase == OR_PHASE_REMOVAL && auxiliary_dim > 0) {    /* :410-414 */
            int64_t a = placement_dim + movement_dim + selection_dim;
            m[a] = 1; set_meta(meta, a, OR_ACT_PROCESS, -1, -1, -1);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Atomic action application, CUDA-kernel semantics (fast_apply_moves_cuda.cu).
 * ---------------------------------------------------------------------------------------------- */
static int count_value(const int8_t *board, int v) {
    int n = 0;
    for (int i = 0; i < OR_CELLS; ++i) if (board[i] == v) ++n;
    return n;
}
static int board_full(const int8_t *board) {
    for (int i = 0; i < OR_CELLS; ++i) if (board[i] == 0) return 0;
    return 1;
}
static int any_marked(const uint8_t *m) {
    for (int i = 0; i < OR_CELLS; ++i) if (m[i]) return 1;
    return 0;
}
/* fast_apply_moves_cuda.cu:183-199: line beats square */
static int detect_shape(const int8_t *board, const uint8_t *marked, int r, int c, int pv) {
    int sq = or_check_squares(board, marked, r, c, pv);
    int ln = or_check_lines(board, marked, r, c, pv);
    if (ln) return 2;
    if (sq) return 1;
    return 0;
}
/* fast_apply_moves_cuda.cu:290-306 */
static int has_unmarked_normal_piece(const int8_t *board, const uint8_t *marked, int pv) {
    for (int idx = 0; idx < OR_CELLS; ++idx) {
        if (board[idx] == pv) {
            if (!or_is_piece_in_shape(board, marked, idx / OR_SIZE, idx % OR_SIZE, pv) && !is_marked(marked, idx))
                return 1;
        }
    }
    return 0;
}

/* :239-288 (move_count bumped only when the placement is applied) */
static int apply_placement(or_state *s, int cell) {
    if (s->phase != OR_PHASE_PLACEMENT) return 0;
    if (cell < 0 || cell >= OR_CELLS) return 0;
    if (s->board[cell] != 0) return 0;
    const uint8_t *opp_marked = s->current_player == 1 ? s->marks_white : s->marks_black;
    if (opp_marked[cell]) return 0;
    s->board[cell] = (int8_t)s->current_player;
    uint8_t *own = s->current_player == 1 ? s->marks_black : s->marks_white;
    if (!own[cell]) {
        int shape = detect_shape(s->board, own, cell / OR_SIZE, cell % OR_SIZE, (int)s->current_player);
        if (shape) {
            s->pending_marks_required = s->pending_marks_remaining = (shape == 2) ? 2 : 1;
            s->phase = OR_PHASE_MARK;
            s->move_count += 1;
            return 1;
        }
    }
    s->pending_marks_required = s->pending_marks_remaining = 0;
    if (board_full(s->board)) {
        s->phase = OR_PHASE_REMOVAL;
    } else {
        s->current_player = -s->current_player;
        s->phase = OR_PHASE_PLACEMENT;
    }
    s->move_count += 1;
    return 1;
}

/* :308-348 */
static int apply_mark(or_state *s, int cell) {
    if (s->phase != OR_PHASE_MARK || s->pending_marks_remaining <= 0) return 0;
    if (cell < 0 || cell >= OR_CELLS) return 0;
    int opp = (int)(-s->current_player);
    uint8_t *opp_marked = opp == -1 ? s->marks_white : s->marks_black;
    if (s->board[cell] != opp || opp_marked[cell]) return 0;
    int in_shape = or_is_piece_in_shape(s->board, opp_marked, cell / OR_SIZE, cell % OR_SIZE, opp);
    if (in_shape && has_unmarked_normal_piece(s->board, opp_marked, opp)) return 0;
    opp_marked[cell] = 1;
    s->pending_marks_remaining -= 1;
    if (s->pending_marks_remaining > 0) return 1;
    s->pending_marks_required = s->pending_marks_remaining = 0;
    if (board_full(s->board)) {
        s->phase = OR_PHASE_REMOVAL;
    } else {
        s->current_player = -s->current_player;
        s->phase = OR_PHASE_PLACEMENT;
    }
    return 1;
}

/* :201-237 -- NB counts a marked EMPTY cell as "removed" (the scalar engine does not; unreachable). */
static int apply_process_removal(or_state *s) {
    /* the CUDA helper does not check the phase at all */
    int ab = any_marked(s->marks_black), aw = any_marked(s->marks_white);
    if (!ab && !aw) {
        s->phase = OR_PHASE_FORCED;
        s->current_player = -1;
        s->forced_removals_done = 0;
        return 1;
    }
    int removed = 0;
    for (int idx = 0; idx < OR_CELLS; ++idx) {
        if (s->marks_black[idx]) { s->board[idx] = 0; ++removed; }
        else if (s->marks_white[idx]) { s->board[idx] = 0; ++removed; }
    }
    memset(s->marks_black, 0, OR_CELLS);
    memset(s->marks_white, 0, OR_CELLS);
    if (removed > 0) {
        s->phase = OR_PHASE_MOVEMENT;
        s->current_player = -1;
    }
    return 1;
}

/* :350-386 */
static int apply_forced(or_state *s, int cell) {
    if (s->phase != OR_PHASE_FORCED || cell < 0 || cell >= OR_CELLS) return 0;
    int r = cell / OR_SIZE, c = cell % OR_SIZE;
    if (s->forced_removals_done == 0) {
        if (s->current_player != -1 || s->board[cell] != 1) return 0;
        if (or_is_piece_in_shape(s->board, 0, r, c, 1)) return 0;
        s->board[cell] = 0;
        s->forced_removals_done = 1;
        s->current_player = 1;
        return 1;
    } else if (s->forced_removals_done == 1) {
        if (s->current_player != 1 || s->board[cell] != -1) return 0;
        if (or_is_piece_in_shape(s->board, 0, r, c, -1)) return 0;
        s->board[cell] = 0;
        s->forced_removals_done = 2;
        s->phase = OR_PHASE_MOVEMENT;
        s->current_player = -1;
        return 1;
    }
    return 0;
}

/* :388-416 */
static int apply_no_moves(or_state *s, int cell) {
    if (s->phase != OR_PHASE_MOVEMENT || cell < 0 || cell >= OR_CELLS) return 0;
    int opp = (int)(-s->current_player);
    if (s->board[cell] != opp) return 0;
    int in_shape = or_is_piece_in_shape(s->board, 0, cell / OR_SIZE, cell % OR_SIZE, opp);
    if (in_shape && has_unmarked_normal_piece(s->board, 0, opp)) return 0;
    s->board[cell] = 0;
    if (count_value(s->board, opp) < 4) return 1;
    s->phase = OR_PHASE_COUNTER;
    s->current_player = -s->current_player;
    return 1;
}

/* :418-456 */
static int apply_capture(or_state *s, int cell) {
    if (s->phase != OR_PHASE_CAPTURE || s->pending_captures_remaining <= 0 || cell < 0 || cell >= OR_CELLS)
        return 0;
    int opp = (int)(-s->current_player);
    const uint8_t *opp_marked = opp == -1 ? s->marks_white : s->marks_black;
    if (s->board[cell] != opp) return 0;
    int in_shape = or_is_piece_in_shape(s->board, opp_marked, cell / OR_SIZE, cell % OR_SIZE, opp);
    if (in_shape && has_unmarked_normal_piece(s->board, opp_marked, opp)) return 0;
    s->board[cell] = 0;
    s->pending_captures_remaining -= 1;
    if (count_value(s->board, opp) < 4 || s->pending_captures_remaining > 0) return 1;
    s->pending_captures_required = s->pending_captures_remaining = 0;
    s->current_player = -s->current_player;
    s->phase = OR_PHASE_MOVEMENT;
    return 1;
}

/* :458-486 */
static int apply_counter(or_state *s, int cell) {
    if (s->phase != OR_PHASE_COUNTER || cell < 0 || cell >= OR_CELLS) return 0;
    int stuck = (int)(-s->current_player);
    if (s->board[cell] != stuck) return 0;
    int in_shape = or_is_piece_in_shape(s->board, 0, cell / OR_SIZE, cell % OR_SIZE, stuck);
    if (in_shape && has_unmarked_normal_piece(s->board, 0, stuck)) return 0;
    s->board[cell] = 0;
    if (count_value(s->board, stuck) < 4) return 1;
    s->phase = OR_PHASE_MOVEMENT;
    s->current_player = -s->current_player;
    return 1;
}

/* :488-546 (from_cell is not range-checked by the reference; out-of-range is UB there, a no-op here) */
static int apply_movement(or_state *s, int from_cell, int dir) {
    if (s->phase != OR_PHASE_MOVEMENT || dir < 0 || dir >= 4) return 0;
    if (from_cell < 0 || from_cell >= OR_CELLS) return 0;
    int rf = from_cell / OR_SIZE, cf = from_cell % OR_SIZE;
    int rt = rf + kDr[dir], ct = cf + kDc[dir];
    if (rt < 0 || rt >= OR_SIZE || ct < 0 || ct >= OR_SIZE) return 0;
    int to = IDX(rt, ct);
    if (s->board[from_cell] != s->current_player || s->board[to] != 0) return 0;
    s->board[to] = s->board[from_cell];
    s->board[from_cell] = 0;
    int shape = detect_shape(s->board, 0, rt, ct, (int)s->current_player);
    if (shape) {
        s->pending_captures_required = s->pending_captures_remaining = (shape == 2) ? 2 : 1;
        s->phase = OR_PHASE_CAPTURE;
        return 1;
    }
    s->pending_captures_required = s->pending_captures_remaining = 0;
    s->current_player = -s->current_player;
    return 1;
}

/* Body of BatchApplyMovesKernel after the parent copy: fast_apply_moves_cuda.cu:624-743 */
int or_apply_action(or_state *s, int32_t kind, int32_t primary, int32_t secondary) {
    const int64_t phase_before = s->phase;
    const int64_t msc_before = s->moves_since_capture;
    int old_total = 0;
    for (int i = 0; i < OR_CELLS; ++i) if (s->board[i] != 0) ++old_total;
    int applied = 0;
    switch (kind) {
        case OR_ACT_PLACE:   applied = apply_placement(s, primary); break;
        case OR_ACT_MARK:    applied = apply_mark(s, primary); s->move_count += 1; break;
        case OR_ACT_PROCESS: applied = apply_process_removal(s); s->move_count += 1; break;
        case OR_ACT_FORCED:  applied = apply_forced(s, primary); s->move_count += 1; break;
        case OR_ACT_MOVE:    applied = apply_movement(s, primary, secondary); s->move_count += 1; break;
        case OR_ACT_NOMOVES: applied = apply_no_moves(s, primary); s->move_count += 1; break;
        case OR_ACT_CAPTURE: applied = apply_capture(s, primary); s->move_count += 1; break;
        case OR_ACT_COUNTER: applied = apply_counter(s, primary); s->move_count += 1; break;
        default: break;
    }
    if (phase_before == OR_PHASE_PLACEMENT || phase_before == OR_PHASE_MARK) {   /* :728-743 */
        s->moves_since_capture = 0;
    } else {
        int new_total = 0;
        for (int i = 0; i < OR_CELLS; ++i) if (s->board[i] != 0) ++new_total;
        s->moves_since_capture = (new_total < old_total) ? 0 : msc_before + 1;
    }
    return applied;
}

void or_batch_apply_moves(int64_t B, const or_batch *in, int64_t N, const int32_t *codes,
                          const int64_t *parents, or_batch *out, uint8_t *applied_flags) {
    for (int64_t i = 0; i < N; ++i) {
        int64_t p = parents[i];
        if (p < 0 || p >= B) { if (applied_flags) applied_flags[i] = 0; continue; }
        or_state s;
        or_load(in, p, &s);
        int ok = or_apply_action(&s, codes[i * 4 + 0], codes[i * 4 + 1], codes[i * 4 + 2]);
        or_store(&s, out, i);
        if (applied_flags) applied_flags[i] = (uint8_t)ok;
    }
}

void or_batch_apply_moves_inplace(int64_t B, or_batch *st, int64_t N, const int32_t *codes,
                                  const int64_t *slots) {
    for (int64_t i = 0; i < N; ++i) {
        int64_t p = slots[i];
        if (p < 0 || p >= B) continue;
        or_state s;
        or_load(st, p, &s);
        or_apply_action(&s, codes[i * 4 + 0], codes[i * 4 + 1], codes[i * 4 + 2]);
        or_store(&s, st, p);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Scalar-engine semantics
 * ---------------------------------------------------------------------------------------------- */
/* game_state.cpp:59-75 */
int or_winner(const or_state *s) {
    if (s->phase != OR_PHASE_MOVEMENT && s->phase != OR_PHASE_CAPTURE && s->phase != OR_PHASE_COUNTER) return 0;
    if (count_value(s->board, 1) < 4) return -1;
    if (count_value(s->board, -1) < 4) return 1;
    return 0;
}
/* game_state.cpp:77-79, game_state.hpp:128-131 */
int or_is_game_over(const or_state *s) {
    return or_winner(s) != 0 || s->move_count >= 144 || s->moves_since_capture >= 36;
}

static void put_code(int32_t *codes, int n, int32_t k, int32_t p, int32_t s_, int32_t e) {
    if (!codes) return;
    codes[n * 4 + 0] = k; codes[n * 4 + 1] = p; codes[n * 4 + 2] = s_; codes[n * 4 + 3] = e;
}

/* move_generator.cpp:242-297 + rule_engine.cpp generators; indices per portable_mcts.cpp:170-198.
 * Each generator emits cells in row-major order, so indices come out ascending. */
int or_legal_actions(const or_state *s, int *indices, int32_t *codes) {
    int n = 0;
    if (or_is_game_over(s)) return 0;                                   /* :243-245 */
    const int8_t *board = s->board;
    const int cur = (int)s->current_player;
    switch ((int)s->phase) {
        case OR_PHASE_PLACEMENT:                                        /* rule_engine.cpp:210-224 */
            for (int idx = 0; idx < OR_CELLS; ++idx)
                if (board[idx] == 0) { put_code(codes, n, OR_ACT_PLACE, idx, -1, -1); indices[n++] = idx; }
            return n;
        case OR_PHASE_MARK: {                                           /* rule_engine.cpp:271-299 */
            if (s->pending_marks_remaining <= 0) return 0;
            const uint8_t *om = cur == 1 ? s->marks_white : s->marks_black;
            int normal_any = 0;
            uint8_t normal[OR_CELLS] = {0};
            for (int idx = 0; idx < OR_CELLS; ++idx)
                if (board[idx] == -cur && !or_is_piece_in_shape(board, om, idx / OR_SIZE, idx % OR_SIZE, -cur)) {
                    normal[idx] = 1; normal_any = 1;
                }
            int pool[OR_CELLS], m = 0;
            if (normal_any) { for (int idx = 0; idx < OR_CELLS; ++idx) if (normal[idx] && !om[idx]) pool[m++] = idx; }
            else { for (int idx = 0; idx < OR_CELLS; ++idx) if (board[idx] == -cur && !om[idx]) pool[m++] = idx; }
            if (m == 0) for (int idx = 0; idx < OR_CELLS; ++idx) if (board[idx] == -cur && !om[idx]) pool[m++] = idx;
            for (int i = 0; i < m; ++i) { put_code(codes, n, OR_ACT_MARK, pool[i], -1, -1); indices[n++] = 180 + pool[i]; }
            return n;
        }
        case OR_PHASE_REMOVAL:                                          /* move_generator.cpp:266-268 */
            put_code(codes, n, OR_ACT_PROCESS, -1, -1, -1); indices[n++] = 216;
            return n;
        case OR_PHASE_FORCED: {                                         /* move_generator.cpp:143-170 (no fallback) */
            int value;
            if (s->forced_removals_done == 0) value = 1;
            else if (s->forced_removals_done == 1) value = -1;
            else return 0;
            for (int idx = 0; idx < OR_CELLS; ++idx)
                if (board[idx] == value && !or_is_piece_in_shape(board, 0, idx / OR_SIZE, idx % OR_SIZE, value)) {
                    put_code(codes, n, OR_ACT_FORCED, idx, -1, -1); indices[n++] = 180 + idx;
                }
            return n;
        }
        case OR_PHASE_MOVEMENT: {                                       /* move_generator.cpp:271-282 */
            for (int r = 0; r < OR_SIZE; ++r) for (int c = 0; c < OR_SIZE; ++c) {
                if (board[IDX(r, c)] != cur) continue;
                for (int d = 0; d < 4; ++d) {
                    int nr = r + kDr[d], nc = c + kDc[d];
                    if (nr >= 0 && nr < OR_SIZE && nc >= 0 && nc < OR_SIZE && board[IDX(nr, nc)] == 0) {
                        put_code(codes, n, OR_ACT_MOVE, IDX(r, c), d, IDX(nr, nc));
                        indices[n++] = 36 + IDX(r, c) * 4 + d;
                    }
                }
            }
            if (n > 0) return n;
            /* GenerateNoMovesOptions: move_generator.cpp:172-204 */
            int all[OR_CELLS], na = 0, nm[OR_CELLS], nn = 0;
            for (int idx = 0; idx < OR_CELLS; ++idx) if (board[idx] == -cur) {
                all[na++] = idx;
                if (!or_is_piece_in_shape(board, 0, idx / OR_SIZE, idx % OR_SIZE, -cur)) nm[nn++] = idx;
            }
            const int *t = nn ? nm : all; int tn = nn ? nn : na;
            for (int i = 0; i < tn; ++i) { put_code(codes, n, OR_ACT_NOMOVES, t[i], -1, -1); indices[n++] = 180 + t[i]; }
            return n;
        }
        case OR_PHASE_CAPTURE: {                                        /* rule_engine.cpp:445-463 */
            if (s->pending_captures_remaining <= 0) return 0;
            const uint8_t *om = cur == 1 ? s->marks_white : s->marks_black;
            int all[OR_CELLS], na = 0, nm[OR_CELLS], nn = 0;
            for (int idx = 0; idx < OR_CELLS; ++idx) if (board[idx] == -cur) {
                all[na++] = idx;
                if (!or_is_piece_in_shape(board, om, idx / OR_SIZE, idx % OR_SIZE, -cur)) nm[nn++] = idx;
            }
            const int *t = nn ? nm : all; int tn = nn ? nn : na;
            for (int i = 0; i < tn; ++i) { put_code(codes, n, OR_ACT_CAPTURE, t[i], -1, -1); indices[n++] = 180 + t[i]; }
            return n;
        }
        case OR_PHASE_COUNTER: {                                        /* move_generator.cpp:206-240 */
            int all[OR_CELLS], na = 0, nm[OR_CELLS], nn = 0;
            for (int idx = 0; idx < OR_CELLS; ++idx) if (board[idx] == -cur) {
                all[na++] = idx;
                if (!or_is_piece_in_shape(board, 0, idx / OR_SIZE, idx % OR_SIZE, -cur)) nm[nn++] = idx;
            }
            const int *t = nn ? nm : all; int tn = nn ? nn : na;
            for (int i = 0; i < tn; ++i) { put_code(codes, n, OR_ACT_COUNTER, t[i], -1, -1); indices[n++] = 180 + t[i]; }
            return n;
        }
        default: return 0;
    }
}

/* Scalar ApplyMove (move_generator.cpp:360-432) on top of the rule_engine.cpp Apply* functions.
 * Legality checks ("throws") are the scalar engine's, which are stricter than the CUDA no-op rules
 * only on unreachable inputs; here: apply with the shared transition code, then fix up the two
 * places where the scalar engine differs (process-removal `removed` counts only real pieces:
 * rule_engine.cpp:344-356; piece totals count +-1 only: move_generator.cpp:419-429). */
int or_apply_move_scalar(const or_state *s, int a, or_state *out) {
    *out = *s;
    int kind, primary = -1, secondary = -1;
    const int phase = (int)s->phase;
    if (a >= 0 && a < 36) { kind = OR_ACT_PLACE; primary = a; if (phase != OR_PHASE_PLACEMENT) return 0; }
    else if (a < 180) { kind = OR_ACT_MOVE; primary = (a - 36) / 4; secondary = (a - 36) % 4; if (phase != OR_PHASE_MOVEMENT) return 0; }
    else if (a < 216) {
        primary = a - 180;
        switch (phase) {
            case OR_PHASE_MARK: kind = OR_ACT_MARK; break;
            case OR_PHASE_CAPTURE: kind = OR_ACT_CAPTURE; break;
            case OR_PHASE_FORCED: kind = OR_ACT_FORCED; break;
            case OR_PHASE_COUNTER: kind = OR_ACT_COUNTER; break;
            case OR_PHASE_MOVEMENT: kind = OR_ACT_NOMOVES; break;
            default: return 0;
        }
    } else if (a == 216) { kind = OR_ACT_PROCESS; if (phase != OR_PHASE_REMOVAL) return 0; }
    else return 0;

    int old_total = count_value(s->board, 1) + count_value(s->board, -1);
    int applied;
    if (kind == OR_ACT_PROCESS) {
        /* rule_engine.cpp:332-366 */
        if (!any_marked(out->marks_black) && !any_marked(out->marks_white)) {
            out->phase = OR_PHASE_FORCED; out->current_player = -1; out->forced_removals_done = 0;
        } else {
            int removed = 0;
            for (int idx = 0; idx < OR_CELLS; ++idx)
                if ((out->marks_black[idx] || out->marks_white[idx]) && out->board[idx] != 0) { out->board[idx] = 0; ++removed; }
            memset(out->marks_black, 0, OR_CELLS);
            memset(out->marks_white, 0, OR_CELLS);
            if (removed > 0) { out->phase = OR_PHASE_MOVEMENT; out->current_player = -1; }
        }
        applied = 1;
    } else {
        or_state tmp = *s;
        applied = or_apply_action(&tmp, kind, primary, secondary);
        if (!applied) return 0;
        *out = tmp;
    }
    out->move_count = s->move_count + 1;                                  /* :418 */
    if (phase == OR_PHASE_PLACEMENT || phase == OR_PHASE_MARK) {
        out->moves_since_capture = 0;
    } else {
        int new_total = count_value(out->board, 1) + count_value(out->board, -1);
        out->moves_since_capture = new_total < old_total ? 0 : s->moves_since_capture + 1;
    }
    return applied;
}

/* ------------------------------------------------------------------------------------------------
 * states_to_model_input: v0/src/net/encoding.cpp:26-79
 * ---------------------------------------------------------------------------------------------- */
void or_states_to_model_input(int64_t B, const or_batch *in, float *out) {
    for (int64_t b = 0; b < B; ++b) {
        float *o = out + b * 11 * OR_CELLS;
        const int8_t *board = in->board + b * OR_CELLS;
        const uint8_t *mb = in->marks_black + b * OR_CELLS;
        const uint8_t *mw = in->marks_white + b * OR_CELLS;
        /* `current` is cast to the board dtype (int8) before the compare: encoding.cpp:51 */
        const int8_t cur = (int8_t)in->current_player[b];
        const int8_t neg = (int8_t)(-cur);
        const int is_black = in->current_player[b] == 1;
        for (int i = 0; i < OR_CELLS; ++i) {
            o[0 * OR_CELLS + i] = board[i] == cur ? 1.0f : 0.0f;
            o[1 * OR_CELLS + i] = board[i] == neg ? 1.0f : 0.0f;
            o[2 * OR_CELLS + i] = (is_black ? mb[i] : mw[i]) ? 1.0f : 0.0f;
            o[3 * OR_CELLS + i] = (is_black ? mw[i] : mb[i]) ? 1.0f : 0.0f;
        }
        for (int p = 1; p <= 7; ++p) {
            float v = in->phase[b] == p ? 1.0f : 0.0f;
            for (int i = 0; i < OR_CELLS; ++i) o[(3 + p) * OR_CELLS + i] = v;
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * root_puct_allocate_visits: v0/src/mcts/root_puct_fused.cu:12-117 (fp32; expression order kept).
 * `volatile` stores keep each intermediate rounded to fp32 exactly like the device code.
 * ---------------------------------------------------------------------------------------------- */
void or_root_puct_allocate_visits(int64_t R, int64_t M, const float *priors, const float *leaf,
                                  const uint8_t *valid, int64_t S, float c_puct, float *visits,
                                  float *value_sum, float *root_values) {
    for (int64_t r = 0; r < R; ++r) {
        const float *p = priors + r * M; const float *lv = leaf + r * M; const uint8_t *vm = valid + r * M;
        float *n = visits + r * M; float *w = value_sum + r * M;
        for (int64_t a = 0; a < M; ++a) { n[a] = 0.0f; w[a] = 0.0f; }
        float total = 0.0f;
        for (int64_t sim = 0; sim < S; ++sim) {
            const float sqrt_total = sqrtf(total + 1.0f);
            float best = -INFINITY; int64_t best_idx = -1;
            for (int64_t a = 0; a < M; ++a) {
                if (!vm[a]) continue;
                const float visit = n[a];
                volatile float q = visit > 0.0f ? (w[a] / fmaxf(visit, 1e-8f)) : 0.0f;
                volatile float t1 = c_puct * p[a];
                volatile float t2 = t1 * sqrt_total;
                volatile float u = t2 / (1.0f + visit);
                volatile float score = q + u;
                if (score > best || (score == best && (best_idx < 0 || a < best_idx))) { best = score; best_idx = a; }
            }
            if (best_idx >= 0) {
                n[best_idx] += 1.0f;
                volatile float nw = w[best_idx] + lv[best_idx];
                w[best_idx] = nw;
                total += 1.0f;
            }
        }
        /* final reduction: the kernel tree-reduces over a power-of-two thread count (:96-116);
         * the pairwise tree below reproduces it for M <= nthreads (one element per thread). */
        int T = 1;
        while (T < (int)M && T < 1024) T <<= 1;
        if (T < 32) T = 32;
        float sv[1024], sw[1024];
        for (int t = 0; t < T; ++t) {
            float pv = 0.0f, pw = 0.0f;
            for (int64_t a = t; a < M; a += T) { pv += n[a]; pw += w[a]; }
            sv[t] = pv; sw[t] = pw;
        }
        for (int off = T / 2; off > 0; off >>= 1)
            for (int t = 0; t < off; ++t) { sv[t] += sv[t + off]; sw[t] += sw[t + off]; }
        root_values[r] = sw[0] / fmaxf(sv[0], 1.0f);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Playout workload helpers (shared definition with liuzhou_b200/csrc/lz_rng.cuh; ours)
 * ---------------------------------------------------------------------------------------------- */
uint64_t or_mix64(uint64_t x) { /* splitmix64 finalizer */
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}
uint32_t or_playout_pick(uint64_t seed, uint64_t game, uint32_t ply, uint32_t n) {
    uint64_t h = or_mix64(or_mix64(seed ^ (game * 0xD1342543DE82EF95ULL)) + (uint64_t)ply);
    return (uint32_t)(((h >> 32) * (uint64_t)n) >> 32);
}

uint64_t or_state_hash(const or_state *s) {
    uint64_t h = 0xCBF29CE484222325ULL;
#define MIXB(v) do { h ^= (uint64_t)(uint8_t)(v); h *= 0x100000001B3ULL; } while (0)
    for (int i = 0; i < OR_CELLS; ++i) MIXB(s->board[i]);
    for (int i = 0; i < OR_CELLS; ++i) MIXB(s->marks_black[i]);
    for (int i = 0; i < OR_CELLS; ++i) MIXB(s->marks_white[i]);
    MIXB(s->phase); MIXB(s->current_player); MIXB(s->pending_marks_required); MIXB(s->pending_marks_remaining);
    MIXB(s->pending_captures_required); MIXB(s->pending_captures_remaining); MIXB(s->forced_removals_done);
    MIXB(s->move_count); MIXB(s->moves_since_capture);
#undef MIXB
    return h;
}

int or_random_playout(uint64_t seed, uint64_t game, int max_plies, or_state *final_state,
                      int *result_from_black, int16_t *trace, uint64_t *state_hash) {
    or_state s, nx;
    or_initial(&s);
    int idx[220];
    int ply = 0;
    uint64_t hh = 0;
    int result = 0;
    for (;;) {
        if (or_is_game_over(&s)) { result = or_winner(&s); break; }
        if (ply >= max_plies) { result = 0; break; }
        int n = or_legal_actions(&s, idx, 0);
        if (n == 0) { result = (int)(-s.current_player); break; }     /* module.cpp:733-735 */
        int a = idx[or_playout_pick(seed, game, (uint32_t)ply, (uint32_t)n)];
        if (trace) trace[ply] = (int16_t)a;
        if (!or_apply_move_scalar(&s, a, &nx)) { result = 2; break; }   /* cannot happen */
        s = nx;
        ++ply;
        hh = or_mix64(hh ^ or_state_hash(&s));
    }
    if (final_state) *final_state = s;
    if (result_from_black) *result_from_black = result;
    if (state_hash) *state_hash = hh;
    return ply;
}

/* Bench helper (cpu_baseline leg): play games [g0, g1) and accumulate plies + outcome counts.
 * out[0] = plies, out[1] = black wins, out[2] = white wins, out[3] = draws, out[4] = xor of state hashes */
void or_random_playouts_range(uint64_t seed, uint64_t g0, uint64_t g1, int max_plies, uint64_t *out) {
    uint64_t plies = 0, bw = 0, ww = 0, dr = 0, hx = 0;
    for (uint64_t g = g0; g < g1; ++g) {
        int res = 0;
        uint64_t h = 0;
        plies += (uint64_t)or_random_playout(seed, g, max_plies, 0, &res, 0, &h);
        if (res > 0) ++bw; else if (res < 0) ++ww; else ++dr;
        hx ^= h;
    }
    out[0] = plies; out[1] = bw; out[2] = ww; out[3] = dr; out[4] = hx;
}

/* Parity helper (full-size config-2 test): play games [g0, g1) and return every game's (plies, result, chained state
 * hash) -- the per-game form of or_random_playouts_range; arrays are indexed by g - g0. */
void or_random_playouts_each(uint64_t seed, uint64_t g0, uint64_t g1, int max_plies, int32_t *plies_out,
                             int8_t *result_out, uint64_t *hash_out) {
    for (uint64_t g = g0; g < g1; ++g) {
        int res = 0;
        uint64_t h = 0;
        plies_out[g - g0] = (int32_t)or_random_playout(seed, g, max_plies, 0, &res, 0, &h);
        result_out[g - g0] = (int8_t)res;
        hash_out[g - g0] = h;
    }
}
