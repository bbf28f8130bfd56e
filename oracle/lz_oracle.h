/*
 * lz_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, byte-per-cell restatement of the reference's (kuailehaha/liuzhou) algorithms for the
 * self-play hot path.  It deliberately keeps the reference's data layout ("SoA of bytes": board
 * int8[B,36], marks bool[B,36], nine int64[B] scalars) and the reference's loop structure so that each
 * function can be read side by side with the file:line it cites.  It shares NO code with
 * liuzhou_b200/csrc (which works on packed bitboards).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call it.
 *
 * Parity status: PINNED -- tests/test_oracle_vs_reference.py checks every function here against the
 * reference's own binaries (oracle/_ref, built from /root/reference by oracle/build_ref.py) and
 * tests/test_oracle_golden.py against committed golden vectors generated from those binaries.
 */
#ifndef LZ_ORACLE_H
#define LZ_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OR_CELLS 36
#define OR_SIZE 6

/* Phase / player / action kind values: v0/include/v0/game_state.hpp:24-37,
 * v0/src/game/fast_legal_mask_common.hpp:13-41 */
enum {
    OR_PHASE_PLACEMENT = 1, OR_PHASE_MARK = 2, OR_PHASE_REMOVAL = 3, OR_PHASE_MOVEMENT = 4,
    OR_PHASE_CAPTURE = 5, OR_PHASE_FORCED = 6, OR_PHASE_COUNTER = 7
};
enum {
    OR_ACT_PLACE = 1, OR_ACT_MOVE = 2, OR_ACT_MARK = 3, OR_ACT_CAPTURE = 4, OR_ACT_FORCED = 5,
    OR_ACT_COUNTER = 6, OR_ACT_NOMOVES = 7, OR_ACT_PROCESS = 8
};

/* One scalar game state (v0/include/v0/game_state.hpp:93-147), reference tensor field order. */
typedef struct {
    int8_t board[OR_CELLS];
    uint8_t marks_black[OR_CELLS];
    uint8_t marks_white[OR_CELLS];
    int64_t phase, current_player;
    int64_t pending_marks_required, pending_marks_remaining;
    int64_t pending_captures_required, pending_captures_remaining;
    int64_t forced_removals_done, move_count, moves_since_capture;
} or_state;

/* SoA view over B states in the reference tensor layout (mcts_gpu.py:40-55). */
typedef struct {
    int8_t *board;         /* [B,36] */
    uint8_t *marks_black;  /* [B,36] */
    uint8_t *marks_white;  /* [B,36] */
    int64_t *phase, *current_player;
    int64_t *pending_marks_required, *pending_marks_remaining;
    int64_t *pending_captures_required, *pending_captures_remaining;
    int64_t *forced_removals_done, *move_count, *moves_since_capture;
} or_batch;

void or_load(const or_batch *b, int64_t i, or_state *s);
void or_store(const or_state *s, or_batch *b, int64_t i);
void or_initial(or_state *s);

/* ---- shape detection: v0/src/rules/rule_engine.cpp:57-154,194-208 ---- */
int or_check_squares(const int8_t *board, const uint8_t *marked, int r, int c, int player_value);
int or_check_lines(const int8_t *board, const uint8_t *marked, int r, int c, int player_value);
int or_is_piece_in_shape(const int8_t *board, const uint8_t *marked, int r, int c, int player_value);

/* ---- tensor-op semantics (what v0_core.encode_actions_fast / batch_apply_moves compute) ---- */

/* v0/src/game/fast_legal_mask.cpp:253-418 (== fast_legal_mask_cuda.cu:282-404).
 * mask u8[B,T] and metadata i32[B,T,4] are fully written (0 / -1 fill included). */
void or_encode_actions(int64_t B, const or_batch *in, int64_t placement_dim, int64_t movement_dim,
                       int64_t selection_dim, int64_t auxiliary_dim, uint8_t *mask, int32_t *metadata);

/* One atomic action with the CUDA kernel's semantics (fast_apply_moves_cuda.cu:239-744): illegal
 * actions are silent no-ops but still bump move_count for every kind except placement.
 * Returns 1 if the action was applied, 0 if it was a no-op where the CPU path
 * (fast_apply_moves.cpp:246-753) would have thrown. */
int or_apply_action(or_state *s, int32_t kind, int32_t primary, int32_t secondary);

/* batch_apply_moves (fast_apply_moves_cuda.cu:548-744). applied_flags may be NULL. Rows whose parent
 * index is out of range are left untouched in `out` (the CUDA kernel returns early). */
void or_batch_apply_moves(int64_t B, const or_batch *in, int64_t N, const int32_t *action_codes,
                          const int64_t *parent_indices, or_batch *out, uint8_t *applied_flags);

/* batch_apply_moves_inplace (fast_apply_moves_cuda.cu:746-917). */
void or_batch_apply_moves_inplace(int64_t B, or_batch *st, int64_t N, const int32_t *action_codes,
                                  const int64_t *slot_indices);

/* ---- scalar-engine semantics (v0::GenerateAllLegalMoves / v0::ApplyMove / GameState) ---- */

/* game_state.cpp:59-79. winner: +1 black, -1 white, 0 none. */
int or_winner(const or_state *s);
int or_is_game_over(const or_state *s);

/* move_generator.cpp:242-297 expressed as 220-d action indices (portable_mcts.cpp:170-198), ascending.
 * Writes up to 220 indices + their (kind, primary, secondary, extra) tensor-style codes
 * (movement: secondary = dir index, extra = dest cell). Returns the count. */
int or_legal_actions(const or_state *s, int *indices, int32_t *codes /* [n][4] or NULL */);

/* move_generator.cpp:360-432 (throws -> returns 0 and leaves *out undefined). */
int or_apply_move_scalar(const or_state *s, int action_index, or_state *out);

/* ---- network-side encodings ---- */

/* v0/src/net/encoding.cpp:26-79 -> f32[B,11,6,6]. */
void or_states_to_model_input(int64_t B, const or_batch *in, float *out);

/* ---- root-PUCT op: v0/src/mcts/root_puct_fused.cu:12-117 (fp32, ties -> lowest index). ---- */
void or_root_puct_allocate_visits(int64_t R, int64_t M, const float *priors, const float *leaf_values,
                                  const uint8_t *valid_mask, int64_t num_simulations,
                                  float exploration_weight, float *visits, float *value_sum,
                                  float *root_values);

/* ---- counter-based RNG shared by the playout workload (ours, not the reference's) ---- */
uint64_t or_mix64(uint64_t x);
uint32_t or_playout_pick(uint64_t seed, uint64_t game, uint32_t ply, uint32_t n);

/* Config-2 workload on the scalar engine: play game `game` with uniform-random legal actions until
 * game over / no legal action / max_plies. Returns plies played; *result_from_black in {+1,-1,0};
 * final state in *final_state; if trace != NULL writes the chosen action index per ply. */
int or_random_playout(uint64_t seed, uint64_t game, int max_plies, or_state *final_state,
                      int *result_from_black, int16_t *trace, uint64_t *state_hash);

void or_random_playouts_range(uint64_t seed, uint64_t g0, uint64_t g1, int max_plies, uint64_t *out);
void or_random_playouts_each(uint64_t seed, uint64_t g0, uint64_t g1, int max_plies, int32_t *plies_out,
                             int8_t *result_out, uint64_t *hash_out);

/* FNV-style hash of a state in canonical field order (used for checksum-of-checksums properties). */
uint64_t or_state_hash(const or_state *s);

#ifdef __cplusplus
}
#endif
#endif
