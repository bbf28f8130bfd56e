// lz_root_ops.cu -- sm_100a kernels for the reference's ATen composite ops around the root search
// (module.cpp:258-871): root_pack_sparse_actions, root_finalize_from_visits, self_play_step_inplace,
// finalize_trajectory_inplace.  In the reference each of these is a chain of 10-40 ATen launches with
// several host syncs; here each is one or two launches (a per-row warp kernel + a single-block stable scan
// where the output order is data dependent).
#include "lz_common.cuh"

using namespace lz;

namespace lzb {
namespace {

constexpr int kScanThreads = 1024;

// Exclusive block-wide scan of one value per thread (blockDim.x == kScanThreads). Returns the exclusive
// prefix; *total receives the block total.
__device__ __forceinline__ long long block_exclusive_scan(long long v, long long* total) {
    __shared__ long long warp_sums[32];
    __shared__ long long block_total;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    long long incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        long long s = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const long long o = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += o;
        }
        warp_sums[lane] = s;
        if (lane == 31) block_total = s;
    }
    __syncthreads();
    const long long base = w > 0 ? warp_sums[w - 1] : 0;
    *total = block_total;
    __syncthreads();
    return base + incl - v;
}

// ------------------------------------------------------------------------------------------------------
// (a8) root_pack_sparse_actions -- module.cpp:258-363
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
pack_count_kernel(const uint8_t* __restrict__ legal, int64_t B, int A, int64_t* __restrict__ row_counts,
                  uint8_t* __restrict__ terminal) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t b = warp; b < B; b += nwarps) {
        int n = 0;
        for (int a = lane; a < A; a += 32) n += legal[b * A + a] != 0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) n += __shfl_xor_sync(0xffffffffu, n, off);
        if (lane == 0) { row_counts[b] = n; terminal[b] = n == 0; }
    }
}

// single block: root_rank (index among valid roots, -1 for terminal rows), flat_offset (exclusive prefix of
// counts), summary = {R, M, N}
__global__ void __launch_bounds__(kScanThreads)
pack_scan_kernel(const int64_t* __restrict__ row_counts, int64_t B, int64_t* __restrict__ root_rank,
                 int64_t* __restrict__ flat_offset, int64_t* __restrict__ summary) {
    const int64_t per = (B + kScanThreads - 1) / kScanThreads;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = min(B, lo + per);
    long long roots = 0, total = 0, mx = 0;
    for (int64_t b = lo; b < hi; ++b) {
        const long long c = row_counts[b];
        roots += c > 0; total += c; mx = max(mx, c);
    }
    long long all_roots, all_total;
    long long r0 = block_exclusive_scan(roots, &all_roots);
    long long t0 = block_exclusive_scan(total, &all_total);
    for (int64_t b = lo; b < hi; ++b) {
        const long long c = row_counts[b];
        root_rank[b] = c > 0 ? r0 : -1;
        flat_offset[b] = t0;
        r0 += c > 0; t0 += c;
    }
    __shared__ long long smax[32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int i = 0; i < kScanThreads / 32; ++i) m = max(m, smax[i]);
        summary[0] = all_roots; summary[1] = m; summary[2] = all_total;
    }
}

__global__ void __launch_bounds__(kThreads)
pack_fill_kernel(const uint8_t* __restrict__ legal, const float* __restrict__ probs,
                 const int32_t* __restrict__ metadata, int64_t B, int A, const int64_t* __restrict__ row_counts,
                 const int64_t* __restrict__ root_rank, const int64_t* __restrict__ flat_offset, int64_t M,
                 int64_t* __restrict__ valid_root_indices, int64_t* __restrict__ counts,
                 uint8_t* __restrict__ valid_mask, int64_t* __restrict__ legal_index_mat,
                 float* __restrict__ priors_mat, int32_t* __restrict__ action_code_mat,
                 int64_t* __restrict__ flat_indices, int32_t* __restrict__ action_codes_all,
                 int64_t* __restrict__ parent_indices_all) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    const int4* meta4 = reinterpret_cast<const int4*>(metadata);
    int4* code_mat4 = reinterpret_cast<int4*>(action_code_mat);
    int4* codes_all4 = reinterpret_cast<int4*>(action_codes_all);
    for (int64_t b = warp; b < B; b += nwarps) {
        const int64_t row = root_rank[b];
        if (row < 0) continue;
        const int64_t cnt = row_counts[b], foff = flat_offset[b];
        // pass 1: row sum of the gathered probabilities
        float sum = 0.0f;
        for (int a = lane; a < A; a += 32) sum += legal[b * A + a] ? probs[b * A + a] : 0.0f;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float denom = fmaxf(sum, 1e-8f);                  // :333-335
        // pass 2: stable compaction, ascending action index
        int running = 0;
        for (int base = 0; base < A; base += 32) {
            const int a = base + lane;
            const bool leg = a < A && legal[b * A + a] != 0;
            const uint32_t bal = __ballot_sync(0xffffffffu, leg);
            if (leg) {
                const int col = running + __popc(bal & ((1u << lane) - 1));
                const int64_t o = row * M + col;
                const int4 code = meta4[b * A + a];
                valid_mask[o] = 1;
                legal_index_mat[o] = a;
                priors_mat[o] = probs[b * A + a] / denom;
                code_mat4[o] = code;
                flat_indices[foff + col] = o;
                codes_all4[foff + col] = code;
                parent_indices_all[foff + col] = b;
            }
            running += __popc(bal);
        }
        for (int64_t col = cnt + lane; col < M; col += 32) {      // padding: mask False, index 0, prior 0, code 0
            const int64_t o = row * M + col;
            valid_mask[o] = 0; legal_index_mat[o] = 0; priors_mat[o] = 0.0f; code_mat4[o] = make_int4(0, 0, 0, 0);
        }
        if (lane == 0) { valid_root_indices[row] = b; counts[row] = cnt; }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a10) root_finalize_from_visits (sample_moves = false) -- module.cpp:441-535
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
finalize_visits_kernel(const int64_t* __restrict__ legal_index_mat, const int32_t* __restrict__ action_code_mat,
                       const uint8_t* __restrict__ valid_mask, const float* __restrict__ visits,
                       const float* __restrict__ value_sum, const int64_t* __restrict__ roots, int64_t R, int M,
                       int64_t A, const float* __restrict__ temps, float* __restrict__ policy_dense,
                       int64_t* __restrict__ chosen_idx, int32_t* __restrict__ chosen_codes,
                       uint8_t* __restrict__ chosen_valid, float* __restrict__ root_value) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t r = warp; r < R; r += nwarps) {
        const float expo = 1.0f / fmaxf(temps[r], 1e-6f);         // :484-486
        const int64_t base = r * M;
        float sum = 0.0f, sv = 0.0f, sw = 0.0f;
        for (int c = lane; c < M; c += 32) {
            const float p = powf(fmaxf(visits[base + c], 1e-8f), expo) * (valid_mask[base + c] ? 1.0f : 0.0f);
            sum += p; sv += visits[base + c]; sw += value_sum[base + c];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            sum += __shfl_xor_sync(0xffffffffu, sum, off);
            sv += __shfl_xor_sync(0xffffffffu, sv, off);
            sw += __shfl_xor_sync(0xffffffffu, sw, off);
        }
        const float denom = fmaxf(sum, 1e-8f);
        const int64_t b = roots[r];
        float* dense = policy_dense + b * A;
        // argmax with torch.max semantics: first maximal element, NaN counts as maximal
        float best = -INFINITY; int best_c = 0x7fffffff; bool best_nan = false;
        for (int c = lane; c < M; c += 32) {
            const bool vm = valid_mask[base + c] != 0;
            const float p = powf(fmaxf(visits[base + c], 1e-8f), expo) * (vm ? 1.0f : 0.0f) / denom;
            if (vm) dense[legal_index_mat[base + c]] = p;        // scatter_add of distinct legal indices
            const bool pn = p != p;
            if (best_c == 0x7fffffff || (pn && !best_nan) || (!best_nan && p > best)) {
                best = p; best_c = c; best_nan = pn;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, off);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, off);
            const bool on = __shfl_xor_sync(0xffffffffu, (int)best_nan, off) != 0;
            bool take;
            if (oc == 0x7fffffff) take = false;
            else if (best_c == 0x7fffffff) take = true;
            else if (on != best_nan) take = on;
            else if (on) take = oc < best_c;
            else take = ob > best || (ob == best && oc < best_c);
            if (take) { best = ob; best_c = oc; best_nan = on; }
        }
        if (lane == 0) {
            const int pick = best_c == 0x7fffffff ? 0 : best_c;
            chosen_idx[b] = legal_index_mat[base + pick];
            reinterpret_cast<int4*>(chosen_codes)[b] = reinterpret_cast<const int4*>(action_code_mat)[base + pick];
            chosen_valid[b] = 1;
            root_value[r] = sw / fmaxf(sv, 1.0f);
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a10) root_sparse_writeback -- module.cpp:365-439: an externally computed legal policy [R,M] and the picked column
// of every root go back to dense [B,A] rows.  Warp per root; `policy_dense` rows are pre-zeroed, the scatter is an
// atomicAdd because the reference's scatter_add_ also sums repeated indices (padding columns carry index 0 with
// weight policy * 0).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
sparse_writeback_kernel(const int64_t* __restrict__ legal_index_mat, const int32_t* __restrict__ action_code_mat,
                        const uint8_t* __restrict__ valid_mask, const float* __restrict__ legal_policy,
                        const int64_t* __restrict__ picks, const int64_t* __restrict__ roots, int64_t R, int M, int64_t A,
                        float* __restrict__ policy_dense, int64_t* __restrict__ chosen_idx,
                        int32_t* __restrict__ chosen_codes, uint8_t* __restrict__ chosen_valid) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t r = warp; r < R; r += nwarps) {
        const int64_t base = r * M, b = roots[r];
        float* dense = policy_dense + b * A;
        for (int c = lane; c < M; c += 32) {
            const float w = legal_policy[base + c] * (valid_mask[base + c] ? 1.0f : 0.0f);
            atomicAdd(dense + legal_index_mat[base + c], w);
        }
        if (lane == 0) {
            const int64_t pick = picks[r];
            chosen_idx[b] = legal_index_mat[base + pick];
            reinterpret_cast<int4*>(chosen_codes)[b] = reinterpret_cast<const int4*>(action_code_mat)[base + pick];
            chosen_valid[b] = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// (a14) self_play_step_inplace -- module.cpp:632-871
// ------------------------------------------------------------------------------------------------------
struct StepScratch { int32_t flag; float result; float soft; int32_t pad; };   // 16 B per active row

__device__ __forceinline__ float soft_value(uint64_t black, uint64_t white, float k) {
    // SoftValueFromBoardBatch, module.cpp:537-545
    const float delta = ((float)popc64(black) - (float)popc64(white)) / 18.0f;
    return tanhf(delta * k);
}

__device__ __forceinline__ void warp_store_state_ll(const lzb_states_out& out, int64_t i, int lane,
                                                    const State<long long>& s, int8_t blo, int8_t bhi) {
    auto cell_byte = [&](int cell, int8_t orig) -> int8_t {
        const uint64_t m = 1ULL << cell;
        return (s.black & m) ? (int8_t)1 : (s.white & m) ? (int8_t)-1 : (s.other & m) ? orig : (int8_t)0;
    };
    out.board[i * 36 + lane] = cell_byte(lane, blo);
    out.marks_black[i * 36 + lane] = (uint8_t)((s.mb >> lane) & 1);
    out.marks_white[i * 36 + lane] = (uint8_t)((s.mw >> lane) & 1);
    if (lane < 4) {
        out.board[i * 36 + 32 + lane] = cell_byte(32 + lane, bhi);
        out.marks_black[i * 36 + 32 + lane] = (uint8_t)((s.mb >> (32 + lane)) & 1);
        out.marks_white[i * 36 + 32 + lane] = (uint8_t)((s.mw >> (32 + lane)) & 1);
    }
    if (lane == 0) {
        out.phase[i] = s.phase; out.current_player[i] = s.player;
        out.pending_marks_required[i] = s.pm_req; out.pending_marks_remaining[i] = s.pm_rem;
        out.pending_captures_required[i] = s.pc_req; out.pending_captures_remaining[i] = s.pc_rem;
        out.forced_removals_done[i] = s.forced; out.move_count[i] = s.move_count;
        out.moves_since_capture[i] = s.msc;
    }
}

__global__ void __launch_bounds__(kThreads)
step_kernel(lzb_states_out st, int64_t B, int64_t* __restrict__ plies, uint8_t* __restrict__ done,
            const int64_t* __restrict__ active_idx, const int32_t* __restrict__ codes,
            const uint8_t* __restrict__ terminal, const uint8_t* __restrict__ chosen_valid, int64_t K,
            int64_t max_plies, float soft_k, StepScratch* __restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    lzb_states_in in;
    in.board = st.board; in.marks_black = st.marks_black; in.marks_white = st.marks_white; in.phase = st.phase;
    in.current_player = st.current_player; in.pending_marks_required = st.pending_marks_required;
    in.pending_marks_remaining = st.pending_marks_remaining;
    in.pending_captures_required = st.pending_captures_required;
    in.pending_captures_remaining = st.pending_captures_remaining;
    in.forced_removals_done = st.forced_removals_done; in.move_count = st.move_count;
    in.moves_since_capture = st.moves_since_capture;
    for (int64_t j = warp; j < K; j += nwarps) {
        const int64_t slot = active_idx[j];
        StepScratch out; out.flag = 0; out.result = 0.0f; out.soft = 0.0f; out.pad = 0;
        if (slot >= 0 && slot < B) {
            State<long long> s;
            int8_t blo, bhi;
            warp_load_state<long long>(in, slot, lane, s, blo, bhi);
            const bool term = terminal[j] != 0, valid = chosen_valid[j] != 0;
            if (term || !valid) {                                   // :730-747 immediate done
                out.flag = 1;
                out.result = term ? -(float)s.player : 0.0f;
                out.soft = soft_value(s.black, s.white, soft_k);
                if (lane == 0) done[slot] = 1;
            } else {
                const int4 code = reinterpret_cast<const int4*>(codes)[j];
                apply_action(s, code.x, code.y, code.z);
                warp_store_state_ll(st, slot, lane, s, blo, bhi);
                long long p = 0;
                if (lane == 0) { p = plies[slot] + 1; plies[slot] = p; }
                p = __shfl_sync(0xffffffffu, p, 0);
                const bool post = s.phase == kMovement || s.phase == kCapture || s.phase == kCounter;
                int win = 0;                                        // :817-824: white<4 overrides black<4
                if (post && popc64(s.black) < kLoseThreshold) win = -1;
                if (post && popc64(s.white) < kLoseThreshold) win = 1;
                const bool draw = s.move_count >= kMaxMoveCount || s.msc >= kNoCaptureLimit;
                if (win != 0 || draw || p >= max_plies) {
                    out.flag = 2;
                    out.result = (float)win;
                    out.soft = soft_value(s.black, s.white, soft_k);
                    if (lane == 0) done[slot] = 1;
                }
            }
        }
        if (lane == 0) scratch[j] = out;
    }
}

// single block: order = immediate rows (flag 1) in active order, then finished rows (flag 2) in active order
__global__ void __launch_bounds__(kScanThreads)
step_compact_kernel(const StepScratch* __restrict__ scratch, const int64_t* __restrict__ active_idx, int64_t K,
                    int64_t* __restrict__ slots, float* __restrict__ result, float* __restrict__ soft,
                    int64_t* __restrict__ num_finalized) {
    const int64_t per = (K + kScanThreads - 1) / kScanThreads;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = min(K, lo + per);
    long long n1 = 0, n2 = 0;
    for (int64_t j = lo; j < hi; ++j) { n1 += scratch[j].flag == 1; n2 += scratch[j].flag == 2; }
    long long t1, t2;
    long long p1 = block_exclusive_scan(n1, &t1);
    long long p2 = block_exclusive_scan(n2, &t2);
    for (int64_t j = lo; j < hi; ++j) {
        const StepScratch s = scratch[j];
        long long o = -1;
        if (s.flag == 1) o = p1++;
        else if (s.flag == 2) o = t1 + p2++;
        if (o >= 0) { slots[o] = active_idx[j]; result[o] = s.result; soft[o] = s.soft; }
    }
    if (threadIdx.x == 0) *num_finalized = t1 + t2;
}

// ------------------------------------------------------------------------------------------------------
// (a15) finalize_trajectory_inplace -- module.cpp:547-630
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
traj_finalize_kernel(float* __restrict__ value_targets, float* __restrict__ soft_targets,
                     const int8_t* __restrict__ signs, const int64_t* __restrict__ step_index_matrix, int64_t G,
                     int64_t T, const int64_t* __restrict__ step_counts, const int64_t* __restrict__ slots,
                     const float* __restrict__ result, const float* __restrict__ soft, int64_t F) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
    for (int64_t f = warp; f < F; f += nwarps) {
        const int64_t g = slots[f];
        if (g < 0 || g >= G) continue;
        const int64_t n = min(step_counts[g], T);
        const float r = result[f], s = soft[f];
        for (int64_t t = lane; t < n; t += 32) {
            const int64_t row = step_index_matrix[g * T + t];
            const float sg = (float)signs[row];
            value_targets[row] = sg * r;
            soft_targets[row] = sg * s;
        }
    }
}

__global__ void __launch_bounds__(kScanThreads)
traj_compact_kernel(const int64_t* __restrict__ step_counts, int64_t G, const int64_t* __restrict__ slots,
                    const float* __restrict__ result, int64_t F, int64_t* __restrict__ final_slots,
                    int64_t* __restrict__ final_counts, int64_t* __restrict__ summary) {
    const int64_t per = (F + kScanThreads - 1) / kScanThreads;
    const int64_t lo = (int64_t)threadIdx.x * per, hi = min(F, lo + per);
    long long kept = 0, bw = 0, ww = 0, dr = 0;
    for (int64_t f = lo; f < hi; ++f) {
        const int64_t g = slots[f];
        const bool keep = g >= 0 && g < G && step_counts[g] > 0;
        if (keep) { ++kept; bw += result[f] > 0.0f; ww += result[f] < 0.0f; dr += result[f] == 0.0f; }
    }
    long long tk, tb, tw, td;
    long long p = block_exclusive_scan(kept, &tk);
    block_exclusive_scan(bw, &tb);
    block_exclusive_scan(ww, &tw);
    block_exclusive_scan(dr, &td);
    for (int64_t f = lo; f < hi; ++f) {
        const int64_t g = slots[f];
        if (g >= 0 && g < G && step_counts[g] > 0) { final_slots[p] = g; final_counts[p] = step_counts[g]; ++p; }
    }
    if (threadIdx.x == 0) { summary[0] = tk; summary[1] = tb; summary[2] = tw; summary[3] = td; }
}

}  // namespace
}  // namespace lzb

using namespace lzb;

extern "C" int lzb_root_pack_count(const uint8_t* legal_mask, int64_t B, int64_t A, int64_t* row_counts,
                                   uint8_t* terminal_mask, int64_t* root_rank, int64_t* flat_offset, int64_t* summary,
                                   void* stream) {
    LZB_REQUIRE(B >= 0 && A >= 0 && A < (1 << 24), "bad shape");
    LZB_REQUIRE(summary, "null summary");
    cudaStream_t s = (cudaStream_t)stream;
    if (B > 0) {
        LZB_REQUIRE(legal_mask && row_counts && terminal_mask && root_rank && flat_offset, "null pointer");
        pack_count_kernel<<<warp_grid(B), kThreads, 0, s>>>(legal_mask, B, (int)A, row_counts, terminal_mask);
        int rc = check_launch("pack_count_kernel");
        if (rc) return rc;
    }
    pack_scan_kernel<<<1, kScanThreads, 0, s>>>(row_counts, B, root_rank, flat_offset, summary);
    return check_launch("pack_scan_kernel");
}

extern "C" int lzb_root_pack_fill(const uint8_t* legal_mask, const float* probs, const int32_t* metadata, int64_t B,
                                  int64_t A, const int64_t* row_counts, const int64_t* root_rank,
                                  const int64_t* flat_offset, int64_t R, int64_t M, int64_t* valid_root_indices,
                                  int64_t* counts, uint8_t* valid_mask, int64_t* legal_index_mat, float* priors_mat,
                                  int32_t* action_code_mat, int64_t* flat_indices, int32_t* action_codes_all,
                                  int64_t* parent_indices_all, void* stream) {
    LZB_REQUIRE(B >= 0 && A >= 0 && R >= 0 && M >= 0, "bad shape");
    if (B == 0 || R == 0) return LZB_OK;
    LZB_REQUIRE(legal_mask && probs && metadata && row_counts && root_rank && flat_offset, "null input");
    LZB_REQUIRE(valid_root_indices && counts && valid_mask && legal_index_mat && priors_mat && action_code_mat &&
                flat_indices && action_codes_all && parent_indices_all, "null output");
    LZB_REQUIRE(((reinterpret_cast<uintptr_t>(metadata) | reinterpret_cast<uintptr_t>(action_code_mat) |
                  reinterpret_cast<uintptr_t>(action_codes_all)) & 15) == 0, "int32[.,4] arrays must be 16-byte aligned");
    pack_fill_kernel<<<warp_grid(B), kThreads, 0, (cudaStream_t)stream>>>(
        legal_mask, probs, metadata, B, (int)A, row_counts, root_rank, flat_offset, M, valid_root_indices, counts,
        valid_mask, legal_index_mat, priors_mat, action_code_mat, flat_indices, action_codes_all, parent_indices_all);
    return check_launch("pack_fill_kernel");
}

extern "C" int lzb_root_finalize_from_visits(const int64_t* legal_index_mat, const int32_t* action_code_mat,
                                             const uint8_t* valid_mask, const float* visits, const float* value_sum,
                                             const int64_t* valid_root_indices, int64_t R, int64_t M,
                                             int64_t batch_size, int64_t total_action_dim,
                                             const float* root_temperatures, float* policy_dense,
                                             int64_t* chosen_action_indices, int32_t* chosen_action_codes,
                                             uint8_t* chosen_valid_mask, float* root_value, void* stream) {
    LZB_REQUIRE(batch_size >= 0, "batch_size must be non-negative");
    LZB_REQUIRE(total_action_dim > 0, "total_action_dim must be positive");
    LZB_REQUIRE(R >= 0 && M >= 0 && M < (1 << 24), "bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    if (batch_size > 0) {
        LZB_REQUIRE(policy_dense && chosen_action_indices && chosen_action_codes && chosen_valid_mask, "null output");
        cudaMemsetAsync(policy_dense, 0, sizeof(float) * batch_size * total_action_dim, s);
        cudaMemsetAsync(chosen_action_indices, 0xFF, sizeof(int64_t) * batch_size, s);
        cudaMemsetAsync(chosen_action_codes, 0xFF, sizeof(int32_t) * 4 * batch_size, s);
        cudaMemsetAsync(chosen_valid_mask, 0, batch_size, s);
    }
    if (R == 0) return LZB_OK;
    LZB_REQUIRE(M > 0, "roots without action columns");
    LZB_REQUIRE(legal_index_mat && action_code_mat && valid_mask && visits && value_sum && valid_root_indices &&
                root_temperatures && root_value, "null input");
    finalize_visits_kernel<<<warp_grid(R), kThreads, 0, s>>>(legal_index_mat, action_code_mat, valid_mask, visits,
                                                            value_sum, valid_root_indices, R, (int)M,
                                                            total_action_dim, root_temperatures, policy_dense,
                                                            chosen_action_indices, chosen_action_codes,
                                                            chosen_valid_mask, root_value);
    return check_launch("finalize_visits_kernel");
}

extern "C" int lzb_root_sparse_writeback(const int64_t* legal_index_mat, const int32_t* action_code_mat,
                                         const uint8_t* valid_mask, const float* legal_policy, const int64_t* local_picks,
                                         const int64_t* valid_root_indices, int64_t R, int64_t M, int64_t batch_size,
                                         int64_t total_action_dim, float* policy_dense, int64_t* chosen_action_indices,
                                         int32_t* chosen_action_codes, uint8_t* chosen_valid_mask, void* stream) {
    LZB_REQUIRE(batch_size >= 0, "batch_size must be non-negative");
    LZB_REQUIRE(total_action_dim > 0, "total_action_dim must be positive");
    LZB_REQUIRE(R >= 0 && M >= 0 && M < (1 << 24), "bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    if (batch_size > 0) {
        LZB_REQUIRE(policy_dense && chosen_action_indices && chosen_action_codes && chosen_valid_mask, "null output");
        cudaMemsetAsync(policy_dense, 0, sizeof(float) * batch_size * total_action_dim, s);
        cudaMemsetAsync(chosen_action_indices, 0xFF, sizeof(int64_t) * batch_size, s);
        cudaMemsetAsync(chosen_action_codes, 0xFF, sizeof(int32_t) * 4 * batch_size, s);
        cudaMemsetAsync(chosen_valid_mask, 0, batch_size, s);
    }
    if (R == 0) return LZB_OK;
    LZB_REQUIRE(M > 0, "roots without action columns");
    LZB_REQUIRE(legal_index_mat && action_code_mat && valid_mask && legal_policy && local_picks && valid_root_indices,
                "null input");
    sparse_writeback_kernel<<<warp_grid(R), kThreads, 0, s>>>(legal_index_mat, action_code_mat, valid_mask, legal_policy,
                                                             local_picks, valid_root_indices, R, (int)M, total_action_dim,
                                                             policy_dense, chosen_action_indices, chosen_action_codes,
                                                             chosen_valid_mask);
    return check_launch("sparse_writeback_kernel");
}

extern "C" int lzb_self_play_step_inplace(const lzb_states_out* states, int64_t B, int64_t* plies, uint8_t* done,
                                          const int64_t* active_idx, const int32_t* chosen_action_codes,
                                          const uint8_t* terminal_mask, const uint8_t* chosen_valid_mask, int64_t K,
                                          int64_t max_game_plies, float soft_value_k, int64_t* finalize_slots,
                                          float* result_from_black, float* soft_value_from_black,
                                          int64_t* num_finalized, void* scratch, void* stream) {
    LZB_REQUIRE(max_game_plies > 0, "max_game_plies must be positive");
    LZB_REQUIRE(states && B >= 0 && K >= 0 && num_finalized, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (K == 0) { cudaMemsetAsync(num_finalized, 0, sizeof(int64_t), s); return LZB_OK; }
    LZB_REQUIRE(plies && done && active_idx && chosen_action_codes && terminal_mask && chosen_valid_mask &&
                finalize_slots && result_from_black && soft_value_from_black && scratch, "null pointer");
    step_kernel<<<warp_grid(K), kThreads, 0, s>>>(*states, B, plies, done, active_idx, chosen_action_codes,
                                                  terminal_mask, chosen_valid_mask, K, max_game_plies, soft_value_k,
                                                  reinterpret_cast<StepScratch*>(scratch));
    int rc = check_launch("step_kernel");
    if (rc) return rc;
    step_compact_kernel<<<1, kScanThreads, 0, s>>>(reinterpret_cast<const StepScratch*>(scratch), active_idx, K,
                                                   finalize_slots, result_from_black, soft_value_from_black,
                                                   num_finalized);
    return check_launch("step_compact_kernel");
}

extern "C" int lzb_finalize_trajectory_inplace(float* value_targets, float* soft_value_targets,
                                               const int8_t* player_signs, const int64_t* step_index_matrix, int64_t G,
                                               int64_t T, const int64_t* step_counts, const int64_t* slots,
                                               const float* result_from_black, const float* soft_value_from_black,
                                               int64_t F, int64_t* final_slots, int64_t* final_counts,
                                               int64_t* summary, void* stream) {
    LZB_REQUIRE(G >= 0 && T >= 0 && F >= 0 && summary, "bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    if (F == 0) { cudaMemsetAsync(summary, 0, 4 * sizeof(int64_t), s); return LZB_OK; }
    LZB_REQUIRE(value_targets && soft_value_targets && player_signs && step_index_matrix && step_counts && slots &&
                result_from_black && soft_value_from_black && final_slots && final_counts, "null pointer");
    traj_finalize_kernel<<<warp_grid(F), kThreads, 0, s>>>(value_targets, soft_value_targets, player_signs,
                                                           step_index_matrix, G, T, step_counts, slots,
                                                           result_from_black, soft_value_from_black, F);
    int rc = check_launch("traj_finalize_kernel");
    if (rc) return rc;
    traj_compact_kernel<<<1, kScanThreads, 0, s>>>(step_counts, G, slots, result_from_black, F, final_slots,
                                                   final_counts, summary);
    return check_launch("traj_compact_kernel");
}
