/*
 * lz_tree_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain-C restatement of the reference's full-tree MCTS protocol
 * (`PortableTreeBatch`, /root/reference/v1/cpp/portable_mcts.cpp:437-977): pointer-linked nodes,
 * fp64 priors / value sums, int visit counts, player-aware backup, lowest-action-index tie-break.
 * Uses the scalar rule engine restated in lz_oracle.c.  Single-threaded.
 *
 * Parity status: PINNED by tests/test_oracle_vs_reference.py against oracle/_ref/_liuzhou_portable_cpp
 * (visit counts, action values and root values identical for identical priors/values).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "lz_oracle.h"

#define ACTIONS 220
#define INPUT_SIZE (11 * OR_CELLS)

typedef struct or_node {
    or_state state;
    struct or_node *parent;
    struct or_node **children;
    int num_children;
    double prior;
    int action_index;
    int visit_count;
    double value_sum;
    int terminal, expanded, no_legal_terminal;
    double initial_value;
} or_node;

typedef struct {
    int tree_index;
    or_node *node;
    or_node **path;
    int path_len;
    int legal[ACTIONS];
    int num_legal;
} or_pending;

typedef struct {
    int num_trees;
    or_node **roots;
    uint8_t *active;
    double c_puct;
    or_pending *pending;
    int num_pending;
    int pending_kind; /* 0 none, 1 roots, 2 leaves */
} or_tree_batch;

/* Node::Node, portable_mcts.cpp:404-415 */
static or_node *node_new(const or_state *s, or_node *parent, double prior, int action) {
    or_node *n = (or_node *)calloc(1, sizeof(or_node));
    n->state = *s;
    n->parent = parent;
    n->prior = prior;
    n->action_index = action;
    n->terminal = or_is_game_over(s);
    return n;
}
static void node_free(or_node *n) {
    if (!n) return;
    for (int i = 0; i < n->num_children; ++i) node_free(n->children[i]);
    free(n->children);
    free(n);
}
static double mean_value(const or_node *n) { /* :430-432 */
    return n->visit_count > 0 ? n->value_sum / (double)n->visit_count : 0.0;
}
/* TerminalValue, :267-273 */
static double terminal_value(const or_state *s) {
    int w = or_winner(s);
    if (w == 0) return 0.0;
    return w == (int)s->current_player ? 1.0 : -1.0;
}

or_tree_batch *or_tree_create(int num_trees, const or_state *states, double c_puct) {
    or_tree_batch *tb = (or_tree_batch *)calloc(1, sizeof(or_tree_batch));
    tb->num_trees = num_trees;
    tb->roots = (or_node **)calloc((size_t)num_trees, sizeof(or_node *));
    tb->active = (uint8_t *)malloc((size_t)num_trees);
    tb->c_puct = c_puct;
    tb->pending = (or_pending *)calloc((size_t)num_trees, sizeof(or_pending));
    for (int i = 0; i < num_trees; ++i) {
        tb->roots[i] = node_new(&states[i], 0, 1.0, -1);
        tb->active[i] = 1;
    }
    return tb;
}
static void clear_pending(or_tree_batch *tb) {
    for (int i = 0; i < tb->num_pending; ++i) free(tb->pending[i].path);
    tb->num_pending = 0;
    tb->pending_kind = 0;
}
void or_tree_free(or_tree_batch *tb) {
    if (!tb) return;
    clear_pending(tb);
    for (int i = 0; i < tb->num_trees; ++i) node_free(tb->roots[i]);
    free(tb->roots); free(tb->active); free(tb->pending); free(tb);
}

/* EncodeModelInput, :241-265 */
static void encode_input(const or_state *s, float *out) {
    memset(out, 0, sizeof(float) * INPUT_SIZE);
    const int player = (int)s->current_player;
    const uint8_t *self_m = player == 1 ? s->marks_black : s->marks_white;
    const uint8_t *opp_m = player == 1 ? s->marks_white : s->marks_black;
    for (int cell = 0; cell < OR_CELLS; ++cell) {
        int v = s->board[cell];
        out[cell] = v == player ? 1.0f : 0.0f;
        out[OR_CELLS + cell] = v == -player ? 1.0f : 0.0f;
        out[2 * OR_CELLS + cell] = self_m[cell] ? 1.0f : 0.0f;
        out[3 * OR_CELLS + cell] = opp_m[cell] ? 1.0f : 0.0f;
    }
    int ch = 3 + (int)s->phase;
    if (ch >= 4 && ch < 11) for (int i = 0; i < OR_CELLS; ++i) out[ch * OR_CELLS + i] = 1.0f;
}

static int export_pending(or_tree_batch *tb, int32_t *tree_indices, float *inputs, uint8_t *masks) {
    for (int p = 0; p < tb->num_pending; ++p) {
        or_pending *pe = &tb->pending[p];
        if (tree_indices) tree_indices[p] = pe->tree_index;
        if (inputs) encode_input(&pe->node->state, inputs + (size_t)p * INPUT_SIZE);
        if (masks) {
            memset(masks + (size_t)p * ACTIONS, 0, ACTIONS);
            for (int i = 0; i < pe->num_legal; ++i) masks[(size_t)p * ACTIONS + pe->legal[i]] = 1;
        }
    }
    return tb->num_pending;
}

/* PrepareRoots, :483-513 */
int or_tree_prepare_roots(or_tree_batch *tb, int32_t *tree_indices, float *inputs, uint8_t *masks) {
    if (tb->pending_kind != 0) return -1;
    for (int i = 0; i < tb->num_trees; ++i) {
        or_node *root = tb->roots[i];
        if (!tb->active[i] || root->terminal) continue;
        if (or_is_game_over(&root->state)) { root->terminal = 1; continue; }
        if (!root->expanded) {
            or_pending *pe = &tb->pending[tb->num_pending++];
            pe->tree_index = i;
            pe->node = root;
            pe->path = (or_node **)malloc(sizeof(or_node *));
            pe->path[0] = root;
            pe->path_len = 1;
            pe->num_legal = or_legal_actions(&root->state, pe->legal, 0);
        }
    }
    tb->pending_kind = 1;
    return export_pending(tb, tree_indices, inputs, masks);
}

/* SelectChild, :832-860 */
static or_node *select_child(const or_tree_batch *tb, or_node *node) {
    const double sqrt_total = sqrt((double)(node->visit_count > 1 ? node->visit_count : 1));
    double best_score = -INFINITY;
    int best_action = 0x7fffffff;
    or_node *best = 0;
    for (int i = 0; i < node->num_children; ++i) {
        or_node *child = node->children[i];
        double q = 0.0;
        if (child->visit_count > 0) {
            q = mean_value(child);
            if (node->state.current_player != child->state.current_player) q = -q;
        }
        const double u = tb->c_puct * child->prior * sqrt_total / (double)(1 + child->visit_count);
        const double score = q + u;
        if (score > best_score || (score == best_score && child->action_index < best_action)) {
            best_score = score;
            best_action = child->action_index;
            best = child;
        }
    }
    return best;
}

/* Backup, :876-892 */
static void backup(or_node **path, int len, double leaf_value) {
    double value = leaf_value;
    for (int rev = len; rev > 0; --rev) {
        or_node *node = path[rev - 1];
        ++node->visit_count;
        node->value_sum += value;
        if (rev > 1) {
            or_node *parent = path[rev - 2];
            if (parent->state.current_player != node->state.current_player) value = -value;
        }
    }
}

/* SelectLeaves, :515-552 (+ SelectPath :862-874) */
int or_tree_select_leaves(or_tree_batch *tb, int32_t *tree_indices, float *inputs, uint8_t *masks) {
    if (tb->pending_kind != 0) return -1;
    for (int i = 0; i < tb->num_trees; ++i) {
        or_node *root = tb->roots[i];
        if (!tb->active[i] || root->terminal) continue;
        int cap = 16, len = 1;
        or_node **path = (or_node **)malloc(sizeof(or_node *) * (size_t)cap);
        path[0] = root;
        or_node *node = root;
        while (node->expanded && node->num_children > 0 && !node->terminal) {
            or_node *child = select_child(tb, node);
            if (!child) break;
            node = child;
            if (len == cap) { cap *= 2; path = (or_node **)realloc(path, sizeof(or_node *) * (size_t)cap); }
            path[len++] = node;
        }
        or_node *leaf = path[len - 1];
        if (leaf->terminal) {
            double v = leaf->no_legal_terminal ? -1.0 : terminal_value(&leaf->state);
            backup(path, len, v);
            free(path);
            continue;
        }
        if (leaf->expanded && leaf->num_children == 0) {
            leaf->terminal = 1;
            leaf->no_legal_terminal = 1;
            backup(path, len, -1.0);
            free(path);
            continue;
        }
        or_pending *pe = &tb->pending[tb->num_pending++];
        pe->tree_index = i;
        pe->node = leaf;
        pe->path = path;
        pe->path_len = len;
        pe->num_legal = or_legal_actions(&leaf->state, pe->legal, 0);
    }
    tb->pending_kind = 2;
    return export_pending(tb, tree_indices, inputs, masks);
}

/* Expand, :894-939 */
static double expand(or_pending *pe, const float *dense_priors, double value) {
    or_node *node = pe->node;
    node->initial_value = value;
    if (pe->num_legal == 0) {
        node->expanded = 1;
        node->terminal = 1;
        node->no_legal_terminal = !or_is_game_over(&node->state);
        node->initial_value = node->no_legal_terminal ? -1.0 : terminal_value(&node->state);
        return node->initial_value;
    }
    double prior_sum = 0.0;
    for (int i = 0; i < pe->num_legal; ++i) prior_sum += (double)dense_priors[pe->legal[i]];
    const int uniform = !isfinite(prior_sum) || prior_sum <= 0.0;
    for (int i = 0; i < node->num_children; ++i) node_free(node->children[i]);
    free(node->children);
    node->children = (or_node **)malloc(sizeof(or_node *) * (size_t)pe->num_legal);
    node->num_children = pe->num_legal;
    for (int i = 0; i < pe->num_legal; ++i) {
        const int action = pe->legal[i];
        const double prior = uniform ? 1.0 / (double)pe->num_legal : (double)dense_priors[action] / prior_sum;
        or_state child;
        or_apply_move_scalar(&node->state, action, &child);
        node->children[i] = node_new(&child, node, prior, action);
    }
    node->expanded = 1;
    return node->initial_value;
}

/* CompletePending, :554-590 */
int or_tree_complete_pending(or_tree_batch *tb, const float *priors, const float *values) {
    if (tb->pending_kind == 0) return -1;
    const int do_backup = tb->pending_kind == 2;
    for (int p = 0; p < tb->num_pending; ++p) {
        or_pending *pe = &tb->pending[p];
        const double v = expand(pe, priors + (size_t)p * ACTIONS, (double)values[p]);
        if (do_backup) backup(pe->path, pe->path_len, v);
    }
    clear_pending(tb);
    return 0;
}

/* RootPriors, :592-624 */
void or_tree_root_priors(const or_tree_batch *tb, float *priors, uint8_t *masks, uint8_t *active) {
    memset(priors, 0, sizeof(float) * (size_t)tb->num_trees * ACTIONS);
    memset(masks, 0, (size_t)tb->num_trees * ACTIONS);
    for (int t = 0; t < tb->num_trees; ++t) {
        const or_node *root = tb->roots[t];
        int usable = tb->active[t] && root->expanded && !root->terminal && root->num_children > 0;
        active[t] = (uint8_t)usable;
        if (!usable) continue;
        for (int i = 0; i < root->num_children; ++i) {
            priors[(size_t)t * ACTIONS + root->children[i]->action_index] = (float)root->children[i]->prior;
            masks[(size_t)t * ACTIONS + root->children[i]->action_index] = 1;
        }
    }
}

/* SetRootPriors, :626-662 */
int or_tree_set_root_priors(or_tree_batch *tb, const float *priors) {
    for (int t = 0; t < tb->num_trees; ++t) {
        or_node *root = tb->roots[t];
        if (!tb->active[t] || root->terminal || root->num_children == 0) continue;
        double sum = 0.0;
        for (int i = 0; i < root->num_children; ++i)
            sum += (double)priors[(size_t)t * ACTIONS + root->children[i]->action_index];
        if (!isfinite(sum) || sum <= 0.0) return -1;
        for (int i = 0; i < root->num_children; ++i)
            root->children[i]->prior =
                (double)priors[(size_t)t * ACTIONS + root->children[i]->action_index] / sum;
    }
    return 0;
}

/* RootOutputs, :664-737 */
void or_tree_root_outputs(const or_tree_batch *tb, uint8_t *masks, int32_t *visits, float *action_values,
                          float *root_values, int32_t *players, uint8_t *terminal, float *inputs) {
    const size_t T = (size_t)tb->num_trees;
    memset(masks, 0, T * ACTIONS);
    memset(visits, 0, sizeof(int32_t) * T * ACTIONS);
    memset(action_values, 0, sizeof(float) * T * ACTIONS);
    for (size_t t = 0; t < T; ++t) {
        const or_node *root = tb->roots[t];
        if (inputs) encode_input(&root->state, inputs + t * INPUT_SIZE);
        players[t] = (int32_t)root->state.current_player;
        terminal[t] = (root->terminal || root->num_children == 0) ? 1 : 0;
        root_values[t] = (float)(root->visit_count > 0 ? mean_value(root)
                                 : (root->no_legal_terminal ? -1.0
                                    : (root->terminal ? terminal_value(&root->state) : root->initial_value)));
        for (int i = 0; i < root->num_children; ++i) {
            const or_node *child = root->children[i];
            masks[t * ACTIONS + child->action_index] = 1;
            visits[t * ACTIONS + child->action_index] = child->visit_count;
            if (child->visit_count > 0) {
                double q = mean_value(child);
                if (root->state.current_player != child->state.current_player) q = -q;
                action_values[t * ACTIONS + child->action_index] = (float)q;
            }
        }
    }
}

/* AdvanceRoots, :739-769 */
int or_tree_advance_roots(or_tree_batch *tb, const int32_t *actions) {
    if (tb->pending_kind != 0) return -1;
    for (int t = 0; t < tb->num_trees; ++t) {
        if (!tb->active[t] || actions[t] < 0) continue;
        or_node *root = tb->roots[t];
        int found = -1;
        for (int i = 0; i < root->num_children; ++i)
            if (root->children[i]->action_index == actions[t]) { found = i; break; }
        if (found < 0) return -2;
        or_node *next = root->children[found];
        root->children[found] = root->children[root->num_children - 1];
        root->num_children -= 1;
        next->parent = 0;
        node_free(root);
        tb->roots[t] = next;
    }
    return 0;
}

void or_tree_deactivate(or_tree_batch *tb, int tree) { tb->active[tree] = 0; }

void or_tree_root_state(const or_tree_batch *tb, int tree, or_state *out) { *out = tb->roots[tree]->state; }
int or_tree_root_visit_count(const or_tree_batch *tb, int tree) { return tb->roots[tree]->visit_count; }

/* test helper: states of the currently pending leaves, in pending order */
int or_tree_pending_states(const or_tree_batch *tb, or_state *out) {
    for (int p = 0; p < tb->num_pending; ++p) out[p] = tb->pending[p].node->state;
    return tb->num_pending;
}
