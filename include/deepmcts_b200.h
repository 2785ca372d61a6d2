/*
 * deepmcts_b200.h -- C ABI of libdeepmcts_b200.so, the B200 (sm_100a) batched MCTS
 * self-play engine that replaces the hot path of henribru/deep-mcts.
 *
 * Reference interfaces replaced (paths relative to the reference's deep_mcts/):
 *   game engines ....... GameManager.{initial_game_state, legal_actions,
 *                        generate_child_state, is_final_state, evaluate_final_state}
 *                        game.py:76-113; othello/game.py:31-141; hex/game.py:31-142;
 *                        hex_with_swap/game.py:9-36; tictactoe/game.py:24-97
 *   tree search ........ MCTS.step / simulation / tree_search / evaluate_leaf /
 *                        expand_node / rollout / backpropagate / add_dirichlet_noise /
 *                        self_play / reset, MCTSAgent.play re-rooting (mcts.py:57-300)
 *   leaf evaluation .... GameNet.evaluate / forward (gamenet.py:61-85, 119-128) over
 *                        ConvolutionalNet.forward (convolutionalnet.py:22-170) and the
 *                        per-game states_to_tensor / policy un-transpose adapters
 *
 * Conventions
 *   - Every function returns 0 on success or a negative dm_status; the message of the
 *     last failure on the calling thread is dm_last_error().  No C++ exception crosses
 *     this boundary.
 *   - Pointers named *_dev are device pointers owned by the caller (e.g. torch tensors
 *     passed by data_ptr()); the library never frees them and never returns memory the
 *     caller must free.  The library owns only what dm_*_create allocates.
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).
 *     All calls are asynchronous on that stream unless documented otherwise.
 *   - A game state is three values: player (1 = FIRST, the first mover and min player;
 *     0 = SECOND, the max player; game.py:9-37) and two 64-bit bitboards whose bit
 *     y*grid + x is set where FIRST / SECOND has a stone (cell values 1 / 0, game.py:40-43).
 *   - Actions are ints x + y*grid; action grid*grid is "pass" (othello) or "swap"
 *     (hex_with_swap).  Values are absolute in [0,1]: FIRST win 0, SECOND win 1, draw .5
 *     (game.py:54-57).
 */
#ifndef DEEPMCTS_B200_H
#define DEEPMCTS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DM_ABI_VERSION 1
#define DM_MAX_ACTIONS 65   /* 8x8 board + pass/swap */
#define DM_MAX_DEPTH 128    /* longest root-to-leaf path / plies per game recorded */
#define DM_MAX_LEAVES 8     /* leaves one tree may have in flight per round (virtual-loss mode) */
#define DM_REPLAY_POS_BITS 38          /* replay counter word: positions written in the low bits, games above */
#define DM_REPLAY_GAME_RING (1 << 18)  /* games whose end positions are remembered (replay window in games) */

typedef enum {
    DM_OK = 0,
    DM_ERR_INVALID = -1,   /* bad argument */
    DM_ERR_CUDA = -2,      /* CUDA runtime failure (message has the cudaError string) */
    DM_ERR_CAPACITY = -3,  /* a tree ran out of node or path capacity */
    DM_ERR_STATE = -4      /* call sequence violated (e.g. expand without select) */
} dm_status;

typedef enum { DM_TICTACTOE = 0, DM_HEX = 1, DM_HEX_SWAP = 2, DM_OTHELLO = 3 } dm_game;

/* built-in evaluators for searches that need no network (mcts.py:202-212, train.py:240-248) */
typedef enum {
    DM_EVAL_NONE = 0,     /* state_evaluator=None: value 0.5, uniform priors, rollout decides */
    DM_EVAL_UNIFORM = 1,  /* train.py:240-248 uniform_state_evaluator */
    DM_EVAL_HASHED = 2,   /* synthetic integer-hash evaluator used by parity tests */
    DM_EVAL_EXTERNAL = 3  /* leaves are handed out by dm_pool_select (network / host callable) */
} dm_eval_kind;

/* rollout policies (mcts.py:219-226); RANDOM = random.choice(legal_actions(s)) */
typedef enum {
    DM_ROLLOUT_NONE = 0, DM_ROLLOUT_RANDOM = 1, DM_ROLLOUT_MIN_ACTION = 2, DM_ROLLOUT_MAX_ACTION = 3
} dm_rollout_kind;

typedef struct {
    int32_t game;                /* dm_game */
    int32_t grid;                /* 3 (tictactoe), 2..8 hex, 4/6/8 othello */
    int32_t n_trees;             /* independent trees (games) in the pool */
    int32_t max_nodes_per_tree;  /* node arena per tree */
    int32_t device;              /* CUDA device ordinal */
    int32_t compaction;          /* 1 = compact a tree's arena in place at a re-root when the next move's search
                                    might not fit (subtree reuse, mcts.py:105-108, without unbounded growth) */
    double c_puct;               /* 1.25 in the reference (mcts.py:42) */
    uint64_t seed;               /* Philox key for device-side sampling */
} dm_pool_config;

/* MCTS constructor arguments that matter on the device (mcts.py:70-95) */
typedef struct {
    int32_t num_simulations;
    int32_t eval_kind;           /* dm_eval_kind */
    int32_t rollout_kind;        /* dm_rollout_kind */
    int32_t sample_move_cutoff;
    double dirichlet_alpha;      /* 0 = off */
    double dirichlet_factor;
    double rollout_share;
    uint64_t eval_salt;          /* DM_EVAL_HASHED only */
    /* 1 (default, 0 means 1) = exact mode: one simulation in flight per tree, visit counts
     * bit-identical to the reference.  2..DM_MAX_LEAVES = throughput mode for small pools
     * (tournaments, single-game play): each dm_pool_select hands out up to that many leaves per
     * tree, steering later descents away from earlier ones with a virtual loss (one extra visit
     * scored as a loss for the player who chose the node) that dm_pool_expand_backup removes. */
    int32_t leaves_per_round;
    int32_t reserved1;
} dm_search_config;

typedef struct dm_pool dm_pool;
typedef struct dm_net dm_net;

int dm_abi_version(void);
const char* dm_last_error(void);

/* ------------------------------------------------------------------ game engines
 * Batched, one state per index, all pointers device.  Replace GameManager methods. */
int dm_game_num_actions(int game, int grid);
/* initial_game_state (othello/game.py:38-48 etc.) */
int dm_game_initial(int game, int grid, int n, uint64_t* first_dev, uint64_t* second_dev,
                    uint8_t* player_dev, void* stream);
/* legal_actions: list_dev[n][DM_MAX_ACTIONS] gets the reference's LIST order
 * (othello/game.py:88-118 = set order of discovery; hex/tictactoe ascending),
 * order_dev[n][DM_MAX_ACTIONS] the order MCTS.expand_node creates children in
 * (iteration order of set(legal), mcts.py:200-215); count_dev[n] entries are valid.
 * Any of list_dev / order_dev may be NULL. */
int dm_game_legal(int game, int grid, int n, const uint64_t* first_dev, const uint64_t* second_dev,
                  const uint8_t* player_dev, uint8_t* list_dev, uint8_t* order_dev,
                  int32_t* count_dev, void* stream);
/* generate_child_state; ok_dev[i] = 0 when action_dev[i] is illegal (the reference asserts) */
int dm_game_child(int game, int grid, int n, const uint64_t* first_dev, const uint64_t* second_dev,
                  const uint8_t* player_dev, const int32_t* action_dev, uint64_t* out_first_dev,
                  uint64_t* out_second_dev, uint8_t* out_player_dev, uint8_t* ok_dev, void* stream);
/* is_final_state + evaluate_final_state(...).value */
int dm_game_status(int game, int grid, int n, const uint64_t* first_dev, const uint64_t* second_dev,
                   const uint8_t* player_dev, uint8_t* is_final_dev, float* outcome_dev, void* stream);
/* MCTS.rollout with the RANDOM / MIN / MAX policy from each state (mcts.py:219-226);
 * seq_dev[i] individualises the Philox stream. */
int dm_game_rollout(int game, int grid, int n, const uint64_t* first_dev, const uint64_t* second_dev,
                    const uint8_t* player_dev, int rollout_kind, uint64_t seed, float* outcome_dev,
                    int32_t* plies_dev, void* stream);

/* uniform_state_evaluator (train.py:240-248) for a leaf batch: value_dev[n] = 0.5 and
 * probs_dev[n][DM_MAX_ACTIONS] = 1/len(legal) on legal actions; count_dev (may be NULL) is the
 * live batch size on the device. */
int dm_eval_uniform(int game, int grid, int n, const int32_t* count_dev, const uint64_t* first_dev,
                    const uint64_t* second_dev, const uint8_t* player_dev, float* value_dev, float* probs_dev,
                    void* stream);

/* the DM_EVAL_HASHED test evaluator as an external one: value_dev[slot] and probs_dev[slot][DM_MAX_ACTIONS] in
 * double for the live leaves; index_dev (may be NULL) lists the slots of a compacted dense batch
 * (dm_pool_leaf_index).  Parity tests use it to drive the leaf-exchange kernels and the evaluation
 * cache with an evaluator whose reference results are pinned. */
int dm_eval_hashed(int game, int grid, int n, const int32_t* count_dev, const int32_t* index_dev,
                   const uint64_t* first_dev, const uint64_t* second_dev, const uint8_t* player_dev, uint64_t salt,
                   double* value_dev, double* probs_dev, void* stream);

/* ------------------------------------------------------------------------ trees */
int dm_pool_create(const dm_pool_config* cfg, dm_pool** out);
int dm_pool_destroy(dm_pool* pool);
int dm_pool_configure(dm_pool* pool, const dm_search_config* cfg);
/* MCTS.reset (mcts.py:243-245) for trees tree_ids_dev[0..n) (NULL = all trees) */
int dm_pool_reset(dm_pool* pool, const int32_t* tree_ids_dev, int n, void* stream);
/* fresh root on an arbitrary state (MCTSAgent.play's Node() branch, mcts.py:274-283) */
int dm_pool_set_root_state(dm_pool* pool, const int32_t* tree_ids_dev, int n, const uint64_t* first_dev,
                           const uint64_t* second_dev, const uint8_t* player_dev, void* stream);
/* MCTSAgent.play re-rooting: move each tree's root to the child whose state equals the
 * observed one, else start a fresh root there (mcts.py:267-283) */
int dm_pool_observe(dm_pool* pool, const int32_t* tree_ids_dev, int n, const uint64_t* first_dev,
                    const uint64_t* second_dev, const uint8_t* player_dev, void* stream);
/* start of MCTS.step (mcts.py:113-117): arm num_simulations on every tree whose root is not
 * final, clear the per-step stats.  Root expansion and Dirichlet noise happen inside the
 * search calls below. */
int dm_pool_begin_step(dm_pool* pool, void* stream);
/* the same for the listed trees only; all other trees sit this step out (batched tournaments:
 * an MCTSAgent searches only when it is its turn, tournament.py:97-108) */
int dm_pool_begin_step_ids(dm_pool* pool, const int32_t* tree_ids_dev, int n, void* stream);
/* host-supplied Dirichlet noise for the next root-noise application, noise_dev[n_trees][DM_MAX_ACTIONS]
 * in child order (parity tests; NULL = sample on device) */
int dm_pool_set_noise(dm_pool* pool, const double* noise_dev);
/* run every armed simulation to completion inside one kernel; only for eval kinds
 * NONE / UNIFORM / HASHED (no external evaluator) */
int dm_pool_search_local(dm_pool* pool, void* stream);
/* external evaluator round, phase 1 (tree_search + terminal handling, mcts.py:150-185):
 * every armed tree runs simulations until one reaches a non-final leaf; those leaves are
 * written compactly to the pool's leaf buffers. */
int dm_pool_select(dm_pool* pool, void* stream);
/* device buffers filled by dm_pool_select: states of the pending leaves and their count
 * (capacity n_trees * DM_MAX_LEAVES entries) */
int dm_pool_leaf_buffers(dm_pool* pool, uint64_t** first_dev, uint64_t** second_dev, uint8_t** player_dev,
                         int32_t** tree_dev, int32_t** count_dev);
/* dense batches (dm_pool_round with selfplay != 0): tree t's leaf lives in slot t; index_dev[0..count) lists the
 * slots whose leaf needs the evaluator this round, ascending (trees served by the evaluation cache, aliased to
 * an equal leaf of the same round, or sitting the round out are left out).  Feed it to
 * dm_net_evaluate_indexed. */
int dm_pool_leaf_index(dm_pool* pool, int32_t** index_dev);
/* cached_state_evaluator (train.py:408-414) on the device: a 2^log2_entries-entry table in HBM keyed by the
 * state -> (priors[DM_MAX_ACTIONS], value) of the external evaluator, probed by every selection of
 * dm_pool_select / dm_selfplay_select / dm_pool_round (exact mode).  A hit needs no evaluation (the leaf is
 * expanded from the stored row at the next expand step), equal leaves of one round are evaluated once, and
 * the evaluator's outputs are inserted when their leaves are expanded.  Result-transparent for a fixed
 * evaluator: visit counts are bit-identical with and without it.  is_f64 selects the row type and must
 * match the evaluator outputs later passed to dm_pool_expand_backup / dm_pool_round.  Call
 * dm_pool_cache_clear after every weight transfer (the reference clears its lru_cache there,
 * train.py:169-171 via a new worker net). */
int dm_pool_cache_enable(dm_pool* pool, int log2_entries, int is_f64);
int dm_pool_cache_clear(dm_pool* pool, void* stream);
/* phase 2 (expand_node + rollout mix + backpropagate, mcts.py:186-231): values_dev[slot]
 * and priors_dev[slot][prior_stride] per pending leaf, as float (is_f64 = 0) or double. */
int dm_pool_expand_backup(dm_pool* pool, const void* priors_dev, int prior_stride, const void* values_dev,
                          int is_f64, void* stream);
/* phase 2 of the previous round and phase 1 of the next in ONE launch (the body of the simulation
 * loop, mcts.py:119-138): every tree with a pending leaf expands it and backs it up from
 * values_dev / priors_dev (indexed by the slot the leaf was handed out in), then selects its next
 * leaf into the leaf buffers.  selfplay != 0 gives the selection dm_selfplay_select's move-making
 * (needs dm_selfplay_enable) and a DENSE leaf batch: tree t's leaf is always in slot t, which removes the
 * shared slot counter from the hot loop; the last block of the launch compacts the slots that need the
 * evaluator into dm_pool_leaf_index and stores their number in the leaf count.  Exact mode only
 * (leaves_per_round 1).
 * The evaluator outputs of the previous batch must stay untouched until this call has run. */
int dm_pool_round(dm_pool* pool, const void* priors_dev, int prior_stride, const void* values_dev, int is_f64,
                  int selfplay, void* stream);
/* result of MCTS.step (mcts.py:139-148): visits_dev[n_trees][DM_MAX_ACTIONS] root child visit
 * counts by action, stats_dev[n_trees][2] = (with expansion, without expansion) */
int dm_pool_root_visits(dm_pool* pool, uint32_t* visits_dev, uint32_t* stats_dev, void* stream);
/* root.children iteration order and the root's own visit count; order_dev[n_trees][DM_MAX_ACTIONS] */
int dm_pool_root_children(dm_pool* pool, uint8_t* order_dev, int32_t* count_dev, uint32_t* root_visits_dev,
                          void* stream);
/* Node.probability (after Dirichlet mixing, mcts.py:233-241) and Node.value_sum (mcts.py:27-54) of the root's
 * children in child order, [n_trees][DM_MAX_ACTIONS] doubles each; either pointer may be NULL */
int dm_pool_root_child_stats(dm_pool* pool, double* prior_dev, double* value_sum_dev, void* stream);
/* self_play's re-rooting with subtree reuse (mcts.py:105-108); action_dev[t] < 0 leaves tree t alone */
int dm_pool_advance(dm_pool* pool, const int32_t* action_dev, void* stream);
/* current root states; any pointer may be NULL */
int dm_pool_states(dm_pool* pool, uint64_t* first_dev, uint64_t* second_dev, uint8_t* player_dev,
                   uint8_t* is_final_dev, float* outcome_dev, void* stream);
/* counters since creation (synchronises).  Returns DM_ERR_CAPACITY (with out[] filled) when [8] grew since the
 * previous call -- each capacity error is reported once, the pool stays usable: [0] simulations, [1] leaf evaluations requested,
 * [2] terminal simulations, [3] nodes allocated high-water, [4] games finished (self-play),
 * [5] sum of selection depths, [6] children scanned, [7] children created, [8] capacity errors,
 * [9] moves played (self-play), [10] rollout plies, [11] invalid actions, [12] arena compactions,
 * [13] leaves served by the evaluation cache, [14] leaves aliased to an equal leaf of the same round,
 * [15] rows inserted into the evaluation cache.  [1] counts only leaves the evaluator really ran. */
int dm_pool_counters(dm_pool* pool, uint64_t out[16]);

/* --------------------------------------------------- continuous batched self-play
 * MCTS.self_play + create_self_play_examples (mcts.py:96-111, train.py:258-283) for every
 * tree at once: when a tree has used its simulations it picks its move (sampled in
 * proportion to visits for the first sample_move_cutoff plies, else argmax), records
 * (state, visit distribution), re-roots, and when the game ends labels the recorded
 * positions with the outcome and appends them to the replay ring, then restarts.
 * dm_selfplay_select replaces dm_pool_select in that loop. */
int dm_selfplay_enable(dm_pool* pool, int replay_capacity_positions);
int dm_selfplay_select(dm_pool* pool, void* stream);
/* replay ring buffers (device, owned by the pool): positions, visit distributions
 * pi[cap][DM_MAX_ACTIONS] float, z[cap] float in [-1,1] from the mover's perspective
 * (train.py:270-282), and the replay counter word *written_dev (monotonic): positions written so far in its low
 * DM_REPLAY_POS_BITS bits, games written so far in the bits above (one atomic hands out both, so games
 * are numbered in the order their positions lie in the ring). */
int dm_selfplay_replay(dm_pool* pool, uint64_t** first_dev, uint64_t** second_dev, uint8_t** player_dev,
                       float** pi_dev, float** z_dev, uint64_t** written_dev);
/* game_end_dev[i % DM_REPLAY_GAME_RING] = positions written up to and including game i: the trainer's window
 * of the last N GAMES (replay_buffer_max_size, train.py:149-151) starts at game_end[games - N - 1]. */
int dm_selfplay_games(dm_pool* pool, uint64_t** game_end_dev);

/* ------------------------------------------------------------------------- net
 * ConvolutionalNet(policy_features = num_actions, grid, in_channels = 3, num_residual,
 * channels, value_head_hidden_units) (convolutionalnet.py:22-51) in eval mode. */
typedef struct {
    int32_t game;
    int32_t grid;
    int32_t channels;
    int32_t num_residual;
    int32_t hidden;
    int32_t device;
    int32_t max_batch;   /* largest number of positions per dm_net_evaluate call */
    int32_t reserved0;
} dm_net_config;

int dm_net_create(const dm_net_config* cfg, dm_net** out);
int dm_net_destroy(dm_net* net);
/* load one state_dict tensor by its reference key name (SURVEY.md 8 a-N); data is a HOST
 * pointer to contiguous fp32 (num_batches_tracked is ignored). */
int dm_net_set_tensor(dm_net* net, const char* key, const float* host_data, int64_t numel);
/* fold BatchNorm into the convolutions and pack weights for the kernels; call after all
 * tensors are set (and again after every weight transfer, train.py:169-171). */
int dm_net_finalize(dm_net* net, void* stream);
/* GameNet.evaluate for a batch (gamenet.py:61-85): states -> value_dev[n] (absolute, [0,1])
 * and probs_dev[n][DM_MAX_ACTIONS] (legal-masked renormalised softmax).  count_dev (may be
 * NULL) is a device int holding the live batch size <= n.  precision: 0 = fp32 CUDA-core
 * path (1e-5), 1 = bf16 tcgen05 path (1e-2). */
int dm_net_evaluate(dm_net* net, int precision, int n, const int32_t* count_dev, const uint64_t* first_dev,
                    const uint64_t* second_dev, const uint8_t* player_dev, float* value_dev,
                    float* probs_dev, void* stream);
/* the same for a compacted dense batch: board i of the batch (i < *count_dev) is the state in slot index_dev[i]
 * and its results go to value_dev[index_dev[i]] / probs_dev[index_dev[i]] (dm_pool_leaf_index) */
int dm_net_evaluate_indexed(dm_net* net, int precision, int n, const int32_t* count_dev, const int32_t* index_dev,
                            const uint64_t* first_dev, const uint64_t* second_dev, const uint8_t* player_dev,
                            float* value_dev, float* probs_dev, void* stream);
/* GameNet.forward (gamenet.py:119-120 + hex adapters): raw value_dev[n] (pre-tanh) and
 * logits_dev[n][DM_MAX_ACTIONS] */
int dm_net_forward(dm_net* net, int precision, int n, const uint64_t* first_dev, const uint64_t* second_dev,
                   const uint8_t* player_dev, float* value_dev, float* logits_dev, void* stream);
/* greedy policy-network rollouts (evaluate_complex_rollouts.py:57-59): play argmax of the
 * masked policy from every state until the game ends; outcome_dev[n] */
int dm_net_rollout_greedy(dm_net* net, int precision, int n, const uint64_t* first_dev,
                          const uint64_t* second_dev, const uint8_t* player_dev, float* outcome_dev,
                          void* stream);
/* CUDA-event timing of the network kernels on their launch stream, per kernel class:
 * [0] encode + first conv, [1] trunk 3x3 convs (the tcgen05 kernels at precision 1), [2] heads (bf16 path: the
 * policy FC only), [3] heads finish (bf16 path).
 * dm_net_profile(net, 1) clears and starts, dm_net_profile_read synchronises and returns the
 * accumulated milliseconds and launch counts.  Used by bench.py for the roofline. */
int dm_net_profile(dm_net* net, int enable);
int dm_net_profile_read(dm_net* net, double ms_out[4], int64_t launches_out[4]);
/* diagnostics: switch on per-warp timing of dm_pool_round and return the buffer it fills every launch,
 * per_tree_dev[n_trees][4] = {cycles of the expand/backup phase, cycles of the selection phase,
 * (moves << 32 | games << 24 | terminal simulations << 16 | levels walked), children scanned} of that launch.
 * Used by scripts/tree_tail_probe.py to see which trees set a launch's duration. */
int dm_pool_debug_timing(dm_pool* pool, uint64_t** per_tree_dev);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
uint64_t dm_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DEEPMCTS_B200_H */
