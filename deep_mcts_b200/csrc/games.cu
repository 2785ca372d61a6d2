// Batched game-engine entry points (dm_game_*): one warp per state for the ordered
// legal-move lists, one thread per state for transitions / status / rollouts.
// Reference interfaces replaced: GameManager.* (game.py:76-113) of the four managers.
#include "games.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;

__global__ void k_game_initial(GameCfg cfg, int n, u64* first, u64* second, u8* player) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GState s = game_initial(cfg);
    first[i] = s.first;
    second[i] = s.second;
    player[i] = (u8)s.player;
}

__global__ void k_game_legal(GameCfg cfg, int n, const u64* first, const u64* second, const u8* player,
                             u8* list_out, u8* order_out, int* count_out) {
    __shared__ u8 s_order[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    __shared__ u8 s_list[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = blockIdx.x * kWarpsPerBlock + warp;
    if (i >= n) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    int cnt = warp_children_order(cfg, s, lane, s_order[warp], s_list[warp]);
    for (int j = lane; j < DM_MAX_ACTIONS; j += 32) {
        if (order_out) order_out[(size_t)i * DM_MAX_ACTIONS + j] = j < cnt ? s_order[warp][j] : 0xff;
        if (list_out) list_out[(size_t)i * DM_MAX_ACTIONS + j] = j < cnt ? s_list[warp][j] : 0xff;
    }
    if (lane == 0 && count_out) count_out[i] = cnt;
}

__global__ void k_game_child(GameCfg cfg, int n, const u64* first, const u64* second, const u8* player,
                             const int* action, u64* ofirst, u64* osecond, u8* oplayer, u8* ok) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    int a = action[i];
    bool legal = game_action_legal(cfg, s, a);
    GState r = s;
    if (legal) r = game_child(cfg, s, a);
    ofirst[i] = r.first;
    osecond[i] = r.second;
    oplayer[i] = (u8)r.player;
    if (ok) ok[i] = legal ? 1 : 0;
}

__global__ void k_game_status(GameCfg cfg, int n, const u64* first, const u64* second, const u8* player,
                              u8* is_final, float* outcome) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    float o;
    bool f = game_is_final(cfg, s, o);
    if (is_final) is_final[i] = f ? 1 : 0;
    if (outcome) outcome[i] = o;
}

__global__ void k_game_rollout(GameCfg cfg, int n, const u64* first, const u64* second, const u8* player,
                               int kind, u64 seed, float* outcome, int* plies) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    Philox rng;
    rng.init(seed, (u32)i, 0x726f6c6cu, 0);
    int p = 0;
    float o = game_rollout(cfg, s, kind, rng, p);
    outcome[i] = o;
    if (plies) plies[i] = p;
}

// uniform_state_evaluator (reference train.py:240-248): value 0.5, 1/len(legal) on legal actions
__global__ void k_eval_uniform(GameCfg cfg, int n, const int* count_dev, const u64* first, const u64* second,
                               const u8* player, float* value, float* probs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int live = count_dev ? min(*count_dev, n) : n;
    if (i >= live) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    bool extra;
    u64 m = game_legal_mask(cfg, s, extra);
    int cnt = __popcll(m) + (extra ? 1 : 0);
    float p = cnt ? (float)(1.0 / (double)cnt) : 0.f;
    float* row = probs + (size_t)i * DM_MAX_ACTIONS;
    for (int a = 0; a < DM_MAX_ACTIONS; ++a)
        row[a] = (a < cfg.cells ? ((m >> a) & 1) : (a == cfg.cells && extra)) ? p : 0.f;
    value[i] = 0.5f;
}

}  // namespace

extern "C" int dm_eval_uniform(int game, int grid, int n, const int32_t* count_dev, const uint64_t* first,
                               const uint64_t* second, const uint8_t* player, float* value, float* probs,
                               void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_eval_uniform: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player && value && probs, "dm_eval_uniform: bad arguments");
    k_eval_uniform<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n, count_dev,
                                                                      (const u64*)first, (const u64*)second, player,
                                                                      value, probs);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_game_num_actions(int game, int grid) {
    if (!game_cfg_valid(game, grid)) {
        dm_set_error("dm_game_num_actions: unsupported game/grid");
        return DM_ERR_INVALID;
    }
    return make_game_cfg(game, grid).num_actions;
}

extern "C" int dm_game_initial(int game, int grid, int n, uint64_t* first, uint64_t* second, uint8_t* player,
                               void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_game_initial: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player, "dm_game_initial: bad arguments");
    k_game_initial<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n, (u64*)first,
                                                                      (u64*)second, player);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_game_legal(int game, int grid, int n, const uint64_t* first, const uint64_t* second,
                             const uint8_t* player, uint8_t* list_dev, uint8_t* order_dev, int32_t* count_dev,
                             void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_game_legal: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player, "dm_game_legal: bad arguments");
    k_game_legal<<<(n + kWarpsPerBlock - 1) / kWarpsPerBlock, kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
        make_game_cfg(game, grid), n, (const u64*)first, (const u64*)second, player, list_dev, order_dev,
        count_dev);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_game_child(int game, int grid, int n, const uint64_t* first, const uint64_t* second,
                             const uint8_t* player, const int32_t* action, uint64_t* ofirst, uint64_t* osecond,
                             uint8_t* oplayer, uint8_t* ok, void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_game_child: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player && action && ofirst && osecond && oplayer,
                 "dm_game_child: bad arguments");
    k_game_child<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n,
                                                                    (const u64*)first, (const u64*)second, player,
                                                                    action, (u64*)ofirst, (u64*)osecond, oplayer, ok);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_game_status(int game, int grid, int n, const uint64_t* first, const uint64_t* second,
                              const uint8_t* player, uint8_t* is_final, float* outcome, void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_game_status: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player, "dm_game_status: bad arguments");
    k_game_status<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n,
                                                                     (const u64*)first, (const u64*)second, player,
                                                                     is_final, outcome);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_game_rollout(int game, int grid, int n, const uint64_t* first, const uint64_t* second,
                               const uint8_t* player, int rollout_kind, uint64_t seed, float* outcome,
                               int32_t* plies, void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_game_rollout: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player && outcome, "dm_game_rollout: bad arguments");
    DM_CHECK_ARG(rollout_kind >= DM_ROLLOUT_RANDOM && rollout_kind <= DM_ROLLOUT_MAX_ACTION,
                 "dm_game_rollout: bad rollout kind");
    k_game_rollout<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n, (const u64*)first,
                                                                   (const u64*)second, player, rollout_kind,
                                                                   (u64)seed, outcome, plies);
    DM_LAUNCH_CHECK();
    return DM_OK;
}
