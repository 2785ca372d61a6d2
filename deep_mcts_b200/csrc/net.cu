// dm_net_*: ConvolutionalNet leaf evaluation on the device (reference
// deep_mcts/convolutionalnet.py:22-170 behind GameNet.evaluate / forward, gamenet.py:61-128).
// precision 0 = fp32 CUDA cores (any architecture the reference can build),
// precision 1 = bf16 tcgen05 trunk (channels == 128, num_residual >= 1).
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "net_fp32.cuh"
// (net_fp32.cuh first: it defines live_count and the shared encode / heads kernels)
#include "net_bf16.cuh"

struct dm_net {
    dm_net_config cfg;
    GameCfg game;
    RowGeom geom;
    std::map<std::string, std::vector<float>> tensors;
    std::vector<void*> allocations;
    bool finalized = false;
    bool bf16_ready = false;
    int NA2 = 0;
    size_t rows_total = 0;
    // fp32 weights
    float *conv1_w = nullptr, *conv1_b = nullptr;
    __half* conv1_tab16 = nullptr;  // k_encode_conv1_table operand: padded fp16 table image
    __nv_bfloat16* stem_w16 = nullptr;  // k_stem_tc B operand: [8 K-chunks (hi 0-3, lo 4-7)][128][8]
    std::vector<float> h_conv1_b;       // folded first-layer bias (passed by value to k_stem_tc)
    bool stem_table = false;            // DM_STEM_TABLE=1 selects the table kernel for the first layer (A/B testing)
    int trunk_cluster = 2;              // fused trunk kernel as 2-CTA clusters with multicast weights; DM_TRUNK_CLUSTER=1: single CTAs (A/B)
    std::vector<float*> trunk_w, trunk_b;
    float *vconv_w = nullptr, *vfc1_w = nullptr, *vfc1_t = nullptr, *vfc1_b = nullptr, *vfc2_w = nullptr, *pfc_w = nullptr, *pfc_b = nullptr;
    float vconv_b = 0.f, vfc2_b = 0.f;
    std::vector<std::vector<float>> h_trunk_b;  // host copies for by-value kernel parameters
    std::vector<float> h_vconv_w;
    // bf16 weights
    std::vector<__nv_bfloat16*> trunk_w16;
    __nv_bfloat16* pfc_w16 = nullptr;  // [g*g*16][NA][8] (k_policy_fc_tc B operand)
    int fc_na = 0, fc_split = 1;
    bool conv_v1 = false;  // DM_CONV_V1=1 selects the non-persistent trunk kernel (A/B testing)
    bool pair_layout = false;  // 8x8 boards: pad-free interleaved layout (DM_NO_PAIR=1 disables, A/B testing)
    bool fused_blocks = false; // g = 8 / g = 7: one fused kernel per residual block (DM_NO_FUSE=1 disables)
    int sm_count = 148;
    float* fc_partial = nullptr;       // [split][max_batch rounded to 128][NA]
    // activations
    float *x0 = nullptr, *x1 = nullptr, *vplane = nullptr;
    __nv_bfloat16 *a0 = nullptr, *a1 = nullptr, *t16 = nullptr;
    // greedy rollout scratch
    u64 *r_first = nullptr, *r_second = nullptr;
    u8* r_player = nullptr;
    float *r_value = nullptr, *r_probs = nullptr;
    int* r_active = nullptr;
    int* r_list[2] = {nullptr, nullptr};   // greedy rollouts: compacted lists of the boards still in play
    // optional CUDA-event timing per kernel class (0 encode+conv1, 1 tcgen05/fp32 trunk conv, 2 heads)
    bool prof = false;
    std::vector<cudaEvent_t> prof_events;  // pairs
    std::vector<int> prof_class;           // class of pair i
    size_t prof_used = 0;                  // pairs in flight
    double prof_ms[4] = {0, 0, 0, 0};
    long long prof_launches[4] = {0, 0, 0, 0};
};

namespace {

int net_alloc_bytes(dm_net* net, void** out, size_t bytes, bool zero) {
    if (*out) return DM_OK;  // already allocated by an earlier dm_net_finalize: sizes depend only on the config
    void* p = nullptr;
    DM_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    net->allocations.push_back(p);
    if (zero) DM_CUDA(cudaMemset(p, 0, bytes ? bytes : 16));
    *out = p;
    return DM_OK;
}

template <typename T>
int net_upload(dm_net* net, T** out, const std::vector<T>& host) {
    void* p = (void*)*out;
    int rc = net_alloc_bytes(net, &p, host.size() * sizeof(T), false);
    if (rc != DM_OK) return rc;
    DM_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = (T*)p;
    return DM_OK;
}

#define DM_TRY(expr)                    \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != DM_OK) return rc__; \
    } while (0)

void prof_fold(dm_net* net) {
    for (size_t i = 0; i < net->prof_used; ++i) {
        float ms = 0.f;
        cudaEventSynchronize(net->prof_events[2 * i + 1]);
        if (cudaEventElapsedTime(&ms, net->prof_events[2 * i], net->prof_events[2 * i + 1]) == cudaSuccess) {
            net->prof_ms[net->prof_class[i]] += ms;
            net->prof_launches[net->prof_class[i]] += 1;
        }
    }
    net->prof_used = 0;
}

// records the start event of a timed launch; returns the pair index or -1
int prof_begin(dm_net* net, int cls, cudaStream_t st) {
    if (!net->prof) return -1;
    if (net->prof_used * 2 + 2 > net->prof_events.size()) {
        if (net->prof_events.size() >= 2 * 16384) prof_fold(net);
        else
            for (int k = 0; k < 512; ++k) {
                cudaEvent_t e;
                cudaEventCreate(&e);
                net->prof_events.push_back(e);
            }
        if (net->prof_class.size() < net->prof_events.size() / 2) net->prof_class.resize(net->prof_events.size() / 2);
    }
    int i = (int)net->prof_used++;
    net->prof_class[i] = cls;
    cudaEventRecord(net->prof_events[2 * i], st);
    return i;
}
void prof_end(dm_net* net, int i, cudaStream_t st) {
    if (i >= 0) cudaEventRecord(net->prof_events[2 * i + 1], st);
}

void free_device(dm_net* net) {
    for (void* p : net->allocations) cudaFree(p);
    net->allocations.clear();
    net->trunk_w.clear();
    net->trunk_b.clear();
    net->trunk_w16.clear();
}

// Greedy policy playouts (evaluate_complex_rollouts.py:57-59) keep a compacted list of the boards still in play: a
// ply's network batch is exactly those boards (dm_net_evaluate_indexed semantics: board i of the batch is state
// list[i], its priors land in row list[i]), finished boards cost nothing more.
// init: final states report their outcome, the others enter the list
__global__ void k_greedy_init(GameCfg cfg, int n, const u64* first, const u64* second, const u8* player, float* outcome,
                              int* list, int* count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    float o;
    if (game_is_final(cfg, s, o)) outcome[i] = o;
    else list[atomicAdd(count, 1)] = i;
}

// one ply: every listed board plays argmax(probs) (np.argmax: first maximum); boards that are still not final
// enter the next list, finished ones report their outcome
__global__ void k_greedy_step(GameCfg cfg, const int* count_in, const int* list_in, u64* first, u64* second, u8* player,
                              const float* probs, float* outcome, int* list_out, int* count_out) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *count_in) return;
    const int i = list_in[k];
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    const float* p = probs + (size_t)i * DM_MAX_ACTIONS;
    int best = 0;
    float bv = p[0];
    for (int a = 1; a < cfg.num_actions; ++a)
        if (p[a] > bv) {
            bv = p[a];
            best = a;
        }
    if (!game_action_legal(cfg, s, best)) {  // degenerate all-zero policy: lowest legal action
        bool extra;
        u64 m = game_legal_mask(cfg, s, extra);
        best = m ? (__ffsll((long long)m) - 1) : cfg.cells;
    }
    s = game_child(cfg, s, best);
    first[i] = s.first;
    second[i] = s.second;
    player[i] = (u8)s.player;
    float o;
    if (game_is_final(cfg, s, o)) outcome[i] = o;
    else list_out[atomicAdd(count_out, 1)] = i;
}

int run_forward(dm_net* net, int precision, int n, const int* count_dev, const int* index, const u64* first,
                const u64* second, const u8* player, int mode, float* value_out, float* probs_out, cudaStream_t st) {
    const GameCfg& g = net->game;
    const int C = net->cfg.channels, R = net->cfg.num_residual, P0 = g.g * g.g, gp = g.g + 2;
    const int L = 2 * R + 1;
    size_t enc_smem = (size_t)(27 * C + C + netk::kEncodeBoards * 3 * gp * gp) * sizeof(float);
    int enc_blocks = (n + netk::kEncodeBoards - 1) / netk::kEncodeBoards;
    netk::HeadParams hp{net->pfc_b, net->vfc1_w, net->vfc1_t, net->vfc1_b, net->vfc2_w, net->vfc2_b, net->cfg.hidden};
    if (precision == 0) {
        int pe = prof_begin(net, 0, st);
        netk::k_encode_conv1<false><<<enc_blocks, 256, enc_smem, st>>>(g, n, count_dev, index, first, second, player,
                                                                       net->conv1_w, net->conv1_b, C, net->x0, nullptr, 0);
        prof_end(net, pe, st);
        DM_LAUNCH_CHECK();
        size_t conv_smem = (size_t)gp * gp * C * sizeof(float);
        for (int i = 0; i < R; ++i) {
            { int pc = prof_begin(net, 1, st); netk::k_conv3x3_f32<<<n, 256, conv_smem, st>>>(g.g, C, n, count_dev, net->x0, net->trunk_w[2 * i],
                                                           net->trunk_b[2 * i], nullptr, net->x1); prof_end(net, pc, st); }
            DM_LAUNCH_CHECK();
            { int pc = prof_begin(net, 1, st); netk::k_conv3x3_f32<<<n, 256, conv_smem, st>>>(g.g, C, n, count_dev, net->x1, net->trunk_w[2 * i + 1],
                                                           net->trunk_b[2 * i + 1], net->x0, net->x0); prof_end(net, pc, st); }
            DM_LAUNCH_CHECK();
        }
        netk::k_vplane_f32<<<(n * P0 + 127) / 128, 128, 0, st>>>(P0, C, n, count_dev, net->x0, net->vconv_w,
                                                                 net->vconv_b, net->vplane);
        DM_LAUNCH_CHECK();
        { int pc = prof_begin(net, 1, st); netk::k_conv3x3_f32<<<n, 256, conv_smem, st>>>(g.g, C, n, count_dev, net->x0, net->trunk_w[L - 1],
                                                       net->trunk_b[L - 1], nullptr, net->x1); prof_end(net, pc, st); }
        DM_LAUNCH_CHECK();
        constexpr int kB = 4;
        size_t hs = (size_t)kB * P0 * C * sizeof(float) + (size_t)kB * (net->NA2 + net->cfg.hidden) * sizeof(float);
        int ph = prof_begin(net, 2, st);
        netk::k_heads<float, kB><<<(n + kB - 1) / kB, 256, hs, st>>>(g, n, count_dev, index, first, second, player, net->x1,
                                                                    net->pfc_w, C, net->NA2, net->vplane, hp, mode,
                                                                    value_out, probs_out);
        prof_end(net, ph, st);
        DM_LAUNCH_CHECK();
        return DM_OK;
    }
    if (!net->bf16_ready) {
        dm_set_error("bf16 path needs channels == 128 and num_residual >= 1");
        return DM_ERR_INVALID;
    }
    int pe = prof_begin(net, 0, st);
    if (net->stem_table) {
        const int enc3_groups = (n + netk::kEncode3Boards - 1) / netk::kEncode3Boards;
        const int enc3_blocks = std::min(enc3_groups, netk::kEncode3CtasPerSm * net->sm_count);
        netk::k_encode_conv1_table<<<enc3_blocks, 256, netk::encode3_smem_bytes(C, g.g), st>>>(g, n, count_dev, index, first, second, player,
                                                                             net->conv1_tab16, C, net->a0,
                                                                             net->rows_total, net->pair_layout ? 1 : 0);
    } else {
        netk::StemArgs sa;
        sa.cfg = g; sa.n = n; sa.count_dev = count_dev; sa.index = index; sa.first = first; sa.second = second;
        sa.player = player; sa.w = net->stem_w16; sa.out = net->a0; sa.rows_total = net->rows_total;
        sa.pair_layout = net->pair_layout ? 1 : 0;
        memcpy(sa.bias_c, net->h_conv1_b.data(), sizeof(sa.bias_c));
        const int stem_items = (n + 1) / 2;
        const int stem_blocks = std::min(stem_items, netk::kStemCtasPerSm * net->sm_count);
        netk::k_stem_tc<<<stem_blocks, netk::kStemThreads, netk::stem_smem_bytes(), st>>>(sa);
    }
    prof_end(net, pe, st);
    DM_LAUNCH_CHECK();
    const bool v1 = net->conv_v1, pair = net->pair_layout;
    const size_t conv_smem = v1 ? netk::conv_smem_bytes(net->geom)
                                : (pair ? netk::conv2_smem_bytes_pair() : netk::conv2_smem_bytes(net->geom));
    const int conv_items = pair ? (n + 3) / 4 : (int)(((size_t)n * net->geom.P + netk::kItemRows - 1) / netk::kItemRows);
    const int conv_blocks = v1 ? (int)(((size_t)n * net->geom.P + netk::kRowsPerCta - 1) / netk::kRowsPerCta)
                               : (conv_items < net->sm_count ? conv_items : net->sm_count);
    const int conv_threads = v1 ? netk::kConvThreads : netk::kConv2Threads;
    // pair layout: 2-CTA clusters sharing the weight stream (needs an even grid and an even item capacity)
    const bool conv_cl2 = pair && !v1 && net->trunk_cluster == 2 && conv_blocks >= 2 && (conv_items % 2) == 0;
    auto conv_kernel = v1 ? netk::k_conv3x3_tc
                          : (pair ? (conv_cl2 ? netk::k_conv3x3_tc2<true, 2> : netk::k_conv3x3_tc2<true, 1>)
                                  : netk::k_conv3x3_tc2<false, 1>);
    const int conv_grid = conv_cl2 ? (conv_blocks & ~1) : conv_blocks;
    netk::ConvArgs a;
    memset(&a, 0, sizeof(a));
    a.rows_total = net->rows_total;
    a.n_boards = n;
    a.count_dev = count_dev;
    a.geom = net->geom;
    if (net->fused_blocks) {
        // one fused launch per residual block: conv1 -> T (shared memory) -> conv2 + skip
        netk::ResBlockArgs ra;
        for (int i0 = 0; i0 < R; i0 += netk::kMaxFusedBlocks) {
            memset(&ra, 0, sizeof(ra));
            ra.x = net->a0;
            ra.rows_total = net->rows_total;
            ra.n_boards = n;
            ra.count_dev = count_dev;
            ra.n_blocks = std::min(netk::kMaxFusedBlocks, R - i0);
            for (int k = 0; k < ra.n_blocks; ++k) {
                const int i = i0 + k;
                ra.w1[k] = net->trunk_w16[2 * i];
                ra.w2[k] = net->trunk_w16[2 * i + 1];
                memcpy(ra.bias1_c[k], net->h_trunk_b[2 * i].data(), sizeof(ra.bias1_c[k]));
                memcpy(ra.bias2_c[k], net->h_trunk_b[2 * i + 1].data(), sizeof(ra.bias2_c[k]));
            }
            if (i0 + ra.n_blocks == R) {
                ra.vplane = net->vplane;
                ra.vhead_b = net->vconv_b;
                memcpy(ra.vw_c, net->h_vconv_w.data(), sizeof(ra.vw_c));
            }
            int pc = prof_begin(net, 1, st);
            // clusters of CL CTAs sharing the weight stream need a grid and an item capacity that are multiples of CL
            int cl = net->trunk_cluster;
            while (cl > 1 && (conv_blocks < cl || (conv_items % cl) != 0)) cl >>= 1;
            const int cblocks = conv_blocks / cl * cl;
            const size_t rsm = g.g == 8 ? netk::resblock_smem_bytes<8>() : netk::resblock_smem_bytes<7>();
#define DM_RESBLOCK(G, CL) netk::k_resblock_tc<G, CL><<<cblocks, netk::kConv2Threads, rsm, st>>>(ra)
            if (g.g == 8) {
                if (cl == 2) DM_RESBLOCK(8, 2);
                else DM_RESBLOCK(8, 1);
            } else {
                if (cl == 2) DM_RESBLOCK(7, 2);
                else DM_RESBLOCK(7, 1);
            }
#undef DM_RESBLOCK
            prof_end(net, pc, st);
            DM_LAUNCH_CHECK();
        }
    }
    for (int i = 0; i < (net->fused_blocks ? 0 : R); ++i) {
        a.in = net->a0; a.out = net->a1; a.out_nhwc = nullptr; a.residual = nullptr;
        a.w = net->trunk_w16[2 * i]; a.bias = net->trunk_b[2 * i]; a.vhead_w = nullptr; a.vplane = nullptr;
        memcpy(a.bias_c, net->h_trunk_b[2 * i].data(), sizeof(a.bias_c));
        {
            int pc = prof_begin(net, 1, st);
            conv_kernel<<<conv_grid, conv_threads, conv_smem, st>>>(a);
            prof_end(net, pc, st);
        }
        DM_LAUNCH_CHECK();
        a.in = net->a1; a.out = net->a0; a.residual = net->a0;
        a.w = net->trunk_w16[2 * i + 1]; a.bias = net->trunk_b[2 * i + 1];
        memcpy(a.bias_c, net->h_trunk_b[2 * i + 1].data(), sizeof(a.bias_c));
        if (i == R - 1) {
            a.vhead_w = net->vconv_w; a.vhead_b = net->vconv_b; a.vplane = net->vplane;
            memcpy(a.vw_c, net->h_vconv_w.data(), sizeof(a.vw_c));
        }
        {
            int pc = prof_begin(net, 1, st);
            conv_kernel<<<conv_grid, conv_threads, conv_smem, st>>>(a);
            prof_end(net, pc, st);
        }
        DM_LAUNCH_CHECK();
    }
    a.in = net->a0; a.out = nullptr; a.out_nhwc = nullptr; a.out_fc = net->t16; a.residual = nullptr;
    a.w = net->trunk_w16[L - 1]; a.bias = net->trunk_b[L - 1]; a.vhead_w = nullptr; a.vplane = nullptr;
    memcpy(a.bias_c, net->h_trunk_b[L - 1].data(), sizeof(a.bias_c));
    {
        int pc = prof_begin(net, 1, st);
        conv_kernel<<<conv_grid, conv_threads, conv_smem, st>>>(a);
        prof_end(net, pc, st);
    }
    DM_LAUNCH_CHECK();
    const int n_cap = (net->cfg.max_batch + 127) / 128 * 128;
    netk::FcArgs fa{net->t16, net->pfc_w16, net->fc_partial, P0, net->fc_na, net->fc_split, n, n_cap, count_dev};
    int ph = prof_begin(net, 2, st);
    netk::k_policy_fc_tc<<<dim3((n + 127) / 128, net->fc_split), netk::kFcThreads, netk::fc_smem_bytes(net->fc_na), st>>>(fa);
    prof_end(net, ph, st);
    ++g_dm_launches;
    ph = prof_begin(net, 3, st);
    {
        const int fb = (n + netk::kFinishBoards - 1) / netk::kFinishBoards, ft = netk::kFinishBoards * 32;
        const size_t fs = (size_t)P0 * net->cfg.hidden * sizeof(float);
#define DM_FINISH(GAME, G)                                                                                             \
    netk::k_heads_finish<GAME, G><<<fb, ft, fs, st>>>(g, n, count_dev, index, first, second, player, net->fc_partial,   \
                                                      net->fc_split, n_cap, net->fc_na, net->vplane, hp, mode, value_out, \
                                                      probs_out)
        if (g.game == DM_OTHELLO && g.g == 8) DM_FINISH(DM_OTHELLO, 8);
        else if (g.game == DM_HEX && g.g == 7) DM_FINISH(DM_HEX, 7);
        else if (g.game == DM_HEX_SWAP && g.g == 7) DM_FINISH(DM_HEX_SWAP, 7);
        else DM_FINISH(-1, 0);
#undef DM_FINISH
    }
    prof_end(net, ph, st);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

}  // namespace

extern "C" int dm_net_create(const dm_net_config* cfg, dm_net** out) {
    DM_CHECK_ARG(cfg && out, "dm_net_create: null argument");
    DM_CHECK_ARG(game_cfg_valid(cfg->game, cfg->grid), "dm_net_create: unsupported game/grid");
    DM_CHECK_ARG(cfg->channels >= 1 && cfg->channels <= 512 && cfg->num_residual >= 0 && cfg->hidden >= 1 &&
                     cfg->max_batch >= 1,
                 "dm_net_create: bad architecture");
    dm_net* net = new (std::nothrow) dm_net();
    DM_CHECK_ARG(net, "dm_net_create: out of host memory");
    net->cfg = *cfg;
    net->game = make_game_cfg(cfg->game, cfg->grid);
    net->geom = RowGeom::make(net->game.g);
    *out = net;
    return DM_OK;
}

extern "C" int dm_net_destroy(dm_net* net) {
    if (!net) return DM_OK;
    cudaSetDevice(net->cfg.device);
    free_device(net);
    for (cudaEvent_t e : net->prof_events) cudaEventDestroy(e);
    delete net;
    return DM_OK;
}

extern "C" int dm_net_set_tensor(dm_net* net, const char* key, const float* host_data, int64_t numel) {
    DM_CHECK_ARG(net && key && host_data && numel >= 0, "dm_net_set_tensor: bad arguments");
    net->tensors[key] = std::vector<float>(host_data, host_data + numel);
    net->finalized = false;
    return DM_OK;
}

extern "C" int dm_net_finalize(dm_net* net, void* stream) {
    DM_CHECK_ARG(net, "dm_net_finalize: null net");
    DM_CUDA(cudaSetDevice(net->cfg.device));
    DM_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const int C = net->cfg.channels, R = net->cfg.num_residual, H = net->cfg.hidden, g = net->game.g;
    const int A = net->game.num_actions, P0 = g * g, n = net->cfg.max_batch;
    HostWeights hw;
    std::string err;
    if (!build_host_weights(net->tensors, C, R, H, g, A, hw, err)) {
        dm_set_error("dm_net_finalize: " + err);
        return DM_ERR_INVALID;
    }
    DM_CUDA(cudaDeviceSynchronize());
    // Buffers are allocated on the first call and only refilled afterwards, so device pointers stay
    // stable across weight transfers (captured CUDA graphs of the search round remain valid).
    net->NA2 = hw.NA2;
    DM_TRY(net_upload(net, &net->conv1_w, hw.conv1_w));
    DM_TRY(net_upload(net, &net->conv1_b, hw.conv1_b));
    const int L = 2 * R + 1;
    if ((int)net->trunk_w.size() != L) {
        net->trunk_w.assign(L, nullptr);
        net->trunk_b.assign(L, nullptr);
    }
    for (int l = 0; l < L; ++l) {
        DM_TRY(net_upload(net, &net->trunk_w[l], hw.trunk_w[l]));
        DM_TRY(net_upload(net, &net->trunk_b[l], hw.trunk_b[l]));
    }
    net->h_trunk_b = hw.trunk_b;
    net->h_vconv_w = hw.vconv_w;
    DM_TRY(net_upload(net, &net->vconv_w, hw.vconv_w));
    net->vconv_b = hw.vconv_b;
    DM_TRY(net_upload(net, &net->vfc1_w, hw.vfc1_w));
    {
        std::vector<float> t((size_t)P0 * H);
        for (int j = 0; j < H; ++j)
            for (int p = 0; p < P0; ++p) t[(size_t)p * H + j] = hw.vfc1_w[(size_t)j * P0 + p];
        DM_TRY(net_upload(net, &net->vfc1_t, t));
    }
    DM_TRY(net_upload(net, &net->vfc1_b, hw.vfc1_b));
    DM_TRY(net_upload(net, &net->vfc2_w, hw.vfc2_w));
    net->vfc2_b = hw.vfc2_b;
    DM_TRY(net_upload(net, &net->pfc_w, hw.pfc_w));
    DM_TRY(net_upload(net, &net->pfc_b, hw.pfc_b));
    size_t act = (size_t)n * P0 * C;
    DM_TRY(net_alloc_bytes(net, (void**)&net->x0, act * sizeof(float), false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->x1, act * sizeof(float), false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->vplane, (size_t)n * P0 * sizeof(float), true));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_first, (size_t)n * 8, false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_second, (size_t)n * 8, false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_player, (size_t)n, false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_value, (size_t)n * 4, false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_probs, (size_t)n * DM_MAX_ACTIONS * 4, false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_active, 16, true));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_list[0], (size_t)n * sizeof(int), false));
    DM_TRY(net_alloc_bytes(net, (void**)&net->r_list[1], (size_t)n * sizeof(int), false));
    net->bf16_ready = false;
    if (C == netk::kC && R >= 1) {
        if ((int)net->trunk_w16.size() != L) net->trunk_w16.assign(L, nullptr);
        for (int l = 0; l < L; ++l) {
            std::vector<__nv_bfloat16> packed;
            pack_trunk_bf16(hw.trunk_w[l], C, packed);
            DM_TRY(net_upload(net, &net->trunk_w16[l], packed));
        }
        {
            // first-conv tables: tab[dy][pattern of 3 cells][C] and base[player][border class][C]
            std::vector<float> tab((size_t)3 * 27 * C, 0.f), basev((size_t)2 * 16 * C, 0.f);
            for (int dy = 0; dy < 3; ++dy)
                for (int pat = 0; pat < 27; ++pat) {
                    const int code[3] = {pat % 3, (pat / 3) % 3, pat / 9};
                    for (int dx = 0; dx < 3; ++dx) {
                        if (!code[dx]) continue;
                        const int tap = dy * 3 + dx, plane = code[dx] - 1;
                        for (int c = 0; c < C; ++c)
                            tab[((size_t)dy * 27 + pat) * C + c] += hw.conv1_w[((size_t)tap * 3 + plane) * C + c];
                    }
                }
            for (int pl = 0; pl < 2; ++pl)
                for (int cls = 0; cls < 16; ++cls) {
                    const int ym = cls & 3, xm = cls >> 2;
                    for (int c = 0; c < C; ++c) {
                        float acc = hw.conv1_b[c];
                        if (pl)
                            for (int tap = 0; tap < 9; ++tap) {
                                const int dy = tap / 3 - 1, dx = tap % 3 - 1;
                                const bool yin = dy == 0 || (dy < 0 ? (ym & 1) : (ym & 2));
                                const bool xin = dx == 0 || (dx < 0 ? (xm & 1) : (xm & 2));
                                if (yin && xin) acc += hw.conv1_w[((size_t)tap * 3 + 2) * C + c];
                            }
                        basev[((size_t)pl * 16 + cls) * C + c] = acc;
                    }
                }
            // one padded fp16 image [tab rows | base rows][C + 8]: the kernel bulk-copies it into shared memory as is
            const int CP = C + 8;
            std::vector<__half> tab16((size_t)netk::kEncode3Rows * CP, __float2half(0.f));
            for (int r = 0; r < 81; ++r)
                for (int c = 0; c < C; ++c) tab16[(size_t)r * CP + c] = __float2half_rn(tab[(size_t)r * C + c]);
            for (int r = 0; r < 32; ++r)
                for (int c = 0; c < C; ++c) tab16[(size_t)(81 + r) * CP + c] = __float2half_rn(basev[(size_t)r * C + c]);
            DM_TRY(net_upload(net, &net->conv1_tab16, tab16));
            // the same layer as a tcgen05 B operand: k = tap * 3 + plane (27 of 32 rows), hi part then lo part
            std::vector<__nv_bfloat16> sw((size_t)8 * C * 8, __float2bfloat16(0.f));
            for (int k = 0; k < 27; ++k)
                for (int c = 0; c < C; ++c) {
                    const float w = hw.conv1_w[(size_t)k * C + c];
                    const __nv_bfloat16 hi = __float2bfloat16_rn(w);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
                    sw[((size_t)(k >> 3) * C + c) * 8 + (k & 7)] = hi;
                    sw[((size_t)(4 + (k >> 3)) * C + c) * 8 + (k & 7)] = lo;
                }
            DM_TRY(net_upload(net, &net->stem_w16, sw));
            net->h_conv1_b = hw.conv1_b;
            const char* st_env = getenv("DM_STEM_TABLE");
            net->stem_table = st_env && st_env[0] == '1';
            DM_CUDA(cudaFuncSetAttribute(netk::k_encode_conv1_table, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)netk::encode3_smem_bytes(C, g)));
        }
        // policy FC weights as the tcgen05 B operand: [(pos*16 + kc)][a][8 channels]
        net->fc_na = netk::fc_padded_actions(A);
        net->fc_split = netk::fc_split(g);
        std::vector<__nv_bfloat16> p16((size_t)P0 * 16 * net->fc_na * 8, __float2bfloat16(0.f));
        for (int pos = 0; pos < P0; ++pos)
            for (int c = 0; c < C; ++c)
                for (int a = 0; a < A; ++a)
                    p16[(((size_t)pos * 16 + c / 8) * net->fc_na + a) * 8 + (c % 8)] =
                        __float2bfloat16(hw.pfc_w[((size_t)pos * C + c) * hw.NA2 + a]);
        DM_TRY(net_upload(net, &net->pfc_w16, p16));
        {
            const char* env = getenv("DM_CONV_V1");
            net->conv_v1 = env && env[0] == '1';
            cudaDeviceGetAttribute(&net->sm_count, cudaDevAttrMultiProcessorCount, net->cfg.device);
        }
        {
            const char* nopair = getenv("DM_NO_PAIR");
            net->pair_layout = (g == 8) && !net->conv_v1 && !(nopair && nopair[0] == '1');
        }
        {
            const char* nofuse = getenv("DM_NO_FUSE");
            // whole-board work items: 8x8 in the pair layout, 7x7 in the flat layout (64 rows per board)
            net->fused_blocks = (net->pair_layout || (g == 7 && !net->conv_v1)) && !(nofuse && nofuse[0] == '1');
        }
        net->rows_total = net->pair_layout
                              ? netk::pair_rows_total(n)
                              : std::max(netk::conv_rows_total(net->geom, n), netk::conv2_rows_total(net->geom, n));
        size_t abytes = (size_t)netk::kChunks * net->rows_total * 16;
        // zeroed once: pad rows are never written afterwards and must stay zero
        DM_TRY(net_alloc_bytes(net, (void**)&net->a0, abytes, true));
        DM_TRY(net_alloc_bytes(net, (void**)&net->a1, abytes, true));
        size_t n128 = ((size_t)n + 127) / 128 * 128;
        DM_TRY(net_alloc_bytes(net, (void**)&net->t16, n128 * P0 * C * sizeof(__nv_bfloat16), true));
        DM_TRY(net_alloc_bytes(net, (void**)&net->fc_partial, (size_t)net->fc_split * n128 * net->fc_na * sizeof(float),
                               true));
        DM_CUDA(cudaFuncSetAttribute(netk::k_conv3x3_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::conv_smem_bytes(net->geom)));
        DM_CUDA(cudaFuncSetAttribute(netk::k_conv3x3_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::conv2_smem_bytes(net->geom)));
        DM_CUDA(cudaFuncSetAttribute(netk::k_conv3x3_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::conv2_smem_bytes_pair()));
        DM_CUDA(cudaFuncSetAttribute(netk::k_conv3x3_tc2<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::conv2_smem_bytes_pair()));
        DM_CUDA(cudaFuncSetAttribute(netk::k_resblock_tc<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::resblock_smem_bytes<8>()));
        DM_CUDA(cudaFuncSetAttribute(netk::k_resblock_tc<7>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::resblock_smem_bytes<7>()));
        DM_CUDA(cudaFuncSetAttribute(netk::k_resblock_tc<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::resblock_smem_bytes<8>()));
        DM_CUDA(cudaFuncSetAttribute(netk::k_resblock_tc<7, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::resblock_smem_bytes<7>()));
        {
            const char* cl = getenv("DM_TRUNK_CLUSTER");
            net->trunk_cluster = (cl && cl[0] == '1') ? 1 : 2;   // clusters of 4 do not all fit at once (7.2 M against 9.7 M sims/s)
        }
        DM_CHECK_ARG((size_t)P0 * H * sizeof(float) <= 200 * 1024, "dm_net_finalize: value head too large");
        {
            const int fs = (int)((size_t)P0 * H * sizeof(float));
            DM_CUDA(cudaFuncSetAttribute(netk::k_heads_finish<-1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
            DM_CUDA(cudaFuncSetAttribute(netk::k_heads_finish<DM_OTHELLO, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
            DM_CUDA(cudaFuncSetAttribute(netk::k_heads_finish<DM_HEX, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
            DM_CUDA(cudaFuncSetAttribute(netk::k_heads_finish<DM_HEX_SWAP, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
        }
        DM_CUDA(cudaFuncSetAttribute(netk::k_policy_fc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)netk::fc_smem_bytes(net->fc_na)));
        net->bf16_ready = true;
    }
    {
        const int gp = g + 2;
        size_t conv_smem = (size_t)gp * gp * C * sizeof(float);
        size_t hs = (size_t)4 * P0 * C * sizeof(float) + (size_t)4 * (hw.NA2 + H) * sizeof(float);
        size_t enc = (size_t)(27 * C + C + netk::kEncodeBoards * 3 * gp * gp) * sizeof(float);
        DM_CHECK_ARG(conv_smem <= 227 * 1024 && hs <= 227 * 1024 && enc <= 227 * 1024,
                     "dm_net_finalize: architecture too large for the shared-memory staged kernels");
        DM_CUDA(cudaFuncSetAttribute(netk::k_conv3x3_f32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)conv_smem));
        DM_CUDA(cudaFuncSetAttribute(netk::k_heads<float, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hs));
        DM_CUDA(cudaFuncSetAttribute(netk::k_encode_conv1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc));
        DM_CUDA(cudaFuncSetAttribute(netk::k_encode_conv1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)enc));
    }
    DM_CUDA(cudaDeviceSynchronize());
    net->finalized = true;
    return DM_OK;
}

static int check_eval_args(dm_net* net, int precision, int n, const void* a, const void* b, const void* c,
                           const void* d, const void* e) {
    DM_CHECK_ARG(net, "dm_net: null net");
    if (!net->finalized) {
        dm_set_error("dm_net: call dm_net_finalize after setting the tensors");
        return DM_ERR_STATE;
    }
    DM_CHECK_ARG(precision == 0 || precision == 1, "dm_net: precision must be 0 (fp32) or 1 (bf16)");
    DM_CHECK_ARG(n >= 0 && n <= net->cfg.max_batch, "dm_net: batch larger than max_batch");
    if (n == 0) return 1;
    DM_CHECK_ARG(a && b && c && d && e, "dm_net: null device pointer");
    return DM_OK;
}

extern "C" int dm_net_evaluate(dm_net* net, int precision, int n, const int32_t* count_dev, const uint64_t* first,
                               const uint64_t* second, const uint8_t* player, float* value, float* probs,
                               void* stream) {
    int rc = check_eval_args(net, precision, n, first, second, player, value, probs);
    if (rc != DM_OK) return rc < 0 ? rc : DM_OK;
    return run_forward(net, precision, n, count_dev, nullptr, (const u64*)first, (const u64*)second, player, 0, value,
                       probs, (cudaStream_t)stream);
}

extern "C" int dm_net_evaluate_indexed(dm_net* net, int precision, int n, const int32_t* count_dev,
                                       const int32_t* index_dev, const uint64_t* first, const uint64_t* second,
                                       const uint8_t* player, float* value, float* probs, void* stream) {
    int rc = check_eval_args(net, precision, n, first, second, player, value, probs);
    if (rc != DM_OK) return rc < 0 ? rc : DM_OK;
    DM_CHECK_ARG(count_dev && index_dev, "dm_net_evaluate_indexed: count_dev and index_dev are required");
    return run_forward(net, precision, n, count_dev, index_dev, (const u64*)first, (const u64*)second, player, 0, value,
                       probs, (cudaStream_t)stream);
}

extern "C" int dm_net_forward(dm_net* net, int precision, int n, const uint64_t* first, const uint64_t* second,
                              const uint8_t* player, float* value, float* logits, void* stream) {
    int rc = check_eval_args(net, precision, n, first, second, player, value, logits);
    if (rc != DM_OK) return rc < 0 ? rc : DM_OK;
    return run_forward(net, precision, n, nullptr, nullptr, (const u64*)first, (const u64*)second, player, 1, value,
                       logits, (cudaStream_t)stream);
}

extern "C" int dm_net_rollout_greedy(dm_net* net, int precision, int n, const uint64_t* first,
                                     const uint64_t* second, const uint8_t* player, float* outcome, void* stream) {
    int rc = check_eval_args(net, precision, n, first, second, player, outcome, outcome);
    if (rc != DM_OK) return rc < 0 ? rc : DM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DM_CUDA(cudaMemcpyAsync(net->r_first, first, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    DM_CUDA(cudaMemcpyAsync(net->r_second, second, (size_t)n * 8, cudaMemcpyDeviceToDevice, st));
    DM_CUDA(cudaMemcpyAsync(net->r_player, player, (size_t)n, cudaMemcpyDeviceToDevice, st));
    // r_active = {count of list 0, count of list 1}; the lists alternate from ply to ply
    int* counts = net->r_active;
    DM_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int), st));
    k_greedy_init<<<(n + 127) / 128, 128, 0, st>>>(net->game, n, net->r_first, net->r_second, net->r_player, outcome,
                                                  net->r_list[0], counts);
    DM_LAUNCH_CHECK();
    const int max_plies = 2 * net->game.cells + 4;
    for (int ply = 0; ply <= max_plies; ++ply) {
        const int cur = ply & 1, nxt = cur ^ 1;
        // the boards still in play are the ply's whole batch; the host looks at their number only every fourth ply
        // (an empty list makes every kernel of a ply return at once)
        if ((ply & 3) == 0) {
            int live = 0;
            DM_CUDA(cudaMemcpyAsync(&live, counts + cur, sizeof(int), cudaMemcpyDeviceToHost, st));
            DM_CUDA(cudaStreamSynchronize(st));
            if (live == 0) return DM_OK;
        }
        rc = run_forward(net, precision, n, counts + cur, net->r_list[cur], net->r_first, net->r_second, net->r_player, 0,
                         net->r_value, net->r_probs, st);
        if (rc != DM_OK) return rc;
        DM_CUDA(cudaMemsetAsync(counts + nxt, 0, sizeof(int), st));
        k_greedy_step<<<(n + 127) / 128, 128, 0, st>>>(net->game, counts + cur, net->r_list[cur], net->r_first,
                                                       net->r_second, net->r_player, net->r_probs, outcome,
                                                       net->r_list[nxt], counts + nxt);
        DM_LAUNCH_CHECK();
    }
    dm_set_error("dm_net_rollout_greedy: games did not finish");
    return DM_ERR_STATE;
}

extern "C" int dm_net_profile(dm_net* net, int enable) {
    DM_CHECK_ARG(net, "dm_net_profile: null net");
    DM_CUDA(cudaSetDevice(net->cfg.device));
    prof_fold(net);
    net->prof = enable != 0;
    if (enable)
        for (int k = 0; k < 4; ++k) {
            net->prof_ms[k] = 0;
            net->prof_launches[k] = 0;
        }
    return DM_OK;
}

extern "C" int dm_net_profile_read(dm_net* net, double ms_out[4], int64_t launches_out[4]) {
    DM_CHECK_ARG(net && ms_out && launches_out, "dm_net_profile_read: null argument");
    DM_CUDA(cudaSetDevice(net->cfg.device));
    prof_fold(net);
    for (int k = 0; k < 4; ++k) {
        ms_out[k] = net->prof_ms[k];
        launches_out[k] = net->prof_launches[k];
    }
    return DM_OK;
}
