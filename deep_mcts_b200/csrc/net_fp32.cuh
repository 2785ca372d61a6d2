// ConvolutionalNet leaf evaluation, CUDA-core kernels shared by both precisions plus the fp32
// trunk (any channel count / board size; this is the 1e-5 path and the one the tiny
// tic-tac-toe net uses).
//
// Replaces (reference deep_mcts/): <game>/convolutionalnet.py states_to_tensor
// (othello 54-73, hex 64-94, hex_with_swap 67-97, tictactoe 51-68), convolutionalnet.py:53-170
// forward, the Hex policy un-transpose (hex/convolutionalnet.py:54-62) and
// GameNet.evaluate's tanh / sign / softmax / mask / renormalise (gamenet.py:61-85).
#pragma once
#include <cuda_bf16.h>

#include "net_pack.cuh"

namespace netk {

constexpr int kEncodeBoards = 4;  // boards per block in the encode + conv1 kernel

__device__ __forceinline__ int live_count(const int* count_dev, int n) {
    if (!count_dev) return n;
    int c = *count_dev;
    return c < n ? c : n;
}

// Input planes for one board: ch0 mover's stones, ch1 opponent's, ch2 constant player id;
// Hex variants transpose the board when SECOND is to move.  Written zero-padded:
// planes[ch][(y+1)*(g+2) + (x+1)].
__device__ __forceinline__ void write_planes(const GameCfg& cfg, u64 first, u64 second, int player, float* planes,
                                             int tid, int nthreads) {
    int g = cfg.g, gp = g + 2;
    bool transpose = (cfg.game == DM_HEX || cfg.game == DM_HEX_SWAP) && player == 0;
    u64 own = player ? first : second, opp = player ? second : first;
    for (int i = tid; i < 3 * gp * gp; i += nthreads) {
        int ch = i / (gp * gp), r = i % (gp * gp);
        int y = r / gp - 1, x = r % gp - 1;
        float v = 0.f;
        if (y >= 0 && y < g && x >= 0 && x < g) {
            int bit = transpose ? (x * g + y) : (y * g + x);
            v = ch == 0 ? (float)((own >> bit) & 1) : (ch == 1 ? (float)((opp >> bit) & 1) : (float)player);
        }
        planes[i] = v;
    }
}

// states -> planes -> conv1 3x3 (3 -> C) + folded BN + ReLU.
// kBf16 = false: out_f32[b][pos][C] (NHWC).  kBf16 = true: padded flat rows, channel chunks of
// 8: out_bf16[(kc*rows_total + row)*8 + j] (see RowGeom).
template <bool kBf16>
__global__ void __launch_bounds__(256) k_encode_conv1(GameCfg cfg, int n, const int* count_dev, const int* index,
                                                      const u64* first,
                                                      const u64* second, const u8* player, const float* __restrict__ w,
                                                      const float* __restrict__ bias, int C, float* out_f32,
                                                      __nv_bfloat16* out_bf16, size_t rows_total) {
    extern __shared__ __align__(16) float smem[];
    const int g = cfg.g, gp = g + 2, P0 = g * g;
    float* s_w = smem;                      // [27][C]
    float* s_b = s_w + 27 * C;              // [C]
    float* s_planes = s_b + C;              // [kEncodeBoards][3][gp*gp]
    int live = live_count(count_dev, n);
    int b0 = blockIdx.x * kEncodeBoards;
    if (b0 >= live) return;
    for (int i = threadIdx.x; i < 27 * C; i += blockDim.x) s_w[i] = w[i];
    for (int i = threadIdx.x; i < C; i += blockDim.x) s_b[i] = bias[i];
    for (int bb = 0; bb < kEncodeBoards; ++bb) {
        int b = b0 + bb;
        if (b < live) {
            const int src = index ? index[b] : b;   // board b of the batch is state index[b] (compacted dense batch)
            write_planes(cfg, first[src], second[src], player[src], s_planes + bb * 3 * gp * gp, threadIdx.x, blockDim.x);
        }
    }
    __syncthreads();
    const RowGeom geom = RowGeom::make(g);
    const int chunks = (C + 7) / 8;
    const int items = kEncodeBoards * P0 * chunks;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        int pos = it % P0, rest = it / P0;
        int kc = rest % chunks, bb = rest / chunks;
        int b = b0 + bb;
        if (b >= live) continue;
        int y = pos / g, x = pos % g;
        const float* pl = s_planes + bb * 3 * gp * gp;
        float acc[8];
        if ((C & 7) == 0) {
            // full chunk: 128-bit shared-memory weight reads (2 per tap and input plane)
            const float4 bias0 = *reinterpret_cast<const float4*>(s_b + kc * 8);
            const float4 bias1 = *reinterpret_cast<const float4*>(s_b + kc * 8 + 4);
            acc[0] = bias0.x; acc[1] = bias0.y; acc[2] = bias0.z; acc[3] = bias0.w;
            acc[4] = bias1.x; acc[5] = bias1.y; acc[6] = bias1.z; acc[7] = bias1.w;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int pp = (y + tap / 3) * gp + (x + tap % 3);
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = pl[ci * gp * gp + pp];
                    const float4 w0 = *reinterpret_cast<const float4*>(s_w + (tap * 3 + ci) * C + kc * 8);
                    const float4 w1 = *reinterpret_cast<const float4*>(s_w + (tap * 3 + ci) * C + kc * 8 + 4);
                    acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
                    acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
                    acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
                    acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = (kc * 8 + j < C) ? s_b[kc * 8 + j] : 0.f;
            for (int tap = 0; tap < 9; ++tap) {
                int dy = tap / 3 - 1, dx = tap % 3 - 1;
                int pp = (y + dy + 1) * gp + (x + dx + 1);
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    float v = pl[ci * gp * gp + pp];
                    const float* wr = s_w + (tap * 3 + ci) * C + kc * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (kc * 8 + j < C) acc[j] = fmaf(v, wr[j], acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
        if (kBf16) {
            size_t row = (size_t)geom.halo + (size_t)b * geom.P + y * geom.Wp + x;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(acc[0], acc[1]), p1 = __floats2bfloat162_rn(acc[2], acc[3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(acc[4], acc[5]), p3 = __floats2bfloat162_rn(acc[6], acc[7]);
            uint4 v;
            v.x = *reinterpret_cast<u32*>(&p0);
            v.y = *reinterpret_cast<u32*>(&p1);
            v.z = *reinterpret_cast<u32*>(&p2);
            v.w = *reinterpret_cast<u32*>(&p3);
            *reinterpret_cast<uint4*>(out_bf16 + ((size_t)kc * rows_total + row) * 8) = v;
        } else {
            float* o = out_f32 + ((size_t)b * P0 + pos) * C + kc * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (kc * 8 + j < C) o[j] = acc[j];
        }
    }
}

// fp32 3x3 conv C -> C + bias (+ residual) + ReLU on NHWC activations, one block per board.
// Thread item = (output channel, strip of up to 8 positions).
__global__ void __launch_bounds__(256) k_conv3x3_f32(int g, int C, int n, const int* count_dev,
                                                     const float* __restrict__ in, const float* __restrict__ w,
                                                     const float* __restrict__ bias, const float* residual,
                                                     float* out) {
    extern __shared__ float s_in[];  // [(g+2)^2][C], zero padded
    int b = blockIdx.x;
    if (b >= live_count(count_dev, n)) return;
    const int gp = g + 2, P0 = g * g;
    for (int i = threadIdx.x; i < gp * gp * C; i += blockDim.x) {
        int pp = i / C, c = i % C;
        int y = pp / gp - 1, x = pp % gp - 1;
        s_in[i] = (y >= 0 && y < g && x >= 0 && x < g) ? in[((size_t)b * P0 + y * g + x) * C + c] : 0.f;
    }
    __syncthreads();
    const int strips = (P0 + 7) / 8;
    for (int it = threadIdx.x; it < strips * C; it += blockDim.x) {
        int co = it % C, strip = it / C;
        float acc[8];
        int base[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int pos = strip * 8 + i;
            if (pos >= P0) pos = P0 - 1;  // duplicate work, result discarded
            base[i] = ((pos / g + 1) * gp + (pos % g + 1)) * C;
            acc[i] = bias[co];
        }
        for (int tap = 0; tap < 9; ++tap) {
            int shift = ((tap / 3 - 1) * gp + (tap % 3 - 1)) * C;
            const float* wt = w + (size_t)tap * C * C + co;
            for (int ci = 0; ci < C; ++ci) {
                float wv = __ldg(wt + (size_t)ci * C);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] = fmaf(s_in[base[i] + shift + ci], wv, acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int pos = strip * 8 + i;
            if (pos >= P0) break;
            size_t o = ((size_t)b * P0 + pos) * C + co;
            float v = acc[i];
            if (residual) v += residual[o];
            out[o] = fmaxf(v, 0.f);
        }
    }
}

// value-head 1x1 conv (C -> 1) + folded BN + ReLU on fp32 NHWC: vplane[b][pos]
__global__ void k_vplane_f32(int P0, int C, int n, const int* count_dev, const float* __restrict__ x,
                             const float* __restrict__ wv, float bv, float* vplane) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= live_count(count_dev, n) * P0) return;
    const float* row = x + (size_t)i * C;
    float acc = bv;
    for (int c = 0; c < C; ++c) acc = fmaf(row[c], wv[c], acc);
    vplane[i] = fmaxf(acc, 0.f);
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }

struct HeadParams {
    const float* pfc_b;   // [A]
    const float* vfc1_w;  // [H][P0]
    const float* vfc1_t;  // [P0][H] (transposed copy for coalesced reads)
    const float* vfc1_b;  // [H]
    const float* vfc2_w;  // [H]
    float vfc2_b;
    int H;
};

// Policy FC + value FCs + GameNet.evaluate post-processing.
//   t_act  [n][K = P0*C] policy-conv output, k = pos*C + c (dense NHWC)
//   pfc_w  [K][NA2]      (same element type as t_act)
//   vplane [n][P0]
// mode 0: evaluate -> value in [0,1], probs masked + renormalised; mode 1: raw value, logits.
// One block = kBoards boards; threads split (action pair, K slice).
template <typename T, int kBoards>
__global__ void __launch_bounds__(256) k_heads(GameCfg cfg, int n, const int* count_dev, const int* index,
                                               const u64* first,
                                               const u64* second, const u8* player, const T* __restrict__ t_act,
                                               const T* __restrict__ pfc_w, int C, int NA2,
                                               const float* __restrict__ vplane, HeadParams hp, int mode,
                                               float* value_out, float* probs_out) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int P0 = cfg.g * cfg.g, K = P0 * C, A = cfg.num_actions;
    T* s_t = reinterpret_cast<T*>(s_raw);                                  // [kBoards][K]
    float* s_logits = reinterpret_cast<float*>(s_raw + (size_t)kBoards * K * sizeof(T));  // [kBoards][NA2]
    float* s_h = s_logits + kBoards * NA2;                                 // [kBoards][H]
    int live = live_count(count_dev, n);
    int b0 = blockIdx.x * kBoards;
    if (b0 >= live) return;
    int nb = min(kBoards, live - b0);
    // stage activations (16-byte vectors when the board size allows)
    {
        size_t bytes = (size_t)nb * K * sizeof(T);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(t_act + (size_t)b0 * K);
        if ((bytes % 16 == 0) && ((reinterpret_cast<size_t>(src) % 16) == 0)) {
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(s_raw);
            for (size_t i = threadIdx.x; i < bytes / 16; i += blockDim.x) d4[i] = s4[i];
        } else {
            for (size_t i = threadIdx.x; i < bytes; i += blockDim.x) s_raw[i] = src[i];
        }
        for (int i = threadIdx.x; i < kBoards * NA2; i += blockDim.x) s_logits[i] = 0.f;
    }
    __syncthreads();
    // ---- policy FC: logits[b][a] = sum_k t[b][k] * W[k][a]
    {
        const int pairs = NA2 / 2;
        const int ksplit = max(1, (int)blockDim.x / pairs);
        int pair = threadIdx.x % pairs, ks = threadIdx.x / pairs;
        if (ks < ksplit) {
            float acc0[kBoards], acc1[kBoards];
#pragma unroll
            for (int b = 0; b < kBoards; ++b) acc0[b] = acc1[b] = 0.f;
            int k_per = (K + ksplit - 1) / ksplit;
            int k_lo = ks * k_per, k_hi = min(K, k_lo + k_per);
            for (int k = k_lo; k < k_hi; ++k) {
                float w0 = to_float(pfc_w[(size_t)k * NA2 + 2 * pair]);
                float w1 = to_float(pfc_w[(size_t)k * NA2 + 2 * pair + 1]);
#pragma unroll
                for (int b = 0; b < kBoards; ++b) {
                    float x = to_float(s_t[(size_t)b * K + k]);
                    acc0[b] = fmaf(x, w0, acc0[b]);
                    acc1[b] = fmaf(x, w1, acc1[b]);
                }
            }
#pragma unroll
            for (int b = 0; b < kBoards; ++b) {
                atomicAdd(&s_logits[b * NA2 + 2 * pair], acc0[b]);
                atomicAdd(&s_logits[b * NA2 + 2 * pair + 1], acc1[b]);
            }
        }
    }
    // ---- value head FC1 (+ReLU): h[b][j]
    for (int i = threadIdx.x; i < nb * hp.H; i += blockDim.x) {
        int b = i / hp.H, j = i % hp.H;
        const float* vp = vplane + (size_t)(b0 + b) * P0;
        const float* wr = hp.vfc1_w + (size_t)j * P0;
        float acc = hp.vfc1_b[j];
        for (int p = 0; p < P0; ++p) acc = fmaf(vp[p], wr[p], acc);
        s_h[b * hp.H + j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    // ---- per board: FC2, Hex un-transpose, tanh/sign, softmax, mask, renormalise (one warp each)
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int b = warp; b < nb; b += nwarps) {
        int gb = b0 + b;
        if (index) gb = index[gb];   // states are read and results written at the leaf's own slot
        GState s;
        s.first = first[gb];
        s.second = second[gb];
        s.player = player[gb];
        float v = 0.f;
        for (int j = lane; j < hp.H; j += 32) v = fmaf(s_h[b * hp.H + j], hp.vfc2_w[j], v);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(DM_FULL_MASK, v, off);
        v += hp.vfc2_b;
        bool transpose = (cfg.game == DM_HEX || cfg.game == DM_HEX_SWAP) && s.player == 0;
        // logits by board action
        float lg[3];
        int na = 0;
        for (int a = lane; a < A; a += 32, ++na) {
            int src = a;
            if (transpose && a < P0) src = (a % cfg.g) * cfg.g + a / cfg.g;
            lg[na] = s_logits[b * NA2 + src] + hp.pfc_b[src];
        }
        if (mode == 1) {
            if (lane == 0) value_out[gb] = v;
            na = 0;
            for (int a = lane; a < DM_MAX_ACTIONS; a += 32, ++na)
                probs_out[(size_t)gb * DM_MAX_ACTIONS + a] = a < A ? lg[na] : 0.f;
            continue;
        }
        // gamenet.py:67-73
        float tv = tanhf(v);
        if (s.player == 1) tv = -tv;
        if (lane == 0) value_out[gb] = (tv + 1.f) / 2.f;
        // softmax over ALL actions, then mask, then renormalise (gamenet.py:74-78)
        float mx = -INFINITY;
        na = 0;
        for (int a = lane; a < A; a += 32, ++na) mx = fmaxf(mx, lg[na]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(DM_FULL_MASK, mx, off));
        float e[3], sum = 0.f;
        na = 0;
        for (int a = lane; a < A; a += 32, ++na) {
            e[na] = expf(lg[na] - mx);
            sum += e[na];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(DM_FULL_MASK, sum, off);
        bool extra;
        u64 mask = game_legal_mask(cfg, s, extra);
        float msum = 0.f;
        na = 0;
        for (int a = lane; a < A; a += 32, ++na) {
            bool legal = (a == cfg.cells) ? extra : (bool)((mask >> a) & 1);
            e[na] = legal ? e[na] / sum : 0.f;
            msum += e[na];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) msum += __shfl_xor_sync(DM_FULL_MASK, msum, off);
        na = 0;
        for (int a = lane; a < DM_MAX_ACTIONS; a += 32, ++na)
            probs_out[(size_t)gb * DM_MAX_ACTIONS + a] = a < A ? e[na] / msum : 0.f;
    }
}

}  // namespace netk
