// bf16 trunk of ConvolutionalNet on the 5th-generation tensor cores: every 3x3, 128 -> 128
// convolution (+ folded BN, + residual, + ReLU) is an implicit GEMM issued with tcgen05.mma,
// accumulating in TMEM.  (reference deep_mcts/convolutionalnet.py:91-129, 132-153; these
// seven layers are 99% of the network's FLOPs.)
//
// Formulation.  Activations live in HBM as padded flat rows (RowGeom): row = one board cell or a
// zero pad, 128 channels split into 16 chunks of 8 bf16 (16 bytes), stored chunk-major:
//     act[(kc * rows_total + row) * 8 + j]
// which is exactly the K-major, no-swizzle UMMA operand layout with 16-byte row pitch
// (SBO = 128 B per 8-row group, LBO = rows * 16 B per K chunk).  Because row pitch is uniform,
// the A operand of tap (dy, dx) is the SAME shared-memory image read from a start address
// shifted by (dy*Wp + dx) rows -- no im2col copies, one activation load per CTA.
//
// Kernels in this file, newest last (all share the PTX wrappers and descriptors below):
//   k_conv3x3_tc      first version: one CTA per 512 rows, load -> 9x16 MMAs -> epilogue in sequence.
//                     Kept selectable (DM_CONV_V1=1) as the A/B baseline of profiles/README.md.
//   k_conv3x3_tc2     persistent CTA per SM, warp-specialised (weight ring, activation producer, MMA
//                     issuer, 4 epilogue warps), double-buffered activation windows and TMEM
//                     accumulators.  Runs the policy-head conv, and the trunk of board sizes other
//                     than 7 and 8.  <kPair = true> reads/writes the pad-free pair layout of 8x8 boards.
//   k_resblock_tc<G>  whole residual blocks (conv -> ReLU -> conv -> + skip -> ReLU) with the
//                     intermediate activation in shared memory, and all blocks of the trunk in one
//                     launch; G = 8 (pair layout) and G = 7 (flat layout, 64 rows per board).  This
//                     is the dominant kernel of the engine.
//   k_policy_fc_tc    policy-head FC as a split-K tcgen05 GEMM;  k_heads_finish: value head FCs,
//                     softmax, legality mask (GameNet.evaluate's post-processing).
// Timing-experiment switches (never set in the shipped build; each removes work to bound what an
// optimisation could gain, results are wrong by construction): DM_EXPERIMENT_ALIGNED,
// DM_EXP_NOEPI, DM_EXP_NOSTORE, DM_EXP_NOWEIGHTS -- the numbers they produced are in
// profiles/README.md.
#pragma once
#include <cuda_bf16.h>

#include "net_pack.cuh"

namespace netk {

constexpr int kC = 128;               // channels (N and the K chunking are specialised for 128)
constexpr int kChunks = kC / 8;       // 16-byte channel chunks
constexpr int kTileM = 128;           // rows per tcgen05.mma (M)
constexpr int kTilesPerCta = 4;       // accumulators per CTA
constexpr int kRowsPerCta = kTileM * kTilesPerCta;
constexpr int kStages = 2;            // weight stages
constexpr int kStageBytes = kC * kC * 2;  // one tap: 32 KB
constexpr int kConvThreads = 192;

struct ConvArgs {
    const __nv_bfloat16* in;        // chunked rows
    __nv_bfloat16* out;             // chunked rows, or null
    __nv_bfloat16* out_nhwc;        // dense [board][pos][128], or null
    __nv_bfloat16* out_fc;          // policy-FC A-operand layout (see k_policy_fc_tc), or null
    const __nv_bfloat16* residual;  // chunked rows, or null (may alias out)
    const __nv_bfloat16* w;         // [9][16][128][8]
    const float* bias;              // [128]
    const float* vhead_w;           // [128] or null: also emit the value-head 1x1 conv
    float vhead_b;
    // Host copies of the two vectors above, passed BY VALUE: kernel parameters live in the constant
    // bank, so the epilogue's per-channel operands become c[0x0][imm] immediates and never touch
    // the shared-memory data pipe, which tcgen05.mma (SS mode, 128 B/clk of operands) needs whole.
    float bias_c[128];
    float vw_c[128];
    float* vplane;                  // [n][g*g]
    size_t rows_total;
    int n_boards;
    const int* count_dev;
    RowGeom geom;
};

__host__ __device__ inline int conv_a_rows(const RowGeom& geom) { return kRowsPerCta + 2 * geom.halo; }
inline size_t conv_smem_bytes(const RowGeom& geom) {
    return 1024 + (size_t)kChunks * conv_a_rows(geom) * 16 + (size_t)kStages * kStageBytes + 2 * kC * sizeof(float) + 128;
}
// rows a buffer must hold for n boards: front halo + whole CTA tiles + back halo
inline size_t conv_rows_total(const RowGeom& geom, int n_boards) {
    size_t body = ((size_t)n_boards * geom.P + kRowsPerCta - 1) / kRowsPerCta * kRowsPerCta;
    return body + 2 * (size_t)geom.halo;
}

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    u32 done = 0;
    unsigned long long spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (++spins > (1ull << 26)) __trap();  // never hang the GPU: fail the launch instead
    }
}
__device__ __forceinline__ void bulk_g2s(u32 dst, const void* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(u32 taddr, u32 (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address, leading (K-chunk) byte offset, stride (8-row group) byte offset, all >> 4; version 1.
__device__ __forceinline__ u64 umma_desc(u32 saddr, u32 lbo_bytes, u32 sbo_bytes) {
    return (u64)((saddr & 0x3FFFFu) >> 4) | ((u64)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((u64)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// The same descriptor split in two 32-bit halves so that advancing along K or M is one 32-bit add
// on the low word (start-address field, units of 16 bytes; never carries into the LBO field).
__device__ __forceinline__ u32 umma_desc_lo(u32 saddr, u32 lbo_bytes) {
    return ((saddr & 0x3FFFFu) >> 4) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ u32 umma_desc_hi(u32 sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ u64 umma_desc_join(u32 lo, u32 hi) { return (u64)lo | ((u64)hi << 32); }

// ---- thread-block clusters: weight stages multicast to every CTA of a cluster -------------------
__device__ __forceinline__ u32 cluster_ctarank() {
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {   // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// bulk copy global -> the same shared-memory offset of every CTA in `mask`; each destination CTA's mbarrier (same
// offset) receives the complete_tx for the bytes that landed in it
__device__ __forceinline__ void bulk_g2s_multicast(u32 dst, const void* src, u32 bytes, u32 bar, unsigned short mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
        "l"(src), "r"(bytes), "r"(bar), "h"(mask)
        : "memory");
}
// tcgen05.commit arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_multicast(u32 bar, unsigned short mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    u32 pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// instruction descriptor: D fp32, A/B bf16, both K-major, N = 128, M = 128
constexpr u32 kIdescBf16 = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(kC >> 3) << 17) | ((u32)(kTileM >> 4) << 24);

__device__ __forceinline__ u32 pack_bf16(float a, float b) {
    __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<u32*>(&p);
}

__global__ void __launch_bounds__(kConvThreads, 1) k_conv3x3_tc(ConvArgs args) {
    extern __shared__ unsigned char smem_raw[];
    const RowGeom geom = args.geom;
    const int a_rows = conv_a_rows(geom);
    const int live = live_count(args.count_dev, args.n_boards);
    const size_t tile_row0 = (size_t)blockIdx.x * kRowsPerCta;  // board-space row of this CTA's first output row
    if ((int)(tile_row0 / geom.P) >= live) return;              // nothing live in this CTA (block-uniform)

    // carve shared memory (1024-byte aligned base)
    unsigned char* base = (unsigned char*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    unsigned char* s_a = base;                                    // [16][a_rows][16 B]
    unsigned char* s_b = s_a + (size_t)kChunks * a_rows * 16;     // [2][32 KB]
    float* s_bias = (float*)(s_b + (size_t)kStages * kStageBytes);
    float* s_vw = s_bias + kC;
    u64* s_bar = (u64*)(s_vw + kC);  // full[2], empty[2], a_full, acc_full
    u32* s_tmem = (u32*)(s_bar + 8);
    const u32 bar_full0 = smem_u32(&s_bar[0]), bar_empty0 = smem_u32(&s_bar[2]);
    const u32 bar_a = smem_u32(&s_bar[4]), bar_acc = smem_u32(&s_bar[5]);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kC; i += blockDim.x) {
        s_bias[i] = args.bias[i];
        s_vw[i] = args.vhead_w ? args.vhead_w[i] : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full0 + 8 * s, 1);
            mbar_init(bar_empty0 + 8 * s, 1);
        }
        mbar_init(bar_a, 1);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = *s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            // ---- producer: activations once, then the nine weight taps through two stages
            const u32 a_chunk_bytes = (u32)a_rows * 16;
            mbar_expect_tx(bar_a, a_chunk_bytes * kChunks);
            for (int kc = 0; kc < kChunks; ++kc)
                bulk_g2s(smem_u32(s_a + (size_t)kc * a_chunk_bytes),
                         args.in + ((size_t)kc * args.rows_total + tile_row0) * 8, a_chunk_bytes, bar_a);
            for (int tap = 0; tap < 9; ++tap) {
                int st = tap % kStages;
                if (tap >= kStages) mbar_wait(bar_empty0 + 8 * st, ((tap / kStages) - 1) & 1);
                mbar_expect_tx(bar_full0 + 8 * st, kStageBytes);
                bulk_g2s(smem_u32(s_b + (size_t)st * kStageBytes), args.w + (size_t)tap * kC * kC, kStageBytes,
                         bar_full0 + 8 * st);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---- MMA issuer
            const u32 lbo_a = (u32)a_rows * 16, lbo_b = kC * 16, sbo = 128;
            mbar_wait(bar_a, 0);
            for (int tap = 0; tap < 9; ++tap) {
                int st = tap % kStages;
                mbar_wait(bar_full0 + 8 * st, (tap / kStages) & 1);
                tc_fence_after();
                const int shift = (tap / 3 - 1) * geom.Wp + (tap % 3 - 1);  // rows
                const u32 b_addr = smem_u32(s_b + (size_t)st * kStageBytes);
#pragma unroll
                for (int tile = 0; tile < kTilesPerCta; ++tile) {
                    const u32 a_addr = smem_u32(s_a) + (u32)(geom.halo + tile * kTileM + shift) * 16;
#pragma unroll
                    for (int k16 = 0; k16 < kC / 16; ++k16) {
                        u64 ad = umma_desc(a_addr + (u32)(2 * k16) * lbo_a, lbo_a, sbo);
                        u64 bd = umma_desc(b_addr + (u32)(2 * k16) * lbo_b, lbo_b, sbo);
                        tc_mma_bf16(tmem_base + (u32)(tile * kC), ad, bd, kIdescBf16, (tap | k16) ? 1u : 0u);
                    }
                }
                tc_commit(bar_empty0 + 8 * st);  // frees this weight stage when the MMAs retire
            }
            tc_commit(bar_acc);
        }
    } else {
        // ---- epilogue: warps 2..5 own TMEM lane quarters (warp % 4)
        const int quarter = warp & 3;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int g = geom.g, P0 = g * g;
        for (int tile = 0; tile < kTilesPerCta; ++tile) {
            const size_t q = tile_row0 + (size_t)tile * kTileM + quarter * 32 + lane;  // board-space row
            const size_t m = q + geom.halo;                                              // buffer row
            const int b = (int)(q / geom.P), rem = (int)(q % geom.P);
            const int y = rem / geom.Wp, x = rem % geom.Wp;
            const bool valid = (b < live) && (y < g) && (x < g);
            float vacc = 0.f;
#pragma unroll 1
            for (int cc = 0; cc < kC / 32; ++cc) {
                u32 v[32];
                tc_ld32(tmem_base + ((u32)(quarter * 32) << 16) + (u32)(tile * kC + cc * 32), v);
                tc_wait_ld();
                if (!valid) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + s_bias[cc * 32 + j];
                if (args.residual) {
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        uint4 r = *reinterpret_cast<const uint4*>(
                            args.residual + ((size_t)(cc * 4 + c4) * args.rows_total + m) * 8);
                        const u32 rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            f[c4 * 8 + 2 * j] += __uint_as_float(rw[j] << 16);
                            f[c4 * 8 + 2 * j + 1] += __uint_as_float(rw[j] & 0xffff0000u);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    f[j] = fmaxf(f[j], 0.f);
                    vacc = fmaf(f[j], s_vw[cc * 32 + j], vacc);
                }
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    uint4 o;
                    o.x = pack_bf16(f[c4 * 8 + 0], f[c4 * 8 + 1]);
                    o.y = pack_bf16(f[c4 * 8 + 2], f[c4 * 8 + 3]);
                    o.z = pack_bf16(f[c4 * 8 + 4], f[c4 * 8 + 5]);
                    o.w = pack_bf16(f[c4 * 8 + 6], f[c4 * 8 + 7]);
                    if (args.out)
                        *reinterpret_cast<uint4*>(args.out + ((size_t)(cc * 4 + c4) * args.rows_total + m) * 8) = o;
                    if (args.out_nhwc)
                        *reinterpret_cast<uint4*>(args.out_nhwc + ((size_t)b * P0 + y * g + x) * kC + cc * 32 + c4 * 8) = o;
                    if (args.out_fc)
                        *reinterpret_cast<uint4*>(args.out_fc +
                                                  ((((size_t)(b >> 7) * P0 + (y * g + x)) * kChunks + (cc * 4 + c4)) * 128 +
                                                   (b & 127)) * 8) = o;
                }
            }
            if (valid && args.vhead_w) args.vplane[(size_t)b * P0 + y * g + x] = fmaxf(vacc + args.vhead_b, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}



// ============================================================================================
// Persistent variant of the trunk convolution (the one the engine runs): one CTA per SM loops
// over work items of 256 rows (two 128-row tiles).  Activation tiles and TMEM accumulators are
// double buffered, so while the epilogue warps drain item j the tensor core already works on
// item j+1 and the producers fetch item j+2:
//   warp 0  weight producer  (9 taps x 32 KB per item through kStages2 stages)
//   warp 1  TMEM allocator + MMA issuer (2 tiles x 9 taps x 8 k-steps = 144 tcgen05.mma per item)
//   warp 2  activation producer (16 chunk copies per item into A buffer j & 1)
//   warps 3-6 epilogue (TMEM lane quarter = warp % 4); residual rows are prefetched into
//           registers before the accumulator is awaited
// TMEM: accumulator (buf, tile) at column (buf*2 + tile) * 128.
constexpr int kItemTiles = 2;
constexpr int kItemRows = kTileM * kItemTiles;  // 256
// weight ring: kStages2 stages of kStageK16 K-steps each (one K-step = 16 input channels of one tap
// = 4 KB; 8 K-steps = one whole tap = 32 KB)
#ifndef DM_STAGE_K16
#define DM_STAGE_K16 4
#endif
#ifndef DM_STAGES
#define DM_STAGES 3
#endif
constexpr int kStageK16 = DM_STAGE_K16;
constexpr int kStage2Bytes = kStageK16 * 2 * kC * 16;
constexpr int kStages2 = DM_STAGES;
constexpr int kStagesPerTap = (kC / 16) / kStageK16;
constexpr int kConv2Threads = 224;

__host__ __device__ inline int conv2_a_rows(const RowGeom& geom) { return kItemRows + 2 * geom.halo; }
inline size_t conv2_smem_bytes(const RowGeom& geom) {
    return 1024 + (size_t)2 * kChunks * conv2_a_rows(geom) * 16 + (size_t)kStages2 * kStage2Bytes + 2 * kC * sizeof(float) +
           512;
}
inline size_t conv2_rows_total(const RowGeom& geom, int n_boards) {
    size_t body = ((size_t)n_boards * geom.P + kItemRows - 1) / kItemRows * kItemRows;
    return body + 2 * (size_t)geom.halo;
}

// ---- pad-free layout for 8x8 boards ("pair layout") -------------------------------------------
// Two boards are interleaved row by row, 9 entries per board row (8 cells + one zero entry):
//     entry(pair p, line l, board b, x) = kPairFront + ((p*9 + l)*2 + b)*9 + x
// line 0 of every pair is a zero line (it is also the bottom pad of the previous pair), lines 1..8
// hold board rows 0..7.  A tap (dy, dx) is a shift of dy*18 + dx entries.  The sixteen 8-row groups
// (line 1..8) x (board 0, 1) of a pair start 9 entries apart, so ONE tcgen05.mma with stride-byte-offset
// 144 B covers exactly the 128 real cells of two boards: the zero entries exist in memory but are
// not rows of the GEMM, which removes the (g+1)^2/g^2 = 1.27x padded rows of the flat layout.
constexpr int kPairLine = 18;                       // entries per line (2 boards x 9)
constexpr int kPairEntries = 9 * kPairLine;         // 162 entries per pair
constexpr int kPairFront = kPairLine;               // zero entries in front of pair 0
constexpr int kPairItemPairs = 2;                   // one work item = 2 pairs = 4 boards = 2 MMA tiles
constexpr int kPairARows = kPairItemPairs * kPairEntries + kPairLine + 1;  // 343: + closing pad line + 1 entry in front
__host__ __device__ inline size_t pair_entry(int board, int y, int x) {
    return (size_t)kPairFront + ((size_t)((board >> 1) * 9 + 1 + y) * 2 + (board & 1)) * 9 + x;
}
inline size_t pair_rows_total(int n_boards) {
    size_t items = ((size_t)n_boards + 3) / 4;
    return (size_t)kPairFront + items * kPairItemPairs * kPairEntries + kPairLine + kPairARows;
}
inline size_t conv2_smem_bytes_pair() {
    return 1024 + (size_t)2 * kChunks * kPairARows * 16 + (size_t)kStages2 * kStage2Bytes + 512;
}

__device__ __forceinline__ void mbar_arrive(u32 bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ============================================================================================
// First layer of the bf16 path: bitboards -> 3 input planes -> 3x3 conv (3 -> 128) + folded BN + ReLU
// -> bf16 rows in the trunk's layout (reference <game>/convolutionalnet.py states_to_tensor +
// convolutionalnet.py:76-89).  Because every input cell is own / opponent / empty, the contribution
// of one board line to an output is a function of just 3 cells = 27 patterns: the host folds the
// weights into  tab[dy][pattern][C]  (pattern = c(x-1) + 3 c(x) + 9 c(x+1)) and
// base[player][border class][C]  (bias + the in-board taps of the constant player plane), so an
// output chunk is four table reads: base + tab[-1] + tab[0] + tab[+1].
// The kernel is bound by shared-memory bandwidth, so the tables are staged as fp16 (one 16-byte read
// = one whole 8-channel chunk; entries are sums of at most 9 folded weights, the fp16 rounding of
// 2^-11 relative per entry stays below the bf16 rounding of the output; sums accumulate in fp32).
// Persistent CTAs: the 30 KB table image (built and padded on the host, rows +16 B so that lanes
// reading DIFFERENT rows at the same column hit different banks) arrives with one bulk copy, then the
// CTA walks groups of kEncode3Boards boards, prefetching the next group's bitboards.
#ifndef DM_ENC_BOARDS
#define DM_ENC_BOARDS 2
#endif
#ifndef DM_ENC_CTAS
#define DM_ENC_CTAS 2
#endif
constexpr int kEncode3Boards = DM_ENC_BOARDS;
constexpr int kEncode3CtasPerSm = DM_ENC_CTAS;
constexpr int kEncode3Rows = 3 * 27 + 2 * 16;                 // table rows: tab then base
__host__ __device__ inline size_t encode3_table_bytes(int C) { return (size_t)kEncode3Rows * (C + 8) * sizeof(__half); }
inline size_t encode3_smem_bytes(int C, int g) {
    return encode3_table_bytes(C) + (size_t)kEncode3Boards * (g + 2) * (g + 2) + 64;
}
__global__ void __launch_bounds__(256) k_encode_conv1_table(GameCfg cfg, int n, const int* count_dev, const int* index,
                                                            const u64* first, const u64* second, const u8* player,
                                                            const __half* __restrict__ tab16,   // [113][C + 8]
                                                            int C, __nv_bfloat16* out_bf16, size_t rows_total,
                                                            int pair_layout) {
    extern __shared__ __align__(16) unsigned char smem_enc[];
    __shared__ __align__(8) u64 s_bar;
    __shared__ int s_player[kEncode3Boards];
    const int g = cfg.g, gp = g + 2, P0 = g * g;
    const int CP = C + 8;                // padded row pitch (halves)
    __half* s_tab = reinterpret_cast<__half*>(smem_enc);   // [3*27][CP]
    __half* s_base = s_tab + 81 * CP;                      // [2*16][CP]
    u8* s_code = reinterpret_cast<u8*>(s_base + 32 * CP);
    const int live = live_count(count_dev, n);
    const int n_groups = (live + kEncode3Boards - 1) / kEncode3Boards;
    if ((int)blockIdx.x >= n_groups) return;
    const u32 bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const u32 bytes = (u32)encode3_table_bytes(C);
        mbar_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(s_tab), tab16, bytes, bar);
    }
    // this thread's slot of the padded code grid never changes: (board in group, padded cell) -> bit index
    const bool fills = (int)threadIdx.x < kEncode3Boards * gp * gp;
    const int my_bb = fills ? (int)threadIdx.x / (gp * gp) : 0;
    int my_bit = -1, my_bit_t = -1;
    if (fills) {
        const int r = (int)threadIdx.x % (gp * gp);
        const int cy = r / gp - 1, cx = r % gp - 1;
        if (cy >= 0 && cy < g && cx >= 0 && cx < g) {
            my_bit = cy * g + cx;
            my_bit_t = cx * g + cy;      // Hex nets see the transposed board when SECOND moves
        }
    }
    const RowGeom geom = RowGeom::make(g);
    const int pos = threadIdx.x & 63, q = threadIdx.x >> 6;
    const int y = pos / g, x = pos % g;
    const int cls = (y > 0) + 2 * (y < g - 1) + 4 * ((x > 0) + 2 * (x < g - 1));
    const bool hex_like = (cfg.game == DM_HEX || cfg.game == DM_HEX_SWAP);
    // bitboards of this thread's board in the first group
    u64 nf = 0, ns = 0;
    int npl = 0;
    {
        const int b = (int)blockIdx.x * kEncode3Boards + my_bb;
        if (fills && b < live) {
            const int src = index ? index[b] : b;   // board b of the batch is state index[b] (compacted dense batch)
            nf = first[src];
            ns = second[src];
            npl = player[src];
        }
    }
    bool tables_ready = false;
    for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const int b0 = grp * kEncode3Boards;
        const u64 f = nf, s2 = ns;
        const int pl = npl;
        {   // prefetch the next group's bitboards
            const int b = (grp + (int)gridDim.x) * kEncode3Boards + my_bb;
            if (fills && b < live) {
                const int src = index ? index[b] : b;
                nf = first[src];
                ns = second[src];
                npl = player[src];
            }
        }
        __syncthreads();   // the previous group's codes are no longer read
        if (fills) {
            u8 code = 0;
            if (my_bit >= 0 && b0 + my_bb < live) {
                const int bit = (hex_like && pl == 0) ? my_bit_t : my_bit;
                const u64 own = pl ? f : s2, opp = pl ? s2 : f;
                code = ((own >> bit) & 1) ? 1 : (((opp >> bit) & 1) ? 2 : 0);
            }
            s_code[threadIdx.x] = code;
            if ((int)threadIdx.x % (gp * gp) == 0) s_player[my_bb] = pl;
        }
        __syncthreads();
        if (!tables_ready) {
            mbar_wait(bar, 0);
            tables_ready = true;
        }
        if (pos >= P0) continue;
#pragma unroll
        for (int bb = 0; bb < kEncode3Boards; ++bb) {
            const int b = b0 + bb;
            if (b >= live) break;
            const u8* code = s_code + bb * gp * gp + y * gp + x;   // top-left of the 3x3 window (padded coords)
            const __half* t0 = s_tab + (0 * 27 + code[0] + 3 * code[1] + 9 * code[2]) * CP;
            const __half* t1 = s_tab + (1 * 27 + code[gp] + 3 * code[gp + 1] + 9 * code[gp + 2]) * CP;
            const __half* t2 = s_tab + (2 * 27 + code[2 * gp] + 3 * code[2 * gp + 1] + 9 * code[2 * gp + 2]) * CP;
            const __half* bs = s_base + ((s_player[bb] ? 16 : 0) + cls) * CP;
            // pair layout (8x8 boards): rows of two boards interleaved, 9 entries per board row
            const size_t row = pair_layout ? pair_entry(b, y, x)
                                           : (size_t)geom.halo + (size_t)b * geom.P + y * geom.Wp + x;
            // C == kC on this path: the four chunks of this thread are unrolled so that all 16 table reads are in flight
#pragma unroll
            for (int k4 = 0; k4 < kChunks / 4; ++k4) {
                const int kc = q + 4 * k4;
                const uint4 r0 = *reinterpret_cast<const uint4*>(bs + kc * 8);
                const uint4 r1 = *reinterpret_cast<const uint4*>(t0 + kc * 8);
                const uint4 r2 = *reinterpret_cast<const uint4*>(t1 + kc * 8);
                const uint4 r3 = *reinterpret_cast<const uint4*>(t2 + kc * 8);
                const u32 w0[4] = {r0.x, r0.y, r0.z, r0.w}, w1[4] = {r1.x, r1.y, r1.z, r1.w};
                const u32 w2[4] = {r2.x, r2.y, r2.z, r2.w}, w3[4] = {r3.x, r3.y, r3.z, r3.w};
                float a[8];
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    const float2 v0 = __half22float2(*reinterpret_cast<const __half2*>(&w0[h]));
                    const float2 v1 = __half22float2(*reinterpret_cast<const __half2*>(&w1[h]));
                    const float2 v2 = __half22float2(*reinterpret_cast<const __half2*>(&w2[h]));
                    const float2 v3 = __half22float2(*reinterpret_cast<const __half2*>(&w3[h]));
                    a[2 * h + 0] = fmaxf((v0.x + v1.x) + (v2.x + v3.x), 0.f);
                    a[2 * h + 1] = fmaxf((v0.y + v1.y) + (v2.y + v3.y), 0.f);
                }
                uint4 v;
                v.x = pack_bf16(a[0], a[1]);
                v.y = pack_bf16(a[2], a[3]);
                v.z = pack_bf16(a[4], a[5]);
                v.w = pack_bf16(a[6], a[7]);
                *reinterpret_cast<uint4*>(out_bf16 + ((size_t)kc * rows_total + row) * 8) = v;
            }
        }
    }
    if (!tables_ready) mbar_wait(bar, 0);   // never leave with the bulk copy in flight
}

// ============================================================================================
// The same first layer on the tensor cores (the one the engine runs; the table kernel above stays selectable with
// DM_STEM_TABLE=1 as its A/B baseline).  A cell's 3x3 window over the three input planes is 27 numbers that are 0 or
// 1, so the layer is a [cells x 27] x [27 x 128] GEMM whose A rows the kernel writes itself, straight from the
// bitboards, into the K-major / no-swizzle operand layout (no gather, no tables): one thread per cell builds its
// row (27 bf16 ones/zeros padded to 32 = four 16-byte chunks), one tcgen05.mma group multiplies 128 cells (two
// boards) by the folded weights, and the same 128 threads drain the accumulator (+bias, ReLU, bf16) into the
// trunk's activation layout.  The weights are split hi + lo (w = bf16(w) + bf16(w - bf16(w))): the B operand has
// K = 64 with both halves, the A descriptor of the lo half simply points at the same rows, so the products are exact
// to 2^-17 relative -- tighter than the fp16 tables of the kernel above -- at no extra operand traffic.
// Per board: ~0.3 k warp instructions for the rows + ~0.7 k for the epilogue (table kernel: 3.2 k), no
// shared-memory bandwidth to speak of; what is left is the 16 KB per board of activation writes.
constexpr int kStemThreads = 128;                 // one thread per GEMM row = TMEM lane
constexpr int kStemCtasPerSm = 4;                 // 4 x 128 TMEM columns
constexpr int kStemABytes = 4 * kTileM * 16;      // 4 K-chunks x 128 rows x 16 B
constexpr int kStemBBytes = 8 * kC * 16;          // 8 K-chunks (hi 0-3, lo 4-7) x 128 couts x 16 B
inline size_t stem_smem_bytes() { return 1024 + kStemABytes + kStemBBytes + 64; }

struct StemArgs {
    GameCfg cfg;
    int n;
    const int* count_dev;
    const int* index;
    const u64* first;
    const u64* second;
    const u8* player;
    const __nv_bfloat16* w;   // [8][128][8]: k = tap * 3 + plane (27 of 32), hi then lo
    __nv_bfloat16* out;       // trunk activation rows
    size_t rows_total;
    int pair_layout;
    float bias_c[128];        // folded bias, by value (constant bank)
};

__global__ void __launch_bounds__(kStemThreads, kStemCtasPerSm) k_stem_tc(StemArgs args) {
    extern __shared__ unsigned char smem_raw[];
    const GameCfg cfg = args.cfg;
    const int g = cfg.g, P0 = g * g;
    const int live = live_count(args.count_dev, args.n);
    const int n_items = (live + 1) / 2;           // one item = one MMA tile = two boards
    if ((int)blockIdx.x >= n_items) return;
    unsigned char* base = (unsigned char*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    unsigned char* s_a = base;
    unsigned char* s_b = s_a + kStemABytes;
    u64* s_bar = (u64*)(s_b + kStemBBytes);       // weights arrived, accumulator ready
    u32* s_tmem = (u32*)(s_bar + 2);
    const u32 bar_w = smem_u32(&s_bar[0]), bar_acc = smem_u32(&s_bar[1]);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int r = threadIdx.x;                    // GEMM row of this thread
    if (threadIdx.x == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_w, kStemBBytes);
        bulk_g2s(smem_u32(s_b), args.w, kStemBBytes, bar_w);
    }
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = *s_tmem;

    // this thread's cell: board (r >> 6) of the item, cell r & 63 in the net's own coordinates
    const int b2 = r >> 6, c = r & 63;
    const bool real = c < P0;
    const int y = real ? c / g : 0, x = real ? c % g : 0;
    const bool hex_like = (cfg.game == DM_HEX || cfg.game == DM_HEX_SWAP);
    const RowGeom geom = RowGeom::make(g);
    bool weights_ready = false;
    int j = 0;
    // index -> state is a chain of two dependent loads: the state of item j+1 and the index of item j+2 are
    // requested while item j is worked on, so neither is ever waited for inside the loop
    const int stride = (int)gridDim.x;
    auto state_index = [&](int item) -> int {
        const int b = 2 * item + b2;
        if (!real || item >= n_items || b >= live) return -1;
        return args.index ? args.index[b] : b;   // board b of the batch is state index[b]
    };
    int src_next = state_index((int)blockIdx.x + stride);
    u64 nf = 0, ns = 0;
    int npl = 0;
    {
        const int src = state_index((int)blockIdx.x);
        if (src >= 0) {
            nf = args.first[src];
            ns = args.second[src];
            npl = args.player[src];
        }
    }
    for (int it = blockIdx.x; it < n_items; it += stride, ++j) {
        const int b = 2 * it + b2;
        const bool valid = real && b < live;
        const u64 f = nf, s2 = ns;
        const int pl = npl;
        if (src_next >= 0) {
            nf = args.first[src_next];
            ns = args.second[src_next];
            npl = args.player[src_next];
        }
        src_next = state_index(it + 2 * stride);
        u32 w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0;
        if (valid) {
            const u64 own = pl ? f : s2, opp = pl ? s2 : f;
            const bool transpose = hex_like && pl == 0;       // Hex nets see the transposed board when SECOND moves
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int ny = y + tap / 3 - 1, nx = x + tap % 3 - 1;
                if (ny < 0 || ny >= g || nx < 0 || nx >= g) continue;   // zero padding
                const int bit = transpose ? nx * g + ny : ny * g + nx;
                const u32 in[3] = {(u32)((own >> bit) & 1), (u32)((opp >> bit) & 1), (u32)pl};
#pragma unroll
                for (int plane = 0; plane < 3; ++plane) {
                    const int k = tap * 3 + plane;            // bf16 1.0 = 0x3f80 in half (k & 1) of word k / 2
                    w[k >> 1] |= in[plane] * ((k & 1) ? 0x3f800000u : 0x00003f80u);
                }
            }
        }
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
            *reinterpret_cast<uint4*>(s_a + ((size_t)kc * kTileM + r) * 16) =
                make_uint4(w[4 * kc], w[4 * kc + 1], w[4 * kc + 2], w[4 * kc + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor core
        __syncthreads();
        if (warp == 0) {
            if (!weights_ready) {
                mbar_wait(bar_w, 0);
                weights_ready = true;
            }
            tc_fence_after();
            if (elect_one()) {
                const u32 lbo = kTileM * 16, sbo = 128;
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {
                    const u64 ad = umma_desc(smem_u32(s_a) + (u32)((k16 & 1) * 2) * lbo, lbo, sbo);   // lo half re-reads the rows
                    const u64 bd = umma_desc(smem_u32(s_b) + (u32)(k16 * 2) * (u32)(kC * 16), kC * 16, sbo);
                    tc_mma_bf16(tmem_base, ad, bd, kIdescBf16, k16 ? 1u : 0u);
                }
                tc_commit(bar_acc);
            }
            __syncwarp();
        }
        mbar_wait(bar_acc, j & 1);
        tc_fence_after();
        const size_t row = args.pair_layout ? pair_entry(b, y, x)
                                            : (size_t)geom.halo + (size_t)b * geom.P + y * geom.Wp + x;
        // two register buffers: the TMEM load of columns cc+1 is in flight while columns cc are converted and stored
        u32 va[32], vb[32];
        const u32 taddr = tmem_base + ((u32)(warp * 32) << 16);
        auto emit = [&](const u32 (&v)[32], int cc) {
            if (!valid) return;
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                float o[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    o[jj] = fmaxf(__uint_as_float(v[c4 * 8 + jj]) + args.bias_c[cc * 32 + c4 * 8 + jj], 0.f);
                uint4 q;
                q.x = pack_bf16(o[0], o[1]);
                q.y = pack_bf16(o[2], o[3]);
                q.z = pack_bf16(o[4], o[5]);
                q.w = pack_bf16(o[6], o[7]);
                *reinterpret_cast<uint4*>(args.out + ((size_t)(cc * 4 + c4) * args.rows_total + row) * 8) = q;
            }
        };
        tc_ld32(taddr, va);
        tc_wait_ld();
        tc_ld32(taddr + 32, vb);
        emit(va, 0);
        tc_wait_ld();
        tc_ld32(taddr + 64, va);
        emit(vb, 1);
        tc_wait_ld();
        tc_ld32(taddr + 96, vb);
        emit(va, 2);
        tc_wait_ld();
        emit(vb, 3);
        tc_fence_before();
        __syncthreads();   // rows and accumulator are free for the next item
    }
    if (!weights_ready && warp == 0) mbar_wait(bar_w, 0);   // never leave with the bulk copy in flight
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

// CL = 2 (pair layout only): clusters of two CTAs share the weight stream by multicast, exactly as in
// k_resblock_tc below (same item count per CTA of a cluster, half a stage fetched by each, 2-arrival empty barriers).
template <bool kPair, int CL = 1>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kConv2Threads, 1) k_conv3x3_tc2(ConvArgs args) {
    static_assert(CL == 1 || kPair, "weight-sharing clusters are built for the pair layout");
    extern __shared__ unsigned char smem_raw[];
    const RowGeom geom = args.geom;
    const int a_rows = kPair ? kPairARows : conv2_a_rows(geom);
    const int live = live_count(args.count_dev, args.n_boards);
    const int n_items = kPair ? ((live + 3) / 4 + CL - 1) / CL * CL   // whole clusters: an item past the live boards stores nothing
                              : (int)(((size_t)live * geom.P + kItemRows - 1) / kItemRows);
    if ((int)blockIdx.x >= n_items) return;
    const u32 cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr unsigned short kClusterMask = (unsigned short)((1u << CL) - 1u);

    unsigned char* base = (unsigned char*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    const u32 a_buf_bytes = (u32)kChunks * a_rows * 16;
    unsigned char* s_a = base;                                  // [2][16][a_rows][16 B]
    unsigned char* s_b = s_a + 2 * (size_t)a_buf_bytes;         // [kStages2][kStage2Bytes]
    u64* s_bar = (u64*)(s_b + (size_t)kStages2 * kStage2Bytes);
    // barriers: b_full[S], b_empty[S], a_full[2], a_empty[2], acc_full[2], acc_empty[2]
    const u32 bar_bfull = smem_u32(&s_bar[0]), bar_bempty = smem_u32(&s_bar[kStages2]);
    const u32 bar_afull = smem_u32(&s_bar[2 * kStages2]), bar_aempty = smem_u32(&s_bar[2 * kStages2 + 2]);
    const u32 bar_accfull = smem_u32(&s_bar[2 * kStages2 + 4]), bar_accempty = smem_u32(&s_bar[2 * kStages2 + 6]);
    u32* s_tmem = (u32*)(s_bar + 2 * kStages2 + 8);

    // warp-uniform role index (the shuffle lets the compiler prove uniformity, so the issue loops
    // below run on the uniform datapath instead of R2UR waterfalls around every UTCHMMA)
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, CL);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_afull + 8 * b, 1);
            mbar_init(bar_aempty + 8 * b, 1);
            mbar_init(bar_accfull + 8 * b, 1);
            mbar_init(bar_accempty + 8 * b, 4);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    tc_fence_after();
    const u32 tmem_base = *s_tmem;

    if (warp == 0) {
        // ---- weight producer (whole warp runs the loop; one elected lane issues)
        int s = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            for (int part = 0; part < 9 * kStagesPerTap; ++part, ++s) {   // the layer's weights are one contiguous image
                const int st = s % kStages2;
                if (s >= kStages2) mbar_wait(bar_bempty + 8 * st, ((s / kStages2) - 1) & 1);
                if (elect_one()) {
                    mbar_expect_tx(bar_bfull + 8 * st, kStage2Bytes);
                    if (CL > 1) {
                        constexpr u32 kShare = kStage2Bytes / CL;
                        bulk_g2s_multicast(smem_u32(s_b + (size_t)st * kStage2Bytes) + cta_rank * kShare,
                                           reinterpret_cast<const unsigned char*>(args.w) + (size_t)part * kStage2Bytes +
                                               cta_rank * kShare,
                                           kShare, bar_bfull + 8 * st, kClusterMask);
                    } else {
                        bulk_g2s(smem_u32(s_b + (size_t)st * kStage2Bytes),
                                 reinterpret_cast<const unsigned char*>(args.w) + (size_t)part * kStage2Bytes, kStage2Bytes,
                                 bar_bfull + 8 * st);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 2) {
        // ---- activation producer
        const u32 a_chunk_bytes = (u32)a_rows * 16;
        int j = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            const int buf = j & 1;
            if (j >= 2) mbar_wait(bar_aempty + 8 * buf, ((j >> 1) - 1) & 1);
            // first entry of the item's activation window (pair layout: one entry before the pad line of its first pair)
            const size_t row0 = kPair ? (size_t)kPairFront + (size_t)it * kPairItemPairs * kPairEntries - 1
                                      : (size_t)it * kItemRows;
            if (elect_one()) {
                mbar_expect_tx(bar_afull + 8 * buf, a_chunk_bytes * kChunks);
                for (int kc = 0; kc < kChunks; ++kc)
                    bulk_g2s(smem_u32(s_a + (size_t)buf * a_buf_bytes + (size_t)kc * a_chunk_bytes),
                             args.in + ((size_t)kc * args.rows_total + row0) * 8, a_chunk_bytes, bar_afull + 8 * buf);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer: the whole warp walks the loop, one elected lane issues each instruction
        const u32 lbo_a = (u32)a_rows * 16, lbo_b = kC * 16;
        const u32 desc_hi_a = umma_desc_hi(kPair ? 144 : 128);   // pair layout: 8-row groups are 9 entries apart
        const u32 desc_hi_b = umma_desc_hi(128);
        const u32 a_k16_step = (2 * lbo_a) >> 4, b_k16_step = (2 * lbo_b) >> 4;  // descriptor units (16 B)
        const u32 a_first = kPair ? (u32)(1 + kPairLine) : (u32)geom.halo;       // first output row within the window
        const u32 a_tile_step = kPair ? (u32)kPairEntries : (u32)kTileM;
        const int line = kPair ? kPairLine : geom.Wp;
        int j = 0, s = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            const int buf = j & 1;
            mbar_wait(bar_afull + 8 * buf, (j >> 1) & 1);
            if (j >= 2) mbar_wait(bar_accempty + 8 * buf, ((j >> 1) - 1) & 1);
            tc_fence_after();
            const u32 a_lo_base = umma_desc_lo(smem_u32(s_a + (size_t)buf * a_buf_bytes), lbo_a) + a_first;
            const u32 d_base = tmem_base + (u32)(buf * kItemTiles * kC);
            for (int tap = 0; tap < 9; ++tap) {
#ifdef DM_EXPERIMENT_ALIGNED
                const int shift = 8 - (int)a_first;   // timing experiment only: every tap reads a 128-byte aligned window (wrong results)
#else
                const int shift = (tap / 3 - 1) * line + (tap % 3 - 1);   // rows == descriptor units
#endif
                const u32 a_lo_tap = a_lo_base + (u32)shift;
                for (int part = 0; part < kStagesPerTap; ++part, ++s) {
                    const int st = s % kStages2;
                    mbar_wait(bar_bfull + 8 * st, (s / kStages2) & 1);
                    tc_fence_after();
                    const u32 b_lo = umma_desc_lo(smem_u32(s_b + (size_t)st * kStage2Bytes), lbo_b);
                    if (elect_one()) {
#pragma unroll
                        for (int tile = 0; tile < kItemTiles; ++tile) {
#pragma unroll
                            for (int kk = 0; kk < kStageK16; ++kk) {
                                const int k16 = part * kStageK16 + kk;
                                tc_mma_bf16(d_base + (u32)(tile * kC),
                                            umma_desc_join(a_lo_tap + (u32)tile * a_tile_step + (u32)k16 * a_k16_step, desc_hi_a),
                                            umma_desc_join(b_lo + (u32)kk * b_k16_step, desc_hi_b), kIdescBf16,
                                            (tap | k16) ? 1u : 0u);
                            }
                        }
                        if (CL > 1) tc_commit_multicast(bar_bempty + 8 * st, kClusterMask);
                        else tc_commit(bar_bempty + 8 * st);
                        if (tap == 8 && part == kStagesPerTap - 1) {
                            tc_commit(bar_accfull + 8 * buf);  // accumulators of this item complete
                            tc_commit(bar_aempty + 8 * buf);   // and its activation buffer may be refilled
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- epilogue warps 3..6
        const int quarter = warp & 3;
        const int g = geom.g, P0 = g * g;
        int j = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            const int buf = j & 1;
            bool waited = false;
            for (int tile = 0; tile < kItemTiles; ++tile) {
                size_t m;
                int b, y, x;
                bool valid;
                if (kPair) {
                    const int r = quarter * 32 + lane;          // accumulator lane = (line, board in pair, x)
                    const int grp = r >> 3;
                    x = r & 7;
                    y = grp >> 1;
                    b = 2 * (it * kPairItemPairs + tile) + (grp & 1);
                    m = pair_entry(b, y, x);
                    valid = b < live;
                } else {
                    const size_t q = (size_t)it * kItemRows + (size_t)tile * kTileM + quarter * 32 + lane;
                    m = q + geom.halo;
                    b = (int)(q / geom.P);
                    const int rem = (int)(q % geom.P);
                    y = rem / geom.Wp;
                    x = rem % geom.Wp;
                    valid = (b < live) && (y < g) && (x < g);
                }
#ifdef DM_EXP_NOEPI
                valid = false;   // timing experiment only
#endif
                // prefetch this row's residual (all 128 channels) before waiting for the tensor core
                uint4 res[kChunks];
                if (args.residual && valid) {
#pragma unroll
                    for (int kc = 0; kc < kChunks; ++kc)
                        res[kc] = *reinterpret_cast<const uint4*>(args.residual + ((size_t)kc * args.rows_total + m) * 8);
                }
                if (!waited) {
                    mbar_wait(bar_accfull + 8 * buf, (j >> 1) & 1);
                    tc_fence_after();
                    waited = true;
                }
                float vacc = 0.f;
#pragma unroll
                for (int cc = 0; cc < kC / 32; ++cc) {
                    u32 v[32];
                    tc_ld32(tmem_base + ((u32)(quarter * 32) << 16) + (u32)((buf * kItemTiles + tile) * kC + cc * 32), v);
                    tc_wait_ld();
                    if (valid) {
                        float f[32];
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) f[jj] = __uint_as_float(v[jj]) + args.bias_c[cc * 32 + jj];
                        if (args.residual) {
#pragma unroll
                            for (int c4 = 0; c4 < 4; ++c4) {
                                const uint4 r = res[cc * 4 + c4];
                                const u32 rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj) {
                                    f[c4 * 8 + 2 * jj] += __uint_as_float(rw[jj] << 16);
                                    f[c4 * 8 + 2 * jj + 1] += __uint_as_float(rw[jj] & 0xffff0000u);
                                }
                            }
                        }
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) f[jj] = fmaxf(f[jj], 0.f);
                        if (args.vhead_w) {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) vacc = fmaf(f[jj], args.vw_c[cc * 32 + jj], vacc);
                        }
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            uint4 o;
                            o.x = pack_bf16(f[c4 * 8 + 0], f[c4 * 8 + 1]);
                            o.y = pack_bf16(f[c4 * 8 + 2], f[c4 * 8 + 3]);
                            o.z = pack_bf16(f[c4 * 8 + 4], f[c4 * 8 + 5]);
                            o.w = pack_bf16(f[c4 * 8 + 6], f[c4 * 8 + 7]);
#ifdef DM_EXP_NOSTORE
                            if (o.x == 0x12345678u && o.y == 0x9abcdef0u)   // timing experiment: keep the math, drop the stores
#endif
                            {
                            if (args.out)
                                *reinterpret_cast<uint4*>(args.out + ((size_t)(cc * 4 + c4) * args.rows_total + m) * 8) = o;
                            if (args.out_nhwc)
                                *reinterpret_cast<uint4*>(args.out_nhwc + ((size_t)b * P0 + y * g + x) * kC + cc * 32 +
                                                          c4 * 8) = o;
                            if (args.out_fc)
                                *reinterpret_cast<uint4*>(
                                    args.out_fc +
                                    ((((size_t)(b >> 7) * P0 + (y * g + x)) * kChunks + (cc * 4 + c4)) * 128 + (b & 127)) * 8) = o;
                            }
                        }
                    }
                }
                if (valid && args.vhead_w) args.vplane[(size_t)b * P0 + y * g + x] = fmaxf(vacc + args.vhead_b, 0.f);
            }
            // this warp is done reading accumulator `buf`
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty + 8 * buf);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ============================================================================================
// Fused residual block for boards whose work items are whole boards: 8x8 in the pair layout and
// 7x7 in the flat layout (P = 64 rows per board, so a 256-row item is exactly 4 boards).  One
// persistent CTA per SM, one work item = 4 boards.  Because items are whole boards, the block's
// intermediate activation T never leaves the SM:
//     X (global) --bulk copy--> bufX --conv1 (tcgen05)--> acc1 --epilogue 1: +b1, ReLU, bf16--> bufT (smem)
//     bufT --conv2 (tcgen05)--> acc2 --epilogue 2: +b2, + X (global, L2 hit), ReLU, bf16--> X' (global)
// (reference ResidualBlock, convolutionalnet.py:109-129).  Compared with two launches of
// k_conv3x3_tc2 this removes the write and re-read of T, i.e. 2 of the 5 activation passes of a block,
// which matters because the trunk runs at the board's power cap.  The weight ring streams both
// layers back to back; conv1 of item j+1 is issued right behind conv2 of item j, so the tensor
// pipe only idles while epilogue 1 writes T.
//   TMEM: acc1 = columns 0..255 (2 tiles), acc2 = columns 256..511.
// Geometry of the two instantiations.  G = 8: pair layout (no pad rows in the GEMM).  G = 7: flat
// rows with Wp = 8; a board's pad column and pad line are rows of the GEMM whose results are
// dropped, and every halo row a real cell reads is a pad of its own or the previous board.
template <int G>
struct ResGeom;
template <>
struct ResGeom<8> {
    static constexpr int kWinRows = kPairARows;          // rows of the activation window in shared memory
    static constexpr int kFirst = 1 + kPairLine;         // first output row inside the window
    static constexpr int kTileStep = kPairEntries;       // window rows between the item's two MMA tiles
    static constexpr int kSbo = 144;                     // bytes between 8-row groups
    static constexpr int kLine = kPairLine;              // window rows per dy
    __device__ static size_t window_row0(int it) { return (size_t)kPairFront + (size_t)it * kPairItemPairs * kPairEntries - 1; }
};
template <>
struct ResGeom<7> {
    static constexpr int kWinRows = kItemRows + 2 * 9;
    static constexpr int kFirst = 9;
    static constexpr int kTileStep = kTileM;
    static constexpr int kSbo = 128;
    static constexpr int kLine = 8;
    __device__ static size_t window_row0(int it) { return (size_t)it * kItemRows; }
};
// accumulator row (tile, r = TMEM lane) of item `it` -> board, cell, window row, global row
template <int G>
struct ResRow {
    int b, y, x;
    u32 widx;
    size_t m;
    bool cell;   // a real board cell (always true in the pair layout)
    __device__ ResRow(int it, int tile, int r) {
        if (G == 8) {
            const int grp = r >> 3;
            x = r & 7;
            y = grp >> 1;
            b = 2 * (it * kPairItemPairs + tile) + (grp & 1);
            widx = (u32)(1 + kPairLine + tile * kPairEntries + grp * 9 + x);
            m = pair_entry(b, y, x);
            cell = true;
        } else {
            const int q = tile * kTileM + r;
            x = q & 7;
            y = (q >> 3) & 7;
            b = it * 4 + (q >> 6);
            widx = (u32)(9 + q);
            m = (size_t)it * kItemRows + q + 9;
            cell = (x < 7) && (y < 7);
        }
    }
};

// Up to kMaxFusedBlocks consecutive residual blocks run in ONE launch: a CTA keeps the same items in
// every block and a 3x3 conv never crosses a board, so block b+1 of an item depends only on what this
// CTA itself wrote for that item in block b -- no grid-wide synchronisation, just an in-CTA "output
// of item j is visible" counter between the epilogue warps and the X producer.
constexpr int kMaxFusedBlocks = 3;
struct ResBlockArgs {
    __nv_bfloat16* x;            // activations, updated in place block after block (input = residual of block 0)
    const __nv_bfloat16* w1[kMaxFusedBlocks];     // [9][16][128][8] per layer
    const __nv_bfloat16* w2[kMaxFusedBlocks];
    float vhead_b;
    float* vplane;               // non-null: value-head 1x1 conv of the last block's output
    size_t rows_total;
    int n_boards;
    int n_blocks;
    const int* count_dev;
    float bias1_c[kMaxFusedBlocks][128];
    float bias2_c[kMaxFusedBlocks][128];
    float vw_c[128];
};

template <int G>
inline size_t resblock_smem_bytes() {
    return 1024 + (size_t)2 * kChunks * ResGeom<G>::kWinRows * 16 + (size_t)kStages2 * kStage2Bytes + 512;
}

// CL = 2: launched as clusters of two CTAs (one TPC) that walk the weight stream in lockstep -- both run the same
// number of items -- and share it: each CTA fetches half of every weight stage and multicasts it into both shared
// memories, so the L2 -> SM weight traffic (4.5x the activation traffic of this kernel) is halved.  A stage is
// refilled only after the MMAs of BOTH CTAs have released it (multicast tcgen05.commit on a 2-arrival barrier).
template <int G, int CL = 1>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(kConv2Threads, 1) k_resblock_tc(ResBlockArgs args) {
    using Geo = ResGeom<G>;
    extern __shared__ unsigned char smem_raw[];
    const int live = live_count(args.count_dev, args.n_boards);
    // CL > 1: rounded up to a whole number of clusters (gridDim.x is one, too), so the CTAs of a cluster run the same
    // number of items and leave together; an item past the live boards computes on stale rows and stores nothing
    const int n_items = ((live + 3) / 4 + CL - 1) / CL * CL;
    if ((int)blockIdx.x >= n_items) return;
    const u32 cta_rank = CL > 1 ? cluster_ctarank() : 0u;
    constexpr unsigned short kClusterMask = (unsigned short)((1u << CL) - 1u);

    unsigned char* base = (unsigned char*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    constexpr u32 kBufBytes = (u32)kChunks * Geo::kWinRows * 16;
    constexpr u32 kLbo = (u32)Geo::kWinRows * 16;
    unsigned char* s_x = base;
    unsigned char* s_t = s_x + kBufBytes;
    unsigned char* s_b = s_t + kBufBytes;
    u64* s_bar = (u64*)(s_b + (size_t)kStages2 * kStage2Bytes);
    // barriers: b_full[S], b_empty[S], x_full, x_free, acc1_full, t_full, acc2_full, acc2_empty
    const u32 bar_bfull = smem_u32(&s_bar[0]), bar_bempty = smem_u32(&s_bar[kStages2]);
    const u32 bar_xfull = smem_u32(&s_bar[2 * kStages2 + 0]), bar_xfree = smem_u32(&s_bar[2 * kStages2 + 1]);
    const u32 bar_acc1 = smem_u32(&s_bar[2 * kStages2 + 2]), bar_tfull = smem_u32(&s_bar[2 * kStages2 + 3]);
    const u32 bar_acc2 = smem_u32(&s_bar[2 * kStages2 + 4]), bar_acc2empty = smem_u32(&s_bar[2 * kStages2 + 5]);
    u32* s_tmem = (u32*)(s_bar + 2 * kStages2 + 8);
    volatile u32* s_done = s_tmem + 1;   // epilogue-warp arrivals: 4 per item whose block output is in global memory
    const int n_local = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // items of this CTA

    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) *s_done = 0;
    // bufT pads must read as zero for conv2: clear it once, the epilogue only ever writes real cells
    for (u32 i = threadIdx.x; i < kBufBytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(s_t)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages2; ++s) {
            mbar_init(bar_bfull + 8 * s, 1);
            mbar_init(bar_bempty + 8 * s, CL);   // released by the MMA warp of every CTA that reads the stage
        }
        mbar_init(bar_xfull, 1);
        mbar_init(bar_xfree, 1);
        mbar_init(bar_acc1, 1);
        mbar_init(bar_tfull, 4);      // one arrival per epilogue warp
        mbar_init(bar_acc2, 1);
        mbar_init(bar_acc2empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill is read by the tensor core
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast at them
    tc_fence_after();
    const u32 tmem_base = *s_tmem;

    if (warp == 0) {
        // ---- weight producer: layer 1 then layer 2 of every item, block after block
        int s = 0;
        for (int blk = 0; blk < args.n_blocks; ++blk)
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            for (int layer = 0; layer < 2; ++layer) {
                const unsigned char* w = reinterpret_cast<const unsigned char*>(layer ? args.w2[blk] : args.w1[blk]);
                for (int part = 0; part < 9 * kStagesPerTap; ++part, ++s) {
                    const int st = s % kStages2;
                    if (s >= kStages2) mbar_wait(bar_bempty + 8 * st, ((s / kStages2) - 1) & 1);
                    if (elect_one()) {
#ifdef DM_EXP_NOWEIGHTS
                        if (s >= kStages2) {   // timing experiment: signal "full" without streaming the weights again
                            mbar_arrive(bar_bfull + 8 * st);
                        } else
#endif
                        {
                        mbar_expect_tx(bar_bfull + 8 * st, kStage2Bytes);
                        if (CL > 1) {
                            constexpr u32 kShare = kStage2Bytes / CL;   // this CTA's share of the stage, sent to all
                            bulk_g2s_multicast(smem_u32(s_b + (size_t)st * kStage2Bytes) + cta_rank * kShare,
                                               w + (size_t)part * kStage2Bytes + cta_rank * kShare, kShare,
                                               bar_bfull + 8 * st, kClusterMask);
                        } else {
                            bulk_g2s(smem_u32(s_b + (size_t)st * kStage2Bytes), w + (size_t)part * kStage2Bytes, kStage2Bytes,
                                     bar_bfull + 8 * st);
                        }
                        }
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 2) {
        // ---- X producer (single buffer: refilled as soon as conv1 of the previous item has retired)
        int j = 0;
        for (int blk = 0; blk < args.n_blocks; ++blk)
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            if (j >= 1) mbar_wait(bar_xfree, (j - 1) & 1);
            if (blk > 0) {
                // this item's output of the previous block (written by this CTA's epilogue warps) must be in
                // global memory before the bulk copy reads it back
                const u32 need = 4u * (u32)(j - n_local + 1);
                u32 spins = 0;
                while (*s_done < need) {
                    __nanosleep(64);
                    if (++spins > (1u << 24)) __trap();
                }
                __threadfence();
                asm volatile("fence.proxy.async;" ::: "memory");
            }
            const size_t row0 = Geo::window_row0(it);
            if (elect_one()) {
                mbar_expect_tx(bar_xfull, kLbo * kChunks);
                for (int kc = 0; kc < kChunks; ++kc)
                    bulk_g2s(smem_u32(s_x + (size_t)kc * kLbo), args.x + ((size_t)kc * args.rows_total + row0) * 8, kLbo,
                             bar_xfull);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer
        const u32 lbo_b = kC * 16;
        const u32 desc_hi_a = umma_desc_hi(Geo::kSbo), desc_hi_b = umma_desc_hi(128);
        const u32 a_k16_step = (2 * kLbo) >> 4, b_k16_step = (2 * lbo_b) >> 4;
        const u32 x_lo = umma_desc_lo(smem_u32(s_x), kLbo) + (u32)Geo::kFirst;
        const u32 t_lo = umma_desc_lo(smem_u32(s_t), kLbo) + (u32)Geo::kFirst;
        int j = 0, s = 0;
        for (int blk = 0; blk < args.n_blocks; ++blk)
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            for (int layer = 0; layer < 2; ++layer) {
                if (layer == 0) {
                    mbar_wait(bar_xfull, j & 1);
                } else {
                    mbar_wait(bar_tfull, j & 1);                              // T of this item is in shared memory
                    if (j >= 1) mbar_wait(bar_acc2empty, (j - 1) & 1);        // epilogue 2 of the previous item drained acc2
                }
                tc_fence_after();
                const u32 a_lo_base = layer ? t_lo : x_lo;
                const u32 d_base = tmem_base + (u32)(layer * kItemTiles * kC);
                for (int tap = 0; tap < 9; ++tap) {
                    const int shift = (tap / 3 - 1) * Geo::kLine + (tap % 3 - 1);
                    const u32 a_lo_tap = a_lo_base + (u32)shift;
                    for (int part = 0; part < kStagesPerTap; ++part, ++s) {
                        const int st = s % kStages2;
                        mbar_wait(bar_bfull + 8 * st, (s / kStages2) & 1);
                        tc_fence_after();
                        const u32 b_lo = umma_desc_lo(smem_u32(s_b + (size_t)st * kStage2Bytes), lbo_b);
                        if (elect_one()) {
#pragma unroll
                            for (int tile = 0; tile < kItemTiles; ++tile) {
#pragma unroll
                                for (int kk = 0; kk < kStageK16; ++kk) {
                                    const int k16 = part * kStageK16 + kk;
                                    tc_mma_bf16(d_base + (u32)(tile * kC),
                                                umma_desc_join(a_lo_tap + (u32)tile * Geo::kTileStep + (u32)k16 * a_k16_step, desc_hi_a),
                                                umma_desc_join(b_lo + (u32)kk * b_k16_step, desc_hi_b), kIdescBf16,
                                                (tap | k16) ? 1u : 0u);
                                }
                            }
                            if (CL > 1) tc_commit_multicast(bar_bempty + 8 * st, kClusterMask);
                            else tc_commit(bar_bempty + 8 * st);
                            if (tap == 8 && part == kStagesPerTap - 1) {
                                if (layer == 0) {
                                    tc_commit(bar_acc1);    // conv1 done: epilogue 1 may start ...
                                    tc_commit(bar_xfree);   // ... and bufX may take the next item
                                } else {
                                    tc_commit(bar_acc2);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ---- epilogue warps 3..6
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;   // accumulator (TMEM) lane of this thread
        constexpr int kCells = G * G;
        int j = 0;
#pragma unroll 1
        for (int blk = 0; blk < args.n_blocks; ++blk)
        for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++j) {
            const float* bias1 = args.bias1_c[blk];
            const float* bias2 = args.bias2_c[blk];
            const bool emit_v = args.vplane != nullptr && blk == args.n_blocks - 1;
            // ---------- epilogue 1: acc1 -> +b1 -> ReLU -> bf16 -> bufT
            mbar_wait(bar_acc1, j & 1);
            tc_fence_after();
#pragma unroll 1
            for (int tile = 0; tile < kItemTiles; ++tile) {
                const ResRow<G> row(it, tile, r);
                unsigned char* trow = s_t + (size_t)row.widx * 16;
#pragma unroll
                for (int cc = 0; cc < kC / 32; ++cc) {
                    u32 v[32];
                    tc_ld32(tmem_base + ((u32)(quarter * 32) << 16) + (u32)(tile * kC + cc * 32), v);
                    tc_wait_ld();
                    if (row.cell) {              // pad rows of T stay zero
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            float f[8];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj)
                                f[jj] = fmaxf(__uint_as_float(v[c4 * 8 + jj]) + bias1[cc * 32 + c4 * 8 + jj], 0.f);
                            uint4 o;
                            o.x = pack_bf16(f[0], f[1]);
                            o.y = pack_bf16(f[2], f[3]);
                            o.z = pack_bf16(f[4], f[5]);
                            o.w = pack_bf16(f[6], f[7]);
                            *reinterpret_cast<uint4*>(trow + (size_t)(cc * 4 + c4) * kLbo) = o;
                        }
                    }
                }
            }
            // publish T to the tensor core (generic-proxy writes -> async proxy) and release acc1
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tfull);

            // ---------- epilogue 2: acc2 -> +b2 + X -> ReLU -> bf16 -> global
            bool waited = false;
#pragma unroll 1
            for (int tile = 0; tile < kItemTiles; ++tile) {
                const ResRow<G> row(it, tile, r);
                const int b = row.b;
                const size_t m = row.m;
                const bool valid = row.cell && b < live;
                uint4 res[kChunks];
                if (valid) {
#pragma unroll
                    for (int kc = 0; kc < kChunks; ++kc)
                        res[kc] = *reinterpret_cast<const uint4*>(args.x + ((size_t)kc * args.rows_total + m) * 8);
                }
                if (!waited) {
                    mbar_wait(bar_acc2, j & 1);
                    tc_fence_after();
                    waited = true;
                }
                float vacc = 0.f;
#pragma unroll
                for (int cc = 0; cc < kC / 32; ++cc) {
                    u32 v[32];
                    tc_ld32(tmem_base + ((u32)(quarter * 32) << 16) + (u32)((kItemTiles + tile) * kC + cc * 32), v);
                    tc_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int c4 = 0; c4 < 4; ++c4) {
                            const uint4 rr = res[cc * 4 + c4];
                            const u32 rw[4] = {rr.x, rr.y, rr.z, rr.w};
                            float f[8];
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                f[2 * jj] = __uint_as_float(v[c4 * 8 + 2 * jj]) + bias2[cc * 32 + c4 * 8 + 2 * jj] +
                                            __uint_as_float(rw[jj] << 16);
                                f[2 * jj + 1] = __uint_as_float(v[c4 * 8 + 2 * jj + 1]) +
                                                bias2[cc * 32 + c4 * 8 + 2 * jj + 1] +
                                                __uint_as_float(rw[jj] & 0xffff0000u);
                            }
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) f[jj] = fmaxf(f[jj], 0.f);
                            if (emit_v) {
#pragma unroll
                                for (int jj = 0; jj < 8; ++jj) vacc = fmaf(f[jj], args.vw_c[cc * 32 + c4 * 8 + jj], vacc);
                            }
                            uint4 o;
                            o.x = pack_bf16(f[0], f[1]);
                            o.y = pack_bf16(f[2], f[3]);
                            o.z = pack_bf16(f[4], f[5]);
                            o.w = pack_bf16(f[6], f[7]);
                            *reinterpret_cast<uint4*>(args.x + ((size_t)(cc * 4 + c4) * args.rows_total + m) * 8) = o;
                        }
                    }
                }
                if (valid && emit_v) args.vplane[(size_t)b * kCells + row.y * G + row.x] = fmaxf(vacc + args.vhead_b, 0.f);
            }
            tc_fence_before();
            // the item's block output is written: make it visible to the bulk copies of the next block
            __threadfence();
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_acc2empty);
                atomicAdd((u32*)s_done, 1u);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ============================================================================================
// Policy-head fully connected layer on tcgen05 (reference convolutionalnet.py:146-153):
//   logits[b][a] = sum_{pos, c} T[b][pos][c] * W[a][c*g*g + pos]
// as a split-K GEMM: M = 128 boards per tile, N = NA (actions padded to a multiple of 16),
// K = g*g*128 cut by board position: one pipeline stage = one position = 16 channel chunks.
//   A: policy-conv output written by k_conv3x3_tc in the layout
//        t_fc[((tile * g*g + pos) * 16 + kc) * 128 + board_in_tile][8]        (32 KB per stage)
//   B: fc_w[(pos * 16 + kc) * NA + a][8]                                      (NA * 256 B per stage)
//   D: 128 lanes x NA fp32 columns in TMEM -> partial[split][board][NA]
// Partials are summed in a fixed order by k_heads_finish, so results are deterministic.
constexpr int kFcThreads = 192;
constexpr int kFcStages = 4;
constexpr int kFcStageA = 128 * kC * 2;  // 32 KB

inline int fc_padded_actions(int A) { return (A + 15) / 16 * 16; }
inline int fc_split(int g) {
    int P0 = g * g;
    for (int s = 8; s >= 1; --s)
        if (P0 % s == 0) return s;
    return 1;
}
inline size_t fc_smem_bytes(int NA) { return 1024 + (size_t)kFcStages * (kFcStageA + (size_t)NA * 256) + 256; }

struct FcArgs {
    const __nv_bfloat16* t_fc;
    const __nv_bfloat16* w;  // [g*g*16][NA][8]
    float* partial;          // [split][n_boards_cap][NA]
    int P0, NA, split, n_boards, n_boards_cap;
    const int* count_dev;
};

__global__ void __launch_bounds__(kFcThreads, 1) k_policy_fc_tc(FcArgs args) {
    extern __shared__ unsigned char smem_raw[];
    const int tile = blockIdx.x, split = blockIdx.y;
    const int live = live_count(args.count_dev, args.n_boards);
    if (tile * 128 >= live) return;
    const int NA = args.NA;
    const u32 stage_b_bytes = (u32)NA * 256;
    unsigned char* base = (unsigned char*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    unsigned char* s_a = base;                                  // [stages][32 KB]
    unsigned char* s_b = s_a + (size_t)kFcStages * kFcStageA;   // [stages][NA*256]
    u64* s_bar = (u64*)(s_b + (size_t)kFcStages * stage_b_bytes);  // full[4], empty[4], acc
    u32* s_tmem = (u32*)(s_bar + 2 * kFcStages + 1);
    const u32 bar_full0 = smem_u32(&s_bar[0]), bar_empty0 = smem_u32(&s_bar[kFcStages]);
    const u32 bar_acc = smem_u32(&s_bar[2 * kFcStages]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFcStages; ++s) {
            mbar_init(bar_full0 + 8 * s, 1);
            mbar_init(bar_empty0 + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)),
                     "r"(128)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem_base = *s_tmem;
    const int pos_per_split = args.P0 / args.split;
    const int pos0 = split * pos_per_split;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < pos_per_split; ++i) {
                int st = i % kFcStages;
                if (i >= kFcStages) mbar_wait(bar_empty0 + 8 * st, ((i / kFcStages) - 1) & 1);
                mbar_expect_tx(bar_full0 + 8 * st, kFcStageA + stage_b_bytes);
                int pos = pos0 + i;
                bulk_g2s(smem_u32(s_a + (size_t)st * kFcStageA),
                         args.t_fc + ((size_t)tile * args.P0 + pos) * (size_t)(16 * 128 * 8), kFcStageA,
                         bar_full0 + 8 * st);
                bulk_g2s(smem_u32(s_b + (size_t)st * stage_b_bytes), args.w + (size_t)pos * 16 * NA * 8, stage_b_bytes,
                         bar_full0 + 8 * st);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const u32 idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((u32)(NA >> 3) << 17) | ((u32)(128 >> 4) << 24);
            const u32 lbo_a = 128 * 16, lbo_b = (u32)NA * 16, sbo = 128;
            for (int i = 0; i < pos_per_split; ++i) {
                int st = i % kFcStages;
                mbar_wait(bar_full0 + 8 * st, (i / kFcStages) & 1);
                tc_fence_after();
                const u32 a_addr = smem_u32(s_a + (size_t)st * kFcStageA);
                const u32 b_addr = smem_u32(s_b + (size_t)st * stage_b_bytes);
#pragma unroll
                for (int k16 = 0; k16 < 8; ++k16) {
                    u64 ad = umma_desc(a_addr + (u32)(2 * k16) * lbo_a, lbo_a, sbo);
                    u64 bd = umma_desc(b_addr + (u32)(2 * k16) * lbo_b, lbo_b, sbo);
                    tc_mma_bf16(tmem_base, ad, bd, idesc, (i | k16) ? 1u : 0u);
                }
                tc_commit(bar_empty0 + 8 * st);
            }
            tc_commit(bar_acc);
        }
    } else {
        const int quarter = warp & 3;
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int board = tile * 128 + quarter * 32 + lane;
        float* out = args.partial + ((size_t)split * args.n_boards_cap + board) * NA;
        for (int c0 = 0; c0 < NA; c0 += 16) {
            u32 v[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(tmem_base + ((u32)(quarter * 32) << 16) + (u32)c0)
                : "memory");
            tc_wait_ld();
            if (board < live) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                           __uint_as_float(v[j + 3]));
                    *reinterpret_cast<float4*>(out + c0 + j) = o;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
    }
}

// Sums the split-K partial logits, runs the value-head FCs and GameNet.evaluate's
// post-processing (gamenet.py:67-85).  One warp per board.
constexpr int kFinishBoards = 16;
// GAME >= 0: instantiated for one (game, board size), so the legality mask's bitboard code has constant shifts.
template <int GAME, int G>
__global__ void __launch_bounds__(kFinishBoards * 32, 2) k_heads_finish(GameCfg cfg_in, int n, const int* count_dev, const int* index,
                                                      const u64* first,
                                                      const u64* second, const u8* player,
                                                      const float* __restrict__ partial, int split, int n_cap, int NA,
                                                      const float* __restrict__ vplane, HeadParams hp, int mode,
                                                      float* value_out, float* probs_out) {
    extern __shared__ __align__(16) float s_w1[];  // [P0][H] transposed value-FC1 weights
    __shared__ float s_vp[kFinishBoards][64];
    const GameCfg cfg = GAME >= 0 ? const_game_cfg<(GAME >= 0 ? GAME : 0), (G > 0 ? G : 8)>() : cfg_in;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gb = blockIdx.x * kFinishBoards + warp;
    const int P0 = cfg.g * cfg.g, A = cfg.num_actions;
    const int live = live_count(count_dev, n);
    if (blockIdx.x * kFinishBoards >= live) return;
    {
        const int n4 = (P0 * hp.H) >> 2;  // 16-byte staging (P0*H is a multiple of 4 for every supported net)
        const float4* src = reinterpret_cast<const float4*>(hp.vfc1_t);
        float4* dst = reinterpret_cast<float4*>(s_w1);
        for (int i = threadIdx.x; i < n4; i += blockDim.x) dst[i] = __ldg(src + i);
        for (int i = (n4 << 2) + threadIdx.x; i < P0 * hp.H; i += blockDim.x) s_w1[i] = hp.vfc1_t[i];
    }
    __syncthreads();
    if (gb >= live) return;
    const int ob = index ? index[gb] : gb;   // the leaf's own slot: where its state is read and its result written
    GState s;
    s.first = first[ob];
    s.second = second[ob];
    s.player = player[ob];
    for (int p = lane; p < P0; p += 32) s_vp[warp][p] = vplane[(size_t)gb * P0 + p];
    __syncwarp();
    float v = 0.f;
    if ((hp.H & 127) == 0) {
        // vfc1_t is the transposed weight [pos][H]: a lane owns four consecutive hidden units, so one
        // 16-byte shared-memory read per position feeds four FMAs (same summation order as below)
        for (int j0 = lane * 4; j0 < hp.H; j0 += 128) {
            float a0 = hp.vfc1_b[j0], a1 = hp.vfc1_b[j0 + 1], a2 = hp.vfc1_b[j0 + 2], a3 = hp.vfc1_b[j0 + 3];
            for (int p = 0; p < P0; ++p) {
                const float x = s_vp[warp][p];
                const float4 w = *reinterpret_cast<const float4*>(s_w1 + p * hp.H + j0);
                a0 = fmaf(x, w.x, a0);
                a1 = fmaf(x, w.y, a1);
                a2 = fmaf(x, w.z, a2);
                a3 = fmaf(x, w.w, a3);
            }
            v = fmaf(fmaxf(a0, 0.f), hp.vfc2_w[j0], v);
            v = fmaf(fmaxf(a1, 0.f), hp.vfc2_w[j0 + 1], v);
            v = fmaf(fmaxf(a2, 0.f), hp.vfc2_w[j0 + 2], v);
            v = fmaf(fmaxf(a3, 0.f), hp.vfc2_w[j0 + 3], v);
        }
    } else {
        for (int j = lane; j < hp.H; j += 32) {
            float acc = hp.vfc1_b[j];
            for (int p = 0; p < P0; ++p) acc = fmaf(s_vp[warp][p], s_w1[p * hp.H + j], acc);
            v = fmaf(fmaxf(acc, 0.f), hp.vfc2_w[j], v);
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(DM_FULL_MASK, v, off);
    v += hp.vfc2_b;
    const bool transpose = (cfg.game == DM_HEX || cfg.game == DM_HEX_SWAP) && s.player == 0;
    float lg[3];
    int na = 0;
    for (int a = lane; a < A; a += 32, ++na) {
        int src = a;
        if (transpose && a < P0) src = (a % cfg.g) * cfg.g + a / cfg.g;
        float acc = 0.f;
        for (int sp = 0; sp < split; ++sp) acc += partial[((size_t)sp * n_cap + gb) * NA + src];
        lg[na] = acc + hp.pfc_b[src];
    }
    if (mode == 1) {
        if (lane == 0) value_out[ob] = v;
        na = 0;
        for (int a = lane; a < DM_MAX_ACTIONS; a += 32, ++na)
            probs_out[(size_t)ob * DM_MAX_ACTIONS + a] = a < A ? lg[na] : 0.f;
        return;
    }
    float tv = tanhf(v);
    if (s.player == 1) tv = -tv;
    if (lane == 0) value_out[ob] = (tv + 1.f) / 2.f;
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
        probs_out[(size_t)ob * DM_MAX_ACTIONS + a] = a < A ? e[na] / msum : 0.f;
}

}  // namespace netk
