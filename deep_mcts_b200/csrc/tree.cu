// PUCT tree search over thousands of independent game trees, one warp per tree.
//
// Replaces MCTS.step / simulation / tree_search / evaluate_leaf / expand_node / rollout /
// backpropagate / add_dirichlet_noise / self_play / reset and MCTSAgent.play's re-rooting
// (reference deep_mcts/mcts.py:57-300) plus create_self_play_examples' labelling
// (deep_mcts/train.py:258-283).
//
// Data layout (HBM, structure of arrays, tree t owns node slots [t*cap, (t+1)*cap)):
//   visits u32 | value_sum f64 | prior f64 | first_child u32 | n_children u8 | action u8
// The children of a node are contiguous and stored in the order the reference's dict
// iterates them (CPython set order, games.cuh), so a warp scores up to 32 children per
// coalesced load and first-extremum ties (mcts.py:167-176) resolve to the lowest child slot.
// All PUCT arithmetic is IEEE double without contraction in the reference's operation
// order, one simulation in flight per tree, which makes visit counts bit-identical to the
// reference for the same evaluator outputs.
#include "games.cuh"

#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr u8 kFlagNeedNoise = 1;
constexpr u8 kFlagStepActive = 2;
constexpr u8 kLeafRoot = 1;    // leaf is a root expansion: only root.visits += 1 (mcts.py:114-116)
constexpr u8 kLeafAlias = 2;   // evaluation cache: another tree's leaf of the same round is the same state; its slot holds the result
constexpr u8 kLeafCached = 4;  // evaluation cache: the result was copied from the table into the tree's hit row

enum Counter {
    kCntSims = 0, kCntLeafEvals = 1, kCntTerminal = 2, kCntNodesHigh = 3, kCntGames = 4, kCntDepth = 5,
    kCntScanned = 6, kCntCreated = 7, kCntCapacity = 8, kCntMoves = 9, kCntRolloutPlies = 10, kCntInvalid = 11,
    kCntCompactions = 12, kCntCacheHits = 13, kCntCacheAliases = 14, kCntCacheInserts = 15
};

// ---- evaluation cache (cached_state_evaluator, train.py:408-414) ---------------------------
// Open-addressing table in HBM keyed by the 17-byte state; a row is the evaluator's output for that
// state (priors by action, then the value).  ctl word: status[63:62] player[61] stamp[60:32] slot[31:0].
constexpr int kCacheWays = 8;           // consecutive entries probed by lanes 0..7
constexpr int kCacheRow = 68;           // DM_MAX_ACTIONS priors + value + 2 pad (16-byte multiple)
constexpr u64 kCtlEmpty = 0, kCtlLocked = 1, kCtlPending = 2, kCtlValid = 3;
constexpr u32 kStampMask = 0x1fffffffu;
constexpr int kCacheMiss = 0, kCacheHit = 1, kCacheAlias = 2;

struct PoolView {
    GameCfg cfg;
    int n_trees;
    u32 cap;
    double c_puct;
    u64 seed;
    // search configuration (mcts.py:70-95)
    int num_simulations, eval_kind, rollout_kind, sample_move_cutoff;
    double dirichlet_alpha, dirichlet_factor, rollout_share;
    u64 eval_salt;
    int leaves_per_round;      // 1 = exact; > 1 = virtual-loss mode
    int level_budget;          // self-play: tree levels a warp may walk per launch before it sits the round out
    const double* host_noise;  // [n_trees][DM_MAX_ACTIONS] in child order, or null
    // nodes
    u32* visits;
    double* value_sum;
    double* prior;
    u32* first_child;
    u8* n_children;
    u8* action;
    // the node's own state and whether it is final (0 = no, else 1 + 2 * outcome): written once at expansion, one child
    // per lane, so the selection walks the tree without replaying a single move (Othello flips, Hex flood fills)
    u64* node_first;
    u64* node_second;
    u8* node_final;
    // per tree
    u32* node_count;
    u32* root;
    u64* root_first;
    u64* root_second;
    u8* root_player;
    int* sims_left;
    u32* stat_with;
    u32* stat_without;
    u8* flags;
    u32* ply;
    u32* rng_ctr;
    u32* path;      // [n_trees][DM_MAX_DEPTH] pending leaf's root-to-leaf path
    int* path_len;  // [n_trees]
    // pending leaf batch (compact)
    u64* leaf_first;
    u64* leaf_second;
    u8* leaf_player;
    u8* leaf_flags;
    int* leaf_tree;
    int* leaf_count;
    // per-tree copy of the pending leaf (slot, state, flags): lets k_round expand it while other warps of the
    // same launch already claim slots of the next batch
    int* pend_slot;
    u64* pend_first;
    u64* pend_second;
    u8* pend_player;
    u8* pend_flags;
    // virtual-loss mode: up to DM_MAX_LEAVES pending paths per tree
    u32* vl_path;   // [n_trees][DM_MAX_LEAVES][DM_MAX_DEPTH]
    int* vl_len;    // [n_trees][DM_MAX_LEAVES]
    int* vl_slot;   // [n_trees][DM_MAX_LEAVES] leaf-batch slot of each pending path
    int* vl_count;  // [n_trees]
    int* armed;     // number of trees that still have simulations left after the last select
    // arena compaction scratch (null = compaction disabled): live-node bitmap and per-word prefix counts
    u32* cmark;     // [n_trees][cap_words]
    u32* cprefix;   // [n_trees][cap_words]
    u32 cap_words;
    unsigned long long* counters;   // [16] global counters (kernels outside the hot loop add here atomically)
    unsigned long long* tcount;     // [n_trees][16] per-tree counter rows: the hot kernels add here without atomics
    u32* scnt;                      // this warp's shared-memory accumulator (set inside the hot kernels), else null
    // evaluation cache (cctl == null: disabled)
    unsigned long long* cctl;   // [centries]
    u64* ckey_first;            // [centries]
    u64* ckey_second;           // [centries]
    void* crow;                 // [centries][kCacheRow] float or double (cache_f64)
    void* hit_row;              // [n_trees][kCacheRow] row copied out of the table for a pending cache hit
    u32 cmask;                  // centries - 1
    int cache_f64;
    u32* cround;                // [0] search-round stamp, [1] blocks finished in the running launch
    int* pend_centry;           // [n_trees] entry claimed for the pending leaf (-1: none)
    unsigned long long* pend_cctl;  // [n_trees] the pending ctl word that claim published
    unsigned long long* dbg_warp;   // [n_trees][4] optional per-warp timing of k_round (dm_pool_debug_timing), else null
    u8* leaf_live;              // [n_trees] dense mode: 1 = slot t holds a leaf the evaluator must run
    int* leaf_index;            // [n_trees] dense mode: compacted list of the live slots (count in leaf_count)
    // self-play recording
    int selfplay;
    u64* ex_first;  // [n_trees][DM_MAX_DEPTH]
    u64* ex_second;
    u8* ex_player;
    float* ex_pi;  // [n_trees][DM_MAX_DEPTH][DM_MAX_ACTIONS]
    u64 replay_cap;
    u64* rp_first;
    u64* rp_second;
    u8* rp_player;
    float* rp_pi;
    float* rp_z;
    unsigned long long* rp_written;   // positions written (low DM_REPLAY_POS_BITS bits) | games written (high bits)
    unsigned long long* rp_game_end;  // [DM_REPLAY_GAME_RING] positions written up to and including game i
};

// ------------------------------------------------------------------ small warp helpers
// Counters.  The hot kernels (one warp per tree, thousands of trees) accumulate in a per-warp shared-memory row and
// add it to the tree's own row of tcount once per launch: ~10 atomics per tree per round on ONE cache line become
// one coalesced read-modify-write of a line nobody else touches.  Everything else adds to the global row atomically.
// Called by one lane of the warp only.
__device__ __forceinline__ void count(const PoolView& P, int which, unsigned long long v) {
#ifndef DM_EXP_NOCOUNT   // timing experiment: how much of the tree kernels is the shared counters
    if (P.scnt) P.scnt[which] += (u32)v;
    else if (v) atomicAdd(&P.counters[which], v);
#endif
}
__device__ __forceinline__ void count_nodes_high(const PoolView& P, u32 nodes) {
    if (P.scnt) P.scnt[kCntNodesHigh] = max(P.scnt[kCntNodesHigh], nodes);
    else atomicMax(&P.counters[kCntNodesHigh], (unsigned long long)nodes);
}
// whole warp: clear / flush this warp's accumulator row
__device__ __forceinline__ void counters_begin(u32* row, int lane) {
    if (lane < 16) row[lane] = 0;
    __syncwarp();
}
__device__ __forceinline__ void counters_flush(const PoolView& P, int t, const u32* row, int lane) {
    __syncwarp();
    if (lane < 16) {
        const u32 v = row[lane];
        if (v) {
            unsigned long long* dst = P.tcount + (size_t)t * 16 + lane;
            *dst = (lane == kCntNodesHigh) ? max(*dst, (unsigned long long)v) : *dst + v;
        }
    }
}

__device__ __forceinline__ GState root_state(const PoolView& P, int t) {
    GState s;
    s.first = P.root_first[t];
    s.second = P.root_second[t];
    s.player = P.root_player[t];
    return s;
}

__device__ __forceinline__ void clear_node(const PoolView& P, size_t idx) {
    P.visits[idx] = 0;
    P.value_sum[idx] = 0.0;
    P.prior[idx] = -1.0;
    P.first_child[idx] = 0;
    P.n_children[idx] = 0;
    P.action[idx] = 0xff;
}

// fresh single-node tree on state s (lane 0 only)
__device__ __forceinline__ void fresh_root(const PoolView& P, int t, const GState& s) {
    size_t base = (size_t)t * P.cap;
    clear_node(P, base);
    P.node_first[base] = s.first;
    P.node_second[base] = s.second;
    P.node_final[base] = (u8)game_final_code(P.cfg, s);
    P.node_count[t] = 1;
    P.root[t] = 0;
    P.root_first[t] = s.first;
    P.root_second[t] = s.second;
    P.root_player[t] = (u8)s.player;
}

// ---- synthetic integer-hash evaluator (oracle/evaluators.py: hashed_salted) -------------
__device__ __forceinline__ u64 mix64(u64 x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ u64 hashed_h0(const GState& s, u64 salt) {
    return mix64(s.first ^ mix64(s.second ^ mix64((u64)s.player + 0x9E3779B97F4A7C15ull * (salt + 1))));
}
__device__ __forceinline__ u32 hashed_weight(u64 h0, int a) {
    return (u32)((mix64(h0 + (u64)a * 0xD6E8FEB86659FD93ull) >> 40) % 997ull) + 1u;
}


// ---- evaluation cache: device side ---------------------------------------------------------
// Result-transparent memo of the external evaluator (cached_state_evaluator, train.py:408-414) plus
// de-duplication of equal leaves inside one round.  Protocol (one search round = one launch of a
// selecting kernel; `now` is the round stamp, bumped by the last block of that launch):
//   select:  probe kCacheWays consecutive entries.
//            valid entry with the leaf's key   -> HIT: the row is copied to the tree's hit row now (rows of
//                                                 valid entries are never written during a launch) and the
//                                                 leaf needs no evaluation;
//            pending entry stamped `now`       -> ALIAS: another tree hands the same state to the evaluator
//                                                 in this round; remember its slot;
//            otherwise                         -> MISS: lock a victim entry (CAS), write the key, publish
//                                                 pending(now, slot); the leaf is evaluated.
//   expand:  the tree that published an entry copies the evaluator's output into the row and flips the
//            entry to valid (only if the ctl word is still the one it published).
// A key is matched seqlock-style (ctl, key, ctl again), rows are written only while their entry is pending
// and read only while it is valid, so a reader never sees a torn key or row; everything else (lost
// claims, replaced entries) only costs a missed hit.  Control words and keys are read through L2.
__device__ __forceinline__ u64 ctl_make(u64 status, int player, u32 stamp, u32 slot) {
    return (status << 62) | ((u64)(player & 1) << 61) | ((u64)(stamp & kStampMask) << 32) | (u64)slot;
}
__device__ __forceinline__ u64 ld_ctl(const unsigned long long* p) { return __ldcv(p); }

struct CacheProbe {
    int kind;    // kCacheMiss / kCacheHit / kCacheAlias
    int entry;   // hit: matching entry; miss: locked victim entry (caller must publish) or -1
    int slot;    // alias: evaluator slot that will hold the result
};

__device__ __forceinline__ CacheProbe cache_probe(const PoolView& P, const GState& s, int lane, u32 now) {
    CacheProbe r;
    r.kind = kCacheMiss;
    r.entry = -1;
    r.slot = -1;
    const u64 h = mix64(s.first ^ mix64(s.second + 0x9E3779B97F4A7C15ull * (u64)(s.player + 1)));
    const u32 idx = ((u32)(h >> 17) + (u32)lane) & P.cmask;
    const bool prober = lane < kCacheWays;
    for (int attempt = 0; attempt < 16; ++attempt) {
        u64 c = 0;
        u32 st = (u32)kCtlLocked;
        bool match = false;
        if (prober) {
            c = ld_ctl(P.cctl + idx);
            st = (u32)(c >> 62);
            if (st >= (u32)kCtlPending && (int)((c >> 61) & 1) == s.player) {
                const u64 kf = __ldcv(P.ckey_first + idx), ks = __ldcv(P.ckey_second + idx);
                __threadfence();
                match = kf == s.first && ks == s.second && ld_ctl(P.cctl + idx) == c;
            }
        }
        const u32 mm = __ballot_sync(DM_FULL_MASK, match);
        if (mm) {
            const int src = __ffs(mm) - 1;
            const u64 cm = __shfl_sync(DM_FULL_MASK, c, src);
            const u32 e = __shfl_sync(DM_FULL_MASK, idx, src);
            if ((cm >> 62) == kCtlValid) {
                r.kind = kCacheHit;
                r.entry = (int)e;
                return r;
            }
            if ((u32)((cm >> 32) & kStampMask) == (now & kStampMask)) {
                r.kind = kCacheAlias;
                r.slot = (int)(u32)cm;
                return r;
            }
            // a pending entry of an earlier round: its result is in flight or lost -- evaluate again, do not insert
            return r;
        }
        const u32 locked = __ballot_sync(DM_FULL_MASK, prober && st == (u32)kCtlLocked);
        if (locked && attempt < 12) {   // somebody is publishing in this window, possibly this very key
            __nanosleep(40);
            continue;
        }
        // victim: empty, else a pending entry nobody will complete, else the oldest valid entry; entries of
        // this and the previous round (live owners) last
        u32 score = 0xffffffffu;
        if (prober && st != (u32)kCtlLocked) {
            const u32 age = (now - (u32)((c >> 32) & kStampMask)) & kStampMask;
            if (st == (u32)kCtlEmpty) score = 0;
            else if (st == (u32)kCtlPending) score = age >= 2 ? 1u : 0xfffffff0u + age;
            else score = 2u + (kStampMask - age);
        }
        const u32 best = __reduce_min_sync(DM_FULL_MASK, score);
        if (best == 0xffffffffu) return r;
        const int src = __ffs(__ballot_sync(DM_FULL_MASK, score == best)) - 1;
        const u32 e = __shfl_sync(DM_FULL_MASK, idx, src);
        const u64 cv = __shfl_sync(DM_FULL_MASK, c, src);
        int ok = 0;
        if (lane == 0) ok = atomicCAS(P.cctl + e, cv, ctl_make(kCtlLocked, 0, 0, 0)) == cv;
        ok = __shfl_sync(DM_FULL_MASK, ok, 0);
        if (ok) {
            r.entry = (int)e;
            return r;
        }
    }
    return r;
}

// after a successful lock: write the key and publish pending(now, slot)   (lane 0)
__device__ __forceinline__ void cache_publish(const PoolView& P, int t, int entry, const GState& s, u32 now, int slot) {
    P.ckey_first[entry] = s.first;
    P.ckey_second[entry] = s.second;
    __threadfence();
    const u64 c = ctl_make(kCtlPending, s.player, now, (u32)slot);
    atomicExch(P.cctl + entry, c);
    P.pend_centry[t] = entry;
    P.pend_cctl[t] = c;
}

// copy the row of a (validated) hit into the tree's hit row; whole warp
template <typename PriorT>
__device__ __forceinline__ void cache_copy_hit(const PoolView& P, int t, int entry, int lane) {
    const PriorT* row = (const PriorT*)P.crow + (size_t)entry * kCacheRow;
    PriorT* dst = (PriorT*)P.hit_row + (size_t)t * kCacheRow;
    for (int j = lane; j < DM_MAX_ACTIONS + 1; j += 32) dst[j] = __ldcg(row + j);
    __syncwarp();
}

// the owner of a pending entry stores the evaluator's output; whole warp
template <typename PriorT>
__device__ __forceinline__ void cache_finalize(const PoolView& P, int t, int lane, const PriorT* pr, PriorT value) {
    const int e = P.pend_centry[t];
    if (e < 0) return;
    const u64 expect = P.pend_cctl[t];
    if (ld_ctl(P.cctl + e) == expect) {
        PriorT* row = (PriorT*)P.crow + (size_t)e * kCacheRow;
        for (int j = lane; j < P.cfg.num_actions; j += 32) row[j] = pr[j];
        if (lane == 0) row[DM_MAX_ACTIONS] = value;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            atomicCAS(P.cctl + e, expect, (expect & ~(3ull << 62)) | (kCtlValid << 62));
            count(P, kCntCacheInserts, 1);
        }
    }
    if (lane == 0) P.pend_centry[t] = -1;
    __syncwarp();
}

// End of a selecting launch.  The last block to finish bumps the round stamp and, in dense mode, compacts the
// live leaf slots into leaf_index / leaf_count for the evaluator (ascending tree order, deterministic).
__device__ __forceinline__ void round_end(const PoolView& P, bool dense) {
    __shared__ u32 s_last;
    __shared__ u32 s_warp_sums[kWarpsPerBlock];
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicInc(P.cround + 1, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) P.cround[0] = (P.cround[0] + 1) & kStampMask;
    if (!dense) return;
    // every thread takes 32 consecutive trees per pass: one 32-byte load of their flags (no dependent-load chain in
    // this serial tail of the launch), a bit mask, one block-wide exclusive scan of the counts
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nthreads = kWarpsPerBlock * 32;
    u32 running = 0;
    for (int base = 0; base < P.n_trees; base += nthreads * 32) {
        const int t0 = base + (int)threadIdx.x * 32;
        u32 mask = 0;
        if (t0 + 32 <= P.n_trees) {
            const uint4* src = reinterpret_cast<const uint4*>(P.leaf_live + t0);
            const uint4 a = __ldcv(src), b = __ldcv(src + 1);
            const u32 w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
            for (int k = 0; k < 8; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if ((w[k] >> (8 * j)) & 0xffu) mask |= 1u << (4 * k + j);
        } else {
            for (int j = 0; j < 32 && t0 + j < P.n_trees; ++j)
                if (__ldcv(P.leaf_live + t0 + j)) mask |= 1u << j;
        }
        const u32 cnt = __popc(mask);
        u32 incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const u32 up = __shfl_up_sync(DM_FULL_MASK, incl, off);
            if (lane >= off) incl += up;
        }
        if (lane == 31) s_warp_sums[warp] = incl;
        __syncthreads();
        u32 pos = running + incl - cnt;
        for (int w2 = 0; w2 < warp; ++w2) pos += s_warp_sums[w2];
        while (mask) {
            P.leaf_index[pos++] = t0 + __ffs(mask) - 1;
            mask &= mask - 1;
        }
        for (int w2 = 0; w2 < kWarpsPerBlock; ++w2) running += s_warp_sums[w2];
        __syncthreads();
    }
    if (threadIdx.x == 0) *P.leaf_count = (int)running;
}

// ------------------------------------------------------------------------------ select
// tree_search (mcts.py:162-179): descend from the root while the node has children.
// s_path (shared, per warp) receives the node indices; returns path length.
__device__ __forceinline__ int descend(const PoolView& P, int t, int lane, GState& s, u32* s_path,
                                       unsigned long long& scanned, bool& overflow, int& final_code) {
    const size_t base = (size_t)t * P.cap;
    u32 node = P.root[t];
    s.player = P.root_player[t];
    int len = 0;
    if (lane == 0) s_path[0] = node;
    len = 1;
    overflow = false;
    // header of the current node; below the root it comes out of the parent's child scan, so one
    // level costs ONE round of (parallel, coalesced) loads instead of a chain of dependent ones
    int n = P.n_children[base + node];
    u32 first = P.first_child[base + node];
    u32 parent_visits = P.visits[base + node];
    double sq = sqrt((double)parent_visits);          // IEEE sqrt, as math.sqrt
    while (n != 0) {
        bool maximise = (s.player == 0);              // SECOND is the max player (game.py:31-33)
        double unvisited = maximise ? 0.0 : 1.0;      // parent.player.loss().value (mcts.py:48-50)
        double best_key = 0.0;
        int best = -1;
        // whole record of this lane's best child
        u32 b_visits = 0, b_first = 0;
        int b_n = 0;
        // children lane and lane + 32 are loaded together so their divisions overlap; a third
        // (only Hex-with-swap 8x8 has 65 actions) takes the loop again
        for (int j0 = lane; j0 < n; j0 += 64) {
            const int j1 = j0 + 32;
            const bool has1 = j1 < n;
            size_t c0 = base + first + j0, c1 = has1 ? c0 + 32 : c0;
            u32 v0 = P.visits[c0], v1 = P.visits[c1];
            double vs0 = P.value_sum[c0], vs1 = P.value_sum[c1];
            double pr0 = P.prior[c0], pr1 = P.prior[c1];
            u32 cf0 = P.first_child[c0], cf1 = P.first_child[c1];
            int cn0 = P.n_children[c0], cn1 = P.n_children[c1];
            double q0 = (v0 == 0) ? unvisited : __ddiv_rn(vs0, (double)v0);
            double q1 = (v1 == 0) ? unvisited : __ddiv_rn(vs1, (double)v1);
            // u = ((c_puct * P) * sqrt(N)) / (1 + n)   (mcts.py:40-43), no FMA contraction
            double u0 = __ddiv_rn(__dmul_rn(__dmul_rn(P.c_puct, pr0), sq), (double)(1u + v0));
            double u1 = __ddiv_rn(__dmul_rn(__dmul_rn(P.c_puct, pr1), sq), (double)(1u + v1));
            // "+ 0.0" folds -0.0 into +0.0 so the integer comparison below orders like the doubles
            double k0 = (maximise ? __dadd_rn(q0, u0) : -__dsub_rn(q0, u0)) + 0.0;
            double k1 = (maximise ? __dadd_rn(q1, u1) : -__dsub_rn(q1, u1)) + 0.0;
            if (best < 0 || k0 > best_key) {
                best_key = k0;
                best = j0;
                b_visits = v0;
                b_first = cf0;
                b_n = cn0;
            }
            if (has1 && k1 > best_key) {
                best_key = k1;
                best = j1;
                b_visits = v1;
                b_first = cf1;
                b_n = cn1;
            }
        }
        // sqrt for the next level, off the critical path of the reduction below
        const double b_sq = sqrt((double)b_visits);
        // warp arg-max (first extremum wins): order-preserving integer image of the key, reduced
        // with three redux.sync instead of five shuffle rounds
        unsigned long long kb = 0;
        if (best >= 0) {
            unsigned long long bits = (unsigned long long)__double_as_longlong(best_key);
            kb = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
        }
        const u32 hi = (u32)(kb >> 32), lo = (u32)kb;
        const u32 mh = __reduce_max_sync(DM_FULL_MASK, hi);
        bool cand = hi == mh;
        const u32 ml = __reduce_max_sync(DM_FULL_MASK, cand ? lo : 0u);
        cand = cand && lo == ml;
        best = (int)__reduce_min_sync(DM_FULL_MASK, cand ? (u32)best : 0xffffffffu);
        // the winning child was scanned by lane (best & 31), whose local best it is
        const int owner = best & 31;
        parent_visits = __shfl_sync(DM_FULL_MASK, b_visits, owner);
        sq = __shfl_sync(DM_FULL_MASK, b_sq, owner);
        const u32 next_first = __shfl_sync(DM_FULL_MASK, b_first, owner);
        const int next_n = __shfl_sync(DM_FULL_MASK, b_n, owner);
        scanned += (unsigned long long)n;
        node = first + (u32)best;
        s.player = 1 - s.player;        // every action, pass and swap included, hands the move over
        if (len >= DM_MAX_DEPTH) {
            overflow = true;
            break;
        }
        if (lane == 0) s_path[len] = node;
        ++len;
        first = next_first;
        n = next_n;
    }
    // the leaf's state and finality were stored when its parent was expanded
    s.first = P.node_first[base + node];
    s.second = P.node_second[base + node];
    final_code = P.node_final[base + node];
    __syncwarp();
    return len;
}

// backpropagate (mcts.py:228-231)
__device__ __forceinline__ void backup(const PoolView& P, int t, int lane, const u32* path, int len, double value) {
    const size_t base = (size_t)t * P.cap;
    for (int i = lane; i < len; i += 32) {
        size_t idx = base + path[i];
        P.visits[idx] += 1;
        P.value_sum[idx] = __dadd_rn(P.value_sum[idx], value);
    }
    __syncwarp();
}

// expand_node (mcts.py:198-217).  Priors come from the built-in evaluators or from an
// external array (ext_f32 / ext_f64 indexed by action).  Returns false on capacity error.
__device__ __forceinline__ bool expand(const PoolView& P, int t, int lane, u32 node, const GState& s, u8* s_order,
                                       int eval_kind, const float* ext_f32, const double* ext_f64,
                                       double& eval_value) {
    const size_t base = (size_t)t * P.cap;
    u64 flips = 0;
    bool have_flips = false;  // Othello: the child order's small-set path already walked every move's rays
    int n = warp_children_order(P.cfg, s, lane, s_order, nullptr, &flips, &have_flips);
    u32 cnt = P.node_count[t];
    if (cnt + (u32)n > P.cap) {
        if (lane == 0) count(P, kCntCapacity, 1);
        return false;
    }
    double uniform_p = __ddiv_rn(1.0, (double)n);
    u64 h0 = 0;
    double total = 1.0;
    eval_value = 0.5;
    if (eval_kind == DM_EVAL_HASHED) {
        h0 = hashed_h0(s, P.eval_salt);
        u32 sum = 0;
        for (int j = lane; j < n; j += 32) sum += hashed_weight(h0, s_order[j]);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(DM_FULL_MASK, sum, off);
        total = (double)sum;
        eval_value = __ddiv_rn((double)((h0 >> 20) % 1001ull), 1000.0);
    }
    for (int j = lane; j < n; j += 32) {
        size_t c = base + cnt + j;
        int a = s_order[j];
        double p;
        if (eval_kind == DM_EVAL_HASHED) p = __ddiv_rn((double)hashed_weight(h0, a), total);
        else if (eval_kind == DM_EVAL_EXTERNAL) p = ext_f64 ? __ldcg(ext_f64 + a) : (double)__ldcg(ext_f32 + a);
        else p = uniform_p;
        P.visits[c] = 0;
        P.value_sum[c] = 0.0;
        P.prior[c] = p;
        P.first_child[c] = 0;
        P.n_children[c] = 0;
        P.action[c] = (u8)a;
        GState child;
        if (have_flips) {  // n <= 18: one pass of this loop, flips belong to s_order[lane]
            const u64 bit = 1ull << a;
            child.player = 1 - s.player;
            child.first = s.player ? (s.first | bit | flips) : (s.first & ~flips);
            child.second = s.player ? (s.second & ~flips) : (s.second | bit | flips);
        } else {
            child = game_child_uniform(P.cfg, s, a);
        }
        P.node_first[c] = child.first;
        P.node_second[c] = child.second;
        P.node_final[c] = (u8)game_child_final_code(P.cfg, s, a, child);
    }
    if (lane == 0) {
        P.first_child[base + node] = cnt;
        P.n_children[base + node] = (u8)n;
        P.node_count[t] = cnt + (u32)n;
        count(P, kCntCreated, (unsigned long long)n);
        count_nodes_high(P, cnt + (u32)n);
    }
    __syncwarp();
    return true;
}

// ---- Dirichlet noise (mcts.py:233-241) --------------------------------------------------
// Gamma(alpha, 1) by Marsaglia-Tsang; alpha < 1 boosted through gamma(alpha + 1) * U^(1/alpha).  Single precision
// on purpose: the sample is only ever used as a random number (it is normalised and mixed into the priors in
// double), and the double-precision log / pow / cospi sequences made the ~2 % of trees that start a new move in a
// launch run 5-7x longer than all the others -- on Hex 7x7 (49 children, alpha 0.33: two boosted samples per lane)
// that single tail set the duration of the whole tree kernel (profiles/tree_tail_probe_r2.jsonl).
__device__ __forceinline__ float uniform_open_f(Philox& rng) {
    return ((float)(rng.next() >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0, 1), 24 bits
}
__device__ __forceinline__ double sample_gamma(Philox& rng, double alpha_d) {
    float alpha = (float)alpha_d;
    float boost = 1.0f;
    if (alpha < 1.0f) {
        boost = powf(uniform_open_f(rng), 1.0f / alpha);
        alpha += 1.0f;
    }
    const float d = alpha - 1.0f / 3.0f, c = rsqrtf(9.0f * d);
    while (true) {
        const float u1 = uniform_open_f(rng), u2 = uniform_open_f(rng);
        const float x = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
        float v = 1.0f + c * x;
        if (v <= 0.0f) continue;
        v = v * v * v;
        const float u = uniform_open_f(rng);
        if (logf(u) < 0.5f * x * x + d - d * v + d * logf(v)) return (double)(d * v * boost);
    }
}

__device__ __forceinline__ void apply_root_noise(const PoolView& P, int t, int lane) {
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    if (n == 0) return;
    u32 first = P.first_child[base + root];
    double noise[3] = {0.0, 0.0, 0.0};
    if (P.host_noise) {
        for (int k = 0, j = lane; j < n; j += 32, ++k) noise[k] = P.host_noise[(size_t)t * DM_MAX_ACTIONS + j];
    } else {
        Philox rng;
        rng.init(P.seed, (u32)t, 0x6e6f6900u + (u32)lane, P.rng_ctr[t]);
        double sum = 0.0;
        for (int k = 0, j = lane; j < n; j += 32, ++k) {
            noise[k] = sample_gamma(rng, P.dirichlet_alpha);
            sum += noise[k];
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(DM_FULL_MASK, sum, off);
        if (sum > 0.0) {
            for (int k = 0; k < 3; ++k) noise[k] /= sum;
        } else {   // every sample underflowed (alpha far below the reference's 0.33): fall back to no preference
            for (int k = 0; k < 3; ++k) noise[k] = 1.0 / (double)n;
        }
        __syncwarp();
        if (lane == 0) P.rng_ctr[t] += 64;
    }
    double keep = 1.0 - P.dirichlet_factor;
    for (int k = 0, j = lane; j < n; j += 32, ++k) {
        size_t c = base + first + j;
        // probability * (1 - factor) + noise[i] * factor
        P.prior[c] = __dadd_rn(__dmul_rn(P.prior[c], keep), __dmul_rn(noise[k], P.dirichlet_factor));
    }
    __syncwarp();
}

// arm one MCTS.step on tree t (lane 0)
__device__ __forceinline__ void arm_step(const PoolView& P, int t) {
    GState s = root_state(P, t);
    float o;
    bool final = game_is_final(P.cfg, s, o);
    P.sims_left[t] = final ? 0 : P.num_simulations;
    P.stat_with[t] = 0;
    P.stat_without[t] = 0;
    u8 f = P.flags[t] & ~(kFlagNeedNoise | kFlagStepActive);
    if (!final) {
        f |= kFlagStepActive;
        if (P.dirichlet_alpha != 0.0) f |= kFlagNeedNoise;
    }
    P.flags[t] = f;
}

// leaf value for the local (no external evaluator) modes: evaluate_leaf, mcts.py:186-196
__device__ __forceinline__ double mix_rollout(const PoolView& P, const GState& s, double eval_value, Philox& rng,
                                              unsigned long long& plies) {
    if (P.eval_kind == DM_EVAL_NONE) {
        int p;
        double r = (double)game_rollout(P.cfg, s, P.rollout_kind, rng, p);
        plies += p;
        return r;
    }
    if (P.rollout_kind == DM_ROLLOUT_NONE || (P.rollout_share < 1.0 && rng.uniform() >= P.rollout_share))
        return eval_value;
    int p;
    double r = (double)game_rollout(P.cfg, s, P.rollout_kind, rng, p);
    plies += p;
    return __ddiv_rn(__dadd_rn(r, eval_value), 2.0);
}

// ---- arena compaction ---------------------------------------------------------------------
// Re-rooting with subtree reuse (mcts.py:105-108) leaves the rest of the old tree unreachable.
// When a tree's arena cannot be sure to hold the next move's search at a re-root, the live subtree is slid to the front
// in place: (1) mark live nodes in one forward pass (children are always allocated after their
// parent, so marks flow forward), (2) prefix-count the marks per 32-node word, (3) move every
// live node to its rank.  Order is preserved, so contiguous child blocks stay contiguous, ranks
// never exceed the old index (in-place safe) and the child order -- hence tie-breaking -- is
// unchanged.  Whole warp.
__device__ __forceinline__ u32 live_rank(const u32* mark, const u32* prefix, u32 x) {
    return prefix[x >> 5] + __popc(mark[x >> 5] & ((1u << (x & 31)) - 1u));
}

__device__ __forceinline__ void compact_tree(const PoolView& P, int t, int lane) {
    if (!P.cmark) return;
    const u32 used = P.node_count[t];
    // compact only when the next move might not fit: a move expands at most (simulations + root) leaves of at most
    // num_actions children each.  (A pass over a 36 k-node Hex arena is ~0.2 ms for that one warp -- longer than the
    // whole launch otherwise takes -- so it must not run "when half full" but when it is needed.)
    const unsigned long long move_worst = (unsigned long long)(P.num_simulations + 2) * (unsigned long long)P.cfg.num_actions;
    const u32 need = (u32)min(move_worst, (unsigned long long)(P.cap / 2));
    if (used + need <= P.cap) return;
    const size_t base = (size_t)t * P.cap;
    u32* mark = P.cmark + (size_t)t * P.cap_words;
    u32* prefix = P.cprefix + (size_t)t * P.cap_words;
    const u32 nw = (used + 31) >> 5;
    const u32 root = P.root[t];
    for (u32 w = lane; w < nw; w += 32) mark[w] = 0;
    __syncwarp();
    if (lane == 0) mark[root >> 5] = 1u << (root & 31);
    __syncwarp();
    // (1) forward marking
    for (u32 w = root >> 5; w < nw; ++w) {
        u32 processed = 0;
        while (true) {
            u32 word = mark[w];
            u32 todo = word & ~processed;
            if (!todo) break;
            if ((todo >> lane) & 1) {
                size_t idx = base + ((size_t)w << 5) + lane;
                u32 n = P.n_children[idx];
                u32 fc = P.first_child[idx];
                if (n && fc != 0xffffffffu) {
                    u32 lo = fc, hi = fc + n;  // [lo, hi)
                    for (u32 ww = lo >> 5; ww <= (hi - 1) >> 5; ++ww) {
                        u32 a = max(lo, ww << 5) - (ww << 5), b = min(hi, (ww + 1) << 5) - (ww << 5);  // bit range [a, b)
                        u32 bits = (b == 32 ? 0xffffffffu : ((1u << b) - 1u)) & ~((1u << a) - 1u);
                        atomicOr(&mark[ww], bits);
                    }
                }
            }
            processed |= todo;
            __syncwarp();
        }
    }
    // (2) exclusive prefix of popcounts per word
    u32 running = 0;
    for (u32 w0 = 0; w0 < nw; w0 += 32) {
        u32 w = w0 + lane;
        u32 c = (w < nw) ? __popc(mark[w]) : 0;
        u32 incl = c;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            u32 v = __shfl_up_sync(DM_FULL_MASK, incl, off);
            if (lane >= off) incl += v;
        }
        if (w < nw) prefix[w] = running + incl - c;
        running += __shfl_sync(DM_FULL_MASK, incl, 31);
    }
    __syncwarp();
    // (3) slide live nodes to their ranks, lowest index first
    for (u32 w = root >> 5; w < nw; ++w) {
        u32 word = mark[w];
        bool live = (word >> lane) & 1;
        size_t src = base + ((size_t)w << 5) + lane;
        u32 v = 0, fc = 0;
        double vs = 0.0, pr = 0.0;
        u8 nc = 0, ac = 0, fin = 0;
        u64 nf = 0, ns = 0;
        u32 dst = 0;
        if (live) {
            v = P.visits[src];
            vs = P.value_sum[src];
            pr = P.prior[src];
            fc = P.first_child[src];
            nc = P.n_children[src];
            ac = P.action[src];
            nf = P.node_first[src];
            ns = P.node_second[src];
            fin = P.node_final[src];
            dst = prefix[w] + __popc(word & ((1u << lane) - 1u));
            if (nc) fc = live_rank(mark, prefix, fc);
        }
        __syncwarp();
        if (live) {
            size_t d = base + dst;
            P.visits[d] = v;
            P.value_sum[d] = vs;
            P.prior[d] = pr;
            P.first_child[d] = fc;
            P.n_children[d] = nc;
            P.action[d] = ac;
            P.node_first[d] = nf;
            P.node_second[d] = ns;
            P.node_final[d] = fin;
        }
        __syncwarp();
    }
    if (lane == 0) {
        P.root[t] = 0;
        P.node_count[t] = running;
        count(P, kCntCompactions, 1);
    }
    __syncwarp();
}

// --------------------------------------------------------------------- kernels: admin
__global__ void k_reset(PoolView P, const int* ids, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = ids ? ids[i] : i;
    if (t < 0 || t >= P.n_trees) return;
    fresh_root(P, t, game_initial(P.cfg));
    P.sims_left[t] = 0;
    P.stat_with[t] = P.stat_without[t] = 0;
    P.flags[t] = 0;
    P.ply[t] = 0;
    P.path_len[t] = 0;
}

__global__ void k_set_root(PoolView P, const int* ids, int n, const u64* first, const u64* second,
                           const u8* player) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = ids ? ids[i] : i;
    if (t < 0 || t >= P.n_trees) return;
    GState s;
    s.first = first[i];
    s.second = second[i];
    s.player = player[i];
    fresh_root(P, t, s);
    P.sims_left[t] = 0;
    P.flags[t] = 0;
}

// MCTSAgent.play (mcts.py:267-283): first child (dict order) whose state equals the observed one
__global__ void k_observe(PoolView P, const int* ids, int n, const u64* first, const u64* second,
                          const u8* player) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = blockIdx.x * kWarpsPerBlock + warp;
    if (i >= n) return;
    int t = ids ? ids[i] : i;
    if (t < 0 || t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    GState want;
    want.first = first[i];
    want.second = second[i];
    want.player = player[i];
    GState cur = root_state(P, t);
    u32 root = P.root[t];
    int nc = P.n_children[base + root];
    u32 fc = P.first_child[base + root];
    int found = -1;
    for (int j0 = 0; j0 < nc && found < 0; j0 += 32) {
        int j = j0 + lane;
        bool match = false;
        if (j < nc) {
            GState c = game_child(P.cfg, cur, P.action[base + fc + j]);
            match = (c.first == want.first && c.second == want.second && c.player == want.player);
        }
        u32 m = __ballot_sync(DM_FULL_MASK, match);
        if (m) found = j0 + __ffs(m) - 1;
    }
    if (lane == 0) {
        if (found >= 0) {
            P.root[t] = fc + (u32)found;
            P.root_first[t] = want.first;
            P.root_second[t] = want.second;
            P.root_player[t] = (u8)want.player;
        } else {
            fresh_root(P, t, want);
        }
        P.sims_left[t] = 0;
    }
    __syncwarp();
    compact_tree(P, t, lane);
}

__global__ void k_begin_step(PoolView P) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n_trees) return;
    arm_step(P, t);
}

// arm only the listed trees; every other tree is parked (sims_left = 0) for this step
__global__ void k_begin_step_ids(PoolView P, const int* ids, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_trees) {
        P.sims_left[i] = 0;
        P.stat_with[i] = 0;
        P.stat_without[i] = 0;
        P.flags[i] &= ~(kFlagNeedNoise | kFlagStepActive);
    }
}
__global__ void k_arm_ids(PoolView P, const int* ids, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int t = ids[i];
    if (t >= 0 && t < P.n_trees) arm_step(P, t);
}

// ------------------------------------------------------- kernel: whole search in one launch
__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_search_local(PoolView P) {
    __shared__ u32 s_path[kWarpsPerBlock][DM_MAX_DEPTH];
    __shared__ u8 s_order[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    int sims = P.sims_left[t];
    if (sims <= 0) return;
    const size_t base = (size_t)t * P.cap;
    u32* path = s_path[warp];
    u8* order = s_order[warp];
    Philox rng;
    rng.init(P.seed, (u32)t, 0x73696d00u, P.rng_ctr[t]);
    unsigned long long scanned = 0, depth = 0, plies = 0, evals = 0, terminal = 0, done = 0;
    bool dead = false;
    double ev;
    u32 root = P.root[t];
    if (P.n_children[base + root] == 0) {
        // step(): expand the root, root.visits += 1 (mcts.py:114-116)
        GState s = root_state(P, t);
        if (expand(P, t, lane, root, s, order, P.eval_kind, nullptr, nullptr, ev)) {
            if (lane == 0) P.visits[base + root] += 1;
            if (P.eval_kind != DM_EVAL_NONE) ++evals;
        } else {
            dead = true;
        }
        __syncwarp();
    }
    u8 flags = P.flags[t];
    if (!dead && (flags & kFlagNeedNoise)) {
        apply_root_noise(P, t, lane);
        flags &= ~kFlagNeedNoise;
    }
    u32 with = P.stat_with[t], without = P.stat_without[t];
    while (!dead && sims > 0) {
        GState s;
        bool overflow;
        int final_code;
        int len = descend(P, t, lane, s, path, scanned, overflow, final_code);
        if (overflow) {
            if (lane == 0) count(P, kCntCapacity, 1);
            break;
        }
        depth += (unsigned long long)(len - 1);
        double value;
        if (final_code) {
            value = 0.5 * (double)(final_code - 1);  // the outcome's value, mcts.py:182-185
            ++without;
            ++terminal;
        } else {
            if (!expand(P, t, lane, path[len - 1], s, order, P.eval_kind, nullptr, nullptr, ev)) break;
            if (P.eval_kind != DM_EVAL_NONE) ++evals;
            value = mix_rollout(P, s, ev, rng, plies);
            ++with;
        }
        backup(P, t, lane, path, len, value);
        --sims;
        ++done;
    }
    if (lane == 0) {
        P.sims_left[t] = dead ? 0 : sims;
        P.stat_with[t] = with;
        P.stat_without[t] = without;
        P.flags[t] = flags;
        P.rng_ctr[t] = rng.c0 + 1;
        count(P, kCntSims, done);
        count(P, kCntLeafEvals, evals);
        count(P, kCntTerminal, terminal);
        count(P, kCntDepth, depth);
        count(P, kCntScanned, scanned);
        count(P, kCntRolloutPlies, plies);
    }
}

// -------------------------------------------------- self-play move / game bookkeeping
// MCTS.self_play's move choice and re-rooting (mcts.py:96-111) and the example labelling
// of create_self_play_examples (train.py:267-283).  Whole warp.
__device__ __forceinline__ void finish_move(const PoolView& P, int t, int lane) {
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    u32 first = P.first_child[base + root];
    GState s = root_state(P, t);
    u32 ply = P.ply[t];
    if (n == 0) {
        // the root was never expanded (its expansion hit the arena's capacity, already counted): there is no
        // move to choose.  Abandon the game without recording it and start over instead of re-rooting on garbage.
        if (lane == 0) {
            fresh_root(P, t, game_initial(P.cfg));
            P.ply[t] = 0;
            arm_step(P, t);
        }
        __syncwarp();
        return;
    }
    // visit totals
    u32 v[3] = {0, 0, 0};
    int a[3] = {0, 0, 0};
    u32 total = 0;
    for (int k = 0, j = lane; j < n; j += 32, ++k) {
        v[k] = P.visits[base + first + j];
        a[k] = P.action[base + first + j];
        total += v[k];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(DM_FULL_MASK, total, off);
    // record (state, pi) (train.py:267-268)
    bool record = ply < DM_MAX_DEPTH;
    if (record) {
        size_t e = (size_t)t * DM_MAX_DEPTH + ply;
        float* pi = P.ex_pi + e * DM_MAX_ACTIONS;
        for (int j = lane; j < DM_MAX_ACTIONS; j += 32) pi[j] = 0.0f;
        __syncwarp();
        for (int k = 0, j = lane; j < n; j += 32, ++k)
            pi[a[k]] = total ? (float)((double)v[k] / (double)total) : 1.0f / (float)n;
        if (lane == 0) {
            P.ex_first[e] = s.first;
            P.ex_second[e] = s.second;
            P.ex_player[e] = (u8)s.player;
        }
    }
    // choose the move
    int chosen = -1;
    if ((int)ply < P.sample_move_cutoff && total > 0) {
        // np.random.choice(len(pi), p=pi): proportional to visit counts
        Philox rng;
        rng.init(P.seed, (u32)t, 0x6d6f7665u, P.rng_ctr[t]);
        u32 r = rng.below(total);
        // inverse CDF over the children in child order (any fixed order gives the same distribution;
        // the device stream is Philox, not numpy's, so only the distribution is comparable anyway)
        u32 run = 0;
        for (int k = 0, j0 = 0; j0 < n && chosen < 0; j0 += 32, ++k) {
            u32 vk = (j0 + lane < n) ? v[k] : 0u;
            u32 incl = vk;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                u32 up = __shfl_up_sync(DM_FULL_MASK, incl, off);
                if (lane >= off) incl += up;
            }
            bool pick = vk > 0 && r >= run + incl - vk && r < run + incl;
            u32 m = __ballot_sync(DM_FULL_MASK, pick);
            if (m) chosen = __shfl_sync(DM_FULL_MASK, a[k], __ffs(m) - 1);
            run += __shfl_sync(DM_FULL_MASK, incl, 31);
        }
        if (lane == 0) P.rng_ctr[t] += 1;
    }
    if (chosen < 0) {
        // np.argmax(pi): most visits, lowest action id on ties
        u32 bv = 0;
        int ba = 0x7fffffff;
        for (int k = 0, j = lane; j < n; j += 32, ++k)
            if (v[k] > bv || (v[k] == bv && a[k] < ba)) {
                bv = v[k];
                ba = a[k];
            }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            u32 ov = __shfl_xor_sync(DM_FULL_MASK, bv, off);
            int oa = __shfl_xor_sync(DM_FULL_MASK, ba, off);
            if (ov > bv || (ov == bv && oa < ba)) {
                bv = ov;
                ba = oa;
            }
        }
        chosen = ba;
    }
    // re-root with subtree reuse
    int child = -1;
    for (int k = 0, j = lane; j < n; j += 32, ++k)
        if (a[k] == chosen) child = j;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) child = max(child, __shfl_xor_sync(DM_FULL_MASK, child, off));
    GState ns = game_child(P.cfg, s, chosen);
    ++ply;
    float outcome;
    bool final = game_is_final(P.cfg, ns, outcome);
    if (final) {
        // label every recorded position with the outcome from its mover's perspective
        u32 cnt = min(ply, (u32)DM_MAX_DEPTH);
        // one atomic hands out the game's positions AND its game number, so games are numbered in ring order
        unsigned long long packed = 0;
        if (lane == 0) packed = atomicAdd(P.rp_written, (1ull << DM_REPLAY_POS_BITS) | (unsigned long long)cnt);
        packed = __shfl_sync(DM_FULL_MASK, packed, 0);
        const unsigned long long slot0 = packed & ((1ull << DM_REPLAY_POS_BITS) - 1ull);
        if (lane == 0) P.rp_game_end[(packed >> DM_REPLAY_POS_BITS) & (DM_REPLAY_GAME_RING - 1)] = slot0 + cnt;
        float z = outcome * 2.0f - 1.0f;  // train.py:270-272
        // The game's positions are one contiguous block in the per-tree record and (up to the ring's wrap-around)
        // in the replay ring.  This runs in the one or two warps per launch whose game just ended and sets the
        // launch's critical path, so it is written for memory-level parallelism: states / player / z one position
        // per lane, the visit distributions as a flat copy with eight loads in flight per lane.
        const size_t e0 = (size_t)t * DM_MAX_DEPTH;
        for (u32 i = lane; i < cnt; i += 32) {
            const size_t r = (size_t)((slot0 + i) % P.replay_cap);
            const u8 pl = P.ex_player[e0 + i];
            P.rp_first[r] = P.ex_first[e0 + i];
            P.rp_second[r] = P.ex_second[e0 + i];
            P.rp_player[r] = pl;
            P.rp_z[r] = (pl == 0) ? z : -z;  // train.py:278-280
        }
        const float* src = P.ex_pi + e0 * DM_MAX_ACTIONS;
        const u32 total_f = cnt * DM_MAX_ACTIONS;
        const size_t ring_f = (size_t)P.replay_cap * DM_MAX_ACTIONS;
        const size_t dst0 = (size_t)(slot0 % P.replay_cap) * DM_MAX_ACTIONS;
        for (u32 k0 = 0; k0 < total_f; k0 += 8 * 32) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const u32 k = k0 + (u32)u * 32 + (u32)lane;
                v[u] = k < total_f ? src[k] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const u32 k = k0 + (u32)u * 32 + (u32)lane;
                if (k < total_f) {
                    size_t d = dst0 + k;
                    while (d >= ring_f) d -= ring_f;    // the block wraps around the end of the ring (once, unless the ring is tiny)
                    P.rp_pi[d] = v[u];
                }
            }
        }
        if (lane == 0) {
            fresh_root(P, t, game_initial(P.cfg));
            P.ply[t] = 0;
            count(P, kCntGames, 1);
        }
    } else if (lane == 0) {
        P.root[t] = first + (u32)child;
        P.root_first[t] = ns.first;
        P.root_second[t] = ns.second;
        P.root_player[t] = (u8)ns.player;
        P.ply[t] = ply;
    }
    if (lane == 0) count(P, kCntMoves, 1);
    __syncwarp();
    compact_tree(P, t, lane);
    if (lane == 0) arm_step(P, t);
    __syncwarp();
}

// --------------------------------------------------- kernel: external-evaluator select
// 7 blocks x 4 warps = 28 warps per SM: 148 SMs hold 4144 trees in ONE wave (ncu showed 1.15
// waves at 80 registers, i.e. a nearly empty second pass over a latency-bound kernel).
constexpr int kLevelBudget = 32;

// kDense: the leaf of tree t always goes to slot t (the batch has n_trees slots, a tree that hands out no
// leaf leaves its previous one there and the evaluator's result for it is ignored).  Continuous
// self-play keeps every tree busy, so nothing is wasted and the per-leaf atomicAdd on ONE counter
// -- 4096 returning atomics on the same address, ~13 % of the kernel's stall samples -- goes away.
template <bool kSelfPlay, bool kDense = false>
__device__ __forceinline__ void select_tree(const PoolView& P, int t, int lane, u32* path, u32 now) {
    const size_t base = (size_t)t * P.cap;
    unsigned long long scanned = 0, depth = 0, terminal = 0, done = 0;
    if (kSelfPlay) {
        if (P.sims_left[t] <= 0) {
            if (P.flags[t] & kFlagStepActive) finish_move(P, t, lane);
            else {
                if (lane == 0) arm_step(P, t);
                __syncwarp();
            }
        }
    }
    int sims = P.sims_left[t];
    u8 flags = P.flags[t];
    u32 without = P.stat_without[t];
    bool emitted = false, live = false;
    int levels = 0;
    while (sims > 0) {
        u32 root = P.root[t];
        GState s;
        int len;
        u8 leaf_flag = 0;
        if (P.n_children[base + root] == 0) {
            // the root itself must be expanded first (mcts.py:114-116)
            s = root_state(P, t);
            if (lane == 0) path[0] = root;
            len = 1;
            leaf_flag = kLeafRoot;
            __syncwarp();
        } else {
            if (flags & kFlagNeedNoise) {
                apply_root_noise(P, t, lane);
                flags &= ~kFlagNeedNoise;
            }
            bool overflow;
            int final_code;
            len = descend(P, t, lane, s, path, scanned, overflow, final_code);
            if (overflow) {
                if (lane == 0) count(P, kCntCapacity, 1);
                sims = 0;
                break;
            }
            depth += (unsigned long long)(len - 1);
            if (final_code) {
                backup(P, t, lane, path, len, 0.5 * (double)(final_code - 1));
                ++without;
                ++terminal;
                ++done;
                --sims;
                // The launch lasts as long as its slowest warp.  A self-play tree that has already
                // walked kLevelBudget levels sits this round out instead of descending again (its
                // result does not depend on which round a simulation runs in); a stepped search
                // keeps going because its caller counts rounds.
                levels += len - 1;
                if (kSelfPlay && levels + (len - 1) > P.level_budget) break;
                continue;
            }
        }
        // hand the leaf to the evaluator -- unless the evaluation cache already knows the state (hit) or another
        // tree hands the same state out in this very round (alias)
        int slot = t;
        int centry = -1;
        if (P.cctl) {
            const CacheProbe probe = cache_probe(P, s, lane, now);
            if (probe.kind == kCacheHit) {
                if (P.cache_f64) cache_copy_hit<double>(P, t, probe.entry, lane);
                else cache_copy_hit<float>(P, t, probe.entry, lane);
                leaf_flag |= kLeafCached;
                live = false;
            } else if (probe.kind == kCacheAlias) {
                slot = probe.slot;
                leaf_flag |= kLeafAlias;
                live = false;
            } else {
                centry = probe.entry;
                live = true;
            }
        } else {
            live = true;
        }
        if (live && !kDense) {
            if (lane == 0) slot = atomicAdd(P.leaf_count, 1);
            slot = __shfl_sync(DM_FULL_MASK, slot, 0);
        }
        for (int i = lane; i < len; i += 32) P.path[(size_t)t * DM_MAX_DEPTH + i] = path[i];
        if (lane == 0) {
            if (centry >= 0) cache_publish(P, t, centry, s, now, slot);
            else if (P.cctl) P.pend_centry[t] = -1;
            P.path_len[t] = len;
            if (live) {
                P.leaf_first[slot] = s.first;
                P.leaf_second[slot] = s.second;
                P.leaf_player[slot] = (u8)s.player;
                P.leaf_flags[slot] = leaf_flag;
                P.leaf_tree[slot] = t;
            }
            P.pend_slot[t] = slot;
            P.pend_first[t] = s.first;
            P.pend_second[t] = s.second;
            P.pend_player[t] = (u8)s.player;
            P.pend_flags[t] = leaf_flag;
        }
        emitted = true;
        break;
    }
    if (lane == 0) {
        P.sims_left[t] = sims;
        P.stat_without[t] = without;
        P.flags[t] = flags;
        if (kDense) P.leaf_live[t] = live ? 1 : 0;
        if (!emitted) P.path_len[t] = 0;
        count(P, kCntSims, done);
        count(P, kCntTerminal, terminal);
        count(P, kCntDepth, depth);
        count(P, kCntScanned, scanned);
    }
}

template <bool kSelfPlay>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 7) k_select(PoolView Pin) {
    __shared__ u32 s_path[kWarpsPerBlock][DM_MAX_DEPTH];
    __shared__ u32 s_cnt[kWarpsPerBlock][16];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    PoolView P = Pin;
    P.scnt = s_cnt[warp];
    const u32 now = __ldcv(P.cround);
    if (t < P.n_trees) {
        counters_begin(P.scnt, lane);
        select_tree<kSelfPlay>(P, t, lane, s_path[warp], now);
        counters_flush(P, t, P.scnt, lane);
    }
    round_end(P, false);
}

// expand_node + rollout mix + backpropagate for one pending leaf (state s, evaluator output in slot `slot`)
template <typename PriorT>
__device__ __forceinline__ void expand_backup_tree(const PoolView& P, int t, int slot, int lane, const GState& s,
                                                   u8 leaf_flag, u8* order, const PriorT* priors, int stride,
                                                   const PriorT* values) {
    const size_t base = (size_t)t * P.cap;
    const int len = P.path_len[t];
    const u32* path = P.path + (size_t)t * DM_MAX_DEPTH;
    double ev;
    // evaluator output for this leaf: the tree's own slot, the slot of the tree it aliased in the same round, or
    // the row the selection copied out of the evaluation cache
    const bool cached = (leaf_flag & kLeafCached) != 0, alias = (leaf_flag & kLeafAlias) != 0;
    const PriorT* pr = cached ? (const PriorT*)P.hit_row + (size_t)t * kCacheRow : priors + (size_t)slot * stride;
    const PriorT leaf_value = cached ? pr[DM_MAX_ACTIONS] : values[slot];
    bool ok;
    if (sizeof(PriorT) == 8) ok = expand(P, t, lane, path[len - 1], s, order, DM_EVAL_EXTERNAL, nullptr, (const double*)pr, ev);
    else ok = expand(P, t, lane, path[len - 1], s, order, DM_EVAL_EXTERNAL, (const float*)pr, nullptr, ev);
    if (!ok) {
        if (lane == 0) {
            P.sims_left[t] = 0;
            P.path_len[t] = 0;
            if (P.cctl) P.pend_centry[t] = -1;
        }
        __syncwarp();
        return;
    }
    if (P.cctl && !cached && !alias) cache_finalize<PriorT>(P, t, lane, pr, leaf_value);
    unsigned long long plies = 0;
    if (leaf_flag & kLeafRoot) {
        if (lane == 0) P.visits[base + path[0]] += 1;  // root.visits += 1, value unused
        __syncwarp();
        u8 flags = P.flags[t];
        if (flags & kFlagNeedNoise) {
            apply_root_noise(P, t, lane);
            if (lane == 0) P.flags[t] = flags & ~kFlagNeedNoise;
        }
    } else {
        double value = (double)leaf_value;
        if (P.rollout_kind != DM_ROLLOUT_NONE) {
            Philox rng;
            rng.init(P.seed, (u32)t, 0x73696d00u, P.rng_ctr[t]);
            value = mix_rollout(P, s, value, rng, plies);
            if (lane == 0) P.rng_ctr[t] = rng.c0 + 1;
        }
        backup(P, t, lane, path, len, value);
        if (lane == 0) {
            P.sims_left[t] -= 1;
            P.stat_with[t] += 1;
            count(P, kCntSims, 1);
        }
    }
    if (lane == 0) {
        P.path_len[t] = 0;
        count(P, cached ? kCntCacheHits : (alias ? kCntCacheAliases : kCntLeafEvals), 1);
        count(P, kCntRolloutPlies, plies);
    }
    __syncwarp();
}

template <typename PriorT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    k_expand_backup(PoolView Pin, const PriorT* priors, int stride, const PriorT* values) {
    __shared__ u8 s_order[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    __shared__ u32 s_cnt[kWarpsPerBlock][16];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= Pin.n_trees || Pin.path_len[t] <= 0) return;
    PoolView P = Pin;
    P.scnt = s_cnt[warp];
    counters_begin(P.scnt, lane);
    // the per-tree copy of the pending leaf: a leaf served by the evaluation cache or aliased to another tree's
    // slot has no slot of its own in the batch
    GState s;
    s.first = P.pend_first[t];
    s.second = P.pend_second[t];
    s.player = P.pend_player[t];
    expand_backup_tree<PriorT>(P, t, P.pend_slot[t], lane, s, P.pend_flags[t], s_order[warp], priors, stride, values);
    counters_flush(P, t, P.scnt, lane);
}

// One launch per search round: expand + back up the tree's pending leaf with the evaluator's output,
// then go straight on to the next selection (mcts.py:119-138: the body of the simulation loop).
// The pending leaf is read from the per-tree copy, because other warps of this launch are already
// claiming slots of the next leaf batch; the evaluator outputs (indexed by the OLD slot) are not
// touched by this kernel.  Saves a kernel boundary per round and finds the path's nodes cache-hot.
// GAME >= 0: instantiated for one (game, board size) with the rules' masks and shifts as constants.
template <bool kSelfPlay, typename PriorT, int GAME, int G>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, 7)
    k_round(PoolView Pin, const PriorT* priors, int stride, const PriorT* values) {
    __shared__ u32 s_path[kWarpsPerBlock][DM_MAX_DEPTH];
    __shared__ u8 s_order[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    PoolView P = Pin;
    if (GAME >= 0) P.cfg = const_game_cfg<(GAME >= 0 ? GAME : 0), (G > 0 ? G : 8)>();
    __shared__ u32 s_cnt[kWarpsPerBlock][16];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    P.scnt = s_cnt[warp];
    const u32 now = __ldcv(P.cround);
    if (t < P.n_trees) {
        const long long t0 = P.dbg_warp ? clock64() : 0;
        counters_begin(P.scnt, lane);
        if (P.path_len[t] > 0) {
            GState s;
            s.first = P.pend_first[t];
            s.second = P.pend_second[t];
            s.player = P.pend_player[t];
            const int slot = P.pend_slot[t];
            const u8 leaf_flag = P.pend_flags[t];
            expand_backup_tree<PriorT>(P, t, slot, lane, s, leaf_flag, s_order[warp], priors, stride, values);
        }
        const long long t1 = P.dbg_warp ? clock64() : 0;
        select_tree<kSelfPlay, kSelfPlay>(P, t, lane, s_path[warp], now);
        if (P.dbg_warp && lane == 0) {   // diagnostics: cycles of the two phases, what the selection did this round
            const long long t2 = clock64();
            unsigned long long* d = P.dbg_warp + (size_t)t * 4;
            d[0] = (unsigned long long)(t1 - t0);
            d[1] = (unsigned long long)(t2 - t1);
            d[2] = ((unsigned long long)P.scnt[kCntMoves] << 32) | ((unsigned long long)P.scnt[kCntGames] << 24) |
                   ((unsigned long long)P.scnt[kCntTerminal] << 16) | (unsigned long long)P.scnt[kCntDepth];
            d[3] = (unsigned long long)P.scnt[kCntScanned];
        }
        counters_flush(P, t, P.scnt, lane);
    }
    round_end(P, kSelfPlay);   // dense leaf batch (self-play): the last block compacts the live slots
}

template <bool kSelfPlay, typename PriorT>
static void launch_round(const PoolView& v, int blocks, int threads, cudaStream_t st, const PriorT* priors, int stride,
                         const PriorT* values) {
    const int game = v.cfg.game, g = v.cfg.g;
    if (game == DM_OTHELLO && g == 8)
        k_round<kSelfPlay, PriorT, DM_OTHELLO, 8><<<blocks, threads, 0, st>>>(v, priors, stride, values);
    else if (game == DM_HEX && g == 7)
        k_round<kSelfPlay, PriorT, DM_HEX, 7><<<blocks, threads, 0, st>>>(v, priors, stride, values);
    else if (game == DM_HEX_SWAP && g == 7)
        k_round<kSelfPlay, PriorT, DM_HEX_SWAP, 7><<<blocks, threads, 0, st>>>(v, priors, stride, values);
    else
        k_round<kSelfPlay, PriorT, -1, 0><<<blocks, threads, 0, st>>>(v, priors, stride, values);
}

// ---------------------------------------------------------------- virtual-loss mode
// Throughput mode for small pools: a tree hands out up to leaves_per_round leaves per round.
// After each descent the path gets a virtual loss -- one extra visit scored as a loss for the
// player who chose the node (the same value the reference gives an unvisited child,
// mcts.py:48-50) -- so the next descent of the same round explores elsewhere.  A leaf that is
// already pending is marked (first_child = kPendingLeaf); reaching it again ends the round for
// that tree.  k_expand_backup_vl removes the virtual losses while backing up the real values.
// Not bit-identical to the reference (simulations are reordered); W = 1 uses the exact kernels.
constexpr u32 kPendingLeaf = 0xffffffffu;

__device__ __forceinline__ double loss_value_for_parent(int root_player, int depth_of_child) {
    // the node at depth d was chosen by the player to move at depth d-1; every action flips the mover
    int parent_player = root_player ^ ((depth_of_child - 1) & 1);
    return parent_player == 1 ? 1.0 : 0.0;  // FIRST (min player) loses at 1.0, SECOND at 0.0
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) k_select_vl(PoolView P) {
    __shared__ u32 s_path[kWarpsPerBlock][DM_MAX_DEPTH];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    u32* path = s_path[warp];
    const int W = P.leaves_per_round;
    unsigned long long scanned = 0, depth = 0, terminal = 0, done = 0;
    int sims = P.sims_left[t];
    u8 flags = P.flags[t];
    u32 without = P.stat_without[t];
    const int root_player = P.root_player[t];
    int pending = 0;
    while (pending < W && pending < sims) {
        u32 root = P.root[t];
        GState s;
        int len;
        u8 leaf_flag = 0;
        if (P.n_children[base + root] == 0) {
            if (pending > 0) break;
            s = root_state(P, t);
            if (lane == 0) path[0] = root;
            len = 1;
            leaf_flag = kLeafRoot;
            __syncwarp();
        } else {
            if (flags & kFlagNeedNoise) {
                apply_root_noise(P, t, lane);
                flags &= ~kFlagNeedNoise;
            }
            bool overflow;
            int final_code;
            len = descend(P, t, lane, s, path, scanned, overflow, final_code);
            if (overflow) {
                if (lane == 0) count(P, kCntCapacity, 1);
                sims = 0;
                break;
            }
            u32 leaf = path[len - 1];
            if (P.first_child[base + leaf] == kPendingLeaf) break;  // collision with a leaf of this round
            depth += (unsigned long long)(len - 1);
            if (final_code) {
                backup(P, t, lane, path, len, 0.5 * (double)(final_code - 1));
                ++without;
                ++terminal;
                ++done;
                --sims;
                continue;
            }
            // virtual loss along the path, mark the leaf
            for (int i = lane; i < len; i += 32) {
                size_t idx = base + path[i];
                P.visits[idx] += 1;
                if (i > 0) P.value_sum[idx] = __dadd_rn(P.value_sum[idx], loss_value_for_parent(root_player, i));
            }
            if (lane == 0) P.first_child[base + leaf] = kPendingLeaf;
            __syncwarp();
        }
        int slot = 0;
        if (lane == 0) slot = atomicAdd(P.leaf_count, 1);
        slot = __shfl_sync(DM_FULL_MASK, slot, 0);
        u32* dst = P.vl_path + ((size_t)t * DM_MAX_LEAVES + pending) * DM_MAX_DEPTH;
        for (int i = lane; i < len; i += 32) dst[i] = path[i];
        if (lane == 0) {
            P.vl_len[t * DM_MAX_LEAVES + pending] = len;
            P.vl_slot[t * DM_MAX_LEAVES + pending] = slot;
            P.leaf_first[slot] = s.first;
            P.leaf_second[slot] = s.second;
            P.leaf_player[slot] = (u8)s.player;
            P.leaf_flags[slot] = leaf_flag;
            P.leaf_tree[slot] = t;
        }
        ++pending;
        if (leaf_flag & kLeafRoot) break;
        __syncwarp();
    }
    if (lane == 0) {
        P.sims_left[t] = sims;
        P.stat_without[t] = without;
        P.flags[t] = flags;
        P.vl_count[t] = pending;
        if (sims > 0) atomicAdd(P.armed, 1);
        count(P, kCntSims, done);
        count(P, kCntTerminal, terminal);
        count(P, kCntDepth, depth);
        count(P, kCntScanned, scanned);
    }
}

template <typename PriorT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
    k_expand_backup_vl(PoolView P, const PriorT* priors, int stride, const PriorT* values) {
    __shared__ u8 s_order[kWarpsPerBlock][DM_MAX_ACTIONS + 3];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    const int pending = P.vl_count[t];
    const int root_player = P.root_player[t];
    int completed = 0;
    for (int w = 0; w < pending; ++w) {
        const int slot = P.vl_slot[t * DM_MAX_LEAVES + w];
        const int len = P.vl_len[t * DM_MAX_LEAVES + w];
        const u32* path = P.vl_path + ((size_t)t * DM_MAX_LEAVES + w) * DM_MAX_DEPTH;
        GState s;
        s.first = P.leaf_first[slot];
        s.second = P.leaf_second[slot];
        s.player = P.leaf_player[slot];
        const u8 leaf_flag = P.leaf_flags[slot];
        const u32 leaf = path[len - 1];
        if (lane == 0 && !(leaf_flag & kLeafRoot)) P.first_child[base + leaf] = 0;  // clear the pending mark
        __syncwarp();
        double ev;
        const PriorT* pr = priors + (size_t)slot * stride;
        bool ok;
        if (sizeof(PriorT) == 8) ok = expand(P, t, lane, leaf, s, s_order[warp], DM_EVAL_EXTERNAL, nullptr,
                                             (const double*)pr, ev);
        else ok = expand(P, t, lane, leaf, s, s_order[warp], DM_EVAL_EXTERNAL, (const float*)pr, nullptr, ev);
        if (!ok) {
            if (lane == 0) P.sims_left[t] = 0;
            break;
        }
        if (leaf_flag & kLeafRoot) {
            if (lane == 0) P.visits[base + path[0]] += 1;
            __syncwarp();
            u8 flags = P.flags[t];
            if (flags & kFlagNeedNoise) {
                apply_root_noise(P, t, lane);
                if (lane == 0) P.flags[t] = flags & ~kFlagNeedNoise;
            }
        } else {
            const double value = (double)values[slot];
            // the visit was already counted by the virtual loss; swap the loss for the real value
            for (int i = lane; i < len; i += 32) {
                size_t idx = base + path[i];
                double delta = (i > 0) ? __dsub_rn(value, loss_value_for_parent(root_player, i)) : value;
                P.value_sum[idx] = __dadd_rn(P.value_sum[idx], delta);
            }
            __syncwarp();
            ++completed;
        }
    }
    if (lane == 0) {
        P.sims_left[t] -= completed;
        P.stat_with[t] += (u32)completed;
        P.vl_count[t] = 0;
        count(P, kCntSims, (unsigned long long)completed);
        count(P, kCntLeafEvals, (unsigned long long)pending);
    }
}

// the DM_EVAL_HASHED evaluator as an EXTERNAL one (double outputs): lets the parity tests drive the
// leaf-exchange kernels, the dense/compacted leaf batch and the evaluation cache with an evaluator whose
// reference results are pinned (tests/golden/mcts_digests.json)
__global__ void k_eval_hashed(GameCfg cfg, int n, const int* count_dev, const int* index, const u64* first,
                              const u64* second, const u8* player, u64 salt, double* value, double* probs) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int live = count_dev ? min(*count_dev, n) : n;
    if (i >= live) return;
    const int src = index ? index[i] : i;
    GState s;
    s.first = first[src];
    s.second = second[src];
    s.player = player[src];
    bool extra;
    const u64 m = game_legal_mask(cfg, s, extra);
    const u64 h0 = hashed_h0(s, salt);
    u32 sum = 0;
    for (int a = 0; a < cfg.num_actions; ++a)
        if (a < cfg.cells ? ((m >> a) & 1) : extra) sum += hashed_weight(h0, a);
    double* row = probs + (size_t)src * DM_MAX_ACTIONS;
    for (int a = 0; a < DM_MAX_ACTIONS; ++a) {
        const bool legal = a < cfg.cells ? ((m >> a) & 1) : (a == cfg.cells && extra);
        row[a] = legal ? __ddiv_rn((double)hashed_weight(h0, a), (double)sum) : 0.0;
    }
    value[src] = __ddiv_rn((double)((h0 >> 20) % 1001ull), 1000.0);
}

// ------------------------------------------------------------------- kernels: read-out
__global__ void k_root_visits(PoolView P, u32* visits_out, u32* stats_out) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    u32 first = P.first_child[base + root];
    if (visits_out) {
        for (int j = lane; j < DM_MAX_ACTIONS; j += 32) visits_out[(size_t)t * DM_MAX_ACTIONS + j] = 0;
        __syncwarp();
        for (int j = lane; j < n; j += 32)
            visits_out[(size_t)t * DM_MAX_ACTIONS + P.action[base + first + j]] = P.visits[base + first + j];
    }
    if (stats_out && lane == 0) {
        stats_out[2 * t] = P.stat_with[t];
        stats_out[2 * t + 1] = P.stat_without[t];
    }
}

__global__ void k_root_children(PoolView P, u8* order_out, int* count_out, u32* root_visits_out) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    u32 first = P.first_child[base + root];
    if (order_out)
        for (int j = lane; j < DM_MAX_ACTIONS; j += 32)
            order_out[(size_t)t * DM_MAX_ACTIONS + j] = j < n ? P.action[base + first + j] : 0xff;
    if (lane == 0) {
        if (count_out) count_out[t] = n;
        if (root_visits_out) root_visits_out[t] = P.visits[base + root];
    }
}

// Node.probability and Node.value_sum of the root's children, in child (dict) order
__global__ void k_root_child_stats(PoolView P, double* prior_out, double* value_sum_out) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    u32 first = P.first_child[base + root];
    for (int j = lane; j < DM_MAX_ACTIONS; j += 32) {
        if (prior_out) prior_out[(size_t)t * DM_MAX_ACTIONS + j] = j < n ? P.prior[base + first + j] : 0.0;
        if (value_sum_out) value_sum_out[(size_t)t * DM_MAX_ACTIONS + j] = j < n ? P.value_sum[base + first + j] : 0.0;
    }
}

__global__ void k_advance(PoolView P, const int* actions) {
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int t = blockIdx.x * kWarpsPerBlock + warp;
    if (t >= P.n_trees) return;
    int a = actions[t];
    if (a < 0) return;
    const size_t base = (size_t)t * P.cap;
    u32 root = P.root[t];
    int n = P.n_children[base + root];
    u32 first = P.first_child[base + root];
    int child = -1;
    for (int j = lane; j < n; j += 32)
        if (P.action[base + first + j] == a) child = j;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) child = max(child, __shfl_xor_sync(DM_FULL_MASK, child, off));
    GState s = root_state(P, t);
    bool moved = true;
    if (child < 0) {
        // root not expanded (or action illegal): follow the reference's Node() fallback
        // only when the action is legal; otherwise flag it
        if (!game_action_legal(P.cfg, s, a)) {
            if (lane == 0) count(P, kCntInvalid, 1);
            moved = false;
        } else if (lane == 0) {
            fresh_root(P, t, game_child(P.cfg, s, a));
        }
    } else if (lane == 0) {
        GState ns = game_child(P.cfg, s, a);
        P.root[t] = first + (u32)child;
        P.root_first[t] = ns.first;
        P.root_second[t] = ns.second;
        P.root_player[t] = (u8)ns.player;
    }
    if (!moved) return;
    if (lane == 0) {
        P.ply[t] += 1;
        P.sims_left[t] = 0;
        P.flags[t] &= ~(kFlagNeedNoise | kFlagStepActive);
    }
    __syncwarp();
    compact_tree(P, t, lane);
}

// counters read-out: global row + the per-tree rows (sum; high-water mark: max) -> out[16]
__global__ void k_sum_counters(PoolView P, unsigned long long* out) {
    __shared__ unsigned long long s_part[32][16];
    const int col = threadIdx.x & 15, grp = threadIdx.x >> 4;   // 512 threads = 32 groups of 16 columns
    unsigned long long acc = 0;
    for (int t = grp; t < P.n_trees; t += 32) {
        const unsigned long long v = P.tcount[(size_t)t * 16 + col];
        acc = (col == kCntNodesHigh) ? max(acc, v) : acc + v;
    }
    s_part[grp][col] = acc;
    __syncthreads();
    if (threadIdx.x < 16) {
        unsigned long long total = P.counters[col];
        for (int g2 = 0; g2 < 32; ++g2)
            total = (col == kCntNodesHigh) ? max(total, s_part[g2][col]) : total + s_part[g2][col];
        out[col] = total;
    }
}

__global__ void k_states(PoolView P, u64* first, u64* second, u8* player, u8* is_final, float* outcome) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P.n_trees) return;
    GState s = root_state(P, t);
    if (first) first[t] = s.first;
    if (second) second[t] = s.second;
    if (player) player[t] = (u8)s.player;
    if (is_final || outcome) {
        float o;
        bool f = game_is_final(P.cfg, s, o);
        if (is_final) is_final[t] = f ? 1 : 0;
        if (outcome) outcome[t] = o;
    }
}

}  // namespace

// =========================================================================== host side
struct dm_pool {
    PoolView v;
    dm_pool_config cfg;
    std::vector<void*> allocations;
    bool configured;
    uint64_t capacity_errors_reported = 0;
};

namespace {

template <typename T>
int pool_alloc(dm_pool* p, T** out, size_t count, bool zero = true) {
    void* ptr = nullptr;
    DM_CUDA(cudaMalloc(&ptr, count * sizeof(T)));
    p->allocations.push_back(ptr);
    if (zero) DM_CUDA(cudaMemset(ptr, 0, count * sizeof(T)));
    *out = (T*)ptr;
    return DM_OK;
}

inline int tree_blocks(int n) { return (n + kWarpsPerBlock - 1) / kWarpsPerBlock; }

}  // namespace

#define DM_TRY(expr)                 \
    do {                             \
        int rc__ = (expr);           \
        if (rc__ != DM_OK) return rc__; \
    } while (0)

extern "C" int dm_pool_create(const dm_pool_config* cfg, dm_pool** out) {
    DM_CHECK_ARG(cfg && out, "dm_pool_create: null argument");
    DM_CHECK_ARG(game_cfg_valid(cfg->game, cfg->grid), "dm_pool_create: unsupported game/grid");
    DM_CHECK_ARG(cfg->n_trees > 0 && cfg->max_nodes_per_tree >= 2, "dm_pool_create: bad sizes");
    DM_CUDA(cudaSetDevice(cfg->device));
    dm_pool* p = new (std::nothrow) dm_pool();
    DM_CHECK_ARG(p, "dm_pool_create: out of host memory");
    p->cfg = *cfg;
    p->configured = false;
    PoolView& v = p->v;
    memset(&v, 0, sizeof(v));
    v.cfg = make_game_cfg(cfg->game, cfg->grid);
    v.n_trees = cfg->n_trees;
    v.cap = (u32)cfg->max_nodes_per_tree;
    v.c_puct = cfg->c_puct;
    v.seed = cfg->seed;
    v.num_simulations = 0;
    v.eval_kind = DM_EVAL_EXTERNAL;
    v.rollout_kind = DM_ROLLOUT_NONE;
    v.dirichlet_factor = 0.25;
    v.rollout_share = 1.0;
    v.level_budget = kLevelBudget;
    if (const char* env = getenv("DM_LEVEL_BUDGET")) v.level_budget = atoi(env) > 0 ? atoi(env) : kLevelBudget;   // A/B switch
    size_t T = (size_t)cfg->n_trees, N = T * (size_t)v.cap;
    int rc = DM_OK;
    auto fail = [&](int code) {
        dm_pool_destroy(p);
        return code;
    };
    // node arrays are not zeroed: fresh_root / expand initialise every slot they publish
    if ((rc = pool_alloc(p, &v.visits, N, false)) || (rc = pool_alloc(p, &v.value_sum, N, false)) ||
        (rc = pool_alloc(p, &v.prior, N, false)) || (rc = pool_alloc(p, &v.first_child, N, false)) ||
        (rc = pool_alloc(p, &v.n_children, N, false)) || (rc = pool_alloc(p, &v.action, N, false)) ||
        (rc = pool_alloc(p, &v.node_first, N, false)) || (rc = pool_alloc(p, &v.node_second, N, false)) ||
        (rc = pool_alloc(p, &v.node_final, N, false)) ||
        (rc = pool_alloc(p, &v.node_count, T)) || (rc = pool_alloc(p, &v.root, T)) ||
        (rc = pool_alloc(p, &v.root_first, T)) || (rc = pool_alloc(p, &v.root_second, T)) ||
        (rc = pool_alloc(p, &v.root_player, T)) || (rc = pool_alloc(p, &v.sims_left, T)) ||
        (rc = pool_alloc(p, &v.stat_with, T)) || (rc = pool_alloc(p, &v.stat_without, T)) ||
        (rc = pool_alloc(p, &v.flags, T)) || (rc = pool_alloc(p, &v.ply, T)) ||
        (rc = pool_alloc(p, &v.rng_ctr, T)) || (rc = pool_alloc(p, &v.path, T * DM_MAX_DEPTH)) ||
        (rc = pool_alloc(p, &v.path_len, T)) || (rc = pool_alloc(p, &v.leaf_first, T * DM_MAX_LEAVES)) ||
        (rc = pool_alloc(p, &v.leaf_second, T * DM_MAX_LEAVES)) || (rc = pool_alloc(p, &v.leaf_player, T * DM_MAX_LEAVES)) ||
        (rc = pool_alloc(p, &v.leaf_flags, T * DM_MAX_LEAVES)) || (rc = pool_alloc(p, &v.leaf_tree, T * DM_MAX_LEAVES)) ||
        (rc = pool_alloc(p, &v.pend_slot, T)) || (rc = pool_alloc(p, &v.pend_first, T)) ||
        (rc = pool_alloc(p, &v.pend_second, T)) || (rc = pool_alloc(p, &v.pend_player, T)) ||
        (rc = pool_alloc(p, &v.pend_flags, T)) ||
        (rc = pool_alloc(p, &v.vl_len, T * DM_MAX_LEAVES)) || (rc = pool_alloc(p, &v.vl_slot, T * DM_MAX_LEAVES)) ||
        (rc = pool_alloc(p, &v.vl_count, T)) || (rc = pool_alloc(p, &v.armed, 1)) ||
        (rc = pool_alloc(p, &v.leaf_count, 1)) || (rc = pool_alloc(p, &v.counters, 32)) ||
        (rc = pool_alloc(p, &v.tcount, T * 16)) ||
        (rc = pool_alloc(p, &v.cround, 2)) || (rc = pool_alloc(p, &v.leaf_live, T)) ||
        (rc = pool_alloc(p, &v.leaf_index, T)))
        return fail(rc);
    if (cfg->compaction) {
        v.cap_words = (v.cap + 31) / 32;
        if ((rc = pool_alloc(p, &v.cmark, T * v.cap_words, false)) || (rc = pool_alloc(p, &v.cprefix, T * v.cap_words, false)))
            return fail(rc);
    }
    k_reset<<<(cfg->n_trees + 127) / 128, 128>>>(v, nullptr, cfg->n_trees);
    ++g_dm_launches;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        dm_set_error(std::string("dm_pool_create: ") + cudaGetErrorString(e));
        return fail(DM_ERR_CUDA);
    }
    *out = p;
    return DM_OK;
}

extern "C" int dm_pool_destroy(dm_pool* p) {
    if (!p) return DM_OK;
    cudaSetDevice(p->cfg.device);
    for (void* ptr : p->allocations) cudaFree(ptr);
    delete p;
    return DM_OK;
}

extern "C" int dm_pool_configure(dm_pool* p, const dm_search_config* sc) {
    DM_CHECK_ARG(p && sc, "dm_pool_configure: null argument");
    DM_CHECK_ARG(sc->num_simulations >= 0, "dm_pool_configure: num_simulations < 0");
    DM_CHECK_ARG(sc->eval_kind >= DM_EVAL_NONE && sc->eval_kind <= DM_EVAL_EXTERNAL, "dm_pool_configure: eval_kind");
    DM_CHECK_ARG(sc->rollout_kind >= DM_ROLLOUT_NONE && sc->rollout_kind <= DM_ROLLOUT_MAX_ACTION,
                 "dm_pool_configure: rollout_kind");
    // mcts.py:82-83
    DM_CHECK_ARG(!(sc->eval_kind == DM_EVAL_NONE && sc->rollout_kind == DM_ROLLOUT_NONE),
                 "Both rollout_policy and state_evaluator cannot be None");
    PoolView& v = p->v;
    v.num_simulations = sc->num_simulations;
    v.eval_kind = sc->eval_kind;
    v.rollout_kind = sc->rollout_kind;
    v.sample_move_cutoff = sc->sample_move_cutoff;
    v.dirichlet_alpha = sc->dirichlet_alpha;
    v.dirichlet_factor = sc->dirichlet_factor;
    v.rollout_share = sc->rollout_share;
    v.eval_salt = sc->eval_salt;
    DM_CHECK_ARG(sc->leaves_per_round >= 0 && sc->leaves_per_round <= DM_MAX_LEAVES,
                 "dm_pool_configure: leaves_per_round out of range");
    v.leaves_per_round = sc->leaves_per_round > 1 ? sc->leaves_per_round : 1;
    if (v.leaves_per_round > 1 && !v.vl_path) {
        DM_CUDA(cudaSetDevice(p->cfg.device));
        DM_TRY(pool_alloc(p, &v.vl_path, (size_t)v.n_trees * DM_MAX_LEAVES * DM_MAX_DEPTH, false));
    }
    p->configured = true;
    return DM_OK;
}

extern "C" int dm_pool_reset(dm_pool* p, const int32_t* ids, int n, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_reset: null pool");
    if (!ids) n = p->v.n_trees;
    DM_CHECK_ARG(n >= 0, "dm_pool_reset: n < 0");
    if (n == 0) return DM_OK;
    k_reset<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v, ids, n);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_set_root_state(dm_pool* p, const int32_t* ids, int n, const uint64_t* first,
                                      const uint64_t* second, const uint8_t* player, void* stream) {
    DM_CHECK_ARG(p && first && second && player, "dm_pool_set_root_state: null argument");
    if (!ids) n = p->v.n_trees;
    if (n <= 0) return DM_OK;
    k_set_root<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v, ids, n, (const u64*)first,
                                                                  (const u64*)second, player);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_observe(dm_pool* p, const int32_t* ids, int n, const uint64_t* first,
                               const uint64_t* second, const uint8_t* player, void* stream) {
    DM_CHECK_ARG(p && first && second && player, "dm_pool_observe: null argument");
    if (!ids) n = p->v.n_trees;
    if (n <= 0) return DM_OK;
    k_observe<<<tree_blocks(n), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v, ids, n, (const u64*)first,
                                                                                (const u64*)second, player);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_begin_step(dm_pool* p, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_begin_step: null pool");
    if (!p->configured) {
        dm_set_error("dm_pool_begin_step: dm_pool_configure has not been called");
        return DM_ERR_STATE;
    }
    k_begin_step<<<(p->v.n_trees + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_begin_step_ids(dm_pool* p, const int32_t* ids, int n, void* stream) {
    DM_CHECK_ARG(p && n >= 0 && (ids || n == 0), "dm_pool_begin_step_ids: bad arguments");
    if (!p->configured) {
        dm_set_error("dm_pool_begin_step_ids: dm_pool_configure has not been called");
        return DM_ERR_STATE;
    }
    k_begin_step_ids<<<(p->v.n_trees + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v, ids, n);
    DM_LAUNCH_CHECK();
    if (n > 0) {
        k_arm_ids<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v, ids, n);
        DM_LAUNCH_CHECK();
    }
    return DM_OK;
}

extern "C" int dm_pool_set_noise(dm_pool* p, const double* noise_dev) {
    DM_CHECK_ARG(p, "dm_pool_set_noise: null pool");
    p->v.host_noise = noise_dev;
    return DM_OK;
}

extern "C" int dm_pool_search_local(dm_pool* p, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_search_local: null pool");
    if (!p->configured || p->v.eval_kind == DM_EVAL_EXTERNAL) {
        dm_set_error("dm_pool_search_local: needs eval_kind NONE / UNIFORM / HASHED");
        return DM_ERR_STATE;
    }
    k_search_local<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

static int launch_select(dm_pool* p, void* stream, bool selfplay) {
    if (!p->configured || p->v.eval_kind != DM_EVAL_EXTERNAL) {
        dm_set_error("dm_pool_select: needs eval_kind EXTERNAL");
        return DM_ERR_STATE;
    }
    DM_CUDA(cudaMemsetAsync(p->v.leaf_count, 0, sizeof(int), (cudaStream_t)stream));
    if (p->v.leaves_per_round > 1) {
        if (selfplay) {
            dm_set_error("virtual-loss mode is not available for continuous self-play (use leaves_per_round = 1)");
            return DM_ERR_STATE;
        }
        DM_CUDA(cudaMemsetAsync(p->v.armed, 0, sizeof(int), (cudaStream_t)stream));
        k_select_vl<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v);
        DM_LAUNCH_CHECK();
        return DM_OK;
    }
    if (selfplay)
        k_select<true><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v);
    else
        k_select<false><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_select(dm_pool* p, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_select: null pool");
    return launch_select(p, stream, false);
}

extern "C" int dm_pool_cache_enable(dm_pool* p, int log2_entries, int is_f64) {
    DM_CHECK_ARG(p, "dm_pool_cache_enable: null pool");
    DM_CHECK_ARG(log2_entries >= 10 && log2_entries <= 28, "dm_pool_cache_enable: log2_entries out of range [10, 28]");
    PoolView& v = p->v;
    if (v.cctl) {
        DM_CHECK_ARG(v.cmask == (1u << log2_entries) - 1u && v.cache_f64 == (is_f64 ? 1 : 0),
                     "dm_pool_cache_enable: the cache already exists with another size or type");
        return DM_OK;
    }
    DM_CUDA(cudaSetDevice(p->cfg.device));
    DM_CUDA(cudaDeviceSynchronize());
    const size_t E = (size_t)1 << log2_entries, T = (size_t)v.n_trees, elem = is_f64 ? 8 : 4;
    unsigned long long* ctl = nullptr;
    char *rows = nullptr, *hits = nullptr;
    DM_TRY(pool_alloc(p, &ctl, E));
    DM_TRY(pool_alloc(p, &v.ckey_first, E, false));
    DM_TRY(pool_alloc(p, &v.ckey_second, E, false));
    DM_TRY(pool_alloc(p, &rows, E * kCacheRow * elem, false));
    DM_TRY(pool_alloc(p, &hits, T * kCacheRow * elem));
    DM_TRY(pool_alloc(p, &v.pend_centry, T, false));
    DM_TRY(pool_alloc(p, &v.pend_cctl, T));
    DM_CUDA(cudaMemset(v.pend_centry, 0xff, T * sizeof(int)));
    v.crow = rows;
    v.hit_row = hits;
    v.cmask = (u32)(E - 1);
    v.cache_f64 = is_f64 ? 1 : 0;
    v.cctl = ctl;   // last: a non-null cctl switches the cache on in the kernels
    return DM_OK;
}

extern "C" int dm_pool_cache_clear(dm_pool* p, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_cache_clear: null pool");
    if (!p->v.cctl) return DM_OK;
    // entries that are pending right now are dropped as well: their owners find another ctl word and skip the store
    DM_CUDA(cudaMemsetAsync(p->v.cctl, 0, ((size_t)p->v.cmask + 1) * sizeof(unsigned long long), (cudaStream_t)stream));
    return DM_OK;
}

extern "C" int dm_pool_debug_timing(dm_pool* p, uint64_t** per_tree) {
    DM_CHECK_ARG(p && per_tree, "dm_pool_debug_timing: null argument");
    if (!p->v.dbg_warp) {
        DM_CUDA(cudaSetDevice(p->cfg.device));
        DM_CUDA(cudaDeviceSynchronize());
        DM_TRY(pool_alloc(p, &p->v.dbg_warp, (size_t)p->v.n_trees * 4));
    }
    *per_tree = (uint64_t*)p->v.dbg_warp;
    return DM_OK;
}

extern "C" int dm_pool_leaf_index(dm_pool* p, int32_t** index) {
    DM_CHECK_ARG(p && index, "dm_pool_leaf_index: null argument");
    *index = p->v.leaf_index;
    return DM_OK;
}

extern "C" int dm_eval_hashed(int game, int grid, int n, const int32_t* count_dev, const int32_t* index_dev,
                              const uint64_t* first, const uint64_t* second, const uint8_t* player, uint64_t salt,
                              double* value, double* probs, void* stream) {
    DM_CHECK_ARG(game_cfg_valid(game, grid), "dm_eval_hashed: unsupported game/grid");
    DM_CHECK_ARG(n >= 0, "negative batch size");
    if (n == 0) return DM_OK;
    DM_CHECK_ARG(first && second && player && value && probs, "dm_eval_hashed: bad arguments");
    k_eval_hashed<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(make_game_cfg(game, grid), n, count_dev, index_dev,
                                                                     (const u64*)first, (const u64*)second, player,
                                                                     (u64)salt, value, probs);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_leaf_buffers(dm_pool* p, uint64_t** first, uint64_t** second, uint8_t** player,
                                    int32_t** tree, int32_t** count) {
    DM_CHECK_ARG(p, "dm_pool_leaf_buffers: null pool");
    if (first) *first = (uint64_t*)p->v.leaf_first;
    if (second) *second = (uint64_t*)p->v.leaf_second;
    if (player) *player = p->v.leaf_player;
    if (tree) *tree = p->v.leaf_tree;
    if (count) *count = p->v.leaf_count;
    return DM_OK;
}

extern "C" int dm_pool_expand_backup(dm_pool* p, const void* priors, int stride, const void* values, int is_f64,
                                     void* stream) {
    DM_CHECK_ARG(p && priors && values, "dm_pool_expand_backup: null argument");
    DM_CHECK_ARG(stride >= p->v.cfg.num_actions, "dm_pool_expand_backup: prior_stride < num_actions");
    DM_CHECK_ARG(!p->v.cctl || p->v.leaves_per_round > 1 || p->v.cache_f64 == (is_f64 ? 1 : 0),
                 "dm_pool_expand_backup: evaluator output type differs from the evaluation cache's");
    if (p->v.leaves_per_round > 1) {
        if (is_f64)
            k_expand_backup_vl<double><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
                p->v, (const double*)priors, stride, (const double*)values);
        else
            k_expand_backup_vl<float><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
                p->v, (const float*)priors, stride, (const float*)values);
        DM_LAUNCH_CHECK();
        return DM_OK;
    }
    if (is_f64)
        k_expand_backup<double><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
            p->v, (const double*)priors, stride, (const double*)values);
    else
        k_expand_backup<float><<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
            p->v, (const float*)priors, stride, (const float*)values);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_round(dm_pool* p, const void* priors, int stride, const void* values, int is_f64, int selfplay,
                             void* stream) {
    DM_CHECK_ARG(p && priors && values, "dm_pool_round: null argument");
    DM_CHECK_ARG(stride >= p->v.cfg.num_actions, "dm_pool_round: prior_stride < num_actions");
    DM_CHECK_ARG(!p->v.cctl || p->v.cache_f64 == (is_f64 ? 1 : 0),
                 "dm_pool_round: evaluator output type differs from the evaluation cache's");
    if (!p->configured || p->v.eval_kind != DM_EVAL_EXTERNAL || p->v.leaves_per_round != 1) {
        dm_set_error("dm_pool_round: needs eval_kind EXTERNAL and leaves_per_round 1");
        return DM_ERR_STATE;
    }
    if (selfplay && !p->v.selfplay) {
        dm_set_error("dm_pool_round: self-play recording is not enabled on this pool");
        return DM_ERR_STATE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!selfplay) DM_CUDA(cudaMemsetAsync(p->v.leaf_count, 0, sizeof(int), st));   // self-play: dense batch, count = n_trees
    const int blocks = tree_blocks(p->v.n_trees), threads = kWarpsPerBlock * 32;
    if (is_f64) {
        if (selfplay) launch_round<true, double>(p->v, blocks, threads, st, (const double*)priors, stride, (const double*)values);
        else launch_round<false, double>(p->v, blocks, threads, st, (const double*)priors, stride, (const double*)values);
    } else {
        if (selfplay) launch_round<true, float>(p->v, blocks, threads, st, (const float*)priors, stride, (const float*)values);
        else launch_round<false, float>(p->v, blocks, threads, st, (const float*)priors, stride, (const float*)values);
    }
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_root_visits(dm_pool* p, uint32_t* visits, uint32_t* stats, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_root_visits: null pool");
    k_root_visits<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v, visits, stats);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_root_children(dm_pool* p, uint8_t* order, int32_t* count_out, uint32_t* root_visits,
                                     void* stream) {
    DM_CHECK_ARG(p, "dm_pool_root_children: null pool");
    k_root_children<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v, order, count_out,
                                                                                                 root_visits);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_root_child_stats(dm_pool* p, double* prior, double* value_sum, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_root_child_stats: null pool");
    k_root_child_stats<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v, prior, value_sum);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_advance(dm_pool* p, const int32_t* actions, void* stream) {
    DM_CHECK_ARG(p && actions, "dm_pool_advance: null argument");
    k_advance<<<tree_blocks(p->v.n_trees), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(p->v, actions);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_states(dm_pool* p, uint64_t* first, uint64_t* second, uint8_t* player, uint8_t* is_final,
                              float* outcome, void* stream) {
    DM_CHECK_ARG(p, "dm_pool_states: null pool");
    k_states<<<(p->v.n_trees + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p->v, (u64*)first, (u64*)second, player,
                                                                         is_final, outcome);
    DM_LAUNCH_CHECK();
    return DM_OK;
}

extern "C" int dm_pool_counters(dm_pool* p, uint64_t out[16]) {
    DM_CHECK_ARG(p && out, "dm_pool_counters: null argument");
    DM_CUDA(cudaSetDevice(p->cfg.device));
    DM_CUDA(cudaDeviceSynchronize());
    k_sum_counters<<<1, 512>>>(p->v, p->v.counters + 16);
    DM_LAUNCH_CHECK();
    DM_CUDA(cudaMemcpy(out, p->v.counters + 16, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (out[kCntCapacity] > p->capacity_errors_reported) {
        // reported once per new error: the counter itself stays in out[8], later calls succeed again
        p->capacity_errors_reported = out[kCntCapacity];
        dm_set_error("a tree ran out of node or path capacity (raise max_nodes_per_tree)");
        return DM_ERR_CAPACITY;
    }
    return DM_OK;
}

extern "C" int dm_selfplay_enable(dm_pool* p, int replay_capacity) {
    DM_CHECK_ARG(p && replay_capacity > 0, "dm_selfplay_enable: bad arguments");
    if (p->v.selfplay) return DM_OK;
    DM_CUDA(cudaSetDevice(p->cfg.device));
    PoolView& v = p->v;
    size_t T = (size_t)v.n_trees, E = T * DM_MAX_DEPTH, R = (size_t)replay_capacity;
    DM_TRY(pool_alloc(p, &v.ex_first, E));
    DM_TRY(pool_alloc(p, &v.ex_second, E));
    DM_TRY(pool_alloc(p, &v.ex_player, E));
    DM_TRY(pool_alloc(p, &v.ex_pi, E * DM_MAX_ACTIONS, false));
    DM_TRY(pool_alloc(p, &v.rp_first, R));
    DM_TRY(pool_alloc(p, &v.rp_second, R));
    DM_TRY(pool_alloc(p, &v.rp_player, R));
    DM_TRY(pool_alloc(p, &v.rp_pi, R * DM_MAX_ACTIONS));
    DM_TRY(pool_alloc(p, &v.rp_z, R));
    DM_TRY(pool_alloc(p, &v.rp_written, 1));
    DM_TRY(pool_alloc(p, &v.rp_game_end, (size_t)DM_REPLAY_GAME_RING));
    v.replay_cap = R;
    v.selfplay = 1;
    return DM_OK;
}

extern "C" int dm_selfplay_select(dm_pool* p, void* stream) {
    DM_CHECK_ARG(p, "dm_selfplay_select: null pool");
    if (!p->v.selfplay) {
        dm_set_error("dm_selfplay_select: call dm_selfplay_enable first");
        return DM_ERR_STATE;
    }
    return launch_select(p, stream, true);
}

extern "C" int dm_selfplay_replay(dm_pool* p, uint64_t** first, uint64_t** second, uint8_t** player, float** pi,
                                  float** z, uint64_t** written) {
    DM_CHECK_ARG(p && p->v.selfplay, "dm_selfplay_replay: self-play not enabled");
    if (first) *first = (uint64_t*)p->v.rp_first;
    if (second) *second = (uint64_t*)p->v.rp_second;
    if (player) *player = p->v.rp_player;
    if (pi) *pi = p->v.rp_pi;
    if (z) *z = p->v.rp_z;
    if (written) *written = (uint64_t*)p->v.rp_written;
    return DM_OK;
}

extern "C" int dm_selfplay_games(dm_pool* p, uint64_t** game_end) {
    DM_CHECK_ARG(p && p->v.selfplay && game_end, "dm_selfplay_games: self-play not enabled");
    *game_end = (uint64_t*)p->v.rp_game_end;
    return DM_OK;
}
