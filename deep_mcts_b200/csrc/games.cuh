// Device-side game engines on 64-bit bitboards: Othello, Hex, Hex-with-swap, tic-tac-toe.
//
// Replaces (reference paths relative to deep_mcts/):
//   othello/game.py:38-141, hex/game.py:37-142, hex_with_swap/game.py:16-36,
//   tictactoe/game.py:28-97
// and the CPython-set child ordering that mcts.py:200-215 inherits from
// `set(legal_actions(state))` (model: SURVEY.md Appendix A).
//
// All functions are written to be executed warp-uniformly (every lane holds the same
// state); the two that use lanes cooperatively say so.
#pragma once
#include "common.cuh"

struct GameCfg {
    int game;         // dm_game
    int g;            // board side
    int cells;        // g*g
    int num_actions;  // cells (+1 pass / swap)
    u64 full;         // all cells
    u64 not_col0;     // cells with x != 0
    u64 not_colL;     // cells with x != g-1
    u64 row0, rowL, col0, colL;
};

static inline GameCfg make_game_cfg(int game, int grid) {
    GameCfg c;
    c.game = game;
    c.g = (game == DM_TICTACTOE) ? 3 : grid;
    c.cells = c.g * c.g;
    c.num_actions = c.cells + ((game == DM_HEX_SWAP || game == DM_OTHELLO) ? 1 : 0);
    c.full = (c.cells == 64) ? ~0ull : ((1ull << c.cells) - 1);
    u64 col0 = 0;
    for (int y = 0; y < c.g; ++y) col0 |= 1ull << (y * c.g);
    c.col0 = col0;
    c.colL = col0 << (c.g - 1);
    c.not_col0 = c.full & ~c.col0;
    c.not_colL = c.full & ~c.colL;
    c.row0 = (1ull << c.g) - 1;
    c.rowL = c.row0 << (c.g * (c.g - 1));
    return c;
}

// The same configuration as compile-time constants: with GAME / G known the compiler folds every
// shift amount and mask of the bitboard code and unrolls the ray loops (the hot tree kernels are
// instantiated for the BASELINE board sizes; the generic kernels read the runtime GameCfg).
template <int GAME, int G>
__device__ __forceinline__ GameCfg const_game_cfg() {
    GameCfg c;
    c.game = GAME;
    c.g = G;
    c.cells = G * G;
    c.num_actions = G * G + ((GAME == DM_HEX_SWAP || GAME == DM_OTHELLO) ? 1 : 0);
    c.full = (G * G == 64) ? ~0ull : ((1ull << (G * G)) - 1);
    u64 col0 = 0;
#pragma unroll
    for (int y = 0; y < G; ++y) col0 |= 1ull << (y * G);
    c.col0 = col0;
    c.colL = col0 << (G - 1);
    c.not_col0 = c.full & ~c.col0;
    c.not_colL = c.full & ~c.colL;
    c.row0 = (1ull << G) - 1;
    c.rowL = c.row0 << (G * (G - 1));
    return c;
}

static inline bool game_cfg_valid(int game, int grid) {
    switch (game) {
        case DM_TICTACTOE: return true;
        case DM_HEX:
        case DM_HEX_SWAP: return grid >= 2 && grid <= 8;
        case DM_OTHELLO: return grid >= 4 && grid <= 8 && (grid % 2 == 0);
        default: return false;
    }
}

struct GState {
    u64 first, second;
    int player;  // 1 FIRST, 0 SECOND
};

// itertools.product([0, 1, -1], repeat=2) without (0,0), as (dx, dy)  (othello/game.py:57, 89)
#define DM_FOR_EACH_OTHELLO_DIR(F) F(0, 1) F(0, -1) F(1, 0) F(1, 1) F(1, -1) F(-1, 0) F(-1, 1) F(-1, -1)

template <int DX, int DY>
__device__ __forceinline__ u64 bb_shift(u64 b, const GameCfg& c) {
    u64 r;
    if (DY > 0 || (DY == 0 && DX > 0)) r = (b << (DY * c.g + DX)) & c.full;
    else r = b >> (-(DY * c.g + DX));
    if (DX > 0) r &= c.not_col0;
    else if (DX < 0) r &= c.not_colL;
    return r;
}

// ------------------------------------------------------------------------------ othello
template <int DX, int DY>
__device__ __forceinline__ u64 othello_moves_dir(u64 own, u64 opp, u64 empty, const GameCfg& c) {
    u64 t = bb_shift<DX, DY>(own, c) & opp;
    for (int i = 0; i < c.g - 3; ++i) t |= bb_shift<DX, DY>(t, c) & opp;
    return bb_shift<DX, DY>(t, c) & empty;
}

__device__ __forceinline__ u64 othello_moves_inline(u64 own, u64 opp, const GameCfg& c) {
    u64 empty = ~(own | opp) & c.full;
    u64 moves = 0;
#define DM_DIR(DX, DY) moves |= othello_moves_dir<DX, DY>(own, opp, empty, c);
    DM_FOR_EACH_OTHELLO_DIR(DM_DIR)
#undef DM_DIR
    return moves;
}
// The 8x8 board (the BASELINE size) as ONE out-of-line copy with constant masks and shifts: the fully unrolled
// eight-direction code is ~500 instructions, and the tree kernels reach it from half a dozen places -- inlined
// everywhere it made k_round a 300 KB kernel that stalled on instruction fetch as much as on memory.
static __device__ __noinline__ u64 othello_moves8(u64 own, u64 opp) {
    return othello_moves_inline(own, opp, const_game_cfg<DM_OTHELLO, 8>());
}
__device__ __forceinline__ u64 othello_moves(u64 own, u64 opp, const GameCfg& c) {
    if (c.g == 8) return othello_moves8(own, opp);
    return othello_moves_inline(own, opp, c);
}

template <int DX, int DY>
__device__ __forceinline__ u64 othello_flips_dir(u64 own, u64 opp, u64 bit, const GameCfg& c) {
    u64 f = bb_shift<DX, DY>(bit, c);
    u64 run = 0;
    while (f & opp) {
        run |= f;
        f = bb_shift<DX, DY>(f, c);
    }
    return (f & own) ? run : 0ull;
}

// stones flipped by playing `bit` (othello/game.py:56-81)
__device__ __forceinline__ u64 othello_flips(u64 own, u64 opp, u64 bit, const GameCfg& c) {
    u64 flips = 0;
#define DM_DIR(DX, DY) flips |= othello_flips_dir<DX, DY>(own, opp, bit, c);
    DM_FOR_EACH_OTHELLO_DIR(DM_DIR)
#undef DM_DIR
    return flips;
}

// the same flips without data-dependent loops: the run of opponent stones next to `bit` in one direction
// (g - 2 stones at most), kept iff the square behind the run is an own stone.  Lanes that evaluate different moves
// stay converged (expand_node computes one child per lane).
template <int DX, int DY>
__device__ __forceinline__ u64 othello_flips_dir_pp(u64 own, u64 opp, u64 bit, const GameCfg& c) {
    u64 t = bb_shift<DX, DY>(bit, c) & opp;
    for (int i = 0; i < c.g - 3; ++i) t |= bb_shift<DX, DY>(t, c) & opp;
    return (bb_shift<DX, DY>(t, c) & own) ? t : 0ull;
}
__device__ __forceinline__ u64 othello_flips_pp_inline(u64 own, u64 opp, u64 bit, const GameCfg& c) {
    u64 flips = 0;
#define DM_DIR(DX, DY) flips |= othello_flips_dir_pp<DX, DY>(own, opp, bit, c);
    DM_FOR_EACH_OTHELLO_DIR(DM_DIR)
#undef DM_DIR
    return flips;
}
static __device__ __noinline__ u64 othello_flips8(u64 own, u64 opp, u64 bit) {
    return othello_flips_pp_inline(own, opp, bit, const_game_cfg<DM_OTHELLO, 8>());
}
__device__ __forceinline__ u64 othello_flips_pp(u64 own, u64 opp, u64 bit, const GameCfg& c) {
    if (c.g == 8) return othello_flips8(own, opp, bit);
    return othello_flips_pp_inline(own, opp, bit, c);
}

// Position of (dx, dy) in itertools.product([0, 1, -1], repeat=2) without the leading (0, 0): the order in which
// legal_actions walks the rays of one stone (othello/game.py:89).
__host__ __device__ constexpr int othello_dir_index(int dx, int dy) {
    return (dx == 0 ? 0 : dx == 1 ? 3 : 6) + (dy == 0 ? 0 : dy == 1 ? 1 : 2) - 1;
}

// One direction seen from the EMPTY square `bit` (a legal move): the run of opponent stones next to it and the
// own stone that brackets the run.  That stone is the one whose ray in the opposite direction reaches `bit` in
// legal_actions' scan (othello/game.py:88-115), so the same walk yields the stones the move flips
// (othello/game.py:56-81) and when the scan discovers the move: `seq` = 8 * stone cell + ray index, the position
// of the (stone, ray) pair in the scan's loop order (stones ascending by cell, rays in product order).
template <int DX, int DY>
__device__ __forceinline__ void othello_probe_dir(u64 own, u64 opp, u64 bit, const GameCfg& c, u64& flips, u32& seq) {
    u64 t = bb_shift<DX, DY>(bit, c) & opp;
    for (int i = 0; i < c.g - 3; ++i) t |= bb_shift<DX, DY>(t, c) & opp;
    const u64 end = bb_shift<DX, DY>(t, c) & own;  // at most one bit: the square right behind the run
    if (end) {
        flips |= t;
        const u32 q = ((u32)(__ffsll((long long)end) - 1) << 3) | (u32)othello_dir_index(-DX, -DY);
        seq = min(seq, q);
    }
}

// ---------------------------------------------------------------------------------- hex
__device__ __forceinline__ u64 hex_neighbours(u64 r, const GameCfg& c) {
    // hex/game.py:131: (0,-1),(1,-1),(1,0),(0,1),(-1,1),(-1,0)
    return bb_shift<0, -1>(r, c) | bb_shift<1, -1>(r, c) | bb_shift<1, 0>(r, c) | bb_shift<0, 1>(r, c) |
           bb_shift<-1, 1>(r, c) | bb_shift<-1, 0>(r, c);
}

__device__ __forceinline__ bool hex_connected(u64 stones, u64 start, u64 goal, const GameCfg& c) {
    u64 reach = stones & start;
    while (reach) {
        if (reach & goal) return true;
        u64 grow = hex_neighbours(reach, c) & stones & ~reach;
        if (!grow) return false;
        reach |= grow;
    }
    return false;
}

// ---------------------------------------------------------------------------- tictactoe
__device__ __forceinline__ bool ttt_line(u64 b) {
    return (b & 0007) == 0007 || (b & 0070) == 0070 || (b & 0700) == 0700 || (b & 0111) == 0111 ||
           (b & 0222) == 0222 || (b & 0444) == 0444 || (b & 0421) == 0421 || (b & 0124) == 0124;
}

// ------------------------------------------------------------------------ generic logic
__device__ __forceinline__ GState game_initial(const GameCfg& c) {
    GState s;
    s.player = 1;
    s.first = s.second = 0;
    if (c.game == DM_OTHELLO) {
        int x = c.g / 2, y = c.g / 2;  // othello/game.py:43-47
        s.second = (1ull << (y * c.g + x)) | (1ull << ((y - 1) * c.g + x - 1));
        s.first = (1ull << (y * c.g + x - 1)) | (1ull << ((y - 1) * c.g + x));
    }
    return s;
}

// evaluate_final_state(...).value
__device__ __forceinline__ float game_outcome(const GameCfg& c, const GState& s) {
    if (c.game == DM_OTHELLO) {
        int f = __popcll(s.first), sc = __popcll(s.second);  // othello/game.py:126-141
        return f > sc ? 0.0f : (sc > f ? 1.0f : 0.5f);
    }
    if (c.game == DM_TICTACTOE) {
        if (ttt_line(s.second)) return 1.0f;  // SECOND tested first, tictactoe/game.py:87-97
        if (ttt_line(s.first)) return 0.0f;
        return 0.5f;
    }
    if (hex_connected(s.second, c.col0, c.colL, c)) return 1.0f;  // hex/game.py:83-93
    if (hex_connected(s.first, c.row0, c.rowL, c)) return 0.0f;   // hex/game.py:94-103
    return 0.5f;
}

// legal cells as a mask plus whether the extra action (pass / swap) is legal
__device__ __forceinline__ u64 game_legal_mask(const GameCfg& c, const GState& s, bool& extra) {
    u64 occ = s.first | s.second;
    extra = false;
    if (c.game == DM_OTHELLO) {
        u64 own = s.player ? s.first : s.second, opp = s.player ? s.second : s.first;
        u64 m = othello_moves(own, opp, c);
        extra = (m == 0);  // othello/game.py:116-117
        return m;
    }
    if (c.game == DM_HEX_SWAP) extra = (s.player == 0) && (__popcll(occ) == 1);  // hex_with_swap/game.py:30-35
    return ~occ & c.full;
}

__device__ __forceinline__ bool game_is_final(const GameCfg& c, const GState& s, float& outcome) {
    outcome = game_outcome(c, s);
    if (c.game == DM_OTHELLO) {
        // othello/game.py:120-124: both sides can only pass
        if (othello_moves(s.player ? s.first : s.second, s.player ? s.second : s.first, c)) return false;
        return othello_moves(s.player ? s.second : s.first, s.player ? s.first : s.second, c) == 0;
    }
    if (c.game == DM_TICTACTOE) return outcome != 0.5f || ((s.first | s.second) == c.full);
    return outcome != 0.5f;  // hex/game.py:74-76
}

__device__ __forceinline__ bool game_action_legal(const GameCfg& c, const GState& s, int action) {
    bool extra;
    u64 m = game_legal_mask(c, s, extra);
    if (action < 0 || action >= c.num_actions) return false;
    if (action == c.cells) return extra;
    return (m >> action) & 1;
}

// generate_child_state for a legal action
__device__ __forceinline__ GState game_child(const GameCfg& c, const GState& s, int action) {
    GState r;
    r.player = 1 - s.player;
    if (action == c.cells) {
        if (c.game == DM_HEX_SWAP) {  // recolour every stone, hex_with_swap/game.py:21-25
            r.first = s.second;
            r.second = s.first;
        } else {  // othello pass, othello/game.py:55-56
            r.first = s.first;
            r.second = s.second;
        }
        return r;
    }
    u64 bit = 1ull << action;
    u64 flips = 0;
    if (c.game == DM_OTHELLO)
        flips = othello_flips(s.player ? s.first : s.second, s.player ? s.second : s.first, bit, c);
    if (s.player) {
        r.first = s.first | bit | flips;
        r.second = s.second & ~flips;
    } else {
        r.first = s.first & ~flips;
        r.second = s.second | bit | flips;
    }
    return r;
}

// generate_child_state for a legal action, loop-free for Othello (identical result to game_child)
__device__ __forceinline__ GState game_child_uniform(const GameCfg& c, const GState& s, int action) {
    if (c.game != DM_OTHELLO || action == c.cells) return game_child(c, s, action);
    GState r;
    r.player = 1 - s.player;
    const u64 bit = 1ull << action;
    const u64 flips = othello_flips_pp(s.player ? s.first : s.second, s.player ? s.second : s.first, bit, c);
    if (s.player) {
        r.first = s.first | bit | flips;
        r.second = s.second & ~flips;
    } else {
        r.first = s.first & ~flips;
        r.second = s.second | bit | flips;
    }
    return r;
}

// is_final_state + outcome of the child reached from the NON-FINAL state `parent` by `action`, as a node code:
// 0 = not final, 1 + 2 * outcome otherwise (1: FIRST wins / value 0, 2: draw, 3: SECOND wins / value 1).
// Hex: the parent has no winner, so only the player who just placed a stone can have connected -- one flood fill
// instead of two; a swap recolours a single stone and cannot end the game.
__device__ __forceinline__ int game_child_final_code(const GameCfg& c, const GState& parent, int action,
                                                     const GState& child) {
    if (c.game == DM_HEX || c.game == DM_HEX_SWAP) {
        if (action == c.cells) return 0;
        if (parent.player) return hex_connected(child.first, c.row0, c.rowL, c) ? 1 : 0;   // FIRST: rows
        return hex_connected(child.second, c.col0, c.colL, c) ? 3 : 0;                      // SECOND: columns
    }
    float o;
    if (!game_is_final(c, child, o)) return 0;
    return 1 + (int)(o * 2.0f);
}
__device__ __forceinline__ int game_final_code(const GameCfg& c, const GState& s) {
    float o;
    if (!game_is_final(c, s, o)) return 0;
    return 1 + (int)(o * 2.0f);
}

// ---------------------------------------------------------------- CPython set ordering
// A 32-slot open-addressing table held one slot per lane (8-slot phase uses lanes 0..7).
// Only sets of <= 18 keys are emulated; >= 19 keys always iterate ascending (the table
// has grown to 128 slots and every key < 128 sits in its own slot).
struct WarpSet {
    int slot;  // this lane's slot (-1 empty)
    int size;  // 8 or 32
    int fill;
    u32 occ;   // occupied slots (warp-uniform copy of "slot >= 0" per lane)

    __device__ __forceinline__ void init() {
        slot = -1;
        size = 8;
        fill = 0;
        occ = 0;
    }
    __device__ __forceinline__ void place(int key, int lane) {
        u32 mask = (u32)size - 1u;
        u32 i = (u32)key & mask, perturb = (u32)key;
        u32 empties = ~occ & (size == 32 ? 0xffffffffu : 0xffu);
        while (true) {
            u32 window = ((i + 9u <= mask) ? 0x3ffu : 1u) << i;  // LINEAR_PROBES = 9
            u32 hit = empties & window;
            if (hit) {
                const int target = __ffs(hit) - 1;
                if (lane == target) slot = key;
                occ |= 1u << target;
                return;
            }
            perturb >>= 5;  // PERTURB_SHIFT
            i = (i * 5u + 1u + perturb) & mask;
        }
    }
    // key must not be in the set yet (callers insert distinct keys)
    __device__ __forceinline__ void add(int key, int lane) {
        place(key, lane);
        ++fill;
        if (size == 8 && fill * 5 >= 21) {  // fill*5 >= mask*3 -> grow to 32 (4*5 = 20 < 32)
            int old = slot;
            u32 old_occ = occ;
            slot = -1;
            occ = 0;
            size = 32;
            while (old_occ) {  // re-insert in slot order (set_table_resize walks the old table)
                const int s = __ffs(old_occ) - 1;
                old_occ &= old_occ - 1;
                place(__shfl_sync(DM_FULL_MASK, old, s), lane);
            }
        }
    }
    // writes keys in slot order to out[0..fill)
    __device__ __forceinline__ void emit(u8* out, int lane) const {
        if (slot >= 0) out[__popc(occ & ((1u << lane) - 1u))] = (u8)slot;
    }
};

__device__ __forceinline__ void emit_ascending(u64 mask, bool extra, int extra_action, u8* out, int lane) {
    // lane l writes the cells l and l+32 it owns
    u32 lo = (u32)mask, hi = (u32)(mask >> 32);
    if ((lo >> lane) & 1) out[__popc(lo & ((1u << lane) - 1u))] = (u8)lane;
    if ((hi >> lane) & 1) out[__popc(lo) + __popc(hi & ((1u << lane) - 1u))] = (u8)(lane + 32);
    if (extra && lane == 0) out[__popcll(mask)] = (u8)extra_action;
}

// A set of 5..18 distinct keys < 64 no two of which agree modulo 32 sits in a 32-slot table with every key in its
// home slot (key & 31) whatever the insertion order was -- the 8-slot phase included, its keys are re-inserted into
// the 32-slot table when the fifth key arrives -- so the set iterates in ascending key & 31.  `keys` has bit k set
// for key k; lo & hi == 0 is exactly "no two keys share a home slot".
__device__ __forceinline__ bool set_order_is_by_home_slot(u64 keys, int n) {
    return n >= 5 && n <= 18 && ((u32)keys & (u32)(keys >> 32)) == 0;
}
__device__ __forceinline__ void emit_by_home_slot(u64 keys, u8* out, u8* out2, int lane) {
    const u32 lo = (u32)keys, hi = (u32)(keys >> 32), any = lo | hi;
    if ((any >> lane) & 1) {
        const int pos = __popc(any & ((1u << lane) - 1u));
        const u8 key = (u8)(((lo >> lane) & 1) ? lane : lane + 32);
        out[pos] = key;
        if (out2) out2[pos] = key;
    }
}

// Warp-cooperative.  Writes the order in which MCTS.expand_node creates children
// (iteration order of set(legal_actions(state)), mcts.py:200-215) to order_out[0..n) and,
// if list_out != nullptr, the reference's legal_actions LIST order to list_out[0..n).
// Both buffers need DM_MAX_ACTIONS + 1 bytes visible to the whole warp.  Returns n.
// If `flips_out` is given and the Othello small-set path ran (warp-uniform `*have_flips`), lane p < n also
// receives the stones flipped by the action order_out[p], a by-product of that path.
__device__ __forceinline__ int warp_children_order(const GameCfg& c, const GState& s, int lane, u8* order_out,
                                                   u8* list_out, u64* flips_out = nullptr,
                                                   bool* have_flips = nullptr) {
    bool extra;
    u64 mask = game_legal_mask(c, s, extra);
    int n = __popcll(mask) + (extra ? 1 : 0);
    if (have_flips) *have_flips = false;
    if (c.game == DM_OTHELLO && mask == 0) {  // only pass
        if (lane == 0) {
            order_out[0] = (u8)c.cells;
            if (list_out) list_out[0] = (u8)c.cells;
        }
        __syncwarp();
        return 1;
    }
    if (c.game != DM_OTHELLO) {
        // legal_actions is the ascending list (+ swap last); children = set(list)
        if (list_out) emit_ascending(mask, extra, c.cells, list_out, lane);
        const u64 keys = mask | ((extra && c.cells < 64) ? (1ull << c.cells) : 0ull);
        if (n >= 19) {
            emit_ascending(mask, extra, c.cells, order_out, lane);
        } else if (!(extra && c.cells >= 64) && set_order_is_by_home_slot(keys, n)) {
            emit_by_home_slot(keys, order_out, nullptr, lane);
        } else {
            WarpSet set;
            set.init();
            u64 m = mask;
            while (m) {
                int k = __ffsll((long long)m) - 1;
                m &= m - 1;
                set.add(k, lane);
            }
            if (extra) set.add(c.cells, lane);
            set.emit(order_out, lane);
        }
        __syncwarp();
        return n;
    }
    // Othello: legal_actions itself returns list(set) built in discovery order
    // (othello/game.py:88-118); expand_node wraps that list in a second set.
    if (n >= 19) {
        emit_ascending(mask, false, c.cells, order_out, lane);
        if (list_out) emit_ascending(mask, false, c.cells, list_out, lane);
        __syncwarp();
        return n;
    }
    const u64 own = s.player ? s.first : s.second, opp = s.player ? s.second : s.first;
    // one legal move per lane (lane j < n: the j-th legal cell, ascending): every direction is walked from the
    // move's square, which gives its flips and the first (stone, ray) pair of the reference's scan that finds it
    emit_ascending(mask, false, c.cells, order_out, lane);
    __syncwarp();
    const int cell = order_out[lane < n ? lane : 0];
    __syncwarp();
    u64 flips = 0;
    u32 seq = 0xffffffffu;
    {
        const u64 bit = 1ull << cell;
#define DM_DIR(DX, DY) othello_probe_dir<DX, DY>(own, opp, bit, c, flips, seq);
        DM_FOR_EACH_OTHELLO_DIR(DM_DIR)
#undef DM_DIR
    }
    if (set_order_is_by_home_slot(mask, n)) {
        // no two moves share a home slot: the scan's set and expand_node's set both iterate by key & 31,
        // in whatever order the scan met the moves
        emit_by_home_slot(mask, order_out, list_out, lane);
    } else {
        if (lane >= n) seq = 0xffffffffu;
        // the scan adds a move when it first meets it: rank the moves by that moment (all seq are distinct)
        int rank = 0;
        for (int i = 0; i < n; ++i) rank += (__shfl_sync(DM_FULL_MASK, seq, i) < seq) ? 1 : 0;
        if (lane < n) order_out[rank] = (u8)cell;
        __syncwarp();
        const int discovered = order_out[lane < n ? lane : 0];  // lane r: the r-th move the scan adds to its set
        __syncwarp();
        WarpSet stage1;
        stage1.init();
        for (int r = 0; r < n; ++r) stage1.add(__shfl_sync(DM_FULL_MASK, discovered, r), lane);
        if (list_out) stage1.emit(list_out, lane);
        // second set: insert stage-1 keys in stage-1 slot order
        WarpSet stage2;
        stage2.init();
        u32 used = stage1.occ;
        while (used) {
            int src = __ffs(used) - 1;
            used &= used - 1;
            stage2.add(__shfl_sync(DM_FULL_MASK, stage1.slot, src), lane);
        }
        stage2.emit(order_out, lane);
    }
    __syncwarp();
    if (flips_out) {
        // lane p takes the flips of the move at child position p from the lane that walked it
        const int a = order_out[lane < n ? lane : 0];
        const int src = __popcll(mask & ((1ull << a) - 1ull));
        *flips_out = __shfl_sync(DM_FULL_MASK, flips, src);
        if (have_flips) *have_flips = true;
    }
    return n;
}

// ----------------------------------------------------------------------------- rollouts
__device__ __forceinline__ int nth_set_bit(u64 m, int n) {
    for (int i = 0; i < n; ++i) m &= m - 1;
    return __ffsll((long long)m) - 1;
}

// MCTS.rollout (mcts.py:219-226) from a non-final state; returns the outcome value
__device__ __forceinline__ float game_rollout(const GameCfg& c, GState s, int kind, Philox& rng, int& plies) {
    float outcome;
    plies = 0;
    if (kind == DM_ROLLOUT_RANDOM && (c.game == DM_HEX || c.game == DM_HEX_SWAP)) {
        // Uniform random alternate play until someone connects has the same winner
        // distribution as filling the board alternately at random and reading the winner
        // once: a completed Hex board has exactly one winner and the first winning chain
        // survives further placements.
        while (true) {
            bool extra;
            u64 m = game_legal_mask(c, s, extra);
            int n = __popcll(m) + (extra ? 1 : 0);
            if (n == 0) break;
            int r = (int)rng.below((u32)n);
            int a = (r == __popcll(m)) ? c.cells : nth_set_bit(m, r);
            s = game_child(c, s, a);
            ++plies;
        }
        return game_outcome(c, s);
    }
    while (!game_is_final(c, s, outcome)) {
        bool extra;
        u64 m = game_legal_mask(c, s, extra);
        int cnt = __popcll(m);
        int a;
        if (cnt == 0) a = c.cells;  // othello pass is the only action
        else if (kind == DM_ROLLOUT_RANDOM) {
            int r = (int)rng.below((u32)(cnt + (extra ? 1 : 0)));
            a = (r == cnt) ? c.cells : nth_set_bit(m, r);
        } else if (kind == DM_ROLLOUT_MIN_ACTION) a = __ffsll((long long)m) - 1;
        else a = extra ? c.cells : (63 - __clzll((long long)m));
        s = game_child(c, s, a);
        ++plies;
    }
    return outcome;
}
