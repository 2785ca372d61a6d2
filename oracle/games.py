"""Bitboard restatement of the reference's four game engines.

TEST INFRASTRUCTURE (see oracle/__init__.py).

A state is ``(player, first, second)``: ``player`` 1 = FIRST (moves first, the
min player), 0 = SECOND (max player) (game.py:9-37); ``first`` / ``second`` are
Python ints whose bit ``y*g + x`` is set when that cell holds a FIRST / SECOND
stone (cells: 1 FIRST, 0 SECOND, -1 empty, game.py:40-43).  Outcome values are
absolute: FIRST win 0.0, SECOND win 1.0, draw 0.5 (game.py:54-57).

Game ids match include/deepmcts_b200.h: 0 tictactoe, 1 hex, 2 hex_with_swap,
3 othello.
"""
from typing import List, Optional, Tuple

from .setorder import IntSet, set_order

TICTACTOE, HEX, HEX_SWAP, OTHELLO = 0, 1, 2, 3
GAME_NAMES = {TICTACTOE: "tictactoe", HEX: "hex", HEX_SWAP: "hex_with_swap", OTHELLO: "othello"}

State = Tuple[int, int, int]  # (player, first_bb, second_bb)

# itertools.product([0, 1, -1], repeat=2) minus (0, 0), as (dx, dy)
# (othello/game.py:57, 89)
OTHELLO_SHIFTS = [(0, 1), (0, -1), (1, 0), (1, 1), (1, -1), (-1, 0), (-1, 1), (-1, -1)]
# hex/game.py:131
HEX_SHIFTS = [(0, -1), (1, -1), (1, 0), (0, 1), (-1, 1), (-1, 0)]

_TTT_LINES = [0o007, 0o070, 0o700, 0o111, 0o222, 0o444, 0o421, 0o124]


class Game:
    """One reference GameManager, restated on bitboards."""

    def __init__(self, game: int, grid_size: int) -> None:
        self.game = game
        self.g = 3 if game == TICTACTOE else grid_size
        g = self.g
        self.cells = g * g
        self.full = (1 << self.cells) - 1
        # num_actions: othello/game.py:34-36, hex/game.py:32-35,
        # hex_with_swap/game.py:12-14, tictactoe/game.py:25-26
        self.num_actions = self.cells + (1 if game in (HEX_SWAP, OTHELLO) else 0)
        self.extra_action = self.cells  # pass (othello) / swap (hex_with_swap)

    # ------------------------------------------------------------------ basics
    def initial_state(self) -> State:
        if self.game == OTHELLO:
            # othello/game.py:38-48
            g = self.g
            x = y = g // 2
            second = (1 << (y * g + x)) | (1 << ((y - 1) * g + x - 1))
            first = (1 << (y * g + x - 1)) | (1 << ((y - 1) * g + x))
            return (1, first, second)
        return (1, 0, 0)  # hex/game.py:37-44, tictactoe/game.py:28-32

    # ----------------------------------------------------------------- othello
    def _othello_discovery(self, state: State) -> List[int]:
        """Squares in the order othello/game.py:88-115 adds them to its set
        (duplicates kept; caller dedups through the set model)."""
        player, first, second = state
        own, opp = (first, second) if player == 1 else (second, first)
        g = self.g
        seq: List[int] = []
        for pos in range(self.cells):  # player_positions: row-major (othello/game.py:160-164)
            if not (own >> pos) & 1:
                continue
            x0, y0 = pos % g, pos // g
            for dx, dy in OTHELLO_SHIFTS:
                x, y = x0, y0
                found = False
                while True:
                    x += dx
                    y += dy
                    if not (0 <= x < g and 0 <= y < g):
                        break
                    if (opp >> (y * g + x)) & 1:
                        found = True
                    else:
                        break
                if found and 0 <= x < g and 0 <= y < g and not ((own | opp) >> (y * g + x)) & 1:
                    seq.append(y * g + x)
        return seq

    def _othello_flips(self, state: State, action: int) -> int:
        # othello/game.py:56-81
        player, first, second = state
        own, opp = (first, second) if player == 1 else (second, first)
        g = self.g
        x0, y0 = action % g, action // g
        flips = 0
        for dx, dy in OTHELLO_SHIFTS:
            x, y = x0, y0
            run = 0
            while True:
                x += dx
                y += dy
                if not (0 <= x < g and 0 <= y < g):
                    break
                if (opp >> (y * g + x)) & 1:
                    run |= 1 << (y * g + x)
                else:
                    break
            if 0 <= x < g and 0 <= y < g and (own >> (y * g + x)) & 1:
                flips |= run
        return flips

    # ------------------------------------------------------------- legal moves
    def legal_list(self, state: State) -> List[int]:
        """What ``manager.legal_actions(state)`` returns, in the reference's list order."""
        player, first, second = state
        occ = first | second
        if self.game == OTHELLO:
            s = IntSet()
            for a in self._othello_discovery(state):
                s.add(a)
            order = s.order()  # list(actions), othello/game.py:118
            return order if order else [self.extra_action]  # othello/game.py:116-117
        acts = [a for a in range(self.cells) if not (occ >> a) & 1]  # hex/game.py:66-72, tictactoe/game.py:55-62
        if self.game == HEX_SWAP and player == 0 and bin(occ).count("1") == 1:
            acts.append(self.extra_action)  # hex_with_swap/game.py:30-35
        return acts

    def children_order(self, state: State) -> List[int]:
        """Order in which MCTS.expand_node creates children: iteration order of
        ``set(legal_actions(state))`` (mcts.py:200, 213-215)."""
        return set_order(self.legal_list(state))

    def legal_mask(self, state: State) -> int:
        """Bit a set <=> action a is legal: the mask of GameNet._mask_illegal_moves (gamenet.py:122-128)."""
        m = 0
        for a in self.legal_list(state):
            m |= 1 << a
        return m

    # ------------------------------------------------------------- transitions
    def child(self, state: State, action: int) -> State:
        player, first, second = state
        if self.game == OTHELLO:
            if action == self.extra_action:
                return (1 - player, first, second)  # othello/game.py:55-56
            flips = self._othello_flips(state, action)
            bit = 1 << action
            if player == 1:
                return (0, first | bit | flips, second & ~flips)
            return (1, first & ~flips, second | bit | flips)
        if self.game == HEX_SWAP and action == self.extra_action:
            return (1 - player, second, first)  # recolour every stone, hex_with_swap/game.py:21-25
        bit = 1 << action
        if player == 1:
            return (0, first | bit, second)
        return (1, first, second | bit)

    # ---------------------------------------------------------------- terminal
    def _hex_connected(self, stones: int, player: int) -> bool:
        # hex/game.py:78-142: SECOND connects x=0 -> x=g-1, FIRST connects y=0 -> y=g-1
        g = self.g
        if player == 0:
            frontier = [y * g for y in range(g) if (stones >> (y * g)) & 1]
        else:
            frontier = [x for x in range(g) if (stones >> x) & 1]
        seen = 0
        for p in frontier:
            seen |= 1 << p
        while frontier:
            p = frontier.pop()
            x, y = p % g, p // g
            if (player == 0 and x == g - 1) or (player == 1 and y == g - 1):
                return True
            for dx, dy in HEX_SHIFTS:
                nx, ny = x + dx, y + dy
                if 0 <= nx < g and 0 <= ny < g:
                    q = ny * g + nx
                    if (stones >> q) & 1 and not (seen >> q) & 1:
                        seen |= 1 << q
                        frontier.append(q)
        return False

    def outcome(self, state: State) -> float:
        """``evaluate_final_state(state).value`` (also defined on non-final states)."""
        player, first, second = state
        if self.game == OTHELLO:
            # othello/game.py:126-141
            f, s = bin(first).count("1"), bin(second).count("1")
            return 0.0 if f > s else (1.0 if s > f else 0.5)
        if self.game == TICTACTOE:
            # SECOND's lines are tested first, tictactoe/game.py:87-97
            if any(second & l == l for l in _TTT_LINES):
                return 1.0
            if any(first & l == l for l in _TTT_LINES):
                return 0.0
            return 0.5
        # hex: SECOND tested first, hex/game.py:83-103
        if self._hex_connected(second, 0):
            return 1.0
        if self._hex_connected(first, 1):
            return 0.0
        return 0.5

    def is_final(self, state: State) -> bool:
        player, first, second = state
        if self.game == OTHELLO:
            # othello/game.py:120-124
            if self._othello_discovery(state):
                return False
            return not self._othello_discovery((1 - player, first, second))
        if self.game == TICTACTOE:
            # tictactoe/game.py:64-68
            return self.outcome(state) != 0.5 or (first | second) == self.full
        return self.outcome(state) != 0.5  # hex/game.py:74-76

    # --------------------------------------------------------------- converters
    def to_grid(self, state: State) -> Tuple[Tuple[int, ...], ...]:
        """Reference-style grid of cell values (-1 / 0 / 1)."""
        _, first, second = state
        g = self.g
        return tuple(
            tuple(
                1 if (first >> (y * g + x)) & 1 else (0 if (second >> (y * g + x)) & 1 else -1)
                for x in range(g)
            )
            for y in range(g)
        )

    def from_grid(self, player: int, grid) -> State:
        g = self.g
        first = second = 0
        for y in range(g):
            for x in range(g):
                c = int(grid[y][x])
                if c == 1:
                    first |= 1 << (y * g + x)
                elif c == 0:
                    second |= 1 << (y * g + x)
        return (int(player), first, second)


def make_game(name_or_id, grid_size: Optional[int] = None) -> Game:
    """Rules object by name or dm_game id (tic-tac-toe is always 3x3, tictactoe/game.py:24-31)."""
    ids = {v: k for k, v in GAME_NAMES.items()}
    gid = ids[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
    return Game(gid, 3 if gid == TICTACTOE else int(grid_size))
