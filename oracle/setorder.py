"""Model of CPython's ``set`` iteration order for small non-negative ints.

TEST INFRASTRUCTURE (see oracle/__init__.py).

Why this exists: the reference creates a node's children by iterating
``set(game_manager.legal_actions(state))`` (mcts.py:200, 213-215) and resolves
PUCT ties with Python's first-extremum ``max``/``min`` over that dict
(mcts.py:167-176).  The child order is therefore the slot order of a CPython
hash set, which this module restates from the published algorithm of
CPython 3.7-3.12 ``Objects/setobject.c`` (set_add_entry / set_insert_clean /
set_table_resize): open addressing, 8 initial slots, LINEAR_PROBES = 9,
PERTURB_SHIFT = 5, grow to the smallest power of two > 4*used when
fill*5 >= mask*3, re-inserting in old slot order.  hash(int k) == k for the
action ids used here (0..64).

Pinned by tests/test_oracle_golden.py against the interpreter's own
``list(set(...))`` and against orders recorded from the reference's MCTS.
"""
from typing import Iterable, List

_LINEAR_PROBES = 9
_PERTURB_SHIFT = 5
_MINSIZE = 8
_EMPTY = -1


def _insert_clean(table: List[int], key: int) -> None:
    mask = len(table) - 1
    perturb = key
    i = key & mask
    while True:
        if table[i] == _EMPTY:
            table[i] = key
            return
        if i + _LINEAR_PROBES <= mask:
            for j in range(1, _LINEAR_PROBES + 1):
                if table[i + j] == _EMPTY:
                    table[i + j] = key
                    return
        perturb >>= _PERTURB_SHIFT
        i = (i * 5 + 1 + perturb) & mask


class IntSet:
    """Incremental model: ``add`` mirrors set_add_entry for int keys (no deletions)."""

    def __init__(self) -> None:
        self.table = [_EMPTY] * _MINSIZE
        self.fill = 0

    def add(self, key: int) -> None:
        table = self.table
        mask = len(table) - 1
        perturb = key
        i = key & mask
        while True:
            probes = _LINEAR_PROBES if i + _LINEAR_PROBES <= mask else 0
            placed = False
            for j in range(probes + 1):
                slot = table[i + j]
                if slot == _EMPTY:
                    table[i + j] = key
                    placed = True
                    break
                if slot == key:
                    return  # already present
            if placed:
                break
            perturb >>= _PERTURB_SHIFT
            i = (i * 5 + 1 + perturb) & mask
        self.fill += 1
        if self.fill * 5 >= mask * 3:
            minused = self.fill * 4  # used == fill (no dummies); used <= 50000
            newsize = _MINSIZE
            while newsize <= minused:
                newsize <<= 1
            old = table
            self.table = [_EMPTY] * newsize
            for k in old:
                if k != _EMPTY:
                    _insert_clean(self.table, k)

    def order(self) -> List[int]:
        return [k for k in self.table if k != _EMPTY]


def set_order(keys: Iterable[int]) -> List[int]:
    """Iteration order of ``set(keys)`` built by successive insertion (list input)."""
    s = IntSet()
    for k in keys:
        s.add(k)
    return s.order()
