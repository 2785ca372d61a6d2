"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol that
include/deepmcts_b200.h declares (no compute calls -- those need a GPU)."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _header_symbols():
    text = (ROOT / "include" / "deepmcts_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_functions():
    syms = _header_symbols()
    assert "dm_pool_create" in syms and "dm_net_evaluate" in syms and len(syms) >= 35


def test_library_exports_every_declared_symbol():
    from deep_mcts_b200 import _lib

    if not _lib.LIB_PATH.exists():
        import __graft_entry__ as entry

        entry.build()
    lib = _lib.load()
    declared = _header_symbols()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared, "ctypes table and header disagree"
    assert lib.dm_abi_version() == 1


def test_invalid_arguments_are_reported_not_crashed():
    from deep_mcts_b200 import _lib

    lib = _lib.load()
    assert lib.dm_game_num_actions(_lib.DM_OTHELLO, 8) == 65
    assert lib.dm_game_num_actions(_lib.DM_HEX, 7) == 49
    assert lib.dm_game_num_actions(_lib.DM_HEX_SWAP, 7) == 50
    assert lib.dm_game_num_actions(_lib.DM_TICTACTOE, 3) == 9
    assert lib.dm_game_num_actions(_lib.DM_OTHELLO, 7) == _lib.DM_ERR_INVALID
    assert b"unsupported" in lib.dm_last_error()
    with pytest.raises(ValueError):
        _lib.call("dm_game_num_actions", 9, 3)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under deep_mcts_b200/ may reference it."""
    for path in (ROOT / "deep_mcts_b200").rglob("*.py"):
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No CPU fallback: if the CUDA library cannot be loaded the product raises, naming the build step."""
    import importlib
    from deep_mcts_b200 import _lib
    monkeypatch.setenv("DM_LIB_PATH", str(tmp_path / "missing.so"))
    fresh = importlib.reload(_lib)
    try:
        with pytest.raises(ImportError) as err:
            fresh.load()
        assert "missing.so" in str(err.value) and "no CPU fallback" in str(err.value)
    finally:
        monkeypatch.delenv("DM_LIB_PATH")
        importlib.reload(_lib)
