"""CPU-side checks of the C-ABI boundary: the library builds/loads, exports every
symbol include/rmpnn_b200.h declares, and the ctypes mirror matches.  No kernels run."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "rmpnn_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rmp_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def library():
    from rmpnn_b200 import _lib
    if not os.path.exists(_lib.lib_path()):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.lib()


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("rmp_knn", "rmp_graph_index", "rmp_edge_list", "rmp_rdesign_raw_features", "rmp_row_mlp",
              "rmp_add_layernorm", "rmp_readout_head", "rmp_last_error_string"):
        assert s in syms


def test_library_exports_every_declared_symbol(library):
    for s in declared_symbols():
        assert hasattr(library, s), f"{s} declared in include/rmpnn_b200.h but not exported"


def test_ctypes_mirror_matches_struct_layout(library):
    from rmpnn_b200 import _lib
    assert library.rmp_abi_struct_size() == ctypes.sizeof(_lib.MlpDesc)
    for name in _lib.SIGNATURES:
        assert name in declared_symbols()


def test_argument_errors_do_not_need_a_gpu(library):
    assert library.rmp_knn(None, None, 1, 1, 1, 0, 1, None, None, None) == -1
    assert b"null" in library.rmp_last_error_string()
    assert library.rmp_row_mlp(None, None) == -1


def test_ops_refuse_cpu_tensors(library):
    from rmpnn_b200 import ops
    X = torch.zeros(1, 4, 6, 3)
    mask = torch.ones(1, 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.knn(X, mask, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.build_graph(X, mask, 2)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from rmpnn_b200 import _lib
    monkeypatch.setattr(_lib, "_LIB", None)
    monkeypatch.setenv("RMPNN_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.LibraryMissing):
        _lib.lib()
