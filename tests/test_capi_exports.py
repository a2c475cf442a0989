"""CPU: the C-ABI library loads and exports every symbol include/cbrs_b200.h declares
(no compute call is made without a GPU)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as entry
from deep_cbrs_amar_renaissance_b200 import _lib as L

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(L.LIB_PATH):
        entry.build()
    return ctypes.CDLL(L.LIB_PATH)


def _declared():
    text = open(os.path.join(REPO, "include", "cbrs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cbrs_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "declared in the header but not exported: " + name


def test_binding_table_matches_header():
    assert sorted(L.exported_symbols()) == _declared()


def test_host_only_queries(lib):
    lib.cbrs_version.restype = ctypes.c_int
    assert lib.cbrs_version() >= 100
    lib.cbrs_sort_workspace_bytes.restype = ctypes.c_size_t
    lib.cbrs_sort_workspace_bytes.argtypes = [ctypes.c_int64]
    assert lib.cbrs_sort_workspace_bytes(1 << 20) > (1 << 20) * 12
    lib.cbrs_graph_build_workspace_bytes.restype = ctypes.c_size_t
    lib.cbrs_graph_build_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int]
    assert lib.cbrs_graph_build_workspace_bytes(1000, 100, 7) > 1000 * 16


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "deep_cbrs_amar_renaissance_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
