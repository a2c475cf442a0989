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


def test_host_only_shape_rules_of_the_tensor_core_entry_points(lib):
    """eligibility / workspace queries need no GPU: they are what the host code dispatches on"""
    i32 = ctypes.c_int32
    lib.cbrs_dense_tf32x3_eligible.argtypes = [i32, i32]
    assert lib.cbrs_dense_tf32x3_eligible(128, 128) and lib.cbrs_dense_tf32x3_eligible(64, 256)
    assert not lib.cbrs_dense_tf32x3_eligible(40, 32) and not lib.cbrs_dense_tf32x3_eligible(64, 24)
    assert not lib.cbrs_dense_tf32x3_eligible(256, 256)          # the two operand images do not fit shared memory
    lib.cbrs_dense_tc_bf16_eligible.argtypes = [i32, i32, i32]
    assert lib.cbrs_dense_tc_bf16_eligible(768, 0, 256) and lib.cbrs_dense_tc_bf16_eligible(128, 64, 64)
    assert not lib.cbrs_dense_tc_bf16_eligible(96, 0, 16) and not lib.cbrs_dense_tc_bf16_eligible(64, 0, 512)
    lib.cbrs_score_catalog_topk_tf32x3_eligible.argtypes = [i32, i32]
    assert lib.cbrs_score_catalog_topk_tf32x3_eligible(64, 64) and lib.cbrs_score_catalog_topk_tf32x3_eligible(32, 24)
    assert not lib.cbrs_score_catalog_topk_tf32x3_eligible(48, 48) and not lib.cbrs_score_catalog_topk_tf32x3_eligible(64, 128)
    lib.cbrs_score_catalog_topk_tf32x3_workspace_bytes.restype = ctypes.c_size_t
    lib.cbrs_score_catalog_topk_tf32x3_workspace_bytes.argtypes = [i32, i32]
    assert lib.cbrs_score_catalog_topk_tf32x3_workspace_bytes(64, 64) >= 2 * 2 * 64 * 128
    assert lib.cbrs_score_catalog_topk_tf32x3_workspace_bytes(48, 48) == 0
    lib.cbrs_score_catalog_topk_bf16_workspace_bytes.restype = ctypes.c_size_t
    lib.cbrs_score_catalog_topk_bf16_workspace_bytes.argtypes = [i32, i32, i32]
    small, wide = (lib.cbrs_score_catalog_topk_bf16_workspace_bytes(1000, 64, 64),
                   lib.cbrs_score_catalog_topk_bf16_workspace_bytes(1000, 64, 128))
    assert small > 1000 * 128 and small - wide >= 2 * 8192 + 2048 - 64 * 128   # the v3 kernel's images ride along for c2 <= 64


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(REPO, "deep_cbrs_amar_renaissance_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
