"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/laplace_b200.h
declares, the ctypes table covers them all, and -- with no GPU -- compute calls fail loudly instead of
falling back to anything."""
import ctypes as C
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "laplace_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lgb_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__
    __graft_entry__.build()
    from laplace_gnn_recommendation_b200 import _lib
    return _lib.load()


def test_header_symbols_exported(lib):
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/laplace_b200.h but not exported by the .so"


def test_ctypes_table_matches_header(lib):
    from laplace_gnn_recommendation_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_functions()
    assert lib.lgb_abi_version() == 4


def test_struct_layouts():
    from laplace_gnn_recommendation_b200._lib import LgbBprArgs, LgbCsr
    assert C.sizeof(LgbCsr) == 3 * 8 + 4 * 8 + 8 + 2 * 8 + 5 * 8 + 2 * 8 + 8 + 4 * 8 + 8 + 2 * 8  # mirrors struct lgb_csr (ABI 2: + hot-column plan, ABI 3: + stage-2 segments, ABI 4: + task_exec, task_seg)
    assert C.sizeof(LgbBprArgs) == 9 * 8 + 8 + 8 + 4 * 4 + 2 * 8 + 8 + 6 * 8 + 2 * 8            # mirrors struct lgb_bpr_args


def test_argument_validation_without_gpu(lib):
    from laplace_gnn_recommendation_b200._lib import LgbCsr
    g = LgbCsr()
    assert lib.lgb_spmm(C.byref(g), None, 64, None, None, None, None, 1.0, 0, None, None) == 1  # LGB_EINVAL
    assert b"lgb_spmm" in lib.lgb_last_error()
    need = C.c_size_t(0)
    assert lib.lgb_csr_build_ws_bytes(1 << 40, 10, C.byref(need)) == 2                       # LGB_ERANGE
    assert lib.lgb_csr_build_ws_bytes(1000, 10, C.byref(need)) == 0 and need.value > 1000 * 24
    assert lib.lgb_bpr_blocks(128) == 16


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    import laplace_gnn_recommendation_b200 as lg
    st = lg.SparseTensor(row=torch.tensor([0, 1]), col=torch.tensor([1, 0]), sparse_sizes=(2, 2))
    with pytest.raises(RuntimeError, match="no CPU"):
        lg.matmul(st, torch.ones(2, 4))
    model = lg.LightGCN(1, 1, 4, 2)
    with pytest.raises(RuntimeError, match="CUDA"):
        model(st)
    with pytest.raises(RuntimeError, match="CUDA"):
        lg.bpr_loss(*[torch.ones(2, 4)] * 6, 0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        lg.sample_mini_batch(2, torch.tensor([[0, 1], [1, 0]]))


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "laplace_gnn_recommendation_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
                # ... nor the CPU emulation of the kernels that tests/ uses (tests/emu/): the product has no CPU path
                assert not re.search(r"tests\.emu|liblaplace_b200_emu|LGB_CPU_EMU|cuda_emu", txt), f
