"""The C-ABI library loads without a GPU, exports every symbol include/mtasr.h declares, and rejects bad arguments
with an error code + message instead of crashing.  No compute is launched here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mtasr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mtasr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    from mtasr_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mtasr.h but not exported by libmtasr.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) <= set(names), set(_lib.SIGNATURES) - set(names)
    assert lib.mtasr_version() >= 100


def test_gemm_desc_layout_matches_header():
    from mtasr_b200._lib import GemmDesc
    src = open(os.path.join(ROOT, "include", "mtasr.h")).read()
    body = src[src.index("typedef struct mtasr_gemm_desc {"):src.index("} mtasr_gemm_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?[A-Za-z0-9_]+\s*\*?", "", decl, count=1)
        fields += [n.strip().lstrip("*") for n in names.split(",")]
    assert fields == [f[0] for f in GemmDesc._fields_]


def test_host_side_queries_and_argument_errors():
    from mtasr_b200 import _lib
    lib = _lib.load()
    assert lib.mtasr_ctc_state_pad(0) == 64 and lib.mtasr_ctc_state_pad(31) == 64 and lib.mtasr_ctc_state_pad(32) == 128
    assert lib.mtasr_ctc_state_pad(255) == 512 and lib.mtasr_ctc_state_pad(256) == -1
    assert lib.mtasr_gemm_n_tiles(128259, 0) == 1004 and lib.mtasr_gemm_n_tiles(64, 0) == 2 and lib.mtasr_gemm_n_tiles(300, 128) == 6
    assert lib.mtasr_gemm_bf16(None, None) == -1
    assert b"null descriptor" in lib.mtasr_last_error_string()
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.batch0, d.batch1 = 0, 8, 8, 1, 1
    assert lib.mtasr_gemm_bf16(C.byref(d), None) == -1 and b"positive" in lib.mtasr_last_error_string()
    assert lib.mtasr_ctc_collapse(None, 1, 1, 0, 0, None, None, None) < 0
    assert lib.mtasr_lstm_fwd(None, None, 0, 1, 1, 16, None, None, None, None, None, None) < 0
    with pytest.raises(_lib.MtasrError):
        _lib.check(-1, "unit-test")


def test_product_path_fails_loudly_without_cuda():
    import torch
    from mtasr_b200 import kernels as K
    from mtasr_b200._lib import MtasrError
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(MtasrError):
        K.cast_bf16(torch.zeros(8))
    with pytest.raises((MtasrError, RuntimeError)):
        K.ctc_collapse(torch.zeros(2, 4, dtype=torch.int64), 1, 0)


def test_sm_budget_precision_switch_and_new_entry_argument_checks():
    """Host-only behaviour of the later additions: SM budget (data-parallel backward), fp32 parity-mode switch, and the
    argument checks of the entry points that need no device to reject bad input."""
    from mtasr_b200 import _lib, precise
    lib = _lib.load()
    total = lib.mtasr_set_sm_budget(0)
    assert total > 0
    assert lib.mtasr_set_sm_budget(total - 8) == total - 8
    assert lib.mtasr_set_sm_budget(total + 100) == total          # a budget above the device is "all SMs"
    assert lib.mtasr_set_sm_budget(-1) < 0
    assert lib.mtasr_set_sm_budget(0) == total
    assert precise.get_precision() == "bf16"
    with precise.precision("fp32"):
        assert precise.get_precision() == "fp32"
    assert precise.get_precision() == "bf16"
    with pytest.raises(ValueError):
        precise.set_precision("fp16")
    assert lib.mtasr_split_bf16(None, 8, 8, 0, 3, None, None) < 0
    assert lib.mtasr_softmax_from_logits(None, None, None, 1, 8, 8, None, None, None) < 0
    assert lib.mtasr_weightnorm_fwd(None, None, 1, 128, None, None, None) < 0
    assert lib.mtasr_attn_softmax_fwd_split(None, None, None, None, 1, 1, 8, 8, 1.0, 3, None, None) < 0
