"""CPU tests of the C-ABI boundary: the library loads and exports every declared symbol, argument
validation fails loudly without a GPU, and the product refuses CPU tensors (no fallback)."""
import ctypes
import os
import re

import pytest
import torch

from util import bn_shadowed_bias, check_summary, close, full_sd, golden, rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol():
    from clskd_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "clskd.h")).read()
    declared = set(re.findall(r"\b(clskd_[a-z0-9_]+)\s*\(", re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)))
    assert declared == set(_lib.EXPORTS) and len(declared) >= 45
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.clskd_abi_version() == 1


def test_argument_errors_are_reported_without_a_gpu():
    from clskd_b200 import _lib
    lib = _lib.load()
    d = _lib.TapConv()          # all-null descriptor
    assert lib.clskd_tapconv_fwd(ctypes.byref(d), None) == -1
    assert b"null" in lib.clskd_last_error()
    assert lib.clskd_lstm_fwd(None, None, 1, 1, 1, 8, 1, 0, 0, 0, 0, 0, 0, None, None, None, None) == -1
    with pytest.raises(RuntimeError, match="clskd_gram_fwd failed"):
        _lib.call("clskd_gram_fwd", None, 0, 4, 16, 16, None, 0, None)


def test_no_cpu_fallback():
    import clskd_b200
    m = clskd_b200.DCCRN(rnn_units=16, use_clstm=True, kernel_num=[4, 8, 8, 16, 16, 16])
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.zeros(1, 1600))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        clskd_b200.tools_for_loss.si_snr(torch.zeros(1, 100), torch.zeros(1, 100))


