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


def test_second_source_padding_rule_and_new_entry_point_validation():
    """Host-side pieces of this round's entry points that run without a GPU: the padded extent of a narrow second source in
    the packed weight of clskd_tapconv_fwd_umma (the library and the CPU model of the ABI must agree: the Python side packs
    with one, the kernel contracts with the other), and loud argument errors of the new calls."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from cabi_emu import EmuLib
    from clskd_b200 import _lib
    lib, emu = _lib.load(), EmuLib()
    for c0 in (8, 16, 24, 32, 48, 64, 96, 128, 192, 256):
        for c1 in (0, 8, 16, 24, 32, 64, 128):
            got = lib.clskd_tapconv_umma_c1p(c0, c1)
            assert got == emu.clskd_tapconv_umma_c1p(c0, c1), (c0, c1)
            assert got >= c1 and got % 16 == 0 and (c1 != 0 or got == 0)
    assert lib.clskd_tapconv_umma_c1p(128, 16) == 64 and lib.clskd_tapconv_umma_c1p(128, 128) == 128
    assert lib.clskd_tapconv_umma_c1p(32, 16) == 32 and lib.clskd_tapconv_umma_c1p(16, 16) == 16
    assert lib.clskd_colgram_supported(_lib.BF16, 16) == 1 and lib.clskd_colgram_supported(_lib.BF16, 128) == 1
    assert lib.clskd_colgram_supported(_lib.F32, 16) == 0 and lib.clskd_colgram_supported(_lib.BF16, 24) == 0
    assert lib.clskd_colgram(None, _lib.BF16, 10, 16, None, None, None) == -1 and b"null" in lib.clskd_last_error()
    assert lib.clskd_abf_fold_stats(None, None, None, 128, 16, None, None, None) == -1
    assert lib.clskd_abf_fold_dgrad(None, None, None, None, None, 10, 1, 128, 16, 64, None, None, None) == -1
    assert lib.clskd_abf_fold_dw1(None, None, None, None, None, None, None, None, 10, 1, 128, 16, None, None) == -1


def test_no_cpu_fallback():
    import clskd_b200
    m = clskd_b200.DCCRN(rnn_units=16, use_clstm=True, kernel_num=[4, 8, 8, 16, 16, 16])
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        m(torch.zeros(1, 1600))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        clskd_b200.tools_for_loss.si_snr(torch.zeros(1, 100), torch.zeros(1, 100))




def test_reference_copy_recipe_is_byte_identical():
    """oracle/make_ref.py copies the reference's hot-path modules unmodified (checked where the reference exists)"""
    import hashlib
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref, dst = "/root/reference", os.path.join(root, "oracle", "_ref")
    if not (os.path.isdir(ref) and os.path.isdir(dst)):
        pytest.skip("needs /root/reference and oracle/_ref (container only)")
    for line in open(os.path.join(dst, "MANIFEST.sha256")):
        h, f = line.split()
        assert hashlib.sha256(open(os.path.join(ref, f), "rb").read()).hexdigest() == h, f
        assert hashlib.sha256(open(os.path.join(dst, f), "rb").read()).hexdigest() == h, f


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints ONE JSON line with the
    contract keys: same metric / unit / config as our arm, impl=reference, cpu_baseline and a zero-copy e2e."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--cpu-batch", "1", "--seconds", "0.5"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True
    assert d["metric"] == "audio-seconds/sec per CLSKD distill step" and d["value"] > 0
    assert d["config"]["workload"].startswith("DCCRN-CL teacher") and d["config"]["per_gpu_batch"] == 64
    # the reference's own modules when a copy is available (oracle/_ref or /root/reference), else the oracle port
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data"):
        assert k in d, k
