"""Timing of the tcgen05 weight-gradient kernel (clskd_tapconv_wgrad_umma) on the step's small-channel shapes.

    python tools/wgbench.py [--shapes 0,1] [--reps 10]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "speech-enhancement-clskd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
from kbench import taps_of  # noqa: E402

# (name, B, T, F, C, N, kind)
SHAPES = [
    ("student enc1 wgrad 16->32 5x2s2", 64, 643, 128, 16, 32, "5x2s2"),
    ("student enc2 wgrad 32->64 5x2s2", 64, 643, 64, 32, 64, "5x2s2"),
    ("abf conv2 wgrad F128 128->32 3x3", 64, 643, 128, 128, 32, "3x3"),
    ("abf conv1 wgrad F128 16->128 1x1", 64, 643, 128, 16, 128, "1x1"),
    ("dec phase wgrad F64 64->16 t2x3", 64, 643, 64, 64, 16, "dec6"),
    ("dec phase wgrad F16 256->64 t2x3", 64, 643, 16, 256, 64, "dec6"),
    ("abf conv2 wgrad F64 128->64 3x3", 64, 643, 64, 128, 64, "3x3"),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default=None)
    ap.add_argument("--reps", type=int, default=10)
    a = ap.parse_args()
    from clskd_b200 import _lib
    _lib.load()
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    shapes = SHAPES if not a.shapes else [SHAPES[int(i)] for i in a.shapes.split(",")]
    for name, B, T, F, C, N, kind in shapes:
        taps, sf = taps_of(kind)
        Fo = F // sf
        g = torch.Generator(device="cpu").manual_seed(1)
        x = (0.5 * torch.randn(B, T, F, C, generator=g)).to(dev).bfloat16()
        dy = (0.5 * torch.randn(B, T, Fo, N, generator=g)).to(dev).bfloat16()
        dw = torch.zeros(len(taps), C, N, dtype=torch.float32, device=dev)
        d = _lib.TapConv()
        d.x0, d.x1 = x.data_ptr(), None
        d.x0_sB, d.x0_sT, d.x0_sF = T * F * C, F * C, C
        d.c0, d.c1 = C, 0
        d.B, d.To, d.Fo, d.Ti, d.Fi = B, T, Fo, T, F
        d.sf, d.ntaps = sf, len(taps)
        for j, (dt, df) in enumerate(taps):
            d.dt[j], d.df[j] = dt, df
        d.w, d.bias, d.N = dw.data_ptr(), None, N
        d.y = dy.data_ptr()
        d.y_sB, d.y_sT, d.y_sF = T * Fo * N, Fo * N, N
        d.x_dtype, d.y_dtype, d.accumulate = _lib.BF16, _lib.BF16, 0
        for _ in range(3):
            _lib.call("clskd_tapconv_wgrad_umma", ctypes.byref(d), st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps):
            _lib.call("clskd_tapconv_wgrad_umma", ctypes.byref(d), st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        M = B * T * Fo
        byts = x.numel() * 2 + dy.numel() * 2
        print("%-36s M=%d  %.3f ms  %.1f TF/s  HBM floor %.3f ms" % (name, M, ms, 2.0 * M * len(taps) * C * N / ms / 1e9,
                                                                  byts / 6.56e9))


if __name__ == "__main__":
    main()
