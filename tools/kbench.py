"""Per-shape A/B timing of the tcgen05 forward kernel (clskd_tapconv_fwd_umma) under its tuning overrides
(clskd_set_tuning): operand-reuse mode (one box per tap / time-grouped patches / full halo patch), resident
weights, CTAs per SM.  Shapes are the step's worst launches (profiles/r01_step_breakdown_v24.json).  Each variant's
output is checked against the legacy mode's (same operands; only the accumulation order of the taps differs).

    python tools/kbench.py [--out gpurun_out/kbench.json] [--quick]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "speech-enhancement-clskd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

# (name, B, T, F, C, N, kind, stats)   kind: "3x3", "1x1", "5x2s2" (stride-2 in f: Fo = F/2), "t2" (2 time taps, sf=1)
SHAPES = [
    ("abf_conv2 F128 128->32", 64, 643, 128, 128, 32, "3x3", True),
    ("abf_conv2_dgrad F128 32->128", 64, 643, 128, 32, 128, "3x3", False),
    ("abf_conv1 F128 16->128", 64, 643, 128, 16, 128, "1x1", True),
    ("abf_conv2 F64 128->64", 64, 643, 64, 128, 64, "3x3", True),
    ("abf_conv2_dgrad F64 64->128", 64, 643, 64, 64, 128, "3x3", False),
    ("abf_conv2 F32 128->128", 64, 643, 32, 128, 128, "3x3", True),
    ("abf_conv2 F16 128->256", 64, 643, 16, 128, 256, "3x3", True),
    ("abf_conv1 F64 32->128", 64, 643, 64, 32, 128, "1x1", True),
    ("abf F256 128->32 1x1", 64, 644, 256, 128, 32, "1x1", False),
    ("student enc1 F128->64 16->32", 64, 643, 128, 16, 32, "5x2s2", True),
    ("student enc2 F64->32 32->64", 64, 643, 64, 32, 64, "5x2s2", True),
    ("teacher enc3 F32->16 128->256", 64, 643, 32, 128, 256, "5x2s2", False),
    ("teacher enc4 F16->8 256->256", 64, 643, 16, 256, 256, "5x2s2", False),
    ("teacher dec phase F8 512->256 t2x3", 64, 643, 8, 512, 256, "dec6", False),
    ("teacher dec phase F16 512->128 t2x3", 64, 643, 16, 512, 128, "dec6", False),
    ("abf_conv1 F256 32->128", 64, 644, 256, 32, 128, "1x1", True),
    ("abf_conv1 F32 64->128", 64, 643, 32, 64, 128, "1x1", True),
]


def taps_of(kind):
    if kind == "3x3":
        return [(dt, df) for df in (-1, 0, 1) for dt in (-1, 0, 1)], 1
    if kind == "1x1":
        return [(0, 0)], 1
    if kind == "5x2s2":
        return [(dt, df) for df in (-2, -1, 0, 1, 2) for dt in (-1, 0)], 2
    if kind == "dec6":          # even sub-pixel phase of the transposed conv: 3 f taps x 2 t taps, stride 1
        return [(dt, df) for df in (-1, 0, 1) for dt in (-1, 0)], 1
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--variants", default=None, help="comma-separated subset of variant names")
    ap.add_argument("--shapes", default=None, help="comma-separated shape indices")
    ap.add_argument("--nostats", action="store_true", help="time every shape without the batch-statistics epilogue")
    a = ap.parse_args()
    from clskd_b200 import _lib
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    # (mode, resident-off, one-cta, v1, split, no-autotune)
    variants = [("v1", (0, 0, 0, 1, 0, 1)), ("legacy", (1, 1, 0, 0, 0, 1)), ("legacy+res", (1, 0, 0, 0, 0, 1)),
                ("time", (2, 1, 0, 0, 0, 1)), ("full", (3, 1, 0, 0, 0, 1)), ("full 1box", (3, 1, 0, 0, 1, 1)),
                ("full+res", (3, 0, 0, 0, 0, 1)), ("auto", (0, 0, 0, 0, 0, 1)), ("tuned", (0, 0, 0, 0, 0, 0))]
    if a.variants:
        variants = [v for v in variants if v[0] in a.variants.split(",")]
    results = []
    st = torch.cuda.current_stream().cuda_stream
    shapes = SHAPES[:4] if a.quick else SHAPES
    if a.shapes:
        shapes = [SHAPES[int(i)] for i in a.shapes.split(",")]
    for name, B, T, F, C, N, kind, stats in shapes:
        taps, sf = taps_of(kind)
        stats = stats and not a.nostats
        Fo = F // sf
        g = torch.Generator(device="cpu").manual_seed(1)
        xs = [(0.5 * torch.randn(B, T, F, C, generator=g, dtype=torch.float32)).to(dev).bfloat16() for _ in range(1)]
        xs += [xs[0].clone(), xs[0].clone()]
        w = (torch.randn(len(taps), N, C, generator=g) / (C * len(taps)) ** 0.5).to(dev).bfloat16()
        y = torch.empty(B, T, Fo, N, dtype=torch.bfloat16, device=dev)
        ssum = torch.zeros(2, N, dtype=torch.float64, device=dev)
        d = _lib.TapConv()
        d.x1 = None
        d.x0_sB, d.x0_sT, d.x0_sF = T * F * C, F * C, C
        d.c0, d.c1 = C, 0
        d.B, d.To, d.Fo, d.Ti, d.Fi = B, T, Fo, T, F
        d.sf, d.ntaps = sf, len(taps)
        for j, (dt, df) in enumerate(taps):
            d.dt[j], d.df[j] = dt, df
        d.w, d.bias, d.N = w.data_ptr(), None, N
        d.y = y.data_ptr()
        d.y_sB, d.y_sT, d.y_sF = T * Fo * N, Fo * N, N
        d.x_dtype, d.y_dtype, d.accumulate = _lib.BF16, _lib.BF16, 0
        if stats:
            d.stats_sum, d.stats_sumsq = ssum[0].data_ptr(), ssum[1].data_ptr()
        M = B * T * Fo
        flops = 2.0 * M * len(taps) * C * N
        byts = B * T * F * C * 2 + M * N * 2
        ref = None
        row = {"shape": name, "M": M, "taps": len(taps), "C": C, "N": N, "Fo": Fo, "sf": sf, "stats": stats,
               "floor_ms": max(flops / (peaks.get("bf16_tflops_sustained", 1416.6) * 1e12),
                               byts / (peaks.get("hbm_gbs", 6560.0) * 1e9)) * 1e3, "variants": {}}
        for vname, tune in variants:
            lib.clskd_set_tuning(0, tune[0])
            lib.clskd_set_tuning(1, tune[1])
            lib.clskd_set_tuning(2, tune[2] if len(tune) > 2 else 0)
            lib.clskd_set_tuning(3, tune[3] if len(tune) > 3 else 0)
            lib.clskd_set_tuning(4, tune[4] if len(tune) > 4 else 0)
            lib.clskd_set_tuning(5, tune[5] if len(tune) > 5 else 0)
            try:
                for i in range(3):
                    d.x0 = xs[i % 3].data_ptr()
                    _lib.call("clskd_tapconv_fwd_umma", ctypes.byref(d), st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(a.reps):
                    d.x0 = xs[i % 3].data_ptr()
                    _lib.call("clskd_tapconv_fwd_umma", ctypes.byref(d), st)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / a.reps
                ssum.zero_()
                d.x0 = xs[0].data_ptr()
                _lib.call("clskd_tapconv_fwd_umma", ctypes.byref(d), st)
                torch.cuda.synchronize()
                cur = (y.float().clone(), ssum.clone())
                if ref is None:
                    ref = cur
                    err, serr = 0.0, 0.0
                else:
                    err = float((cur[0] - ref[0]).abs().max() / ref[0].abs().max())
                    serr = float(((cur[1] - ref[1]).abs() / (ref[1].abs() + 1e-6)).max()) if stats else 0.0
                row["variants"][vname] = {"ms": ms, "tflops": flops / ms / 1e9, "gbs": byts / ms / 1e6,
                                          "max_rel_diff_vs_legacy": err, "stats_rel_diff": serr}
            except RuntimeError as e:
                row["variants"][vname] = {"error": str(e)[:200]}
        for k in range(6):
            lib.clskd_set_tuning(k, 0)
        results.append(row)
        print("%-38s floor %.3f ms | " % (name, row["floor_ms"]) + " | ".join(
            "%s %.3f%s" % (k, v.get("ms", -1), "" if v.get("max_rel_diff_vs_legacy", 0) < 2e-2 else " !!DIFF %.2g" % v["max_rel_diff_vs_legacy"])
            for k, v in row["variants"].items()), flush=True)
        del xs, y, w
        torch.cuda.empty_cache()
    if a.out:
        with open(a.out, "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
