"""Warm per-entry-point GPU time of one distillation step (CUDA events around every C-ABI call) and the
host enqueue time of the step - says which kernels own the step and whether the host is the bottleneck.

    python tools/step_breakdown.py [--batch 64] [--student half] [--steps 3] [--out gpurun_out/breakdown.json]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "speech-enhancement-clskd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--student", default="half")
    ap.add_argument("--mode", default="clskd")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--out", default=None)
    ap.add_argument("--top", type=int, default=60)
    a = ap.parse_args()
    import clskd_b200
    from clskd_b200 import _lib
    from clskd_b200.distill import DistillTrainer
    dev = torch.device("cuda", 0)
    clskd_b200.set_precision("bf16")
    _lib.load()
    torch.manual_seed(1)
    teacher = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **bench.WIDTHS["teacher"]).to(dev)
    torch.manual_seed(2)
    student = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **bench.WIDTHS[a.student]).to(dev)
    L = int(a.seconds * bench.SR)
    X = 0.1 * torch.randn(a.batch, L, device=dev)
    y = 0.1 * torch.randn(a.batch, L, device=dev)
    torch.manual_seed(3)
    tr = DistillTrainer(teacher, student, mode=a.mode, example_input=X[:2])
    if a.no_overlap:
        tr.step_fn.overlap_teacher = False
        tr.step_fn.overlap_abf = False
    for _ in range(3):
        tr.train_step(X, y)
    torch.cuda.synchronize()
    # 1) host enqueue time vs GPU time, no instrumentation
    host, gpu = [], []
    for _ in range(a.steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        tr.train_step(X, y)
        e1.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3)
        gpu.append(e0.elapsed_time(e1))
    # 2) events around every C-ABI call
    orig = _lib.call
    recs = []
    phase = {"name": "fwd"}

    def call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = orig(name, *args)
        e1.record()
        shp = None
        if name.startswith("clskd_tapconv") and args and hasattr(args[0], "_obj"):
            d = args[0]._obj
            shp = "M=%d taps=%d C=%d+%d N=%d Fo=%d sf=%d x=%s y=%s" % (
                d.B * d.To * d.Fo, d.ntaps, d.c0, d.c1, d.N, d.Fo, d.sf, "bf16" if d.x_dtype else "f32",
                "bf16" if d.y_dtype else "f32")
        if name == "clskd_colstats":
            shp = "M=%d C=%d %s" % (args[2], args[3], "bf16" if args[1] else "f32")
        elif name == "clskd_bn_act_fwd":
            shp = "M=%d C=%d %s->%s" % (args[2], args[3], "bf16" if args[1] else "f32", "bf16" if args[10] else "f32")
        elif name in ("clskd_bn_act_bwd_stats", "clskd_bn_act_bwd_apply"):
            shp = "M=%d C=%d %s" % (args[4], args[5], "bf16" if args[1] else "f32")
        elif name in ("clskd_gram_fwd_umma", "clskd_gram_bwd_umma", "clskd_gram_fwd", "clskd_gram_bwd"):
            shp = "B=%d K=%d %s" % (args[2], args[3], "bf16" if args[1] else "f32")
        elif name in ("clskd_abf_mid_fwd",):
            shp = "B=%d T=%d F=%d Fy=%d C=%d" % tuple(args[3:8])
        elif name in ("clskd_abf_mid_bwd", "clskd_abf_mid_bwd_fold"):
            shp = "B=%d T=%d F=%d Fy=%d C=%d" % tuple(args[4:9])
        elif name == "clskd_colgram":
            shp = "M=%d C=%d %s" % (args[2], args[3], "bf16" if args[1] else "f32")
        elif name == "clskd_abf_mid_xs_fwd":
            shp = "B=%d T=%d F=%d Fy=%d C=%d" % tuple(args[4:9])
        elif name == "clskd_abf_mid_xs_bwd":
            shp = "B=%d T=%d F=%d Fy=%d C=%d" % tuple(args[5:10])
        elif name == "clskd_cbn_moments":
            shp = "M=%d Cc=%d" % (args[2], args[3])
        elif name in ("clskd_lstm_fwd",):
            shp = "T=%d R=%d H=%d" % (args[2], args[3], args[5])
        if name == "clskd_sum_n":
            shp = "k=%d n=%d %s" % (args[4], args[6], "bf16" if args[5] else "f32")
        elif name in ("clskd_tapsum_fwd", "clskd_tapsum_bwd"):
            shp = "B=%d Ti=%d Fi=%d To=%d Fo=%d sf=%d Zc=%d %s" % (tuple(args[2:9]) + ("bf16" if args[1] else "f32",))
        if name == "clskd_strided_copy4d":
            shp = "shape=%s src=%s/%s dst=%s/%s" % (list(args[6]), "bf16" if args[1] else "f32", list(args[2]),
                                                   "bf16" if args[4] else "f32", list(args[5]))
        recs.append((name, e0, e1, shp))
        return rc
    mods = [_lib, clskd_b200.ops, clskd_b200.clstm, clskd_b200.tools_for_model, clskd_b200.distill, clskd_b200.framework]
    for m in mods:
        if getattr(m, "call", None) is orig:
            m.call = call
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tr.train_step(X, y)
    e1.record()
    torch.cuda.synchronize()
    inst_ms = e0.elapsed_time(e1)
    tot = {}
    shaped = []
    for name, a0, a1, shp in recs:
        if shp:
            shaped.append((a0.elapsed_time(a1), name, shp))
        d = tot.setdefault(name, [0, 0.0])
        d[0] += 1
        d[1] += a0.elapsed_time(a1)
    rows = sorted(tot.items(), key=lambda kv: -kv[1][1])
    ssum = sum(v[1] for v in tot.values())
    res = {"host_enqueue_ms": host, "gpu_ms": gpu, "instrumented_step_ms": inst_ms, "sum_of_call_ms": ssum,
           "overlap_teacher": not a.no_overlap,
           "by_entry_point": [{"name": k, "launches": v[0], "ms": round(v[1], 3)} for k, v in rows]}
    print(json.dumps({k: res[k] for k in ("host_enqueue_ms", "gpu_ms", "instrumented_step_ms", "sum_of_call_ms")}))
    for k, v in rows:
        print("%-32s %5d %9.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / ssum))
    print("---- tapconv launches by time")
    for ms, name, shp in sorted(shaped, reverse=True)[:a.top]:
        if not name.startswith("clskd_tapconv"):
            print("%8.3f ms               %-26s %s" % (ms, name, shp))
            continue
        d = dict(kv.split("=") for kv in shp.split())
        c0, c1 = d["C"].split("+")
        fl = 2.0 * int(d["M"]) * int(d["taps"]) * (int(c0) + int(c1)) * int(d["N"])
        print("%8.3f ms %7.1f TF/s  %-26s %s" % (ms, fl / ms / 1e9, name, shp))
    res["tapconv_launches"] = [{"ms": ms, "name": n, "shape": sh} for ms, n, sh in sorted(shaped, reverse=True)]
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
