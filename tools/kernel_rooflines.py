"""Per-kernel-class roofline table of one step from a tools/step_breakdown.py JSON (--top 1000 so that every shaped
launch is listed).  Algorithmic bytes = each operand read once + each result written once; FLOP = 2*M*K*N.

    python tools/kernel_rooflines.py gpurun_out/breakdown.json profiles/r02_kernel_rooflines.md [title]
"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kv(shape):
    out = {}
    for tok in shape.split():
        if "=" in tok:
            k, v = tok.split("=", 1)
            out[k] = v
    return out


def es(tag):
    return 2 if tag == "bf16" else 4


def work(name, shape):
    """-> (algorithmic bytes, flops, class name)"""
    s = kv(shape)
    if name.startswith("clskd_tapconv"):
        c0, c1 = (int(v) for v in s["C"].split("+"))
        M, taps, N, sf = int(s["M"]), int(s["taps"]), int(s["N"]), int(s["sf"])
        C = c0 + c1
        fl = 2.0 * M * taps * C * N
        by = M * sf * C * es(s["x"]) + M * N * es(s["y"])
        base = name.replace("clskd_", "")
        bound = "compute" if fl / 1416.6e12 > by / 6560e9 else "memory"
        return by, fl, "%s (%s-bound shapes)" % (base, bound)
    if name == "clskd_abf_mid_bwd_fold":
        B, T, F, Fy, C = (int(s[k]) for k in ("B", "T", "F", "Fy", "C"))
        rows, yrows = B * T * F, B * T * Fy
        # g, z1, logits, y_prev read once; dxp and dy_prev written
        return rows * (2 * C + 2 * C + 8) + yrows * 2 * C + rows * 2 * C + yrows * 2 * C, 0.0, "abf_mid_bwd_fold (one pass)"
    if name == "clskd_colgram":
        M, C = int(s["M"]), int(s["C"])
        return M * C * 2, 0.0, "colgram"
    if name in ("clskd_abf_mid_fwd", "clskd_abf_mid_bwd"):
        B, T, F, Fy, C = (int(s[k]) for k in ("B", "T", "F", "Fy", "C"))
        rows, yrows = B * T * F, B * T * Fy
        if name.endswith("fwd"):
            return rows * (2 * C + 2 * C + 8) + yrows * 2 * C, 0.0, "abf_mid_fwd"
        one = rows * (2 * C + 2 * C + 8) + yrows * 2 * C          # g, z1, logits, y_prev
        return 2 * one + rows * 2 * C + yrows * 2 * C, 0.0, "abf_mid_bwd (stats + apply passes)"
    if name in ("clskd_bn_act_fwd", "clskd_bn_act_bwd_stats", "clskd_bn_act_bwd_apply", "clskd_colstats"):
        M, C = int(s["M"]), int(s["C"])
        e = 2 if "bf16" in shape else 4
        k = {"clskd_bn_act_fwd": 2, "clskd_bn_act_bwd_stats": 2, "clskd_bn_act_bwd_apply": 3, "clskd_colstats": 1}[name]
        return k * M * C * e, 0.0, name.replace("clskd_", "")
    if name in ("clskd_gram_fwd_umma", "clskd_gram_fwd", "clskd_gram_bwd_umma", "clskd_gram_bwd"):
        B, K = int(s["B"]), int(s["K"])
        e = 2 if "bf16" in shape else 4
        k = 1 if "fwd" in name else 2
        return k * B * K * e, 2.0 * B * B * K * k, name.replace("clskd_", "")
    if name == "clskd_sum_n":
        return (int(s["k"]) + 1) * int(s["n"]) * (2 if "bf16" in shape else 4), 0.0, "sum_n (gradient fan-in)"
    if name in ("clskd_tapsum_fwd", "clskd_tapsum_bwd"):
        B, Ti, Fi, To, Fo, Zc = (int(s[k]) for k in ("B", "Ti", "Fi", "To", "Fo", "Zc"))
        return B * Ti * Fi * Zc * 2 + B * To * Fo * 2 * 2, 0.0, name.replace("clskd_", "")
    if name == "clskd_strided_copy4d":
        dims = [int(v) for v in shape.split("shape=[")[1].split("]")[0].split(",")]
        n = 1
        for v in dims:
            n *= v
        src = 2 if "src=bf16" in shape else 4
        dst = 2 if "dst=bf16" in shape else 4
        return n * (src + dst), 0.0, "strided_copy4d (layout / dtype adapters)"
    return None


def main():
    src, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(src)
    d = json.load(open(src))
    peaks = {"hbm_gbs": 6560.0, "bf16_tflops_sustained": 1416.6}
    try:
        peaks.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6560.0)) * 1e9
    tc = float(peaks.get("bf16_tflops_sustained", 1416.6)) * 1e12
    cls = collections.OrderedDict()
    shaped = collections.Counter()
    for r in d["tapconv_launches"]:
        w = work(r["name"], r["shape"])
        if w is None:
            continue
        by, fl, cname = w
        a = cls.setdefault(cname, {"n": 0, "ms": 0.0, "by": 0.0, "fl": 0.0})
        a["n"] += 1
        a["ms"] += r["ms"]
        a["by"] += by
        a["fl"] += fl
        shaped[r["name"]] += 1
    listed = sum(a["ms"] for a in cls.values())
    total = d["sum_of_call_ms"]
    by_entry = {e["name"]: e for e in d["by_entry_point"]}
    lines = ["# Per-kernel-class roofline table of one CLSKD step (%s)" % title, "",
             "Source: `%s` (`tools/step_breakdown.py --no-overlap --top 1000`: single-stream pass, CUDA events around every C-ABI"
             % os.path.basename(src),
             "call; 64 x 4 s, half student, bf16 policy).  Algorithmic bytes = each operand read once + each result written once;",
             "FLOP = 2*M*K*N.  Peaks: %.0f GB/s HBM copy, %.0f TFLOP/s sustained bf16 (`MEASURED_PEAKS.json`).  `frac` = achieved /"
             % (hbm / 1e9, tc / 1e12),
             "peak of the bounding resource.  Step: %.1f ms of kernel time (sum over calls), %.1f ms instrumented." %
             (total, d["instrumented_step_ms"]), "",
             "| kernel class | launches | ms | algorithmic GB | GB/s | TFLOP | TFLOP/s | bound | frac |",
             "|---|---:|---:|---:|---:|---:|---:|---|---:|"]
    for cname, a in sorted(cls.items(), key=lambda kv_: -kv_[1]["ms"]):
        t = a["ms"] * 1e-3
        gbs, tfs = a["by"] / t / 1e9, a["fl"] / t / 1e12
        compute = "compute-bound" in cname
        frac = (a["fl"] / t) / tc if compute else (a["by"] / t) / hbm
        lines.append("| %s | %d | %.2f | %.2f | %.0f | %.2f | %.0f | %s | %.2f |" % (
            cname, a["n"], a["ms"], a["by"] / 1e9, gbs, a["fl"] / 1e12, tfs, "tensor" if compute else "hbm", frac))
    for nm in ("clskd_lstm_fwd", "clskd_lstm_bwd_policy", "clskd_abf_xs2_bwd", "clskd_abf_xs2_fwd"):
        if nm in by_entry:
            e = by_entry[nm]
            note = {"clskd_lstm_fwd": "latency: T sequential steps (mma.sync recurrence)",
                    "clskd_lstm_bwd_policy": "latency: T sequential steps (mma.sync BPTT)",
                    "clskd_abf_xs2_bwd": "hbm: 8.4 GB (one pass over gout / y_prev + 24 B per row)",
                    "clskd_abf_xs2_fwd": "hbm: 4.2 GB"}[nm]
            extra = ""
            if nm == "clskd_abf_xs2_bwd":
                extra = " | 8.40 | %.0f | | | hbm | %.2f |" % (8.4e9 / (e["ms"] * 1e-3) / 1e9, 8.4e9 / (e["ms"] * 1e-3) / hbm)
            elif nm == "clskd_abf_xs2_fwd":
                extra = " | 4.18 | %.0f | | | hbm | %.2f |" % (4.18e9 / (e["ms"] * 1e-3) / 1e9, 4.18e9 / (e["ms"] * 1e-3) / hbm)
            else:
                extra = " | | | | | latency | — |"
            lines.append("| %s (%s) | %d | %.2f%s" % (nm.replace("clskd_", ""), note, e["launches"], e["ms"], extra))
            listed += e["ms"]
    lines.append("| everything else (pack / unpack tables, losses, mask, overlap-add, small statistics kernels, Adam, ...) | | %.2f | | | | | | |"
                 % (total - listed))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
