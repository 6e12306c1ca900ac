"""Per-kernel launch list of ONE full benchmark step from an ncu CSV log
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none
 --profile-from-start off --csv --log-file launches.csv python bench.py --profile-step ...`):
launches, time share and DRAM traffic of every kernel, plus the JSON that bench.py reads for `roofline.traffic`
(mean DRAM bytes per launch of the tcgen05 forward kernel over ALL its launches of the step).

Usage: python tools/ncu_step_traffic.py launches.csv out.md out.json
"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    src, out_md, out_json = sys.argv[1:4]
    lines = [l for l in open(src, errors="replace") if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    launches = {}
    for r in rows:
        key = (r["ID"], r["Kernel Name"])
        launches.setdefault(key, {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", "")) * UNIT.get(r["Metric Unit"], 1.0)
    tot_t = 0.0
    for (_, name), m in launches.items():
        short = name.split("(")[0].split("::")[-1]
        a = per.setdefault(short, {"launches": 0, "time_s": 0.0, "bytes": 0.0})
        a["launches"] += 1
        a["time_s"] += m.get("gpu__time_duration.sum", 0.0)
        a["bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        tot_t += m.get("gpu__time_duration.sum", 0.0)
    order = sorted(per.items(), key=lambda kv: -kv[1]["time_s"])
    with open(out_md, "w") as f:
        f.write("# ncu launch list of one CLSKD step (64 x 4 s, half student, bf16 policy; cold-cache, serialised: read the SHARES)\n\n")
        f.write("%d launches, %.1f ms of kernel time under ncu.\n\n" % (len(launches), tot_t * 1e3))
        f.write("| kernel | launches | ms | share | DRAM GB | GB/s |\n|---|---:|---:|---:|---:|---:|\n")
        for name, a in order[:45]:
            f.write("| `%s` | %d | %.3f | %.1f %% | %.3f | %.0f |\n" % (
                name, a["launches"], a["time_s"] * 1e3, 100 * a["time_s"] / tot_t, a["bytes"] / 1e9,
                a["bytes"] / a["time_s"] / 1e9 if a["time_s"] else 0))
    um = [v for k, v in per.items() if k.startswith("tapconv_umma")]
    n = sum(v["launches"] for v in um)
    b = sum(v["bytes"] for v in um)
    t = sum(v["time_s"] for v in um)
    own = sum(v["launches"] for k, v in per.items() if not k.startswith(("at", "vectorized", "elementwise", "Memset", "Memcpy")))
    json.dump({"entry_point": "clskd_tapconv_fwd_umma", "launches_in_step": n, "traffic_bytes_per_launch": b / max(n, 1),
               "traffic_bytes_step": b, "share_of_step_under_ncu": t / tot_t if tot_t else None,
               "source": "ncu dram__bytes_read+write of every tapconv_umma launch of one full step (tools/ncu_step_traffic.py, %s)" % src,
               "kernels_in_step": len(launches), "own_kernel_launches": own}, open(out_json, "w"), indent=1)
    print(open(out_md).read()[:2500])


if __name__ == "__main__":
    main()
