"""Summarise an `ncu --set full` report of one kernel: per-launch DRAM traffic, duration, tensor-pipe
and memory-throughput percentages.  Writes a markdown summary and a small JSON that bench.py reads
for the `roofline.traffic` field.
Usage: python tools/ncu_traffic.py report.ncu-rep|raw.csv kernel_tag out.md out.json"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "rd",
    "dram__bytes_write.sum": "wr",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_pct",
    "launch__registers_per_thread": "regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def main():
    rep, tag, out_md, out_json = sys.argv[1:5]
    if rep.endswith(".csv"):        # already exported with `ncu -i rep --page raw --csv`
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    tens = [h for h in hdr if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h]
    recs = []
    for r in rows[2:]:
        rec = {}
        for k, name in WANT.items():
            if k in idx and r[idx[k]] not in ("", "n/a", "no data"):
                v = float(r[idx[k]].replace(",", ""))
                rec[name] = v * UNIT.get(units[idx[k]], 1.0)
        for h in tens:
            try:
                rec.setdefault("tensor_any", 0.0)
                rec["tensor_any"] = max(rec["tensor_any"], float(r[idx[h]].replace(",", "")))
            except ValueError:
                pass
        recs.append(rec)
    n = len(recs)
    tot_dur = sum(r.get("dur", 0) for r in recs)
    tot_traffic = sum(r.get("rd", 0) + r.get("wr", 0) for r in recs)
    wavg = lambda key: sum(r.get(key, 0) * r.get("dur", 0) for r in recs) / tot_dur if tot_dur else 0.0
    summary = {"kernel": tag, "launches": n, "traffic_bytes_per_launch": tot_traffic / max(n, 1),
               "avg_duration_us": tot_dur / max(n, 1) * 1e6, "dram_pct_time_weighted": wavg("dram_pct"),
               "l2_pct_time_weighted": wavg("l2_pct"), "tensor_pipe_pct_time_weighted": wavg("tensor_any"),
               "sm_pct_time_weighted": wavg("sm_pct"), "registers": recs[0].get("regs") if recs else None}
    with open(out_json, "w") as f:
        json.dump(summary, f, indent=1)
    with open(out_md, "w") as f:
        f.write("# ncu --set full summary: %s\n\n" % tag)
        for k, v in summary.items():
            f.write("- %s: %s\n" % (k, v))
        f.write("\n| # | dur us | dram MB | dram %% | L2 %% | tensor %% | sm %% |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for i, r in enumerate(recs):
            f.write("| %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f |\n" % (
                i, r.get("dur", 0) * 1e6, (r.get("rd", 0) + r.get("wr", 0)) / 1e6, r.get("dram_pct", 0),
                r.get("l2_pct", 0), r.get("tensor_any", 0), r.get("sm_pct", 0)))
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
