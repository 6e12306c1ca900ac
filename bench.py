"""bench.py - audio-seconds/second of one CLSKD distillation step (DCCRN-CL teacher -> half-width
DCCRN student) on N B200s, plus the e2e / roofline / cpu_baseline evidence the driver asks for.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = teacher forward (frozen, eval) + student forward + ABF cross-layer fusion (encoder and
decoder sides) + 14 SPKD terms + log-STFT-magnitude base loss + backward + gradient all-reduce (N>1)
+ Adam, on a synthetic batch of B x 4 s of 16 kHz audio per GPU (BASELINE.json configs[1]; weak
scaling: the per-GPU batch is fixed, so N=8 is configs[3]'s global batch 512).
`--impl reference` times the CPU oracle's restatement of the same step (the reference is a Python
package that does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "speech-enhancement-clskd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

SR = 16000
WIDTHS = {
    "teacher": dict(kernel_num=[32, 64, 128, 256, 256, 256], rnn_units=256),
    "half": dict(kernel_num=[16, 32, 64, 128, 128, 128], rnn_units=128),
    "quarter": dict(kernel_num=[8, 16, 32, 64, 64, 64], rnn_units=64),
}
# algorithmic forward GFLOP per 4 s utterance; scaled linearly in seconds
FWD_GFLOP = {"teacher": 69.11, "half": 17.75, "quarter": 4.87}
ABF_GFLOP = {"half": 60.6, "quarter": 29.5}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--student", default="half", choices=["half", "quarter"])
    ap.add_argument("--mode", default="clskd", choices=["clskd", "spkd_all", "spkd", "mse", "stft"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=2, help="utterances in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap-abf", type=int, default=None, help="1/0: encoder-side ABF chain on a second stream")
    ap.add_argument("--dump-launches", default=None, help="write per-launch shapes/times of the profiled kernel here")
    ap.add_argument("--profile-kernel", default="auto",
                    help="C-ABI entry point timed with CUDA events for the roofline object")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU arm
def cpu_step_fn(student_w, batch, seconds, mode):
    """One distillation step of the ORACLE on the host cores (test infrastructure used as the
    reported CPU baseline only)."""
    from oracle import dccrn_oracle as D
    from oracle import losses_oracle as LO
    torch.set_num_threads(os.cpu_count() or 1)
    t_sd = D.make_state_dict(**WIDTHS["teacher"], seed=1, randomize_bn=False)
    s_sd = D.make_state_dict(**WIDTHS[student_w], seed=2, randomize_bn=False)
    params = {k: v.clone().requires_grad_(True) for k, v in s_sd.items()
              if v.is_floating_point() and "running" not in k and not k.startswith(("stft.", "istft."))}
    s_live = dict(s_sd)
    s_live.update(params)
    g = torch.Generator().manual_seed(0)
    L = int(seconds * SR)
    X, y = 0.1 * torch.randn(batch, L, generator=g), 0.1 * torch.randn(batch, L, generator=g)
    abf_e = abf_d = None
    if mode == "clskd":
        abf_e, abf_d = _oracle_abf_sds(t_sd, s_sd, X[:1, :SR // 2], g)
        for sd in (abf_e, abf_d):
            for k in list(sd):
                if sd[k].is_floating_point() and "running" not in k:
                    sd[k] = sd[k].requires_grad_(True)
                    params["abf." + str(id(sd)) + k] = sd[k]
    opt = torch.optim.Adam(list(params.values()), lr=6e-4)

    def step():
        opt.zero_grad()
        loss, _ = LO.clskd_step_loss(t_sd, s_live, X, y, abf_e, abf_d, mode=mode)
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step, batch * seconds


def _oracle_abf_sds(t_sd, s_sd, x, g):
    """random ABF weights (kaiming_uniform(a=1) like framework.py:194-195) shaped for these models"""
    from oracle import dccrn_oracle as D
    tt, st = {}, {}
    with torch.no_grad():
        D.dccrn_forward(t_sd, x, taps=tt)
        D.dccrn_forward(s_sd, x, taps=st)
    out = []
    for kind in ("encoder", "decoder"):
        smaps, tmaps = st[kind], tt[kind]
        in_ch = [m.shape[1] for m in smaps]
        out_ch = [m.shape[1] for m in tmaps]
        if kind == "encoder":           # deepest first
            in_ch, out_ch = in_ch[::-1], out_ch[::-1]
        mid = min(512, in_ch[0])          # framework.py:238: channels of the deepest student map
        sd = {}
        for i, (ci, co) in enumerate(zip(in_ch, out_ch)):
            p = "abfs.%d." % i
            b1, b2 = (6.0 / (2 * ci)) ** 0.5 * 1.0, (6.0 / (2 * mid * 9)) ** 0.5
            sd[p + "conv1.0.weight"] = (torch.rand(mid, ci, 1, 1, generator=g) * 2 - 1) * b1
            sd[p + "conv2.0.weight"] = (torch.rand(co, mid, 3, 3, generator=g) * 2 - 1) * b2
            for c, n in (("conv1.1.", mid), ("conv2.1.", co)):
                sd[p + c + "weight"], sd[p + c + "bias"] = torch.ones(n), torch.zeros(n)
                sd[p + c + "running_mean"], sd[p + c + "running_var"] = torch.zeros(n), torch.ones(n)
            if i > 0:
                sd[p + "att_conv.0.weight"] = (torch.rand(2, 2 * mid, 1, 1, generator=g) * 2 - 1) * (1.0 / (2 * mid)) ** 0.5
                sd[p + "att_conv.0.bias"] = torch.zeros(2)
        out.append(sd)
    return out


def time_cpu(step, n, warm=1):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, audio_s = cpu_step_fn(args.student, args.cpu_batch, args.seconds, args.mode)
    dt = time_cpu(step, max(1, args.steps), warm=max(1, min(args.warmup, 1)))
    val = audio_s / dt
    cores = os.cpu_count() or 1
    sample = ("oracle (CPU port of the reference step) on a bounded sample of the workload: %d x %.0f s utterances "
              "per step instead of %d, %d threads") % (args.cpu_batch, args.seconds, args.batch, cores)
    line = {
        "impl": "reference", "metric": "audio-seconds/sec per CLSKD distill step", "value": val,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),       # the arm's workload; each CPU step is a bounded sample of it
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    return {"workload": "DCCRN-CL teacher -> %s-width DCCRN student %s distill step, %d x %.0f s 16 kHz per GPU"
                        % (args.student, args.mode.upper(), batch, args.seconds),
            "per_gpu_batch": batch, "segment_s": args.seconds, "sample_rate": SR, "student": args.student,
            "mode": args.mode, "precision_policy": args.precision,
            "l2": "working set (GBs of activations per step) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


class KernelTimer:
    """CUDA-event timing of ONE C-ABI entry point inside the timed region (for the roofline)."""

    def __init__(self, lib_mod, name):
        self.mod, self.name, self.pairs, self.flops, self.bytes = lib_mod, name, [], 0.0, 0.0
        self.enabled = False
        self._orig = lib_mod.call
        self.counts = {}
        self.shapes = []

    def install(self, work_fn, shape_fn=None):
        orig, me = self._orig, self

        def call(name, *a):
            if me.enabled:
                me.counts[name] = me.counts.get(name, 0) + 1
            if me.enabled and name == me.name:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rc = orig(name, *a)
                e1.record()
                me.pairs.append((e0, e1))
                f, b = work_fn(a)
                me.shapes.append(shape_fn(a) if shape_fn else None)
                me.flops += f
                me.bytes += b
                return rc
            return orig(name, *a)
        import clskd_b200
        for m in (self.mod, clskd_b200.ops, clskd_b200.clstm, clskd_b200.tools_for_model, clskd_b200.distill):
            if getattr(m, "call", None) is orig:
                m.call = call

    def total_ms(self):
        return sum(a.elapsed_time(b) for a, b in self.pairs)


def tapconv_work(a):
    """algorithmic FLOPs / bytes of one tapconv launch from its descriptor"""
    if not hasattr(a[0], "_obj"):
        return 0.0, 0.0          # not a tap-list contraction (only its time is reported)
    d = a[0]._obj
    M = d.B * d.To * d.Fo
    K = d.ntaps * (d.c0 + d.c1)
    xe = 2 if d.x_dtype == 1 else 4
    ye = 2 if d.y_dtype == 1 else 4
    flops = 2.0 * M * K * d.N
    byts = float(d.B * d.Ti * d.Fi * (d.c0 + d.c1) * xe + M * d.N * ye + K * d.N * xe)
    return flops, byts


def tapconv_shape(a):
    if not hasattr(a[0], "_obj"):
        return {"M": 0, "taps": 0, "C": 0, "N": 0, "Fo": 0, "sf": 0, "out": "-"}
    d = a[0]._obj
    return {"M": d.B * d.To * d.Fo, "taps": d.ntaps, "C": d.c0 + d.c1, "N": d.N, "Fo": d.Fo, "sf": d.sf,
            "out": "bf16" if d.y_dtype == 1 else "f32"}


def run_ours(args):
    import torch.distributed as dist
    import clskd_b200
    from clskd_b200 import _lib
    from clskd_b200.distill import DistillTrainer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    clskd_b200.set_precision(args.precision)
    _lib.load()

    torch.manual_seed(1)
    teacher = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS["teacher"]).to(dev)
    torch.manual_seed(2)
    student = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS[args.student]).to(dev)
    L = int(args.seconds * SR)
    B = args.batch
    g = torch.Generator().manual_seed(100 + rank)
    X_h = (0.1 * torch.randn(B, L, generator=g)).pin_memory()
    y_h = (0.1 * torch.randn(B, L, generator=g)).pin_memory()
    X, y = X_h.to(dev), y_h.to(dev)
    torch.manual_seed(3)                       # identical ABF init on every rank
    tr = DistillTrainer(teacher, student, mode=args.mode, example_input=X[:2])

    if args.overlap_abf is not None:
        tr.step_fn.overlap_abf = bool(args.overlap_abf)
    kname = args.profile_kernel
    if kname == "auto":
        kname = "clskd_tapconv_fwd_umma" if args.precision == "bf16" else "clskd_tapconv_fwd"
    kt = KernelTimer(_lib, kname)
    kt.install(tapconv_work, tapconv_shape)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tr.train_step(X, y)
    barrier()
    # ---------------- timed region 1: inputs resident in HBM
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = tr.train_step(X, y)
    e1.record()
    barrier()
    launches = _lib.launch_count - launches0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    # ---------------- timed region 2: end to end through the public API with host buffers
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e2.record()
    for _ in range(args.steps):
        Xd = X_h.to(dev, non_blocking=True)
        yd = y_h.to(dev, non_blocking=True)
        loss_host = float(tr.train_step(Xd, yd))      # device->host read of the step's loss
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # ---------------- region 3: kernel timing for the roofline object.  Per-launch CUDA events only measure the
    # kernel when nothing else shares the SMs, so this pass runs the SAME steps with the side streams (frozen
    # teacher, encoder-side ABF chain) disabled; `value` / `e2e` above come from the untouched regions.
    ov = (tr.step_fn.overlap_teacher, tr.step_fn.overlap_abf)
    tr.step_fn.overlap_teacher = tr.step_fn.overlap_abf = False
    tr.train_step(X, y)
    e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    kt.enabled = True
    e4.record()
    for _ in range(args.steps):
        tr.train_step(X, y)
    e5.record()
    barrier()
    kt.enabled = False
    ms_serial = e4.elapsed_time(e5)
    tr.step_fn.overlap_teacher, tr.step_fn.overlap_abf = ov
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    audio_s = world * B * args.seconds * args.steps
    value = audio_s / (ms / 1e3)
    e2e = audio_s / (ms_e2e / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        n_launch = max(1, len(kt.pairs))
        k_ms = kt.total_ms()
        achieved = kt.flops / (k_ms / 1e3) / 1e12 if k_ms > 0 else 0.0
        traffic = None
        try:      # dram bytes per launch of the same kernel from the committed `ncu --set full` capture
            tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
            if tj.get("entry_point") == kname:
                traffic = tj.get("traffic_bytes_per_launch")
        except Exception:
            pass
        roofline = {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic,
                    "launches_timed": len(kt.pairs), "avg_launch_ms": k_ms / n_launch,
                    "share_of_step": k_ms / ms_serial if ms_serial > 0 else None,
                    "timed_in": "separate pass of the same %d steps with the side streams disabled (single-stream "
                                "execution, %.2f ms/step): per-launch CUDA events are distorted when two streams "
                                "share the SMs" % (args.steps, ms_serial / args.steps),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF (of fallback)",
                    "algorithmic_gflop_per_launch": kt.flops / n_launch / 1e9}
        step_gflop = B * args.seconds / 4.0 * (FWD_GFLOP["teacher"] + 3 * FWD_GFLOP[args.student] +
                                               (3 * ABF_GFLOP[args.student] if args.mode == "clskd" else 0.0))
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            step, a_s = cpu_step_fn(args.student, args.cpu_batch, args.seconds, args.mode)
            dt = time_cpu(step, 1, warm=1)
            cores = os.cpu_count() or 1
            cpu = {"value": a_s / dt, "unit": "audio-s/s", "cores": cores, "kind": "port",
                   "sample": "oracle (CPU port of the reference step) on %d x %.0f s utterances, 1 warm-up + 1 timed step, %d threads"
                             % (args.cpu_batch, args.seconds, cores)}
        line = {
            "metric": "audio-seconds/sec per CLSKD distill step", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": 2 * B * L * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "launches_by_entry_point": kt.counts,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "step_tflops": step_gflop / 1e3 / (ms / args.steps / 1e3) if ms > 0 else None,
            "loss": float(loss), "loss_e2e": loss_host,
            "umma_launches": clskd_b200.ops.umma_launches, "core_launches": clskd_b200.ops.core_launches,
            "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
        }
        print(json.dumps(line), flush=True)
        if args.dump_launches:
            rows = []
            for (e_a, e_b), shp in zip(kt.pairs, kt.shapes):
                ms_l = e_a.elapsed_time(e_b)
                fl = 2.0 * shp["M"] * shp["taps"] * shp["C"] * shp["N"]
                rows.append(dict(shp, ms=ms_l, tflops=fl / (ms_l / 1e3) / 1e12 if ms_l > 0 else 0.0))
            with open(args.dump_launches, "w") as f:
                json.dump(rows, f)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
