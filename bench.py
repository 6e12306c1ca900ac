"""bench.py - audio-seconds/second of one CLSKD distillation step (DCCRN-CL teacher -> half-width
DCCRN student) on N B200s, plus the e2e / roofline / cpu_baseline evidence the driver asks for.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = teacher forward (frozen, eval) + student forward + ABF cross-layer fusion (encoder and
decoder sides) + 14 SPKD terms + log-STFT-magnitude base loss + backward + gradient all-reduce (N>1)
+ Adam, on a synthetic batch of B x 4 s of 16 kHz audio per GPU (BASELINE.json configs[1]).

Default scaling is WEAK (64 utterances per GPU; N = 8 is configs[3]'s global batch 512).  At N = 2 / 4 the
line also carries `extra.global_batch_512`: the same step at 512 / N utterances per GPU (configs[3] as
written); `--global-batch G` runs that strong-scaling form as the main measurement.

At N = 1 the line also carries `extra` objects, each measured by a child process of this script on the same
GPU after the main measurement (own clock record each; skip with --no-extras):
  spkd_all   - BASELINE configs[2]: SPKD on all encoder / decoder / LSTM layer pairs, no ABF
  faithful   - the reference-faithful CLSKD step (train-mode teacher BatchNorm under autograd, student
               forward twice: distill.py:49-50,77,85,100)
  quarter    - CLSKD with the reference's own quarter-width student (config.py:47-48)
  rtf        - BASELINE configs[4]: student inference RTF, 256 x 60 s, 16 kHz and the 8 kHz variant
  config1    - BASELINE configs[0]: teacher forward + SI-SNR, batch 8 x 4 s, on the host cores

`--impl reference` times the reference's OWN unmodified modules (oracle/_ref, copied by oracle/make_ref.py)
assembled into the same step on the host cores (kind "reference"); if oracle/_ref is absent it times the
oracle port (kind "port").
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
EXTRAS = ("spkd_all", "faithful", "quarter", "rtf", "config1")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="utterances per GPU")
    ap.add_argument("--global-batch", type=int, default=None,
                    help="total utterances over all GPUs (strong scaling: per-GPU batch = G / N)")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--student", default="half", choices=["half", "quarter"])
    ap.add_argument("--mode", default="clskd", choices=["clskd", "spkd_all", "spkd", "mse", "stft"])
    ap.add_argument("--faithful", action="store_true",
                    help="reference-faithful step: train-mode teacher BatchNorm under autograd, student forward twice")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="utterances in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--only", default=None, choices=list(EXTRAS), help="child mode: measure one extra object")
    ap.add_argument("--overlap-abf", type=int, default=None, help="1/0: encoder-side ABF chain on a second stream")
    ap.add_argument("--dump-launches", default=None, help="write per-launch shapes/times of the profiled kernel here")
    ap.add_argument("--profile-step", action="store_true",
                    help="after the warm-up run ONE step between cudaProfilerStart/Stop and exit (for `ncu --profile-from-start off`)")
    ap.add_argument("--profile-kernel", default="auto",
                    help="C-ABI entry point timed with CUDA events for the roofline object")
    return ap.parse_args()


# ------------------------------------------------------------------------------------ CPU arm
def _cpu_inputs(batch, seconds):
    g = torch.Generator().manual_seed(0)
    L = int(seconds * SR)
    return 0.1 * torch.randn(batch, L, generator=g), 0.1 * torch.randn(batch, L, generator=g)


def cpu_step_fn(student_w, batch, seconds, mode, faithful=False):
    """One distillation step on the host cores -> (step, audio seconds per step, kind).
    kind "reference": the reference's own unmodified modules (oracle/_ref or /root/reference);
    kind "port": the oracle's restatement (when no copy of the reference is available)."""
    torch.set_num_threads(os.cpu_count() or 1)
    X, y = _cpu_inputs(batch, seconds)
    from oracle import ref_shim
    if ref_shim.available() and mode in ("clskd", "spkd_all"):
        from oracle import ref_step
        return ref_step.make_step(WIDTHS["teacher"], WIDTHS[student_w], X, y, mode=mode, faithful=faithful), \
            batch * seconds, "reference"
    return _port_step_fn(student_w, X, y, mode), batch * seconds, "port"


def _port_step_fn(student_w, X, y, mode):
    from oracle import dccrn_oracle as D
    from oracle import losses_oracle as LO
    t_sd = D.make_state_dict(**WIDTHS["teacher"], seed=1, randomize_bn=False)
    s_sd = D.make_state_dict(**WIDTHS[student_w], seed=2, randomize_bn=False)
    params = {k: v.clone().requires_grad_(True) for k, v in s_sd.items()
              if v.is_floating_point() and "running" not in k and not k.startswith(("stft.", "istft."))}
    s_live = dict(s_sd)
    s_live.update(params)
    abf_e = abf_d = None
    if mode == "clskd":
        tt, st = {}, {}
        with torch.no_grad():
            D.dccrn_forward(t_sd, X[:1, :SR // 2], taps=tt)
            D.dccrn_forward(s_sd, X[:1, :SR // 2], taps=st)
        e_in, e_out = [m.shape[1] for m in st["encoder"]], [m.shape[1] for m in tt["encoder"]]
        d_in, d_out = [m.shape[1] for m in st["decoder"]][::-1], [m.shape[1] for m in tt["decoder"]][::-1]
        abf_e, abf_d = LO.make_abf_state_dict(e_in, e_out, 7, False), LO.make_abf_state_dict(d_in, d_out, 8, False)
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
    return step


def time_cpu(step, n, warm=1):
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    return (time.perf_counter() - t0) / n


def cpu_sample_text(kind, batch, seconds, full_batch, cores, steps, warm):
    who = ("the reference's own unmodified modules (oracle/_ref) assembled into the step" if kind == "reference"
           else "oracle (CPU port of the reference step)")
    return ("%s on a bounded sample of the workload: %d x %.0f s utterances per step instead of %d, %d warm-up + %d "
            "timed steps, %d threads" % (who, batch, seconds, full_batch, warm, steps, cores))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, audio_s, kind = cpu_step_fn(args.student, args.cpu_batch, args.seconds, args.mode, args.faithful)
    warm = 1
    t0 = time.perf_counter()
    step()                                           # warm-up (also sizes the run)
    first = time.perf_counter() - t0
    steps = max(1, args.steps)
    if first * steps > 240.0:                        # keep the whole arm within a few minutes
        steps = max(1, int(240.0 / first))
    dt = time_cpu(step, steps, warm=0)
    val = audio_s / dt
    cores = os.cpu_count() or 1
    batch = per_gpu_batch(args, int(os.environ.get("WORLD_SIZE", "1")))
    line = {
        "impl": "reference", "metric": "audio-seconds/sec per CLSKD distill step", "value": val,
        "unit": "audio-s/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "steps_timed": steps, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, batch),       # the arm's workload; each CPU step is a bounded sample of it
        "cpu_baseline": {"value": val, "unit": "audio-s/s", "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, args.cpu_batch, args.seconds, batch, cores, steps, warm)},
        "e2e": {"value": val, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def per_gpu_batch(args, world):
    if args.global_batch:
        if args.global_batch % world:
            raise SystemExit("--global-batch must be a multiple of the number of GPUs")
        return args.global_batch // world
    return args.batch


def workload_config(args, batch):
    return {"workload": "DCCRN-CL teacher -> %s-width DCCRN student %s distill step%s, %d x %.0f s 16 kHz per GPU"
                        % (args.student, args.mode.upper(), " (reference-faithful)" if args.faithful else "", batch,
                           args.seconds),
            "per_gpu_batch": batch, "global_batch": args.global_batch, "segment_s": args.seconds, "sample_rate": SR,
            "student": args.student, "mode": args.mode, "faithful": bool(args.faithful),
            "precision_policy": args.precision,
            "l2": "working set (GBs of activations per step) exceeds the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region.  nvidia-smi takes up to a second to print its
    first line (longer on a multi-GPU box), so start() is called BEFORE the warm-up steps and begin() marks the start
    of the timed region; stop() keeps the samples whose own timestamps fall inside [begin, stop]."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.t0 = [], None, index, None

    def start(self):
        try:
            self.t_start = time.time()
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    @staticmethod
    def _ts(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except Exception:
            return None

    def _read(self):
        for line in self.proc.stdout:
            cols = [c.strip() for c in line.split(",")]
            self.rows.append((self._ts(cols[0]) if cols else None, cols[1:]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.12)                       # let the sample that covers the end of the region arrive
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.th.join(timeout=2)
        t0 = self.t0 if self.t0 is not None else self.t_start
        inside = [r for ts, r in self.rows if ts is not None and t0 <= ts <= t1 + 0.1]
        note = None
        if not inside and self.rows:
            # region shorter than the sampling period: the sample nearest to it (the GPU was under the same load
            # during the warm-up steps just before)
            near = min((abs((ts if ts is not None else t1) - t1), i) for i, (ts, _) in enumerate(self.rows))[1]
            inside = [self.rows[near][1]]
            note = "no sample inside the %.0f ms region: nearest sample used" % ((t1 - t0) * 1e3)
        sm = sorted(int(r[0]) for r in inside if r and r[0].isdigit())
        mx = [int(r[1]) for r in inside if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].startswith("Active") for r in inside)]
        out = {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
               "reasons": reasons, "samples": len(inside)}
        if note:
            out["note"] = note
        return out


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


def _alg_k(d):
    """algorithmic contraction depth of one launch.  The split-bf16 launches of the fp32-input GEMMs stage
    [x_hi | x_lo | x_hi] (both sources are the same tensor, c0 = 2 Kp, c1 = Kp): they EXECUTE 3 Kp but the
    algorithm is one K-deep fp32 GEMM, so only Kp counts."""
    if d.c1 and d.x0 == d.x1 and d.c0 == 2 * d.c1 and d.ntaps == 1:
        return d.c1
    return d.ntaps * (d.c0 + d.c1)


def tapconv_work(a):
    """algorithmic FLOPs / bytes of one tapconv launch from its descriptor"""
    if not hasattr(a[0], "_obj"):
        return 0.0, 0.0          # not a tap-list contraction (only its time is reported)
    d = a[0]._obj
    M = d.B * d.To * d.Fo
    K = _alg_k(d)
    xe = 2 if d.x_dtype == 1 else 4
    ye = 2 if d.y_dtype == 1 else 4
    flops = 2.0 * M * K * d.N
    cin = K // d.ntaps
    byts = float(d.B * d.Ti * d.Fi * cin * xe + M * d.N * ye + K * d.N * xe)
    return flops, byts


def tapconv_shape(a):
    if not hasattr(a[0], "_obj"):
        return {"M": 0, "taps": 0, "C": 0, "N": 0, "Fo": 0, "sf": 0, "out": "-"}
    d = a[0]._obj
    return {"M": d.B * d.To * d.Fo, "taps": d.ntaps, "C": _alg_k(d) // d.ntaps, "N": d.N, "Fo": d.Fo, "sf": d.sf,
            "out": "bf16" if d.y_dtype == 1 else "f32", "split": bool(_alg_k(d) != d.ntaps * (d.c0 + d.c1))}


def _load_traffic(kname):
    """measured DRAM bytes per launch of the roofline kernel: `ncu` capture of every launch of that kernel in
    ONE full benchmark step (tools/ncu_traffic.py -> profiles/r02_roofline_traffic.json)"""
    for fn in ("r02_roofline_traffic.json",):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
            if tj.get("entry_point") == kname:
                return tj.get("traffic_bytes_per_launch"), tj.get("source")
        except Exception:
            pass
    return None, None


def _tuning_stats(_lib):
    """shapes the forward kernel's per-shape autotuner timed in this process / how many kept the round-1 configuration"""
    import ctypes
    try:
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        _lib.load().clskd_tuning_stats(ctypes.byref(a), ctypes.byref(b))
        return {"shapes_tuned": a.value, "kept_round1_config": b.value}
    except Exception:
        return None


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
    B = per_gpu_batch(args, world)

    def host_batch(n):
        g = torch.Generator().manual_seed(100 + rank)
        return (0.1 * torch.randn(n, L, generator=g)).pin_memory(), (0.1 * torch.randn(n, L, generator=g)).pin_memory()
    X_h, y_h = host_batch(B)
    X, y = X_h.to(dev), y_h.to(dev)
    torch.manual_seed(3)                       # ABF init (rank 0's is broadcast to every rank by the trainer)
    if args.faithful:
        teacher.train()
    tr = DistillTrainer(teacher, student, mode=args.mode, faithful=args.faithful, example_input=X[:2])

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

    def timed(n, Xd, yd):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            out = tr.train_step(Xd, yd)
        e1.record()
        barrier()
        return e0.elapsed_time(e1), out

    sampler = ClockSampler(local)
    if rank == 0 and not args.profile_step:
        sampler.start()                      # (before the warm-up: nvidia-smi needs up to a second to produce its first line)
    for _ in range(args.warmup):
        tr.train_step(X, y)
    barrier()
    if args.profile_step:
        tr.step_fn.overlap_teacher = tr.step_fn.overlap_abf = False      # ncu serialises kernels anyway
        tr.train_step(X, y)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        tr.train_step(X, y)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"profiled": "one %s step, %d x %.0f s" % (args.mode, B, args.seconds)}), flush=True)
        return
    # ---------------- timed region 1: inputs resident in HBM
    if rank == 0:
        sampler.begin()
    launches0 = _lib.launch_count
    ms, loss = timed(args.steps, X, y)
    launches = _lib.launch_count - launches0
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
    kt.enabled = True
    ms_serial, _ = timed(args.steps, X, y)
    kt.enabled = False
    tr.step_fn.overlap_teacher, tr.step_fn.overlap_abf = ov
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    audio_s = world * B * args.seconds * args.steps
    value = audio_s / (ms / 1e3)
    e2e = audio_s / (ms_e2e / 1e3)
    max_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---------------- configs[3] as written: global batch 512 on 2 / 4 GPUs (256 / 128 utterances per GPU)
    extra = {}
    if world in (2, 4) and not args.global_batch and not args.no_extras and args.mode == "clskd":
        Bg = 512 // world
        try:
            del X, y
            torch.cuda.empty_cache()
            Xg_h, yg_h = host_batch(Bg)
            Xg, yg = Xg_h.to(dev), yg_h.to(dev)
            samp = ClockSampler(local)
            if rank == 0:
                samp.start()
            for _ in range(2):
                tr.train_step(Xg, yg)
            if rank == 0:
                samp.begin()
            nst = 3
            msg, _ = timed(nst, Xg, yg)
            ck = samp.stop() if rank == 0 else None
            tg = torch.tensor([msg], dtype=torch.float64, device=dev)
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
            extra["global_batch_512"] = {
                "config": "BASELINE configs[3]: CLSKD step, global batch 512 x %.0f s over %d GPUs (%d per GPU)"
                          % (args.seconds, world, Bg),
                "value": 512 * args.seconds * nst / (float(tg[0]) / 1e3), "unit": "audio-s/s", "ms_per_step": float(tg[0]) / nst,
                "steps": nst, "warmup": 2, "scaling": "strong", "clocks": ck,
                "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
            del Xg, yg
        except Exception as e:      # noqa: BLE001  (an out-of-memory here must not lose the main measurement)
            extra["global_batch_512"] = {"error": repr(e)[:300]}

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
        traffic, traffic_src = _load_traffic(kname)
        roofline = {"kernel": kname, "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf if peak_tf else None, "traffic": traffic, "traffic_source": traffic_src,
                    "launches_timed": len(kt.pairs), "avg_launch_ms": k_ms / n_launch,
                    "share_of_step": k_ms / ms_serial if ms_serial > 0 else None,
                    "timed_in": "separate pass of the same %d steps with the side streams disabled (single-stream "
                                "execution, %.2f ms/step): per-launch CUDA events are distorted when two streams "
                                "share the SMs" % (args.steps, ms_serial / args.steps),
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PF (of fallback)",
                    "flop_accounting": "algorithmic 2*M*K*N per launch; split-bf16 launches of the fp32-input GEMMs "
                                       "count K (not the 3K they execute)",
                    "algorithmic_gflop_per_launch": kt.flops / n_launch / 1e9,
                    "algorithmic_mb_per_launch": kt.bytes / n_launch / 1e6}
        step_gflop = B * args.seconds / 4.0 * (FWD_GFLOP["teacher"] + 3 * FWD_GFLOP[args.student] +
                                               (3 * ABF_GFLOP[args.student] if args.mode == "clskd" else 0.0))
        line = {
            "metric": "audio-seconds/sec per CLSKD distill step", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": workload_config(args, B),
            "e2e": {"value": e2e, "unit": "audio-s/s", "h2d_bytes_per_step": 2 * B * L * 4, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "launches_by_entry_point": kt.counts,
            "clocks": clocks, "roofline": roofline, "cpu_baseline": None,
            "step_tflops": step_gflop / 1e3 / (ms / args.steps / 1e3) if ms > 0 else None,
            "loss": float(loss), "loss_e2e": loss_host,
            "umma_launches": clskd_b200.ops.umma_launches, "core_launches": clskd_b200.ops.core_launches,
            "max_mem_gb": max_mem, "autotune": _tuning_stats(_lib),
        }
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
    if rank != 0:
        return
    # ---------------- free the GPU, then the CPU baseline and the extra objects (children of this process)
    del tr, teacher, student
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline and world == 1:
        step, a_s, kind = cpu_step_fn(args.student, args.cpu_batch, args.seconds, args.mode, args.faithful)
        dt = time_cpu(step, 1, warm=1)
        cores = os.cpu_count() or 1
        line["cpu_baseline"] = {"value": a_s / dt, "unit": "audio-s/s", "cores": cores, "kind": kind,
                                "sample": cpu_sample_text(kind, args.cpu_batch, args.seconds, B, cores, 1, 1)}
    if world == 1 and not args.no_extras and not args.only and args.mode == "clskd" and not args.faithful \
            and args.student == "half" and not args.global_batch:
        for name in EXTRAS:
            extra[name] = run_child(name, args)
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ extra objects
def run_child(name, args):
    """measure one extra object in a child process (a failure there cannot lose the main line)"""
    cmd = [sys.executable, os.path.abspath(__file__), "--only", name, "--precision", args.precision,
           "--seconds", str(args.seconds)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": "no JSON from child (rc %d): %s" % (r.returncode, (r.stderr or "")[-300:])}
    except Exception as e:      # noqa: BLE001
        return {"error": repr(e)[:300]}


def _short_step_bench(args, mode, student, faithful, steps=5, warmup=3, batch=64):
    import clskd_b200
    from clskd_b200 import _lib
    from clskd_b200.distill import DistillTrainer
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    clskd_b200.set_precision(args.precision)
    _lib.load()
    torch.manual_seed(1)
    teacher = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS["teacher"]).to(dev)
    torch.manual_seed(2)
    stu = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS[student]).to(dev)
    L = int(args.seconds * SR)
    g = torch.Generator().manual_seed(100)
    X_h, y_h = (0.1 * torch.randn(batch, L, generator=g)).pin_memory(), (0.1 * torch.randn(batch, L, generator=g)).pin_memory()
    X, y = X_h.to(dev), y_h.to(dev)
    torch.manual_seed(3)
    if faithful:
        teacher.train()
    tr = DistillTrainer(teacher, stu, mode=mode, faithful=faithful, example_input=X[:2])
    samp = ClockSampler(0)
    samp.start()
    for _ in range(warmup):
        tr.train_step(X, y)
    torch.cuda.synchronize()
    samp.begin()
    n0 = _lib.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.train_step(X, y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(steps):
        loss = float(tr.train_step(X_h.to(dev, non_blocking=True), y_h.to(dev, non_blocking=True)))
    e3.record()
    torch.cuda.synchronize()
    ck = samp.stop()
    audio = batch * args.seconds * steps
    return {"value": audio / (ms / 1e3), "unit": "audio-s/s", "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "e2e": {"value": audio / (e2.elapsed_time(e3) / 1e3), "unit": "audio-s/s",
                    "h2d_bytes_per_step": 2 * batch * L * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": (_lib.launch_count - n0), "clocks": ck, "loss": loss,
            "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
            "umma_launches": clskd_b200.ops.umma_launches, "core_launches": clskd_b200.ops.core_launches}


def run_only(args):
    name = args.only
    if name == "spkd_all":
        out = _short_step_bench(args, "spkd_all", "half", False)
        out["config"] = "BASELINE configs[2]: SPKD step on all encoder/decoder/LSTM layer pairs (no ABF), 64 x %.0f s, half student" % args.seconds
    elif name == "faithful":
        out = _short_step_bench(args, "clskd", "half", True)
        out["config"] = ("reference-faithful CLSKD step (train-mode teacher BatchNorm under autograd, student forward twice; "
                         "distill.py:49-50,77,85,100), 64 x %.0f s, half student" % args.seconds)
    elif name == "quarter":
        out = _short_step_bench(args, "clskd", "quarter", False)
        out["config"] = "CLSKD step with the reference's quarter-width student (config.py:47-48), 64 x %.0f s" % args.seconds
    elif name == "rtf":
        out = run_rtf(args)
    else:
        out = run_config1(args)
    print(json.dumps(out), flush=True)


def run_rtf(args, batch=256, seconds=60.0, chunk_s=None):
    """BASELINE configs[4]: student inference RTF (eval, no_grad) on batch x 60 s, 16 kHz and the 8 kHz variant
    (same STFT / model hyper-parameters on half as many samples), quarter and half students."""
    import clskd_b200
    clskd_b200.set_precision(args.precision)
    dev = torch.device("cuda", 0)
    out = {"config": "BASELINE configs[4]: DCCRN-CL student inference, %d x %.0f s utterances, eval/no_grad" % (batch, seconds),
           "unit": "RTF = seconds of compute per second of audio (lower is better)", "results": []}
    samp = ClockSampler(0)
    samp.start()
    for student in ("quarter", "half"):
        torch.manual_seed(0)
        model = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS[student]).to(dev).eval()
        for variant, sr in (("16k", 16000), ("8k", 8000)):
            Ls = int(seconds * sr)
            x_h = (0.1 * torch.randn(batch, Ls)).pin_memory()
            x = x_h.to(dev)
            run = (lambda inp: model.enhance_streaming(inp)) if hasattr(model, "enhance_streaming") else \
                (lambda inp: model(inp, is_feat=True))
            with torch.no_grad():
                for _ in range(2):
                    run(x)
                torch.cuda.synchronize()
                torch.cuda.reset_peak_memory_stats()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(3):
                    yv = run(x)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 3
                e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e2.record()
                for _ in range(2):
                    y_h = run(x_h.to(dev, non_blocking=True)).float().cpu()
                e3.record()
                torch.cuda.synchronize()
                ms_e2e = e2.elapsed_time(e3) / 2
            audio = batch * seconds
            out["results"].append({"student": student, "variant": variant, "samples": Ls, "ms_per_batch": ms,
                                   "rtf": ms / 1e3 / audio, "audio_s_per_s": audio / (ms / 1e3),
                                   "rtf_e2e": ms_e2e / 1e3 / audio, "streaming": hasattr(model, "enhance_streaming"),
                                   "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30})
            del x, yv, y_h
            torch.cuda.empty_cache()
        del model
    out["clocks"] = samp.stop()
    return out


def run_config1(args):
    """BASELINE configs[0]: DCCRN-CL teacher forward + SI-SNR, batch 8 x 4 s, eval/no_grad, on the host cores
    through the reference's own modules (oracle/_ref), next to the same workload through the C ABI on the GPU."""
    from oracle import ref_shim
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    X, y = _cpu_inputs(8, args.seconds)
    out = {"config": "BASELINE configs[0]: DCCRN-CL teacher forward + SI-SNR, 8 x %.0f s, eval, no_grad" % args.seconds,
           "unit": "audio-s/s"}
    if ref_shim.available():
        from oracle import ref_step
        step = ref_step.make_config1(WIDTHS["teacher"], X, y)
        kind = "reference"
    else:
        from oracle import dccrn_oracle as D
        from oracle import losses_oracle as LO
        sd = D.make_state_dict(**WIDTHS["teacher"], seed=1)

        def step():
            with torch.no_grad():
                return float(-LO.si_snr(D.dccrn_forward(sd, X)[-1], y))
        kind = "port"
    dt = min(time_cpu(step, 1, warm=1), time_cpu(step, 1, warm=0), time_cpu(step, 1, warm=0))
    out["cpu"] = {"value": 8 * args.seconds / dt, "ms_per_batch": dt * 1e3, "cores": cores, "kind": kind,
                  "sample": "the full configuration (8 utterances), 1 warm-up, best of 3"}
    if torch.cuda.is_available():
        import clskd_b200
        dev = torch.device("cuda", 0)
        res = {}
        for pol in ("fp32", "bf16"):
            clskd_b200.set_precision(pol)
            torch.manual_seed(1)
            m = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS["teacher"]).to(dev).eval()
            Xp, yp = X.pin_memory(), y.pin_memory()
            with torch.no_grad():
                for _ in range(3):
                    m.loss(m(Xp.to(dev), is_feat=True), yp.to(dev), loss_mode="SI-SNR")
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    v = float(m.loss(m(Xp.to(dev, non_blocking=True), is_feat=True), yp.to(dev, non_blocking=True),
                                     loss_mode="SI-SNR"))
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res[pol] = {"value": 8 * args.seconds / (ms / 1e3), "ms_per_batch": ms, "loss": v,
                        "timed": "end to end: H2D of the batch + forward + SI-SNR + D2H of the loss"}
        out["gpu_e2e"] = res
    return out


def main():
    args = parse()
    if args.only:
        run_only(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
