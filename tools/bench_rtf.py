"""Student-inference real-time factor (BASELINE.json configs[4]): DCCRN-CL student, eval mode, no_grad,
B utterances x S seconds at 16 kHz (and the "8 kHz variant" = same STFT/model on half as many samples,
SURVEY 8d).  RTF = wall time / audio seconds; prints one JSON line per setting.
Usage: python tools/bench_rtf.py [--batch 256] [--seconds 60] [--student half|quarter] [--precision bf16]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "speech-enhancement-clskd_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402

WIDTHS = {"half": dict(kernel_num=[16, 32, 64, 128, 128, 128], rnn_units=128),
          "quarter": dict(kernel_num=[8, 16, 32, 64, 64, 64], rnn_units=64)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--student", default="quarter", choices=list(WIDTHS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import clskd_b200
    clskd_b200.set_precision(args.precision)
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = clskd_b200.DCCRN(masking_mode="E", use_clstm=True, **WIDTHS[args.student]).to(dev).eval()
    for name, sr in (("16k", 16000), ("8k", 8000)):
        L = int(args.seconds * sr)
        x = 0.1 * torch.randn(args.batch, L, device=dev)
        with torch.no_grad():
            for _ in range(args.warmup):
                model(x, is_feat=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                y = model(x, is_feat=True)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        audio_s = args.batch * args.seconds
        print(json.dumps({"metric": "DCCRN student inference RTF", "variant": name, "student": args.student,
                          "batch": args.batch, "seconds": args.seconds, "samples": L, "ms_per_batch": ms,
                          "rtf": (ms / 1e3) / audio_s, "audio_s_per_s": audio_s / (ms / 1e3),
                          "precision": args.precision, "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}),
              flush=True)
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
