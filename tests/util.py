"""Shared helpers of the test-suite."""
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")


def golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def full_sd(sd, win_len=400, hop=100, fft_len=512, win_type="hamming"):
    """golden state_dicts omit the deterministic STFT buffers: re-create them (oracle.init_kernels)."""
    from oracle.dccrn_oracle import init_kernels
    sd = dict(sd)
    sd["stft.weight"], _ = init_kernels(win_len, hop, fft_len, win_type)
    sd["istft.weight"], sd["istft.window"] = init_kernels(win_len, hop, fft_len, win_type, invers=True)
    sd["istft.enframe"] = torch.eye(win_len)[:, None, :]
    return sd


def summ(t, n=512):
    t = t.detach().double().cpu().reshape(-1)
    step = max(1, t.numel() // n)
    return {"numel": t.numel(), "sum": float(t.sum()), "asum": float(t.abs().sum()),
            "sample": t[::step][:n].float().clone(), "step": step}


def check_summary(t, ref, rtol=1e-4, atol=1e-5, what=""):
    """compare a tensor with a golden summary (numel, sum, abs-sum, strided sample)"""
    n = ref["sample"].numel()
    t = t.detach().double().cpu().reshape(-1)
    assert t.numel() == ref["numel"], "%s: numel %d vs %d" % (what, t.numel(), ref["numel"])
    sample = t[::ref["step"]][:n].float()
    scale = max(ref["asum"] / max(ref["numel"], 1), 1e-12)          # mean |x|
    err = (sample - ref["sample"]).abs().max().item() if n else 0.0
    assert err <= atol + rtol * max(scale, ref["sample"].abs().max().item()), \
        "%s: sample max-abs err %.3e (mean|x| %.3e)" % (what, err, scale)
    asum = float(t.abs().sum())
    assert abs(asum - ref["asum"]) <= rtol * ref["asum"] + atol * ref["numel"], \
        "%s: abs-sum %.6e vs %.6e" % (what, asum, ref["asum"])
    assert abs(float(t.sum()) - ref["sum"]) <= rtol * ref["asum"] + atol * ref["numel"], \
        "%s: sum %.6e vs %.6e" % (what, float(t.sum()), ref["sum"])


def _f(v):
    return float(v.detach()) if torch.is_tensor(v) else float(v)


def rel_err(a, b):
    a, b = _f(a), _f(b)
    return abs(a - b) / max(abs(b), 1e-12)



def close(a, b, rtol=1e-5, atol=1e-5):
    """scalar comparison for values that may sit near zero (dB-valued objectives)"""
    a, b = _f(a), _f(b)
    return abs(a - b) <= atol + rtol * abs(b)


def bn_shadowed_bias(name):
    """conv biases that feed a train-mode BatchNorm have a mathematically zero gradient (the batch
    mean is subtracted): the reference's autograd value is rounding noise and is not compared."""
    return name.endswith("_conv.bias") and not name.startswith("decoder.5.")


def sample_rel_l2(t, ref):
    """relative L2 error of a tensor against a golden summary, evaluated on the stored strided sample
    (ref = {"sample", "step", "numel", ...} from make_golden_real.summ)"""
    n = ref["sample"].numel()
    t = t.detach().double().cpu().reshape(-1)
    assert t.numel() == ref["numel"], "numel %d vs %d" % (t.numel(), ref["numel"])
    s = t[::ref["step"]][:n]
    r = ref["sample"].double()
    return float((s - r).norm() / max(float(r.norm()), 1e-30))


class ParityLog:
    """Collects measured parity errors of the GPU tests into gpurun_out/parity_<tag>.json (merged over
    tests) so that the stated bounds can be read next to what was measured."""

    def __init__(self, tag="r02"):
        root = os.path.dirname(HERE)
        self.path = os.path.join(root, "gpurun_out", "parity_%s.json" % tag)

    def add(self, test, **vals):
        import json
        os.makedirs(os.path.dirname(self.path), exist_ok=True)
        data = {}
        if os.path.exists(self.path):
            try:
                data = json.load(open(self.path))
            except Exception:
                data = {}
        data.setdefault(test, {}).update(vals)
        with open(self.path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
