"""Time-domain objectives (reference: tools_for_loss.py:16-108) on fused reduction kernels:
each loss is one or two passes over the waveforms with fp64 accumulation plus a scalar epilogue,
instead of ~10 elementwise/reduce launches."""
import torch

from .ops import WaveLossFn, dense


def remove_dc(data):
    return data - torch.mean(data, -1, keepdim=True)


def l2_norm(s1, s2):
    return torch.sum(s1 * s2, -1, keepdim=True)


def si_snr(s1, s2, eps=1e-8):
    """mean_b 10 log10(|a s2|^2 / (|s1 - a s2|^2 + eps) + eps), a = <s1,s2>/(<s2,s2>+eps);
    s1 = estimate, s2 = reference; no mean removal (tools_for_loss.py:37-47)."""
    return WaveLossFn.apply(s1, s2, "si_snr", eps)


def sdr(s1, s2, eps=1e-8):
    """mean_b 10 log10(<s1,s1>^2 / (<s1-s2,s1-s2>^2 + eps))  (tools_for_loss.py:30-34)."""
    return WaveLossFn.apply(s1, s2, "sdr", eps)


def si_sdr(reference, estimation, eps=1e-8):
    """10 log10(mean_b(|proj|^2/|noise|^2 + eps) + eps) (tools_for_loss.py:50-97)."""
    return WaveLossFn.apply(reference, estimation, "si_sdr", eps)


def mse(a, b):
    """F.mse_loss(a, b, reduction='mean') (DCCRN.py:260-261)."""
    return WaveLossFn.apply(a, b, "mse", 0.0)


class rmse(torch.nn.Module):
    def forward(self, y_true, y_pred):
        m = torch.mean((y_pred - y_true) ** 2, axis=-1)
        return torch.mean(torch.sqrt(m + 1e-7))
