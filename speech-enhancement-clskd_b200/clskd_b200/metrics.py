"""Batched validation metrics on the GPU (reference: KnowledgeDistillation.validation_step, distill.py:150-200,
which loops over the batch and calls asteroid's `get_metrics` on the CPU for every utterance).

One library pass (clskd_pair_moments) produces the second moments of each (estimate, clean) and (mixture, clean)
pair; SI-SDR (scale-invariant SDR with mean removal, the definition asteroid's get_metrics takes from
pb_bss_eval.evaluation.si_sdr), SNR and their improvements over the mixture follow in closed form per utterance.
STOI / PESQ / BSS-eval SDR-SIR-SAR are CPU libraries (pystoi, pesq, mir_eval) outside this path.
"""
import math

import torch

from . import ops
from .ops import call


def pair_moments(a, b):
    """fp64 [B, 5] = per-utterance (sum a, sum b, sum a^2, sum a*b, sum b^2) of two [B, L] waveform batches"""
    ops._require_cuda(a, b)
    if a.dim() == 3:
        a = a.squeeze(1)
    if b.dim() == 3:
        b = b.squeeze(1)
    if a.shape != b.shape or a.dim() != 2:
        raise ValueError("pair_moments: expected two [B, L] batches, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    a, b = a.float(), b.float()
    if a.stride(1) != 1:
        a = ops.dense(a)
    if b.stride(1) != 1:
        b = ops.dense(b)
    B, L = a.shape
    out = torch.empty((B, 5), dtype=torch.float64, device=a.device)
    call("clskd_pair_moments", a.data_ptr(), b.data_ptr(), B, L, a.stride(0), b.stride(0), out.data_ptr(), ops._stream())
    return out


def _si_sdr_from_moments(m, L, eps=0.0):
    """rows of (sum e, sum r, sum e^2, sum e*r, sum r^2) -> SI-SDR [dB] of estimate e against reference r with mean
    removal: alpha = <e0, r0> / <r0, r0>, 10 log10(|alpha r0|^2 / |e0 - alpha r0|^2)"""
    se, sr, see, ser, srr = (m[:, i] for i in range(5))
    cee = see - se * se / L
    cer = ser - se * sr / L
    crr = srr - sr * sr / L
    target = cer * cer / crr.clamp_min(1e-300)
    noise = (cee - target).clamp_min(0.0)
    return 10.0 * torch.log10((target + eps) / (noise + eps).clamp_min(1e-300))


def _snr_from_moments(m):
    """plain SNR [dB] of estimate against reference: 10 log10(|r|^2 / |r - e|^2)"""
    see, ser, srr = m[:, 2], m[:, 3], m[:, 4]
    return 10.0 * torch.log10(srr.clamp_min(1e-300) / (srr - 2 * ser + see).clamp_min(1e-300))


def batch_metrics(mix, clean, est):
    """Per-utterance and mean metrics of a validation batch, all computed on the device:
    si_sdr / input_si_sdr / si_sdr_imp and snr / input_snr / snr_imp (dB).  Returns (dict of means as python floats,
    dict of per-utterance fp64 tensors)."""
    L = clean.shape[-1]
    m_est = pair_moments(est, clean)
    m_mix = pair_moments(mix, clean)
    per = {"si_sdr": _si_sdr_from_moments(m_est, L), "input_si_sdr": _si_sdr_from_moments(m_mix, L),
           "snr": _snr_from_moments(m_est), "input_snr": _snr_from_moments(m_mix)}
    per["si_sdr_imp"] = per["si_sdr"] - per["input_si_sdr"]
    per["snr_imp"] = per["snr"] - per["input_snr"]
    stacked = torch.stack([per[k].mean() for k in sorted(per)]).cpu()          # ONE device->host read
    means = {k: float(v) for k, v in zip(sorted(per), stacked)}
    return means, per
