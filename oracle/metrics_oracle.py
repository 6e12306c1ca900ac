"""ORACLE (test infrastructure only) - validation metrics of the reference's training loop.

distill.py:150-200 calls asteroid.metrics.get_metrics(mix, clean, estimate, sample_rate=16000), which is NOT under
/root/reference (asteroid fork 0.6.1dev, un-vendored) and delegates SI-SDR to
pb_bss_eval.evaluation.si_sdr.  Its published algorithm is restated here in numpy - PARITY UNPINNED for this
metric (no asteroid / pb_bss_eval in this image to run against):

    reference, estimation zero-mean;  alpha = <estimation, reference> / <reference, reference>
    si_sdr = 10 log10( |alpha reference|^2 / |alpha reference - estimation|^2 )
"""
import numpy as np


def si_sdr(reference, estimation):
    reference = np.asarray(reference, dtype=np.float64)
    estimation = np.asarray(estimation, dtype=np.float64)
    reference = reference - reference.mean(-1, keepdims=True)
    estimation = estimation - estimation.mean(-1, keepdims=True)
    alpha = (estimation * reference).sum(-1, keepdims=True) / (reference ** 2).sum(-1, keepdims=True)
    proj = alpha * reference
    noise = estimation - proj
    return 10 * np.log10((proj ** 2).sum(-1) / (noise ** 2).sum(-1))


def snr(reference, estimation):
    reference = np.asarray(reference, dtype=np.float64)
    estimation = np.asarray(estimation, dtype=np.float64)
    return 10 * np.log10((reference ** 2).sum(-1) / ((reference - estimation) ** 2).sum(-1))


def batch_metrics(mix, clean, est):
    out = {"si_sdr": si_sdr(clean, est), "input_si_sdr": si_sdr(clean, mix), "snr": snr(clean, est),
           "input_snr": snr(clean, mix)}
    out["si_sdr_imp"] = out["si_sdr"] - out["input_si_sdr"]
    out["snr_imp"] = out["snr"] - out["input_snr"]
    return out
