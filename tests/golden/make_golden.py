"""Generate tests/golden/*.pt by running the UNMODIFIED reference modules from /root/reference.

Container-only (the reference tree does not exist on the GPU box); the produced fixtures are
committed.  Run:  python tests/golden/make_golden.py

What is pinned (all through the reference's own classes, imported by oracle/ref_shim.py):
  dccrn.pt   - local DCCRN (DCCRN.py:14-257) forward in eval and train mode on fixed weights/input:
               full enhanced waveform, summaries of masks / spectra / every feature tap
               (feature_extraction.py:3-50), BatchNorm running-stat updates;
               DCCRN.loss modes 'SI-SNR', 'MSE', 'SDR', 'SI-SDR' (DCCRN.py:259-267).
  losses.pt  - si_snr / sdr / si_sdr (tools_for_loss.py), SPKDLoss, STFTLoss /
               MultiResolutionSTFTLoss (framework.py; torch.stft shimmed with return_complex=True
               because torch>=2 rejects the reference's call), ABF / ReviewKD outputs on fixed
               weights (framework.py:176-263; `.cuda()` neutralised on this CPU box).
  step.pt    - the CLSKD / SPKD-all / SPKD / MSE / STFT training-step losses
               (distill.py:72-148 etc. restated over the local DCCRN as in SURVEY 3.1, built from
               the reference's modules) and the student-gradient summaries from torch autograd.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.dccrn_oracle import make_state_dict  # noqa: E402

TEACHER = dict(kernel_num=[8, 16, 32, 64, 64, 64], rnn_units=64)      # reference student width
STUDENT = dict(kernel_num=[4, 8, 16, 32, 32, 32], rnn_units=32)
B, L = 2, 8000


def summ(t, n=512):
    t = t.detach().double().reshape(-1)
    step = max(1, t.numel() // n)
    return {"numel": t.numel(), "sum": float(t.sum()), "asum": float(t.abs().sum()),
            "sample": t[::step][:n].float().clone(), "step": step}


def main():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    mods = ref_shim.load()
    RefDCCRN = mods["DCCRN"].DCCRN
    fw, fe, tl = mods["framework"], mods["feature_extraction"], mods["tools_for_loss"]

    # --- torch-version shims (documented in the module docstring)
    _stft = torch.stft

    def stft_compat(*a, **k):
        k.setdefault("return_complex", True)
        out = _stft(*a, **k)
        return torch.view_as_real(out) if out.is_complex() else out
    torch.stft = stft_compat
    torch.nn.Module.cuda = lambda self, device=None: self
    torch.Tensor.cuda = lambda self, *a, **k: self

    def build(cfg, seed):
        sd = make_state_dict(cfg["kernel_num"], cfg["rnn_units"], seed=seed)
        m = RefDCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
        # keep the reference's OWN stft/istft buffers (they pin the oracle's init_kernels)
        own = {k: v.clone() for k, v in m.state_dict().items() if k.startswith("stft.") or k.startswith("istft.")}
        sd.update(own)
        m.load_state_dict(sd, strict=True)
        return m, sd

    teacher, t_sd = build(TEACHER, 1)
    student, s_sd = build(STUDENT, 2)
    g = torch.Generator().manual_seed(123)
    X = 0.1 * torch.randn(B, L, generator=g)
    y = 0.1 * torch.randn(B, L, generator=g)

    # ------------------------------------------------------------------ dccrn.pt
    # the STFT buffers are deterministic (init_kernels): store summaries of the reference's own
    # buffers instead of 3 MB of copies; loaders re-create them with oracle.init_kernels
    def strip(sd):
        return {k: v for k, v in sd.items() if not (k.startswith("stft.") or k.startswith("istft."))}
    ref_sd = teacher.state_dict()
    out = {"teacher_cfg": TEACHER, "student_cfg": STUDENT, "t_sd": strip(t_sd), "s_sd": strip(s_sd), "X": X, "y": y,
           "buffers": {k: summ(ref_sd[k], 2048) for k in ("stft.weight", "istft.weight", "istft.window", "istft.enframe")}}
    for name, model in (("teacher", teacher), ("student", student)):
        for mode in ("eval", "train"):
            model.train(mode == "train")
            model.load_state_dict(t_sd if name == "teacher" else s_sd)
            ext = fe.DCCRN(model)
            with torch.no_grad():
                mr, mi, re, im, wav = model(X)
            ext.remove_hook()
            fm = ext.feature_maps
            rec = {"wav": wav.clone(), "mask_real": summ(mr), "mask_imag": summ(mi), "real": summ(re),
                   "imag": summ(im),
                   "encoder": [summ(t) for t in fm["encoder"]], "decoder": [summ(t) for t in fm["decoder"]],
                   "encoder_shapes": [tuple(t.shape) for t in fm["encoder"]],
                   "decoder_shapes": [tuple(t.shape) for t in fm["decoder"]],
                   "clstm": [summ(t) for t in fm["clstm"][0]],
                   "clstm_shapes": [tuple(t.shape) for t in fm["clstm"][0]]}
            if mode == "train":
                rec["running"] = {k: v.clone() for k, v in model.state_dict().items() if "running_" in k}
            else:
                rec["loss"] = {lm: float(model.loss(wav, y, re, im, loss_mode=lm)) for lm in ("SI-SNR", "MSE", "SDR", "SI-SDR")}
            out["%s_%s" % (name, mode)] = rec
    torch.save(out, os.path.join(HERE, "dccrn.pt"))

    # ------------------------------------------------------------------ losses.pt
    lo = {}
    a = 0.3 * torch.randn(4, 3000, generator=g)
    b = a + 0.2 * torch.randn(4, 3000, generator=g)
    lo["a"], lo["b"] = a, b
    lo["si_snr"] = float(tl.si_snr(a, b))
    lo["sdr"] = float(tl.sdr(a, b))
    lo["si_sdr"] = float(tl.si_sdr(a, b))
    zs = torch.randn(6, 5, 7, 11, generator=g)
    zt = torch.randn(6, 9, 7, 11, generator=g)
    lo["zs"], lo["zt"] = zs, zt
    lo["spkd_batchmean"] = float(fw.SPKDLoss(zs, zt, "batchmean")())
    lo["spkd_sum"] = float(fw.SPKDLoss(zs, zt, "sum")())
    sx = 0.1 * torch.randn(3, 4000, generator=g)
    sy = sx + 0.05 * torch.randn(3, 4000, generator=g)
    lo["sx"], lo["sy"] = sx, sy
    sc, mag = fw.STFTLoss(512, 100, 400)(sx, sy)
    lo["stft_512_100_400"] = (float(sc), float(mag))
    sc, mag = fw.MultiResolutionSTFTLoss([512], [100], [400])(sx, sy)
    lo["mrstft_distill"] = (float(sc), float(mag))                       # distill.py:59
    sc, mag = fw.MultiResolutionSTFTLoss([512], [16], [32])(sx, sy)
    lo["mrstft_reviewkd"] = (float(sc), float(mag))                      # distill_ReviewKD.py:56
    sc, mag = fw.MultiResolutionSTFTLoss([256, 512, 128], [30, 60, 12], [150, 300, 60])(sx, sy)
    lo["mrstft_3res"] = (float(sc), float(mag))
    lo["stft_mag_512"] = summ(fw.stft(sx, 512, 100, 400, torch.hann_window(400)))

    # ABF / ReviewKD on the student's taps, lifted to the teacher's channel counts
    student.load_state_dict(s_sd)
    student.train()
    ext = fe.DCCRN(student)
    with torch.no_grad():
        student(X)
    ext.remove_hook()
    s_enc, s_dec = ext.feature_maps["encoder"], ext.feature_maps["decoder"]
    t_enc_ch = TEACHER["kernel_num"]
    t_dec_ch = [64, 64, 32, 16, 8, 2]
    torch.manual_seed(7)
    e_shapes = [m.shape[2] for m in s_enc][::-1]
    rk_enc = fw.ReviewKD([m.shape[1] for m in s_enc], t_enc_ch, e_shapes, e_shapes, s_enc, "encoder")
    d_shapes = [m.shape[2] for m in s_dec]
    rk_dec = fw.ReviewKD([m.shape[1] for m in s_dec][::-1], t_dec_ch[::-1], d_shapes, d_shapes, s_dec, "decoder")

    def abf_sd(rk):     # `abfs` is a plain list slice in the reference (not registered): collect by hand
        sd = {}
        for i, abf in enumerate(rk.abfs):
            for k, v in abf.state_dict().items():
                sd["abfs.%d.%s" % (i, k)] = v.clone()
        return sd
    for rk in (rk_enc, rk_dec):        # randomise BN affine so it is exercised
        for abf in rk.abfs:
            for seq in (abf.conv1, abf.conv2):
                seq[1].weight.data = 1 + 0.2 * torch.randn(seq[1].weight.shape, generator=g)
                seq[1].bias.data = 0.1 * torch.randn(seq[1].bias.shape, generator=g)
    lo["abf_enc_sd"], lo["abf_dec_sd"] = abf_sd(rk_enc), abf_sd(rk_dec)
    with torch.no_grad():
        f_enc, f_dec = rk_enc(X), rk_dec(X)
    lo["abf_enc_out"] = [summ(t) for t in f_enc]
    lo["abf_dec_out"] = [summ(t) for t in f_dec]
    lo["abf_enc_shapes"] = [tuple(t.shape) for t in f_enc]
    lo["abf_dec_shapes"] = [tuple(t.shape) for t in f_dec]
    torch.save(lo, os.path.join(HERE, "losses.pt"))

    # ------------------------------------------------------------------ step.pt
    st = {}
    stft_loss = fw.MultiResolutionSTFTLoss(fft_sizes=[512], win_lengths=[400], hop_sizes=[100])   # distill.py:59

    def taps(model, grad):
        ext = fe.DCCRN(model)
        with torch.set_grad_enabled(grad):
            wav = model(X, is_feat=True)
        ext.remove_hook()
        fm = ext.feature_maps
        re, im = fm["clstm"][0]
        return wav, fm["encoder"], fm["decoder"], re.transpose(0, 1), im.transpose(0, 1)

    for mode in ("clskd", "spkd_all", "spkd", "mse", "stft"):
        teacher.load_state_dict(t_sd)
        student.load_state_dict(s_sd)
        teacher.eval()
        student.train()
        for p in teacher.parameters():
            p.requires_grad = False
        student.zero_grad()
        abf_params = []
        t_wav, t_enc, t_dec, t_re, t_im = taps(teacher, False)
        s_wav, s_enc, s_dec, s_re, s_im = taps(student, True)
        terms = {"base": stft_loss(s_wav, y)[1]}
        if mode in ("clskd", "spkd_all"):
            if mode == "clskd":
                rk_enc.feature_maps, rk_dec.feature_maps = s_enc, s_dec
                for rk in (rk_enc, rk_dec):
                    for abf in rk.abfs:
                        abf.train()
                        abf.zero_grad()
                        abf_params += list(abf.named_parameters())
                f_enc, f_dec = rk_enc(X), rk_dec(X)
            else:
                f_enc, f_dec = s_enc, s_dec
            terms["encoder"] = sum(fw.SPKDLoss(a_, b_, "batchmean")() for a_, b_ in zip(f_enc, t_enc))
            terms["decoder"] = sum(fw.SPKDLoss(a_, b_, "batchmean")() for a_, b_ in zip(f_dec, t_dec))
            terms["clstm_real"] = fw.SPKDLoss(s_re, t_re, reduction="batchmean")()
            terms["clstm_img"] = fw.SPKDLoss(s_im, t_im, reduction="batchmean")()
        elif mode == "spkd":
            terms["kd"] = fw.SPKDLoss(s_wav.unsqueeze(1), t_wav.unsqueeze(1), reduction="batchmean")()
        elif mode == "mse":
            terms["kd"] = torch.nn.functional.mse_loss(s_wav, t_wav)
        elif mode == "stft":
            terms["kd"] = stft_loss(s_wav, t_wav)[1]
        loss = sum(terms.values())
        loss.backward()
        rec = {"loss": float(loss), "terms": {k: float(v) for k, v in terms.items()},
               "grads": {n: summ(p.grad) for n, p in student.named_parameters() if p.grad is not None},
               "grad_full": {n: p.grad.clone() for n, p in student.named_parameters()
                             if p.grad is not None and p.numel() <= 4096}}
        if mode == "clskd":
            rec["abf_enc_grads"] = {}
            rec["abf_dec_grads"] = {}
            for key, rk in (("abf_enc_grads", rk_enc), ("abf_dec_grads", rk_dec)):
                for i, abf in enumerate(rk.abfs):
                    for n, p in abf.named_parameters():
                        if p.grad is not None:
                            rec[key]["abfs.%d.%s" % (i, n)] = summ(p.grad)
        st[mode] = rec
    torch.save(st, os.path.join(HERE, "step.pt"))
    for f in ("dccrn.pt", "losses.pt", "step.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
