"""Generate tests/golden/real_width.pt by running the UNMODIFIED reference modules from /root/reference
at the BENCHMARKED widths and sequence length: teacher kernel_num [32,64,128,256,256,256] / rnn 256
(config.py:31-35), half-width student [16,32,64,128,128,128] / rnn 128 (BASELINE configs[1]) and the
reference's own quarter-width student [8,16,32,64,64,64] / rnn 64 (config.py:47-48), B = 4 utterances
of 4 s (L = 64000, T = 643 frames).

Container-only (the reference tree does not exist on the GPU box).  Weights and inputs are NOT stored:
they are regenerated from seeds by oracle.make_state_dict / oracle.make_abf_state_dict / a seeded
generator (checksums are stored and verified by the tests), so the fixture holds outputs only:
  * enhanced waveforms of the first two utterances: teacher eval, teacher train-mode BN, students train;
  * summaries of every teacher / student feature tap;
  * the CLSKD training-step loss (distill.py:72-148 restated over the local DCCRN as in SURVEY 3.1,
    built from the reference's own modules), its 5 terms, strided samples + L2 norms of every
    student and ABF gradient from torch autograd - for the eval-mode teacher (default) and for the
    reference-faithful train-mode teacher; and the SPKD-all step (configs[2]).
Run:  python tests/golden/make_golden_real.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle.dccrn_oracle import make_state_dict  # noqa: E402
from oracle.losses_oracle import make_abf_state_dict  # noqa: E402

TEACHER = dict(kernel_num=[32, 64, 128, 256, 256, 256], rnn_units=256)
STUDENTS = {"half": dict(kernel_num=[16, 32, 64, 128, 128, 128], rnn_units=128),
            "quarter": dict(kernel_num=[8, 16, 32, 64, 64, 64], rnn_units=64)}
SEEDS = dict(teacher=1, student=2, abf_enc=7, abf_dec=8, data=123)
B, L = 4, 64000
NS = 2048          # samples kept per tensor


def summ(t, n=NS):
    t = t.detach().double().reshape(-1)
    step = max(1, t.numel() // n)
    return {"numel": t.numel(), "sum": float(t.sum()), "asum": float(t.abs().sum()), "l2": float(t.norm()),
            "sample": t[::step][:n].float().clone(), "step": step}


def inputs():
    g = torch.Generator().manual_seed(SEEDS["data"])
    return 0.1 * torch.randn(B, L, generator=g), 0.1 * torch.randn(B, L, generator=g)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    mods = ref_shim.load()
    RefDCCRN = mods["DCCRN"].DCCRN
    fw, fe = mods["framework"], mods["feature_extraction"]
    _stft = torch.stft

    def stft_compat(*a, **k):          # torch >= 2 rejects the reference's torch.stft call (framework.py:27)
        k.setdefault("return_complex", True)
        out = _stft(*a, **k)
        return torch.view_as_real(out) if out.is_complex() else out
    torch.stft = stft_compat
    torch.nn.Module.cuda = lambda self, device=None: self            # framework.py:198-202 on a CPU box
    torch.Tensor.cuda = lambda self, *a, **k: self

    def build(cfg, seed):
        sd = make_state_dict(cfg["kernel_num"], cfg["rnn_units"], seed=seed)
        m = RefDCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
        m.load_state_dict(sd, strict=True)
        return m, sd

    X, y = inputs()
    teacher, t_sd = build(TEACHER, SEEDS["teacher"])
    out = {"teacher_cfg": TEACHER, "student_cfgs": STUDENTS, "seeds": SEEDS, "B": B, "L": L,
           "X_sum": summ(X, 64), "y_sum": summ(y, 64),
           "t_sd_sum": {k: summ(t_sd[k], 16) for k in ("encoder.3.0.real_conv.weight", "enhance.0.real_lstm.weight_hh_l0",
                                                      "decoder.1.0.imag_conv.weight", "encoder.2.1.weight")}}
    stft_loss = fw.MultiResolutionSTFTLoss(fft_sizes=[512], win_lengths=[400], hop_sizes=[100])   # distill.py:59

    def taps(model, grad):
        ext = fe.DCCRN(model)
        with torch.set_grad_enabled(grad):
            wav = model(X, is_feat=True)
        ext.remove_hook()
        fm = ext.feature_maps
        re, im = fm["clstm"][0]
        return wav, fm["encoder"], fm["decoder"], re.transpose(0, 1), im.transpose(0, 1)

    def tap_rec(wav, enc, dec, re, im):
        return {"wav": wav[:2].detach().clone(), "wav_sum": summ(wav),
                "encoder": [summ(t) for t in enc], "decoder": [summ(t) for t in dec],
                "clstm": [summ(re), summ(im)]}

    for p in teacher.parameters():
        p.requires_grad = False
    t_taps = {}
    for tmode in ("eval", "train"):
        teacher.load_state_dict(t_sd)
        teacher.train(tmode == "train")
        t_taps[tmode] = taps(teacher, False)
        out["teacher_" + tmode] = tap_rec(*t_taps[tmode])
        print("teacher", tmode, float(t_taps[tmode][0].abs().max()))

    # eval-mode teacher whose BatchNorm running statistics MATCH the data (what a trained teacher has): one
    # train-mode pass with momentum 1 sets running_mean / running_var to the batch statistics; stored so that
    # the tests load exactly these buffers
    teacher.load_state_dict(t_sd)
    bns = [m for m in teacher.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    for m in bns:
        m.momentum = 1.0
    teacher.train()
    with torch.no_grad():
        teacher(X)
    for m in bns:
        m.momentum = 0.1
    out["teacher_cal_running"] = {k: v.clone() for k, v in teacher.state_dict().items() if "running_" in k}
    teacher.eval()
    out["teacher_eval_cal"] = tap_rec(*taps(teacher, False))
    print("teacher eval (calibrated running stats)", float(out["teacher_eval_cal"]["wav"].abs().max()))

    for sname, scfg in STUDENTS.items():
        student, s_sd = build(scfg, SEEDS["student"])
        rec = {"s_sd_sum": {k: summ(s_sd[k], 16) for k in ("encoder.3.0.real_conv.weight", "decoder.1.0.imag_conv.weight")}}
        student.train()
        s_wav, s_enc, s_dec, s_re, s_im = taps(student, False)
        rec["student_train"] = tap_rec(s_wav, s_enc, s_dec, s_re, s_im)
        enc_in = [m.shape[1] for m in s_enc]
        dec_in = [m.shape[1] for m in s_dec][::-1]
        t_enc_ch = [m.shape[1] for m in t_taps["eval"][1]]
        t_dec_ch = [m.shape[1] for m in t_taps["eval"][2]][::-1]
        e_shapes = [m.shape[2] for m in s_enc][::-1]
        d_shapes = [m.shape[2] for m in s_dec]
        rk_enc = fw.ReviewKD(enc_in, t_enc_ch, e_shapes, e_shapes, s_enc, "encoder")
        rk_dec = fw.ReviewKD(dec_in, t_dec_ch, d_shapes, d_shapes, s_dec, "decoder")
        abf_sds = {"enc": make_abf_state_dict(enc_in, t_enc_ch, SEEDS["abf_enc"]),
                   "dec": make_abf_state_dict(dec_in, t_dec_ch, SEEDS["abf_dec"])}
        rec["abf_args"] = {"enc_in": enc_in, "enc_out": t_enc_ch, "dec_in": dec_in, "dec_out": t_dec_ch}

        def load_abf(rk, sd):          # `abfs` is a plain python list in the reference: load block by block
            for i, abf in enumerate(rk.abfs):
                pre = "abfs.%d." % i
                abf.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=True)

        for mode, tmode in (("clskd", "eval"), ("clskd", "train"), ("spkd_all", "eval")):
            if sname == "quarter" and (mode, tmode) != ("clskd", "eval"):
                continue
            student.load_state_dict(s_sd)
            student.train()
            student.zero_grad()
            t_wav, t_enc, t_dec, t_re, t_im = t_taps[tmode]
            s_wav, s_enc, s_dec, s_re, s_im = taps(student, True)
            terms = {"base": stft_loss(s_wav, y)[1]}
            if mode == "clskd":
                load_abf(rk_enc, abf_sds["enc"])
                load_abf(rk_dec, abf_sds["dec"])
                rk_enc.feature_maps, rk_dec.feature_maps = s_enc, s_dec
                for rk in (rk_enc, rk_dec):
                    for abf in rk.abfs:
                        abf.train()
                        abf.zero_grad()
                f_enc, f_dec = rk_enc(X), rk_dec(X)
            else:
                f_enc, f_dec = s_enc, s_dec
            enc_terms = [fw.SPKDLoss(a_, b_, "batchmean")() for a_, b_ in zip(f_enc, t_enc)]
            dec_terms = [fw.SPKDLoss(a_, b_, "batchmean")() for a_, b_ in zip(f_dec, t_dec)]
            terms["encoder"] = sum(enc_terms)
            terms["decoder"] = sum(dec_terms)
            terms["clstm_real"] = fw.SPKDLoss(s_re, t_re, reduction="batchmean")()
            terms["clstm_img"] = fw.SPKDLoss(s_im, t_im, reduction="batchmean")()
            loss = sum(terms.values())
            loss.backward()
            r = {"loss": float(loss), "terms": {k: float(v) for k, v in terms.items()},
                 "enc_terms": [float(v) for v in enc_terms], "dec_terms": [float(v) for v in dec_terms],
                 "grads": {n: summ(p.grad) for n, p in student.named_parameters() if p.grad is not None}}
            if mode == "clskd":
                r["fused_enc"] = [summ(t) for t in f_enc]
                r["fused_dec"] = [summ(t) for t in f_dec]
                for key, rk in (("abf_enc_grads", rk_enc), ("abf_dec_grads", rk_dec)):
                    r[key] = {}
                    for i, abf in enumerate(rk.abfs):
                        for n, p in abf.named_parameters():
                            if p.grad is not None:
                                r[key]["abfs.%d.%s" % (i, n)] = summ(p.grad)
            rec["%s_%s" % (mode, tmode)] = r
            print(sname, mode, tmode, r["loss"], r["terms"])
        out[sname] = rec
    path = os.path.join(HERE, "real_width.pt")
    torch.save(out, path)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
