"""ORACLE harness (test infrastructure; bench.py's CPU arm only): one CLSKD / SPKD-all training step
(distill.py:72-148 + optimizer :202-204) assembled from the reference's OWN unmodified modules - DCCRN,
feature_extraction.DCCRN hooks, framework.ReviewKD / SPKDLoss / MultiResolutionSTFTLoss - with the
adaptations of SURVEY 3.1 for the local model (is_feat=True, batch-first LSTM taps, ReviewKD lists read off
the maps).  Also config 1 of BASELINE.json: teacher forward + SI-SNR (DCCRN.py:266-267)."""
import torch

from . import ref_shim


def _models(widths_t, widths_s):
    mods = ref_shim.load()
    ref_shim.apply_torch_compat()
    R = mods["DCCRN"].DCCRN
    torch.manual_seed(1)
    teacher = R(rnn_units=widths_t["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=widths_t["kernel_num"])
    torch.manual_seed(2)
    student = R(rnn_units=widths_s["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=widths_s["kernel_num"])
    for p in teacher.parameters():                 # distill.py:49-50
        p.requires_grad = False
    return mods, teacher, student


def make_step(widths_t, widths_s, X, y, mode="clskd", faithful=False):
    """-> step() running one optimisation step of the reference's modules on CPU; returns the loss."""
    mods, teacher, student = _models(widths_t, widths_s)
    fw, fe = mods["framework"], mods["feature_extraction"]
    stft_loss = fw.MultiResolutionSTFTLoss(fft_sizes=[512], win_lengths=[400], hop_sizes=[100])    # distill.py:59
    teacher.train(faithful)                        # the reference never calls .eval()
    student.train()

    def taps(model, grad):
        ext = fe.DCCRN(model)
        with torch.set_grad_enabled(grad):
            wav = model(X, is_feat=True)
        ext.remove_hook()
        fm = ext.feature_maps
        re, im = fm["clstm"][0]
        return wav, fm["encoder"], fm["decoder"], re.transpose(0, 1), im.transpose(0, 1)

    rk = {}
    params = list(student.parameters())
    if mode == "clskd":                            # persistent, trainable ABF blocks (the benchmarked configuration)
        with torch.no_grad():
            _, t_enc, t_dec, _, _ = taps(teacher, False)
            _, s_enc, s_dec, _, _ = taps(student, False)
        torch.manual_seed(3)
        e_shapes = [m.shape[2] for m in s_enc][::-1]
        d_shapes = [m.shape[2] for m in s_dec]
        rk["enc"] = fw.ReviewKD([m.shape[1] for m in s_enc], [m.shape[1] for m in t_enc], e_shapes, e_shapes, s_enc, "encoder")
        rk["dec"] = fw.ReviewKD([m.shape[1] for m in s_dec][::-1], [m.shape[1] for m in t_dec][::-1], d_shapes, d_shapes,
                                s_dec, "decoder")
        for r in rk.values():
            for abf in r.abfs:
                params += list(abf.parameters())
    opt = torch.optim.Adam(params, lr=6e-4)        # distill.py:203

    def step():
        opt.zero_grad()
        t_wav, t_enc, t_dec, t_re, t_im = taps(teacher, faithful)
        s_wav, s_enc, s_dec, s_re, s_im = taps(student, True)
        if faithful:
            s_wav = student(X, is_feat=True)       # distill.py:100
        loss = stft_loss(s_wav, y)[1]
        if mode == "clskd":
            rk["enc"].feature_maps, rk["dec"].feature_maps = s_enc, s_dec
            f_enc, f_dec = rk["enc"](X), rk["dec"](X)
        else:
            f_enc, f_dec = s_enc, s_dec
        for a, b in zip(f_enc, t_enc):
            loss = loss + fw.SPKDLoss(a, b, "batchmean")()
        for a, b in zip(f_dec, t_dec):
            loss = loss + fw.SPKDLoss(a, b, "batchmean")()
        loss = loss + fw.SPKDLoss(s_re, t_re, reduction="batchmean")() + fw.SPKDLoss(s_im, t_im, reduction="batchmean")()
        loss.backward()
        opt.step()
        return float(loss.detach())
    return step


def make_config1(widths_t, X, y):
    """BASELINE configs[0]: DCCRN-CL teacher forward + SI-SNR, eval mode, no_grad (DCCRN.py:266-267)."""
    mods, teacher, _ = _models(widths_t, dict(kernel_num=[8, 16, 32, 64, 64, 64], rnn_units=64))
    teacher.eval()

    def step():
        with torch.no_grad():
            _, _, real, imag, wav = teacher(X)
            return float(teacher.loss(wav, y, real, imag, loss_mode="SI-SNR"))
    return step
