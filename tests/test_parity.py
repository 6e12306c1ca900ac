"""Parity tests of the product against the reference's golden outputs and the oracle.

Every test runs twice through the `dev` fixture:
  * `emu`  (CPU, always): the numpy model of the C ABI (tests/cabi_emu.py) stands in for the CUDA
    kernels, so the host-side logic - convolution plans, packing tables, stride arithmetic, the
    autograd wiring, the ABF chain, the distillation steps, the flat-bucket optimizer - is checked
    without a GPU;
  * `cuda` (`-m gpu`, on a B200): the same assertions against the real sm_100a kernels through
    the C ABI, fp32 policy (waveform tolerance 1e-5, loss tolerance 1e-3 relative or tighter).
"""
import pytest
import torch

from util import bn_shadowed_bias, check_summary, close, full_sd, golden, rel_err


@pytest.fixture(params=["emu", pytest.param("cuda", marks=pytest.mark.gpu)])
def dev(request, monkeypatch):
    import clskd_b200
    clskd_b200.set_precision("fp32")
    if request.param == "emu":
        import cabi_emu
        cabi_emu.install(monkeypatch)
        return torch.device("cpu")
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    clskd_b200._lib.load()          # raises if the CUDA library is missing: no fallback
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------ model
def _build(cfg, sd, dev, masking_mode="E"):
    import clskd_b200
    m = clskd_b200.DCCRN(rnn_units=cfg["rnn_units"], masking_mode=masking_mode, use_clstm=True,
                         kernel_num=cfg["kernel_num"])
    m.load_state_dict(full_sd(sd))
    return m.to(dev)


def test_state_dict_keys_match_reference_layout():
    import clskd_b200
    from oracle.dccrn_oracle import make_state_dict
    kn, ru = [8, 16, 32, 64, 64, 64], 64
    m = clskd_b200.DCCRN(rnn_units=ru, use_clstm=True, kernel_num=kn)
    ref = make_state_dict(kn, ru)
    ours = m.state_dict()
    assert set(ours) == set(ref)
    for k in ref:
        assert tuple(ours[k].shape) == tuple(ref[k].shape), k
    assert sum(p.numel() for p in m.parameters()) == 231565


@pytest.mark.parametrize("name", ["teacher", "student"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_dccrn_forward_golden(dev, name, mode):
    import clskd_b200
    g = golden("dccrn.pt")
    ref = g["%s_%s" % (name, mode)]
    m = _build(g[name + "_cfg"], g["t_sd" if name == "teacher" else "s_sd"], dev)
    m.train(mode == "train")
    ext = clskd_b200.feature_extraction.DCCRN(m)
    with torch.no_grad():
        mr, mi, re, im, wav = m(g["X"].to(dev))
    ext.remove_hook()
    assert (wav.cpu() - ref["wav"]).abs().max().item() <= 1e-5          # north_star fp32 tolerance
    for t, key in ((mr, "mask_real"), (mi, "mask_imag"), (re, "real"), (im, "imag")):
        check_summary(t, ref[key], what=key)
    fm = ext.feature_maps
    for kind in ("encoder", "decoder"):
        assert [tuple(t.shape) for t in fm[kind]] == ref[kind + "_shapes"]
        for i, t in enumerate(fm[kind]):
            check_summary(t, ref[kind][i], what="%s[%d]" % (kind, i))
    assert [tuple(t.shape) for t in fm["clstm"][0]] == ref["clstm_shapes"]
    for i, t in enumerate(fm["clstm"][0]):
        check_summary(t, ref["clstm"][i], what="clstm[%d]" % i)
    if mode == "train":
        sd = m.state_dict()
        for k, v in ref["running"].items():
            assert torch.allclose(sd[k].cpu(), v, rtol=1e-5, atol=1e-6), k
    else:
        y = g["y"].to(dev)
        for lm in ("SI-SNR", "MSE", "SDR", "SI-SDR"):
            assert close(m.loss(wav, y, loss_mode=lm), ref["loss"][lm], atol=1e-5 if lm != "MSE" else 1e-9), lm


@pytest.mark.parametrize("masking_mode", ["E", "C", "R"])
def test_dccrn_forward_vs_oracle_all_mask_modes(dev, masking_mode):
    from oracle import dccrn_oracle as D
    kn, ru = [4, 8, 8, 16, 16, 16], 16
    sd = D.make_state_dict(kn, ru, seed=11)
    m = _build(dict(kernel_num=kn, rnn_units=ru), sd, dev, masking_mode)
    m.eval()
    x = 0.1 * torch.randn(3, 3000, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        ref = D.dccrn_forward(sd, x, masking_mode=masking_mode)
        out = m(x.to(dev))
        wav_only = m(x.to(dev), is_feat=True)
    for a, b in zip(out, ref):
        assert (a.cpu() - b).abs().max().item() < 1e-5
    assert torch.equal(wav_only, out[-1])


def test_empty_and_ragged_inputs(dev):
    """zero-length batch and lengths that are not a multiple of the hop (T = L//100 + 3 frames)"""
    from oracle import dccrn_oracle as D
    kn, ru = [4, 8, 8, 16, 16, 16], 16
    sd = D.make_state_dict(kn, ru, seed=4)
    m = _build(dict(kernel_num=kn, rnn_units=ru), sd, dev)
    m.eval()
    for L in (1237, 400, 99):
        x = 0.1 * torch.randn(1, L, generator=torch.Generator().manual_seed(L))
        with torch.no_grad():
            ref = D.dccrn_forward(sd, x)[-1]
            out = m(x.to(dev), is_feat=True)
        assert out.shape == ref.shape == (1, (L // 100) * 100)
        assert ref.numel() == 0 or (out.cpu() - ref).abs().max().item() < 1e-5, L


def test_public_block_api(dev):
    """the drop-in modules called through their public (logical NCHW) forward, like the reference's"""
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 16, 9, generator=g)
    conv = tm.ComplexConv2d(6, 10, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1))
    conv.real_conv.bias.data.normal_(generator=g)
    conv.imag_conv.bias.data.normal_(generator=g)
    ref = D.complex_conv2d(x, conv.real_conv.weight, conv.real_conv.bias, conv.imag_conv.weight, conv.imag_conv.bias)
    with torch.no_grad():
        assert torch.allclose(conv.to(dev)(x.to(dev)).cpu(), ref, atol=1e-5)
    dec = tm.ComplexConvTranspose2d(6, 4, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0), output_padding=(1, 0))
    dec.real_conv.bias.data.normal_(generator=g)
    ref = D.complex_deconv2d(x, dec.real_conv.weight, dec.real_conv.bias, dec.imag_conv.weight, dec.imag_conv.bias)
    with torch.no_grad():
        assert torch.allclose(dec.to(dev)(x.to(dev)).cpu(), ref, atol=1e-5)
    a, b = torch.randn(2, 4, 3, 5, generator=g), torch.randn(2, 6, 3, 5, generator=g)
    assert torch.equal(tm.complex_cat([a, b], 1), D.complex_cat([a, b], 1))
    stft = tm.ConvSTFT(400, 100, 512, 'hamming', 'complex').to(dev)
    istft = tm.ConviSTFT(400, 100, 512, 'hamming', 'complex').to(dev)
    wav = 0.1 * torch.randn(2, 1600, generator=g)
    spec = stft(wav.to(dev))
    assert spec.shape == (2, 514, 19)
    assert torch.allclose(spec.cpu(), D.conv_stft(wav, stft.weight.cpu(), 400, 100), atol=1e-5)
    rec = istft(spec)
    assert rec.shape == (2, 1, 1600) and (rec.squeeze(1).cpu() - wav).abs().max() < 1e-5     # round trip
    bn = tm.BatchNorm2d(6).to(dev)
    ref_bn = torch.nn.BatchNorm2d(6)
    with torch.no_grad():
        assert torch.allclose(bn(x.to(dev)).cpu(), ref_bn(x), atol=1e-5)
        assert torch.allclose(bn.running_var.cpu(), ref_bn.running_var, atol=1e-6)
        pr = tm.PReLU().to(dev)
        assert torch.allclose(pr(x.to(dev)).cpu(), torch.nn.functional.prelu(x, torch.tensor([0.25])), atol=1e-6)


def test_conv_block_gradients_vs_autograd(dev):
    """conv / deconv(+skip) / BN+PReLU backward (dgrad, wgrad, bias, BN, slope) against torch autograd
    of the oracle"""
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(2)
    blk = tm.ConvBNAct(tm.ComplexConv2d(8, 12, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1)),
                       tm.BatchNorm2d(12), tm.PReLU()).to(dev)
    dblk = tm.ConvBNAct(tm.ComplexConvTranspose2d(24, 6, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0),
                                                  output_padding=(1, 0)), tm.BatchNorm2d(6), tm.PReLU()).to(dev)
    for b_ in (blk, dblk):
        for n, p in b_.named_parameters():
            if n.endswith("bias") or n.startswith("1."):
                p.data = (p.data.cpu() + 0.3 * torch.randn(p.shape, generator=g)).to(dev)
    x = torch.randn(2, 8, 16, 7, generator=g)
    xd = x.clone().to(dev).requires_grad_(True)
    h = blk(xd)                                   # [2,12,8,7]
    y = dblk(h, h)                                # skip = same tensor: [2,6,16,8]
    w = torch.randn(y.shape, generator=g)
    (y * w.to(dev)).sum().backward()
    # oracle
    sd = {}
    for pre, b_ in (("e.", blk), ("d.", dblk)):
        for k, v in b_.state_dict().items():
            sd[pre + k] = v.detach().cpu().clone().requires_grad_(v.is_floating_point() and "running" not in k)
    xo = x.clone().requires_grad_(True)
    ho = D.complex_conv2d(xo, sd["e.0.real_conv.weight"], sd["e.0.real_conv.bias"], sd["e.0.imag_conv.weight"],
                          sd["e.0.imag_conv.bias"])
    ho = D._bn_prelu(ho, sd, "e.", True)
    yo = D.complex_deconv2d(D.complex_cat([ho, ho], 1), sd["d.0.real_conv.weight"], sd["d.0.real_conv.bias"],
                            sd["d.0.imag_conv.weight"], sd["d.0.imag_conv.bias"])
    yo = D._bn_prelu(yo, sd, "d.", True)
    assert torch.allclose(y.detach().cpu(), yo.detach(), atol=2e-5)
    (yo * w).sum().backward()
    assert torch.allclose(xd.grad.cpu(), xo.grad, rtol=1e-3, atol=1e-4)
    for pre, b_ in (("e.", blk), ("d.", dblk)):
        for n, p in b_.named_parameters():
            if n.endswith("_conv.bias"):
                continue                      # shadowed by train-mode BN: zero up to rounding
            ref = sd[pre + n].grad
            assert torch.allclose(p.grad.cpu(), ref, rtol=2e-3, atol=2e-4 * ref.abs().max().item() + 1e-6), pre + n


@pytest.mark.parametrize("B", [3, 83])
def test_complex_lstm(dev, B):
    """B = 3: two batch rows per CTA; B = 83: 2*B*2 row/set pairs exceed the SM count, so the recurrence
    kernel switches to four rows per CTA (the large-batch / inference configuration); ragged last CTA."""
    from clskd_b200.clstm import NavieComplexLSTM
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(1)
    lstm = NavieComplexLSTM(input_size=24, hidden_size=16, projection_dim=24)
    r, i = torch.randn(7, B, 12, generator=g), torch.randn(7, B, 12, generator=g)
    sd = {"x." + k: v.clone() for k, v in lstm.state_dict().items()}
    rr, ri = D.complex_lstm(r, i, sd, "x.", True)
    lstm = lstm.to(dev)
    rd = r.clone().to(dev).requires_grad_(True)
    out = lstm([rd, i.to(dev)])
    assert torch.allclose(out[0].cpu(), rr, atol=1e-5) and torch.allclose(out[1].cpu(), ri, atol=1e-5)
    # BPTT against torch autograd of the oracle
    (out[0].sum() + (out[1] ** 2).sum()).backward()
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    r2 = r.detach().clone().requires_grad_(True)
    o2 = D.complex_lstm(r2, i, params, "x.", True)
    (o2[0].sum() + (o2[1] ** 2).sum()).backward()
    assert torch.allclose(rd.grad.cpu(), r2.grad, atol=1e-4)
    for k, p in lstm.named_parameters():
        assert torch.allclose(p.grad.cpu(), params["x." + k].grad, rtol=1e-3, atol=1e-4), k


# ------------------------------------------------------------------------------------ losses
def test_losses_golden(dev):
    from clskd_b200 import framework as fw
    from clskd_b200 import tools_for_loss as tl
    L = golden("losses.pt")
    a, b = L["a"].to(dev), L["b"].to(dev)
    assert close(tl.si_snr(a, b), L["si_snr"]) and close(tl.sdr(a, b), L["sdr"]) and close(tl.si_sdr(a, b), L["si_sdr"])
    zs, zt = L["zs"].to(dev), L["zt"].to(dev)
    assert rel_err(fw.SPKDLoss(zs, zt, "batchmean")(), L["spkd_batchmean"]) < 1e-5
    assert rel_err(fw.SPKDLoss(zs, zt, "sum")(), L["spkd_sum"]) < 1e-5
    sx, sy = L["sx"].to(dev), L["sy"].to(dev)
    sc, mag = fw.STFTLoss(512, 100, 400).to(dev)(sx, sy)
    assert rel_err(sc, L["stft_512_100_400"][0]) < 1e-5 and rel_err(mag, L["stft_512_100_400"][1]) < 1e-5
    for key, cfg in (("mrstft_distill", ([512], [100], [400])), ("mrstft_reviewkd", ([512], [16], [32])),
                     ("mrstft_3res", ([256, 512, 128], [30, 60, 12], [150, 300, 60]))):
        sc, mag = fw.MultiResolutionSTFTLoss(*cfg).to(dev)(sx, sy)
        assert rel_err(sc, L[key][0]) < 1e-5 and rel_err(mag, L[key][1]) < 1e-5, key
    check_summary(fw.stft(sx, 512, 100, 400, torch.hann_window(400).to(dev)), L["stft_mag_512"], what="stft_mag")


def test_loss_gradients_vs_autograd(dev):
    """autograd Functions of the objectives against torch autograd of the oracle"""
    from clskd_b200 import framework as fw
    from clskd_b200 import tools_for_loss as tl
    from oracle import losses_oracle as LO
    g = torch.Generator().manual_seed(3)
    a = (0.3 * torch.randn(3, 700, generator=g))
    b = a + 0.2 * torch.randn(3, 700, generator=g)
    for ours, ref, wrt in ((tl.si_snr, LO.si_snr, 0), (tl.sdr, LO.sdr, 0), (tl.sdr, LO.sdr, 1), (tl.si_sdr, LO.si_sdr, 1),
                           (tl.mse, torch.nn.functional.mse_loss, 0)):
        args1 = [a.clone().to(dev), b.clone().to(dev)]
        args2 = [a.clone(), b.clone()]
        args1[wrt].requires_grad_(True)
        args2[wrt].requires_grad_(True)
        ours(*args1).backward()
        ref(*args2).backward()
        assert torch.allclose(args1[wrt].grad.cpu(), args2[wrt].grad, rtol=1e-3, atol=1e-6), (ours.__name__, wrt)
    zs = torch.randn(5, 3, 4, 6, generator=g)
    zt = torch.randn(5, 7, 4, 6, generator=g)
    z1, z2 = zs.clone().to(dev).requires_grad_(True), zs.clone().requires_grad_(True)
    fw.SPKDLoss(z1, zt.to(dev), "batchmean")().backward()
    LO.spkd(z2, zt).backward()
    assert torch.allclose(z1.grad.cpu(), z2.grad, rtol=1e-3, atol=1e-7)
    x1, x2 = a.clone().to(dev).requires_grad_(True), a.clone().requires_grad_(True)
    sc, mag = fw.STFTLoss(128, 30, 100).to(dev)(x1, b.to(dev))
    (sc + 2 * mag).backward()
    sc, mag = LO.stft_loss(x2, b, 128, 30, 100)
    (sc + 2 * mag).backward()
    assert torch.allclose(x1.grad.cpu(), x2.grad, rtol=1e-3, atol=1e-6)


def test_hcl_vs_oracle(dev):
    from clskd_b200 import framework as fw
    from oracle import losses_oracle as LO
    g = torch.Generator().manual_seed(9)
    fs = [torch.randn(2, 6, 8, 11, generator=g), torch.randn(2, 4, 4, 11, generator=g), torch.randn(2, 4, 2, 5, generator=g)]
    ft = [torch.randn(f.shape, generator=g) for f in fs]
    f1 = [f.clone().to(dev).requires_grad_(True) for f in fs]
    f2 = [f.clone().requires_grad_(True) for f in fs]
    l1 = fw.hcl(f1, [t.to(dev) for t in ft])
    l2 = LO.hcl(f2, ft)
    assert rel_err(l1.detach(), l2.detach()) < 1e-5
    l1.backward()
    l2.backward()
    for a, b in zip(f1, f2):
        assert torch.allclose(a.grad.cpu(), b.grad, rtol=1e-3, atol=1e-7)


def _load_abf(rk, sd):
    own = rk.state_dict()
    assert set(own) == set(sd), (sorted(set(own) ^ set(sd)))
    rk.load_state_dict(sd)


def test_review_kd_golden(dev):
    import clskd_b200
    from clskd_b200 import framework as fw
    g, L = golden("dccrn.pt"), golden("losses.pt")
    m = _build(g["student_cfg"], g["s_sd"], dev)
    m.train()
    ext = clskd_b200.feature_extraction.DCCRN(m)
    X = g["X"].to(dev)
    with torch.no_grad():
        m(X)
    ext.remove_hook()
    s_enc, s_dec = ext.feature_maps["encoder"], ext.feature_maps["decoder"]
    rk_enc = fw.build_review_kd(s_enc, "encoder", out_channels=g["teacher_cfg"]["kernel_num"])
    rk_dec = fw.build_review_kd(s_dec, "decoder", out_channels=[64, 64, 32, 16, 8, 2][::-1])
    _load_abf(rk_enc, L["abf_enc_sd"])
    _load_abf(rk_dec, L["abf_dec_sd"])
    with torch.no_grad():
        f_enc, f_dec = rk_enc(X), rk_dec(X)
    assert [tuple(t.shape) for t in f_enc] == L["abf_enc_shapes"]
    assert [tuple(t.shape) for t in f_dec] == L["abf_dec_shapes"]
    for i, t in enumerate(f_enc):
        check_summary(t, L["abf_enc_out"][i], what="abf_enc[%d]" % i)
    for i, t in enumerate(f_dec):
        check_summary(t, L["abf_dec_out"][i], what="abf_dec[%d]" % i)


# ------------------------------------------------------------------------------------ steps
@pytest.mark.parametrize("mode", ["clskd", "spkd_all", "spkd", "mse", "stft"])
def test_distill_step_golden(dev, mode):
    from clskd_b200.distill import DistillStep
    g, L, S = golden("dccrn.pt"), golden("losses.pt"), golden("step.pt")
    ref = S[mode]
    teacher = _build(g["teacher_cfg"], g["t_sd"], dev)
    student = _build(g["student_cfg"], g["s_sd"], dev)
    student.train()
    X, y = g["X"].to(dev), g["y"].to(dev)
    step = DistillStep(teacher, student, mode=mode)
    if mode == "clskd":
        step.materialize(X)
        student.load_state_dict(full_sd(g["s_sd"]))      # materialize ran a train-mode forward
        _load_abf(step.abf_encoder, L["abf_enc_sd"])
        _load_abf(step.abf_decoder, L["abf_dec_sd"])
    loss = step(X, y)
    assert rel_err(loss.detach(), ref["loss"]) < 1e-4             # north_star: loss rel. error <= 1e-3
    for k, v in ref["terms"].items():
        assert rel_err(step.last_terms[k].detach(), v) < 1e-4, k
    loss.backward()
    params = dict(student.named_parameters())
    for name, gref in ref["grads"].items():
        if bn_shadowed_bias(name):
            continue
        assert params[name].grad is not None, name
        check_summary(params[name].grad, gref, rtol=2e-3, atol=5e-7, what="grad " + name)   # fp32 reduction-order noise
    if mode == "clskd":
        for key, rk in (("abf_enc_grads", step.abf_encoder), ("abf_dec_grads", step.abf_decoder)):
            ps = dict(rk.named_parameters())
            for name, gref in ref[key].items():
                check_summary(ps[name].grad, gref, rtol=2e-3, atol=1e-7, what=key + " " + name)


def test_faithful_step_matches_oracle_with_train_mode_teacher(emu):
    """DistillStep(faithful=True): teacher BatchNorm in train mode with autograd enabled and the student
    forward run twice, as distill.py:49-50,77,85,100 does (host logic; CPU model of the C ABI)."""
    import clskd_b200
    from clskd_b200.distill import DistillStep
    from oracle import losses_oracle as LO
    clskd_b200.set_precision("fp32")
    dev = torch.device("cpu")
    g, L = golden("dccrn.pt"), golden("losses.pt")
    teacher = _build(g["teacher_cfg"], g["t_sd"], dev)
    student = _build(g["student_cfg"], g["s_sd"], dev)
    teacher.train()
    student.train()
    X, y = g["X"][:, :3200], g["y"][:, :3200]
    step = DistillStep(teacher, student, mode="spkd_all", faithful=True)
    rm0 = student.encoder[0][1].running_mean.clone()
    loss = step(X, y)
    ref, terms = LO.clskd_step_loss(full_sd(g["t_sd"]), full_sd(g["s_sd"]), X, y, mode="spkd_all",
                                    teacher_training=True)
    assert rel_err(loss, ref) < 1e-4
    for k, v in terms.items():
        assert rel_err(step.last_terms[k], v) < 1e-4, k
    assert teacher.training and int(teacher.encoder[0][1].num_batches_tracked) == 1     # train-mode teacher BN
    assert int(student.encoder[0][1].num_batches_tracked) == 2                           # two student forwards
    assert not torch.equal(student.encoder[0][1].running_mean, rm0)
    loss.backward()
    assert student.encoder[0][0].real_conv.weight.grad is not None
    assert all(p.grad is None for p in teacher.parameters())


def test_fresh_abf_step_rebuilds_untrained_fusion_blocks(emu):
    """DistillStep(fresh_abf=True): new random ABF weights on every step, never trainable
    (distill.py:92-96: build_review_kd inside training_step); the student still receives gradients."""
    import clskd_b200
    from clskd_b200.distill import DistillStep
    clskd_b200.set_precision("fp32")
    dev = torch.device("cpu")
    g = golden("dccrn.pt")
    teacher = _build(g["teacher_cfg"], g["t_sd"], dev)
    student = _build(g["student_cfg"], g["s_sd"], dev)
    student.train()
    X, y = g["X"][:, :3200], g["y"][:, :3200]
    step = DistillStep(teacher, student, mode="clskd", fresh_abf=True)
    torch.manual_seed(0)
    l1 = step(X, y)
    l1.backward()
    assert step.abf_encoder is None and step.abf_decoder is None           # nothing persistent
    assert student.encoder[2][0].real_conv.weight.grad.abs().sum() > 0
    n_train = len(step.trainable_parameters())
    assert n_train == sum(1 for p in student.parameters() if p.requires_grad)
    student.load_state_dict(full_sd(g["s_sd"]))
    torch.manual_seed(1)
    l2 = step(X, y)
    assert abs(float(l1.detach()) - float(l2.detach())) > 0                 # different random projections
    assert rel_err(step.last_terms["base"], float(l2.detach()) - sum(
        float(v.detach()) for k, v in step.last_terms.items() if k != "base")) < 1e-4


def test_flat_adam_matches_torch_adam(dev):
    from clskd_b200.distill import FlatAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(5, 3), (7,), (2, 2)]
    ps = [torch.nn.Parameter(torch.randn(s, generator=g).to(dev)) for s in shapes]
    qs = [torch.nn.Parameter(p.detach().cpu().clone()) for p in ps]
    ours = FlatAdam(ps, lr=6e-4, weight_decay=5e-4)
    ref = torch.optim.Adam(qs, lr=6e-4, weight_decay=5e-4)
    for it in range(3):
        for p, q in zip(ps, qs):
            gr = torch.randn(p.shape, generator=g)
            p.grad, q.grad = gr.clone().to(dev), gr.clone()
        if it == 1:                                     # a parameter without a gradient is skipped like torch does
            ps[2].grad = None
            qs[2].grad = None
        ours.pack_grads()
        ours.step()
        ref.step()
        for p, q in zip(ps, qs):
            assert torch.allclose(p.detach().cpu(), q.detach(), rtol=1e-5, atol=1e-6)


def test_flat_adam_state_dict_round_trip(dev):
    from clskd_b200.distill import FlatAdam
    g = torch.Generator().manual_seed(1)
    mk = lambda: [torch.nn.Parameter(torch.randn(s, generator=torch.Generator().manual_seed(4)).to(dev)) for s in [(6, 2), (3,)]]
    grads = [[torch.randn(p.shape, generator=g) for p in mk()] for _ in range(4)]
    a_params, b_params = mk(), mk()
    a, b = FlatAdam(a_params, lr=1e-3, weight_decay=1e-2), FlatAdam(b_params, lr=1e-3, weight_decay=1e-2)

    def run(opt, params, its):
        for it in its:
            for p, gr in zip(params, grads[it]):
                p.grad = gr.clone().to(dev)
            opt.pack_grads()
            opt.step()
    run(a, a_params, range(4))
    run(b, b_params, range(2))
    sd = b.state_dict()
    c_params = mk()
    c = FlatAdam(c_params, lr=5.0)
    c.flat_p.copy_(b.flat_p)
    c.load_state_dict(sd)
    run(c, c_params, range(2, 4))
    assert torch.allclose(c.flat_p, a.flat_p, rtol=1e-6, atol=1e-7)
    c_params[0].data = c_params[0].data.clone()           # detached from the bucket (what model.to() would do)
    with pytest.raises(RuntimeError):
        c.check_views()


def test_distill_step_with_plain_lstm_bottleneck(dev):
    """use_clstm=False (the DCCRN constructor default, DCCRN.py:101-109): the feature modes tap the plain LSTM's
    output; the step loss equals the oracle's with the LSTM output split into halves."""
    import clskd_b200
    from clskd_b200.distill import DistillStep
    from oracle import losses_oracle as LO
    torch.manual_seed(0)
    kw_t, kw_s = dict(kernel_num=[4, 8, 8, 16, 16, 16], rnn_units=16), dict(kernel_num=[2, 4, 4, 8, 8, 8], rnn_units=8)
    teacher = clskd_b200.DCCRN(masking_mode="E", use_clstm=False, **kw_t)
    student = clskd_b200.DCCRN(masking_mode="E", use_clstm=False, **kw_s)
    t_sd = {k: v.detach().clone() for k, v in teacher.state_dict().items()}
    s_sd = {k: v.detach().clone() for k, v in student.state_dict().items()}
    teacher, student = teacher.to(dev), student.to(dev)
    student.train()
    g = torch.Generator().manual_seed(0)
    X, y = 0.1 * torch.randn(3, 2000, generator=g), 0.1 * torch.randn(3, 2000, generator=g)
    step = DistillStep(teacher, student, mode="spkd_all")
    loss = step(X.to(dev), y.to(dev))
    loss.backward()
    assert student.enhance.weight_hh_l0.grad is not None
    tt, st = {}, {}
    from oracle.dccrn_oracle import dccrn_forward
    with torch.no_grad():
        dccrn_forward(t_sd, X, training=False, taps=tt)
        s_wav = dccrn_forward(s_sd, X, training=True, taps=st)[-1]
        ref = LO.mr_stft_loss(s_wav, y, [512], [100], [400])[1]
        ref = ref + sum(LO.spkd(a, b) for a, b in zip(st["encoder"], tt["encoder"]))
        ref = ref + sum(LO.spkd(a, b) for a, b in zip(st["decoder"], tt["decoder"]))
        ys, yt = st["clstm"][0].transpose(0, 1), tt["clstm"][0].transpose(0, 1)
        hs, ht = ys.shape[-1] // 2, yt.shape[-1] // 2
        ref = ref + LO.spkd(ys[..., :hs], yt[..., :ht]) + LO.spkd(ys[..., hs:], yt[..., ht:])
    assert rel_err(loss.detach(), ref) < 1e-4


def test_trainer_step_decreases_loss(dev):
    """three optimizer steps of the SPKD-all recipe on a fixed batch"""
    from clskd_b200.distill import DistillTrainer
    from oracle import dccrn_oracle as D
    cfg_t, cfg_s = dict(kernel_num=[4, 8, 8, 16, 16, 16], rnn_units=16), dict(kernel_num=[2, 4, 4, 8, 8, 8], rnn_units=8)
    teacher = _build(cfg_t, D.make_state_dict(cfg_t["kernel_num"], cfg_t["rnn_units"], seed=1), dev)
    student = _build(cfg_s, D.make_state_dict(cfg_s["kernel_num"], cfg_s["rnn_units"], seed=2), dev)
    g = torch.Generator().manual_seed(0)
    X, y = (0.1 * torch.randn(2, 2000, generator=g)).to(dev), (0.1 * torch.randn(2, 2000, generator=g)).to(dev)
    tr = DistillTrainer(teacher, student, mode="spkd_all", lr=1e-3)
    losses = [float(tr.train_step(X, y)) for _ in range(4)]
    assert all(torch.isfinite(torch.tensor(losses))) and losses[-1] < losses[0]


@pytest.mark.parametrize("cin,cout,ks", [(128, 2, 3), (256, 2, 1), (2, 128, 1), (16, 1, 3), (2, 16, 3)])
def test_narrow_real_convs_vs_torch(dev, cin, cout, ks):
    """ABF's convolutions with very few output or input channels (the 2-channel mask map and the
    2-logit attention conv) take dedicated GEMV / outer-product kernels: forward, dgrad, wgrad."""
    from clskd_b200 import framework as fw
    g = torch.Generator().manual_seed(cin * 7 + cout)
    conv = fw.RealConv2d(cin, cout, ks, padding=ks // 2, bias=True)
    x = torch.randn(2, cin, 16, 70, generator=g)
    up = torch.randn(2, cout, 16, 70, generator=g)
    xr = x.clone().requires_grad_(True)
    wr, br = conv.weight.detach().clone().requires_grad_(True), conv.bias.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, br, padding=ks // 2)
    (ref * up).sum().backward()
    gw_ref, gb_ref = wr.grad, br.grad
    conv = conv.to(dev)
    xd = x.clone().to(dev).requires_grad_(True)
    y = conv(xd)
    (y * up.to(dev)).sum().backward()
    assert torch.allclose(y.detach().cpu(), ref.detach(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(xd.grad.cpu(), xr.grad, atol=1e-4, rtol=1e-4)
    assert torch.allclose(conv.weight.grad.cpu(), gw_ref, atol=2e-3, rtol=1e-3)
    assert torch.allclose(conv.bias.grad.cpu(), gb_ref, atol=2e-3, rtol=1e-3)


@pytest.mark.parametrize("cin,cout,ks", [(16, 2, 3), (8, 1, 3)])
def test_tap_in_channel_decomposition_vs_torch(dev, cin, cout, ks):
    """the narrow-conv decomposition (pointwise GEMM onto (tap, n) channels + tap gather-sum and its
    adjoint) forced on under the fp32 policy"""
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(cin + ks)
    conv = fw.RealConv2d(cin, cout, ks, padding=ks // 2, bias=False)
    x = torch.randn(2, cin, 12, 19, generator=g)
    up = torch.randn(2, cout, 12, 19, generator=g)
    xr = x.clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, padding=ks // 2)
    (ref * up).sum().backward()
    conv = conv.to(dev)
    ops.policy.narrow = "always"
    try:
        xp = x.permute(0, 3, 2, 1).contiguous().to(dev).requires_grad_(True)
        assert conv._use_narrow(xp, None)
        y = conv.forward_phys(xp)
        (y * up.permute(0, 3, 2, 1).to(dev)).sum().backward()
    finally:
        ops.policy.narrow = "auto"
    assert torch.allclose(y.detach().permute(0, 3, 2, 1).cpu(), ref.detach(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(xp.grad.permute(0, 3, 2, 1).cpu(), xr.grad, atol=1e-4, rtol=1e-4)
    assert torch.allclose(conv.weight.grad.cpu(), wr.grad, atol=1e-3, rtol=1e-3)


def test_mask_layer_decomposition_vs_oracle(dev):
    """last decoder layer (complex transposed conv onto ONE complex channel, with skip input) through
    the tap-in-channel path (forced on under the fp32 policy): forward, both data gradients, weights, bias"""
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(21)
    dec = tm.ComplexConvTranspose2d(16, 2, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0), output_padding=(1, 0))
    dec.real_conv.bias.data.normal_(generator=g)
    dec.imag_conv.bias.data.normal_(generator=g)
    a = torch.randn(2, 8, 12, 9, generator=g)
    sk = torch.randn(2, 8, 12, 9, generator=g)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in dec.state_dict().items()}
    ar, sr = a.clone().requires_grad_(True), sk.clone().requires_grad_(True)
    ref = D.complex_deconv2d(D.complex_cat([ar, sr], 1), sd["real_conv.weight"], sd["real_conv.bias"],
                             sd["imag_conv.weight"], sd["imag_conv.bias"])
    up = torch.randn(ref.shape, generator=g)
    (ref * up).sum().backward()
    dec = dec.to(dev)
    ap = a.permute(0, 3, 2, 1).contiguous().to(dev).requires_grad_(True)
    sp = sk.permute(0, 3, 2, 1).contiguous().to(dev).requires_grad_(True)
    ops.policy.narrow = "always"
    try:
        assert dec._use_narrow(ap, sp)
        y = dec.forward_phys(ap, sp, torch.float32)
        (y * up.permute(0, 3, 2, 1).to(dev)).sum().backward()
    finally:
        ops.policy.narrow = "auto"
    assert torch.allclose(y.detach().permute(0, 3, 2, 1).cpu(), ref.detach(), atol=1e-4, rtol=1e-4)
    assert torch.allclose(ap.grad.permute(0, 3, 2, 1).cpu(), ar.grad, atol=1e-4, rtol=1e-4)
    assert torch.allclose(sp.grad.permute(0, 3, 2, 1).cpu(), sr.grad, atol=1e-4, rtol=1e-4)
    for k, p in dec.named_parameters():
        assert torch.allclose(p.grad.cpu(), sd[k].grad, atol=1e-3, rtol=1e-3), k


def test_reviewkd_step_vs_oracle(dev):
    """ReviewKD recipe (distill_ReviewKD.py:69-139 on the local DCCRN, hcl on 4-D maps) against the
    oracle with injected ABF weights: loss terms and student gradients"""
    from clskd_b200.distill import DistillStep
    from oracle import losses_oracle as LO
    g, L = golden("dccrn.pt"), golden("losses.pt")
    teacher = _build(g["teacher_cfg"], g["t_sd"], dev)
    student = _build(g["student_cfg"], g["s_sd"], dev)
    student.train()
    X, y = g["X"].to(dev), g["y"].to(dev)
    step = DistillStep(teacher, student, mode="reviewkd")
    step.materialize(X)
    student.load_state_dict(full_sd(g["s_sd"]))
    _load_abf(step.abf_encoder, L["abf_enc_sd"])
    _load_abf(step.abf_decoder, L["abf_dec_sd"])
    loss = step(X, y)
    loss.backward()
    s_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                and not k.startswith(("stft.", "istft.")) else v) for k, v in full_sd(g["s_sd"]).items()}
    ref, terms = LO.clskd_step_loss(full_sd(g["t_sd"]), s_sd, g["X"], g["y"], L["abf_enc_sd"], L["abf_dec_sd"],
                                    mode="reviewkd")
    ref.backward()
    for k, v in terms.items():
        assert rel_err(step.last_terms[k].detach(), v.detach()) < 1e-4, k
    assert rel_err(loss.detach(), ref.detach()) < 1e-4
    for name, p in student.named_parameters():
        if bn_shadowed_bias(name) or s_sd[name].grad is None:
            continue
        gr = s_sd[name].grad
        assert torch.allclose(p.grad.cpu(), gr, rtol=5e-3, atol=2e-3 * gr.abs().max().item() + 1e-9), name


@pytest.mark.parametrize("training", [True, False])
def test_complex_batch_norm_vs_oracle(dev, training):
    """ComplexBatchNorm (tools_for_model.py:335-512; use_cbn=True) forward, train and eval statistics"""
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(3)
    m = tm.ComplexBatchNorm(12)
    for n_, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(0.5 + torch.rand(b.shape, generator=g) if ("RV" in n_ and "ri" not in n_)
                    else 0.1 * torch.randn(b.shape, generator=g))
    m.Br.data.normal_(generator=g)
    m.Bi.data.normal_(generator=g)
    p = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    x = torch.randn(3, 12, 6, 7, generator=g)
    upd = {}
    ref = D.complex_batch_norm(x, p, training, update=upd)
    m = m.to(dev)
    m.train(training)
    with torch.no_grad():
        out = m(x.to(dev))
    assert torch.allclose(out.cpu(), ref, atol=1e-5, rtol=1e-5)
    if training:
        for k, v in upd.items():
            assert torch.allclose(getattr(m, k).cpu(), v, atol=1e-6), k
        assert int(m.num_batches_tracked) == 1


@pytest.mark.parametrize("training", [True, False])
def test_complex_batch_norm_backward_vs_oracle_autograd(dev, training):
    """gradients of ComplexBatchNorm wrt the input (through the batch mean and 2x2 covariance when
    training) and wrt Wrr/Wri/Wii/Br/Bi against autograd of the oracle restatement"""
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(11)
    m = tm.ComplexBatchNorm(10)
    for n_, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(0.5 + torch.rand(b.shape, generator=g) if ("RV" in n_ and "ri" not in n_)
                    else 0.1 * torch.randn(b.shape, generator=g))
    for prm in (m.Wrr, m.Wii):
        prm.data.add_(0.3 * torch.randn(prm.shape, generator=g))
    m.Br.data.normal_(generator=g)
    m.Bi.data.normal_(generator=g)
    p = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    names = ["Wrr", "Wri", "Wii", "Br", "Bi"]
    for k in names:
        p[k].requires_grad_(True)
    x = (torch.randn(4, 10, 5, 9, generator=g) * torch.linspace(0.5, 2.0, 10).view(1, 10, 1, 1) + 0.3)
    gy = torch.randn(4, 10, 5, 9, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = D.complex_batch_norm(xr, p, training)
    (ref * gy).sum().backward()
    m = m.to(dev)
    m.train(training)
    xd = x.to(dev).requires_grad_(True)
    out = m(xd)
    (out * gy.to(dev)).sum().backward()
    assert torch.allclose(out.detach().cpu(), ref.detach(), atol=1e-5, rtol=1e-5)
    sx = xr.grad.abs().max().item()
    assert (xd.grad.cpu() - xr.grad).abs().max().item() < 2e-5 * max(sx, 1.0), "dx"
    for k in names:
        ref_g = p[k].grad
        got = getattr(m, k).grad.cpu()
        assert (got - ref_g).abs().max().item() < 1e-4 * max(ref_g.abs().max().item(), 1.0), k


def test_dccrn_with_complex_batch_norm_trains(dev):
    """use_cbn=True model variant: one SI-SNR training step runs end to end (gradients reach every
    ComplexBatchNorm parameter and the first encoder convolution)"""
    import clskd_b200
    torch.manual_seed(0)
    m = clskd_b200.DCCRN(rnn_units=16, use_clstm=True, use_cbn=True, kernel_num=[4, 8, 8, 16, 16, 16]).to(dev).train()
    gen = torch.Generator().manual_seed(0)
    x, y = 0.1 * torch.randn(2, 1600, generator=gen), 0.1 * torch.randn(2, 1600, generator=gen)
    wav = m(x.to(dev))[-1]
    loss = m.loss(wav, y.to(dev), loss_mode='SI-SNR')
    loss.backward()
    assert torch.isfinite(loss).all()
    for n_, prm in m.named_parameters():
        if n_.startswith("encoder.") and (".1." in n_ or n_.endswith("0.real_conv.weight")):
            assert prm.grad is not None and torch.isfinite(prm.grad).all(), n_
    assert m.encoder[0][1].Wrr.grad.abs().sum() > 0


@pytest.mark.parametrize("training", [False, True])
def test_dccrn_with_complex_batch_norm_vs_oracle(dev, training):
    """use_cbn=True model (DCCRN.py:80-81): enhanced waveform and SI-SNR gradients against the oracle"""
    import clskd_b200
    from oracle import dccrn_oracle as D
    torch.manual_seed(4)
    m = clskd_b200.DCCRN(rnn_units=16, use_clstm=True, use_cbn=True, kernel_num=[4, 8, 8, 16, 16, 16])
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    live = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and ".R" not in k
                                                   and not k.startswith(("stft.", "istft."))) else v)
            for k, v in sd.items()}
    gen = torch.Generator().manual_seed(1)
    x, y = 0.1 * torch.randn(3, 2400, generator=gen), 0.1 * torch.randn(3, 2400, generator=gen)
    ref_wav = D.dccrn_forward(live, x, training=training)[-1]
    from oracle import losses_oracle as LO
    ref_loss = -LO.si_snr(ref_wav, y)
    ref_loss.backward()
    m = m.to(dev).train(training)
    wav = m(x.to(dev))[-1]
    loss = m.loss(wav, y.to(dev), loss_mode='SI-SNR')
    loss.backward()
    assert (wav.detach().cpu() - ref_wav.detach()).abs().max().item() < 2e-5
    assert rel_err(loss, ref_loss) < 1e-4
    for name in ("encoder.0.1.Wri", "encoder.2.1.Wrr", "decoder.1.1.Bi", "encoder.1.0.real_conv.weight",
                 "decoder.4.1.Wii", "enhance.0.real_lstm.weight_ih_l0"):
        got, want = dict(m.named_parameters())[name].grad.cpu(), live[name].grad
        assert (got - want).abs().max().item() < 2e-3 * max(want.abs().max().item(), 1e-6) + 1e-6, name


@pytest.mark.parametrize("training", [False, True])
def test_dccrn_plain_lstm_bottleneck_vs_oracle(dev, training):
    """DCCRN(use_clstm=False) - the reference constructor's default: 2-layer nn.LSTM + Linear `tranform`
    (DCCRN.py:100-110, 193-199): state_dict keys, enhanced waveform and SI-SNR gradients vs the oracle"""
    import clskd_b200
    from oracle import dccrn_oracle as D
    from oracle import losses_oracle as LO
    torch.manual_seed(8)
    m = clskd_b200.DCCRN(rnn_units=24, use_clstm=False, kernel_num=[4, 8, 8, 16, 16, 16])
    keys = set(m.state_dict())
    for k in ("enhance.weight_ih_l0", "enhance.weight_hh_l1", "enhance.bias_ih_l1", "tranform.weight", "tranform.bias"):
        assert k in keys, k
    assert tuple(m.enhance.weight_ih_l0.shape) == (96, 64) and tuple(m.tranform.weight.shape) == (64, 24)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    live = {k: (v.clone().requires_grad_(True) if (v.is_floating_point() and "running" not in k
                                                   and not k.startswith(("stft.", "istft."))) else v)
            for k, v in sd.items()}
    gen = torch.Generator().manual_seed(2)
    x, y = 0.1 * torch.randn(3, 2400, generator=gen), 0.1 * torch.randn(3, 2400, generator=gen)
    ref_wav = D.dccrn_forward(live, x, training=training)[-1]
    ref_loss = -LO.si_snr(ref_wav, y)
    ref_loss.backward()
    m = m.to(dev).train(training)
    wav = m(x.to(dev))[-1]
    loss = m.loss(wav, y.to(dev), loss_mode='SI-SNR')
    loss.backward()
    assert (wav.detach().cpu() - ref_wav.detach()).abs().max().item() < 2e-5
    assert rel_err(loss, ref_loss) < 1e-4
    for name in ("enhance.weight_ih_l0", "enhance.weight_hh_l0", "enhance.bias_hh_l1", "enhance.weight_ih_l1",
                 "tranform.weight", "tranform.bias", "encoder.1.0.real_conv.weight", "decoder.0.0.imag_conv.weight"):
        got, want = dict(m.named_parameters())[name].grad.cpu(), live[name].grad
        assert (got - want).abs().max().item() < 2e-3 * max(want.abs().max().item(), 1e-6) + 1e-6, name


@pytest.mark.parametrize("name", ["cbn_model", "lstm_model"])
def test_model_variants_golden(dev, name):
    """product vs fixtures generated from the unmodified reference (tests/golden/variants.pt):
    DCCRN(use_cbn=True) and DCCRN(use_clstm=False) - waveform in eval and train mode, -SI-SNR, gradients"""
    import clskd_b200
    V = golden("variants.pt")[name]
    m = clskd_b200.DCCRN(masking_mode="E", kernel_num=V["kernel_num"], **V["kw"])
    m.load_state_dict(full_sd(V["sd"]))
    m = m.to(dev)
    x, y = V["x"].to(dev), V["y"].to(dev)
    m.eval()
    with torch.no_grad():
        assert (m(x)[-1].cpu() - V["wav_eval"]).abs().max().item() < 2e-5
    m.load_state_dict(full_sd(V["sd"]))          # eval left the running statistics untouched; be explicit anyway
    m.train()
    wav = m(x)[-1]
    assert (wav.detach().cpu() - V["wav_train"]).abs().max().item() < 2e-5
    loss = m.loss(wav, y, loss_mode='SI-SNR')
    assert rel_err(loss, V["loss_train"]) < 1e-4
    loss.backward()
    params = dict(m.named_parameters())
    for k, g in V["grads"].items():
        got = params[k].grad.cpu()
        assert (got - g).abs().max().item() < 2e-3 * max(g.abs().max().item(), 1e-6) + 1e-6, k


def test_dccrn_with_complex_batch_norm_runs(dev):
    """use_cbn=True model variant (DCCRN.py:80-81): forward under no_grad produces the reference shapes"""
    import clskd_b200
    m = clskd_b200.DCCRN(rnn_units=16, use_clstm=True, use_cbn=True, kernel_num=[4, 8, 8, 16, 16, 16]).to(dev).eval()
    with torch.no_grad():
        out = m(0.1 * torch.randn(2, 1600, generator=torch.Generator().manual_seed(0)).to(dev))
    assert out[-1].shape == (2, 1600) and torch.isfinite(out[-1]).all()
    assert out[0].shape == (2, 257, 19)


@pytest.mark.parametrize("L,chunk", [(6000, 16), (9037, 23), (4100, 30)])
def test_streaming_inference_equals_whole_utterance(dev, L, chunk):
    """Time-chunked streaming inference (carried LSTM state, 1-frame encoder halo, 6-frame decoder look-ahead,
    3-frame overlap-add halo) reproduces the whole-utterance forward and the oracle (fp32: <= 1e-5)."""
    from oracle import dccrn_oracle as D
    kn, ru = [4, 8, 8, 16, 16, 16], 16
    sd = D.make_state_dict(kn, ru, seed=21)
    m = _build(dict(kernel_num=kn, rnn_units=ru), sd, dev)
    m.eval()
    x = 0.1 * torch.randn(3, L, generator=torch.Generator().manual_seed(L))
    with torch.no_grad():
        ref = D.dccrn_forward(sd, x)[-1]
        whole = m(x.to(dev), is_feat=True)
        stream = m.enhance_streaming(x.to(dev), chunk_frames=chunk)
    assert stream.shape == whole.shape == ref.shape
    assert (stream.cpu() - whole.cpu()).abs().max().item() < 1e-5
    assert (stream.cpu() - ref).abs().max().item() < 1e-5
    m.train()
    with pytest.raises(RuntimeError):
        m.enhance_streaming(x.to(dev), chunk_frames=chunk)


# ------------------------------------------------------------------------------------ training module / metrics
def test_batched_validation_metrics_vs_oracle(dev):
    """SI-SDR (mean-removed, pb_bss_eval definition) / SNR and their improvements over the mixture for a whole batch
    from one library pass, against the numpy restatement"""
    from clskd_b200 import metrics
    from oracle import metrics_oracle as MO
    g = torch.Generator().manual_seed(3)
    clean = 0.1 * torch.randn(5, 4001, generator=g) + 0.01
    mix = clean + 0.05 * torch.randn(5, 4001, generator=g)
    est = 0.7 * clean + 0.01 * torch.randn(5, 4001, generator=g) - 0.02
    means, per = metrics.batch_metrics(mix.to(dev), clean.to(dev), est.to(dev))
    ref = MO.batch_metrics(mix.numpy(), clean.numpy(), est.numpy())
    for k, v in ref.items():
        assert torch.allclose(per[k].cpu(), torch.from_numpy(v), rtol=1e-6, atol=1e-6), k
        assert abs(means[k] - float(v.mean())) < 1e-5, k


def test_knowledge_distillation_module_hooks(dev):
    """the reference's LightningModule hook names over the fused step: training_step lowers the loss, validation_step
    returns batched metrics, configure_optimizers returns an optimizer; `fit` drives them without Lightning"""
    from clskd_b200.lightning import KnowledgeDistillation
    from oracle import dccrn_oracle as D
    cfg_t, cfg_s = dict(kernel_num=[4, 8, 8, 16, 16, 16], rnn_units=16), dict(kernel_num=[2, 4, 4, 8, 8, 8], rnn_units=8)
    teacher = _build(cfg_t, D.make_state_dict(cfg_t["kernel_num"], cfg_t["rnn_units"], seed=1), dev)
    student = _build(cfg_s, D.make_state_dict(cfg_s["kernel_num"], cfg_s["rnn_units"], seed=2), dev)
    g = torch.Generator().manual_seed(0)
    y = (0.1 * torch.randn(3, 2000, generator=g)).to(dev)
    X = y + (0.05 * torch.randn(3, 2000, generator=g)).to(dev)
    kd = KnowledgeDistillation(teacher, student, mode="spkd_all", lr=1e-3)
    hist = kd.fit([(X.unsqueeze(1), y.unsqueeze(1))] * 3, [(X, y)], epochs=2)
    assert hist[1]["train_loss"] < hist[0]["train_loss"]
    for k in ("si_sdr", "input_si_sdr", "si_sdr_imp", "snr", "snr_imp"):
        assert k in hist[1] and hist[1][k] == hist[1][k]
    assert abs(hist[1]["si_sdr_imp"] - (hist[1]["si_sdr"] - hist[1]["input_si_sdr"])) < 1e-6
    assert isinstance(kd.configure_optimizers(), torch.optim.Adam)
    assert "train_base" in kd.logged and student.training
