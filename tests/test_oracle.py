"""CPU tests: pin the oracle (oracle/*.py) against the golden fixtures that were produced by the
UNMODIFIED reference modules (tests/golden/make_golden.py), against the reference's own
known-answer docstring values, and - when /root/reference is present - against the live reference."""
import numpy as np
import pytest
import torch

from oracle import dccrn_oracle as D
from oracle import losses_oracle as LO
from oracle import ref_shim
from util import bn_shadowed_bias, check_summary, close, full_sd, golden, rel_err


@pytest.fixture(scope="module")
def g_dccrn():
    return golden("dccrn.pt")


@pytest.fixture(scope="module")
def g_losses():
    return golden("losses.pt")


@pytest.fixture(scope="module")
def g_step():
    return golden("step.pt")


def test_init_kernels_match_reference_buffers(g_dccrn):
    sd = full_sd({})
    for k, ref in g_dccrn["buffers"].items():
        check_summary(sd[k], ref, rtol=1e-6, atol=1e-7, what=k)


def test_stft_shapes_from_reference_notebook():
    # test_shape.ipynb cell 7: 3 s input -> 483 frames (T = L/100 + 3)
    w, _ = D.init_kernels(400, 100, 512, "hamming")
    assert D.conv_stft(torch.zeros(1, 48000), w, 400, 100).shape == (1, 514, 483)


@pytest.mark.parametrize("name", ["teacher", "student"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_dccrn_forward_matches_reference(g_dccrn, name, mode):
    ref = g_dccrn["%s_%s" % (name, mode)]
    sd = full_sd(g_dccrn["t_sd" if name == "teacher" else "s_sd"])
    taps, upd = {}, {}
    with torch.no_grad():
        mr, mi, re, im, wav = D.dccrn_forward(sd, g_dccrn["X"], training=(mode == "train"), taps=taps, update=upd)
    assert (wav - ref["wav"]).abs().max().item() <= 2e-6
    for t, key in ((mr, "mask_real"), (mi, "mask_imag"), (re, "real"), (im, "imag")):
        check_summary(t, ref[key], what=key)
    for kind in ("encoder", "decoder"):
        assert [tuple(t.shape) for t in taps[kind]] == ref[kind + "_shapes"]
        for i, t in enumerate(taps[kind]):
            check_summary(t, ref[kind][i], what="%s[%d]" % (kind, i))
    assert [tuple(t.shape) for t in taps["clstm"]] == ref["clstm_shapes"]
    for i, t in enumerate(taps["clstm"]):
        check_summary(t, ref["clstm"][i], what="clstm[%d]" % i)
    if mode == "train":
        for k, v in ref["running"].items():
            assert torch.allclose(upd[k], v, rtol=1e-5, atol=1e-6), k
    else:
        y = g_dccrn["y"]
        assert close(-LO.si_snr(wav, y), ref["loss"]["SI-SNR"])
        assert close(torch.nn.functional.mse_loss(wav, y), ref["loss"]["MSE"], atol=1e-9)
        assert close(-LO.sdr(y, wav), ref["loss"]["SDR"])
        assert close(-LO.si_sdr(y, wav), ref["loss"]["SI-SDR"])


def test_feature_shapes_of_reference_notebook():
    """test_shape.ipynb cells 1-7: quarter-width student on 8 x 3 s (checked on 1 x 3 s)."""
    sd = D.make_state_dict([8, 16, 32, 64, 64, 64], 64, seed=3)
    taps = {}
    with torch.no_grad():
        D.dccrn_forward(sd, torch.zeros(1, 48000), taps=taps)
    assert [tuple(t.shape[1:]) for t in taps["encoder"]] == [(8, 128, 483), (16, 64, 483), (32, 32, 483),
                                                             (64, 16, 483), (64, 8, 483), (64, 4, 483)]
    assert [tuple(t.shape[1:]) for t in taps["decoder"]] == [(64, 8, 484), (64, 16, 484), (32, 32, 484),
                                                             (16, 64, 484), (8, 128, 484), (2, 256, 484)]
    assert [tuple(t.shape) for t in taps["clstm"]] == [(483, 1, 128)] * 2


def test_si_sdr_known_answers():
    """docstring values of tools_for_loss.py:60-77"""
    np.random.seed(0)
    reference = np.random.randn(100)
    t = lambda a: torch.from_numpy(np.asarray(a)).float()[None]
    assert abs(float(LO.si_sdr(t(reference), t(reference * 2 + 1))) - 6.3704606) < 2e-4   # sanity: see below
    np.random.seed(0)
    reference = np.random.randn(100)
    assert abs(float(LO.si_sdr(t(reference), t(np.flip(reference).copy()))) - (-25.127672346460717)) < 1e-3
    assert abs(float(LO.si_sdr(t(reference), t(reference + np.flip(reference)))) - 0.481070445785553) < 1e-3
    assert abs(float(LO.si_sdr(t(reference), t(reference + 0.5))) - 6.3704606032577304) < 1e-3
    # si_snr(ref + 0.5, ref) gives the same value
    assert abs(float(LO.si_snr(t(reference + 0.5), t(reference))) - 6.3704606) < 1e-3


def test_objectives_match_reference(g_losses):
    a, b = g_losses["a"], g_losses["b"]
    assert close(LO.si_snr(a, b), g_losses["si_snr"])
    assert close(LO.sdr(a, b), g_losses["sdr"])
    assert close(LO.si_sdr(a, b), g_losses["si_sdr"])
    assert rel_err(LO.spkd(g_losses["zs"], g_losses["zt"], "batchmean"), g_losses["spkd_batchmean"]) < 1e-5
    assert rel_err(LO.spkd(g_losses["zs"], g_losses["zt"], "sum"), g_losses["spkd_sum"]) < 1e-5


def test_stft_losses_match_reference(g_losses):
    sx, sy = g_losses["sx"], g_losses["sy"]
    sc, mag = LO.stft_loss(sx, sy, 512, 100, 400)
    assert rel_err(sc, g_losses["stft_512_100_400"][0]) < 1e-5 and rel_err(mag, g_losses["stft_512_100_400"][1]) < 1e-5
    for key, cfg in (("mrstft_distill", ([512], [100], [400])), ("mrstft_reviewkd", ([512], [16], [32])),
                     ("mrstft_3res", ([256, 512, 128], [30, 60, 12], [150, 300, 60]))):
        sc, mag = LO.mr_stft_loss(sx, sy, *cfg)
        assert rel_err(sc, g_losses[key][0]) < 1e-5 and rel_err(mag, g_losses[key][1]) < 1e-5, key
    check_summary(LO.stft_mag(sx, 512, 100, 400, torch.hann_window(400)), g_losses["stft_mag_512"], what="stft_mag")


def _student_taps(g_dccrn):
    sd = full_sd(g_dccrn["s_sd"])
    taps = {}
    with torch.no_grad():
        D.dccrn_forward(sd, g_dccrn["X"], training=True, taps=taps)
    return taps


def test_review_kd_matches_reference(g_dccrn, g_losses):
    taps = _student_taps(g_dccrn)
    e_shapes = [m.shape[2] for m in taps["encoder"]][::-1]
    d_shapes = [m.shape[2] for m in taps["decoder"]]
    with torch.no_grad():
        f_enc = LO.review_kd_forward(taps["encoder"], g_losses["abf_enc_sd"], e_shapes, e_shapes, "encoder")
        f_dec = LO.review_kd_forward(taps["decoder"], g_losses["abf_dec_sd"], d_shapes, d_shapes, "decoder")
    assert [tuple(t.shape) for t in f_enc] == g_losses["abf_enc_shapes"]
    assert [tuple(t.shape) for t in f_dec] == g_losses["abf_dec_shapes"]
    for i, t in enumerate(f_enc):
        check_summary(t, g_losses["abf_enc_out"][i], what="abf_enc[%d]" % i)
    for i, t in enumerate(f_dec):
        check_summary(t, g_losses["abf_dec_out"][i], what="abf_dec[%d]" % i)


@pytest.mark.parametrize("mode", ["clskd", "spkd_all", "spkd", "mse", "stft"])
def test_step_loss_and_grads_match_reference(g_dccrn, g_losses, g_step, mode):
    ref = g_step[mode]
    t_sd = full_sd(g_dccrn["t_sd"])
    s_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                and not k.startswith(("stft.", "istft.")) else v) for k, v in full_sd(g_dccrn["s_sd"]).items()}
    loss, terms = LO.clskd_step_loss(t_sd, s_sd, g_dccrn["X"], g_dccrn["y"], g_losses["abf_enc_sd"],
                                     g_losses["abf_dec_sd"], mode=mode)
    assert rel_err(loss, ref["loss"]) < 1e-4
    for k, v in ref["terms"].items():
        assert rel_err(terms[k], v) < 1e-4, k
    loss.backward()
    for name, gref in ref["grads"].items():
        if bn_shadowed_bias(name):
            continue
        g = s_sd[name].grad
        assert g is not None, name
        check_summary(g, gref, rtol=2e-3, atol=1e-7, what="grad " + name)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_oracle_against_live_reference():
    mods = ref_shim.load()
    kn, ru = [4, 8, 8, 16, 16, 16], 16
    sd = D.make_state_dict(kn, ru, seed=11)
    m = mods["DCCRN"].DCCRN(rnn_units=ru, masking_mode="E", use_clstm=True, kernel_num=kn)
    m.load_state_dict(sd)
    m.eval()
    x = 0.1 * torch.randn(3, 3000, generator=torch.Generator().manual_seed(5))
    for mode in ("E", "C", "R"):
        m.masking_mode = mode
        with torch.no_grad():
            ref = m(x)
            out = D.dccrn_forward(sd, x, masking_mode=mode)
        for a, b in zip(out, ref):
            assert (a - b).abs().max().item() < 1e-5, mode


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("training", [True, False])
def test_complex_batch_norm_oracle_against_live_reference(training):
    import warnings
    mods = ref_shim.load()
    g = torch.Generator().manual_seed(3)
    m = mods["tools_for_model"].ComplexBatchNorm(12)
    for n_, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(0.5 + torch.rand(b.shape, generator=g) if "RV" in n_ and "ri" not in n_ else 0.1 * torch.randn(b.shape, generator=g))
    m.Br.data.normal_(generator=g)
    m.Bi.data.normal_(generator=g)
    p = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    x = torch.randn(3, 12, 5, 7, generator=g)
    m.train(training)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with torch.no_grad():
            ref = m(x)
    upd = {}
    out = D.complex_batch_norm(x, p, training, update=upd)
    assert torch.allclose(out, ref, atol=1e-5, rtol=1e-5)
    if training:
        for k, v in upd.items():
            assert torch.allclose(v, getattr(m, k), atol=1e-6), k


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("training", [True, False])
def test_complex_batch_norm_gradients_oracle_against_live_reference(training):
    """pins the oracle's ComplexBatchNorm BACKWARD: autograd of the restatement == autograd of the
    unmodified reference module (input and affine-parameter gradients)"""
    import warnings
    mods = ref_shim.load()
    g = torch.Generator().manual_seed(9)
    m = mods["tools_for_model"].ComplexBatchNorm(8)
    for n_, b in m.named_buffers():
        if b.is_floating_point():
            b.copy_(0.5 + torch.rand(b.shape, generator=g) if "RV" in n_ and "ri" not in n_ else 0.1 * torch.randn(b.shape, generator=g))
    m.Br.data.normal_(generator=g)
    m.Bi.data.normal_(generator=g)
    p = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    for k in ("Wrr", "Wri", "Wii", "Br", "Bi"):
        p[k].requires_grad_(True)
    x = torch.randn(3, 8, 4, 6, generator=g) + 0.2
    gy = torch.randn(3, 8, 4, 6, generator=g)
    m.train(training)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        (m(xa) * gy).sum().backward()
    (D.complex_batch_norm(xb, p, training) * gy).sum().backward()
    assert torch.allclose(xb.grad, xa.grad, atol=1e-5, rtol=1e-4)
    for k in ("Wrr", "Wri", "Wii", "Br", "Bi"):
        assert torch.allclose(p[k].grad, getattr(m, k).grad, atol=1e-4, rtol=1e-4), k


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_use_cbn_model_oracle_against_live_reference():
    """DCCRN(use_cbn=True) (DCCRN.py:80-81): the oracle forward follows the live reference in eval and
    train mode (the norm slot holds ComplexBatchNorm, inferred from the state_dict keys)"""
    import warnings
    mods = ref_shim.load()
    kn, ru = [4, 8, 8, 16, 16, 16], 16
    torch.manual_seed(21)
    m = mods["DCCRN"].DCCRN(rnn_units=ru, masking_mode="E", use_clstm=True, use_cbn=True, kernel_num=kn)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = 0.1 * torch.randn(2, 2400, generator=torch.Generator().manual_seed(6))
    for training in (False, True):
        m.train(training)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with torch.no_grad():
                ref = m(x)
                out = D.dccrn_forward(sd, x, training=training)
        for a, b in zip(out, ref):
            assert (a - b).abs().max().item() < 2e-5, training


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present (GPU box)")
def test_plain_lstm_model_oracle_against_live_reference():
    """DCCRN(use_clstm=False) (DCCRN.py:100-110, 193-199): oracle forward == live reference"""
    mods = ref_shim.load()
    kn, ru = [4, 8, 8, 16, 16, 16], 24
    torch.manual_seed(33)
    m = mods["DCCRN"].DCCRN(rnn_units=ru, masking_mode="E", use_clstm=False, kernel_num=kn).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = 0.1 * torch.randn(2, 2400, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        ref = m(x)
        out = D.dccrn_forward(sd, x)
    for a, b in zip(out, ref):
        assert (a - b).abs().max().item() < 1e-5


# ---------------------------------------------------------------- golden fixtures of the model variants
# (tests/golden/variants.pt, generated from the UNMODIFIED reference by tests/golden/make_golden_variants.py;
#  these run everywhere, also where /root/reference does not exist)
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_complex_batch_norm_oracle_vs_golden(mode):
    V = golden("variants.pt")["cbn"]
    p = {k: v.clone() for k, v in V["state"].items()}
    for k in ("Wrr", "Wri", "Wii", "Br", "Bi"):
        p[k].requires_grad_(True)
    x = V["x"].clone().requires_grad_(True)
    upd = {}
    y = D.complex_batch_norm(x, p, mode == "train", update=upd)
    (y * V["gy"]).sum().backward()
    ref = V[mode]
    assert torch.allclose(y.detach(), ref["y"], atol=1e-5, rtol=1e-5)
    assert torch.allclose(x.grad, ref["dx"], atol=1e-5, rtol=1e-4)
    for k, g in ref["grads"].items():
        assert torch.allclose(p[k].grad, g, atol=1e-4, rtol=1e-4), k
    if mode == "train":
        for k, v in upd.items():
            assert torch.allclose(v, ref["running"][k], atol=1e-6), k


@pytest.mark.parametrize("name", ["cbn_model", "lstm_model"])
def test_model_variants_oracle_vs_golden(name):
    """DCCRN(use_cbn=True) and DCCRN(use_clstm=False): waveform (eval, train), -SI-SNR and gradients"""
    from oracle import losses_oracle as LO
    V = golden("variants.pt")[name]
    sd = full_sd(V["sd"])
    with torch.no_grad():
        assert (D.dccrn_forward(sd, V["x"])[-1] - V["wav_eval"]).abs().max().item() < 2e-5
    live = {k: (v.clone().requires_grad_(True) if k in V["grads"] else v) for k, v in sd.items()}
    wav = D.dccrn_forward(live, V["x"], training=True)[-1]
    assert (wav.detach() - V["wav_train"]).abs().max().item() < 2e-5
    loss = -LO.si_snr(wav, V["y"])
    assert rel_err(loss, V["loss_train"]) < 1e-4
    loss.backward()
    for k, g in V["grads"].items():
        assert (live[k].grad - g).abs().max().item() < 2e-3 * max(g.abs().max().item(), 1e-6) + 1e-6, k


def test_oracle_step_at_benchmarked_widths_vs_reference_fixture():
    """tests/golden/real_width.pt (unmodified reference, teacher [32..256]/256 -> half student, B = 4 x 4 s,
    T = 643): the oracle's CLSKD step loss, its 5 terms and the student / ABF gradients at the benchmarked widths."""
    from oracle import dccrn_oracle as D
    from oracle import losses_oracle as LO
    from util import sample_rel_l2
    G = golden("real_width.pt")
    g = torch.Generator().manual_seed(G["seeds"]["data"])
    X, y = 0.1 * torch.randn(G["B"], G["L"], generator=g), 0.1 * torch.randn(G["B"], G["L"], generator=g)
    tc, sc, rec = G["teacher_cfg"], G["student_cfgs"]["half"], G["half"]
    t_sd = D.make_state_dict(tc["kernel_num"], tc["rnn_units"], seed=G["seeds"]["teacher"])
    s_sd = D.make_state_dict(sc["kernel_num"], sc["rnn_units"], seed=G["seeds"]["student"])
    a = rec["abf_args"]
    e_sd = LO.make_abf_state_dict(a["enc_in"], a["enc_out"], G["seeds"]["abf_enc"])
    d_sd = LO.make_abf_state_dict(a["dec_in"], a["dec_out"], G["seeds"]["abf_dec"])
    leaf = lambda sd: {k: (v.clone().requires_grad_(True) if v.is_floating_point() and "running" not in k
                           and not k.startswith(("stft.", "istft.")) else v) for k, v in sd.items()}
    s_sd, e_sd, d_sd = leaf(s_sd), leaf(e_sd), leaf(d_sd)
    loss, terms = LO.clskd_step_loss(t_sd, s_sd, X, y, abf_enc_sd=e_sd, abf_dec_sd=d_sd, mode="clskd")
    ref = rec["clskd_eval"]
    assert rel_err(loss.detach(), ref["loss"]) < 1e-5
    for k, v in ref["terms"].items():
        assert rel_err(terms[k].detach(), v) < 1e-4, k
    loss.backward()
    for name, gref in ref["grads"].items():
        if bn_shadowed_bias(name):
            continue
        assert sample_rel_l2(s_sd[name].grad, gref) < 2e-3, name
    for key, sd in (("abf_enc_grads", e_sd), ("abf_dec_grads", d_sd)):
        for name, gref in ref[key].items():
            if gref["l2"] > 1e-12:
                assert sample_rel_l2(sd[name].grad, gref) < 2e-3, (key, name)
