"""GPU-only parity tests of the bf16 / tcgen05 policy: the tensor-core implicit GEMM against the
fp32 CUDA-core kernel and the CPU oracle on identical (bf16-representable) operands, and the full
model / distillation step against the oracle within the north_star bf16 tolerances
(enhanced-waveform max-abs error <= 1e-3, loss relative error <= 1e-3)."""
import pytest
import torch

from util import full_sd, golden, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda_dev():
    import clskd_b200
    assert torch.cuda.is_available()
    clskd_b200._lib.load()
    yield torch.device("cuda:0")
    clskd_b200.set_precision("fp32")


def _round_params(mod):
    for p in mod.parameters():
        p.data = p.data.bfloat16().float()


@pytest.mark.parametrize("cin,cout,F,T,B", [(32, 64, 64, 37, 3), (64, 128, 16, 130, 2), (256, 256, 8, 65, 2),
                                            (16, 32, 128, 9, 1)])
def test_umma_complex_conv_vs_fp32_kernel_and_oracle(cuda_dev, cin, cout, F, T, B):
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(cin + F)
    conv = tm.ComplexConv2d(cin, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1))
    conv.real_conv.bias.data.normal_(generator=g)
    conv.imag_conv.bias.data.normal_(generator=g)
    _round_params(conv)
    x = torch.randn(B, cin, F, T, generator=g).bfloat16().float()
    ref = D.complex_conv2d(x, conv.real_conv.weight, conv.real_conv.bias, conv.imag_conv.weight, conv.imag_conv.bias)
    conv = conv.to(cuda_dev)
    xp = x.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16()        # physical [B,T,F,C]
    with torch.no_grad():
        clskd_b200.set_precision("bf16")
        n0 = ops.umma_launches
        y_umma = conv.forward_phys(xp, torch.float32)
        assert ops.umma_launches == n0 + 1, "tcgen05 path was not taken"
        clskd_b200.set_precision("fp32")
        y_core = conv.forward_phys(xp, torch.float32)
    y_umma, y_core = y_umma.permute(0, 3, 2, 1).cpu(), y_core.permute(0, 3, 2, 1).cpu()
    scale = ref.abs().max().item()
    assert (y_core - ref).abs().max().item() < 2e-5 * max(scale, 1)
    assert (y_umma - ref).abs().max().item() < 2e-5 * max(scale, 1)      # same operands, fp32 accumulation


@pytest.mark.parametrize("cin,cskip,cout,F,T,B", [(64, 64, 32, 8, 33, 2), (256, 256, 256, 4, 70, 2), (32, 32, 16, 64, 11, 1)])
def test_umma_complex_deconv_with_skip_vs_oracle(cuda_dev, cin, cskip, cout, F, T, B):
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(cin + F)
    dec = tm.ComplexConvTranspose2d(cin + cskip, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0),
                                    output_padding=(1, 0))
    dec.real_conv.bias.data.normal_(generator=g)
    _round_params(dec)
    a = torch.randn(B, cin, F, T, generator=g).bfloat16().float()
    s = torch.randn(B, cskip, F, T, generator=g).bfloat16().float()
    ref = D.complex_deconv2d(D.complex_cat([a, s], 1), dec.real_conv.weight, dec.real_conv.bias,
                             dec.imag_conv.weight, dec.imag_conv.bias)
    dec = dec.to(cuda_dev)
    ap = a.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16()
    sp = s.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16()
    with torch.no_grad():
        clskd_b200.set_precision("bf16")
        n0 = ops.umma_launches
        y = dec.forward_phys(ap, sp, torch.float32)
        assert ops.umma_launches == n0 + 2, "two sub-pixel tcgen05 launches expected"
    y = y.permute(0, 3, 2, 1).cpu()
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() < 2e-5 * max(ref.abs().max().item(), 1)


def test_umma_bf16_output_and_3x3(cuda_dev):
    """ABF's 3x3 real conv (9 taps, stride 1) with bf16 output"""
    import clskd_b200
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(0)
    conv = fw.RealConv2d(64, 128, 3, padding=1, bias=False)
    _round_params(conv)
    x = torch.randn(2, 64, 16, 21, generator=g).bfloat16().float()
    ref = torch.nn.functional.conv2d(x, conv.weight, padding=1)
    conv = conv.to(cuda_dev)
    xp = x.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16()
    with torch.no_grad():
        clskd_b200.set_precision("bf16")
        n0 = ops.umma_launches
        y = conv.forward_phys(xp)
        assert ops.umma_launches == n0 + 1 and y.dtype == torch.bfloat16
    y = y.float().permute(0, 3, 2, 1).cpu()
    assert (y - ref).abs().max().item() < 1e-2 * ref.abs().max().item()      # bf16 output rounding (2^-8)


@pytest.mark.parametrize("width", ["golden_teacher", "half"])
def test_bf16_forward_waveform_tolerance(cuda_dev, width):
    import clskd_b200
    from clskd_b200 import ops
    from oracle import dccrn_oracle as D
    if width == "golden_teacher":
        g = golden("dccrn.pt")
        cfg, sd, x = g["teacher_cfg"], full_sd(g["t_sd"]), g["X"]
    else:
        cfg = dict(kernel_num=[16, 32, 64, 128, 128, 128], rnn_units=128)
        sd = D.make_state_dict(cfg["kernel_num"], cfg["rnn_units"], seed=5)
        x = 0.1 * torch.randn(2, 6000, generator=torch.Generator().manual_seed(1))
    m = clskd_b200.DCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
    m.load_state_dict(sd)
    m = m.to(cuda_dev).eval()
    with torch.no_grad():
        ref = D.dccrn_forward(sd, x)[-1]
        clskd_b200.set_precision("bf16")
        n0 = ops.umma_launches
        wav = m(x.to(cuda_dev), is_feat=True).float().cpu()
    assert ops.umma_launches > n0
    assert (wav - ref).abs().max().item() <= 1e-3


def test_bf16_distill_step_loss_tolerance(cuda_dev):
    import clskd_b200
    from clskd_b200.distill import DistillStep
    from oracle import dccrn_oracle as D
    from oracle import losses_oracle as LO
    cfg_t = dict(kernel_num=[32, 64, 64, 64, 64, 64], rnn_units=64)
    cfg_s = dict(kernel_num=[16, 32, 32, 32, 32, 32], rnn_units=32)
    t_sd = D.make_state_dict(cfg_t["kernel_num"], cfg_t["rnn_units"], seed=1)
    s_sd = D.make_state_dict(cfg_s["kernel_num"], cfg_s["rnn_units"], seed=2)

    def mk(cfg, sd):
        m = clskd_b200.DCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
        m.load_state_dict(sd)
        return m.to(cuda_dev)
    g = torch.Generator().manual_seed(0)
    X, y = 0.1 * torch.randn(4, 4000, generator=g), 0.1 * torch.randn(4, 4000, generator=g)
    ref, terms = LO.clskd_step_loss(t_sd, s_sd, X, y, mode="spkd_all")
    clskd_b200.set_precision("bf16")
    student = mk(cfg_s, s_sd)
    student.train()
    step = DistillStep(mk(cfg_t, t_sd), student, mode="spkd_all")
    loss = step(X.to(cuda_dev), y.to(cuda_dev))
    loss.backward()
    for k, v in terms.items():
        assert rel_err(step.last_terms[k].detach(), v.detach()) < 2e-2, k     # per-term (small SPKD terms are noisier)
    assert rel_err(loss.detach(), ref.detach()) < 1e-3
    assert all(torch.isfinite(p.grad).all() for p in student.parameters() if p.grad is not None)


def _wgrad_both_policies(cuda_dev, mod, inputs, ops):
    """weight gradients of `mod.forward_phys(*inputs)` under the bf16 (tcgen05) and fp32 (CUDA-core)
    policies for the same bf16-representable operands and upstream gradient"""
    import clskd_b200
    g = torch.Generator().manual_seed(7)
    res = {}
    up = None
    for pol in ("bf16", "fp32"):
        clskd_b200.set_precision(pol)
        mod.zero_grad()
        n0 = ops.umma_launches
        y = mod.forward_phys(*inputs)
        if up is None:
            up = torch.randn(y.shape, generator=g).bfloat16().float().to(cuda_dev)
        (y.float() * up).sum().backward()
        res[pol] = ({k: p.grad.detach().clone() for k, p in mod.named_parameters() if p.grad is not None},
                    ops.umma_launches - n0)
    return res


@pytest.mark.parametrize("wmode", [3, 4, 0])
@pytest.mark.parametrize("case", ["conv_64_128", "conv_32_64", "conv_16_32", "deconv_skip_128", "deconv_skip_64",
                                  "abf3x3_128_256", "abf3x3_128_32", "abf1x1_32_128", "abf3x3F128_128_32",
                                  "abf3x3F128_32_128", "abf3x3F256_64_48", "abf3x3F64_128_64",
                                  "conv_8_16", "conv_16_8", "deconv_skip_16", "abf1x1_8_64", "abf3x3_24_40"])
def test_umma_wgrad_vs_cuda_core(cuda_dev, case, wmode):
    """wmode 3: operand reuse wherever possible (full halo patch at F >= 128, time-grouped patches below); 1: one box per tap"""
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    g = torch.Generator().manual_seed(len(case))

    def act(B, T, F, C):
        return torch.randn(B, T, F, C, generator=g).bfloat16().to(cuda_dev)
    if case.startswith("conv_"):
        cin, cout = (int(v) for v in case.split("_")[1:])
        mod = tm.ComplexConv2d(cin, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1))
        inputs = (act(3, 37, 32, cin),)
    elif case.startswith("deconv_skip_"):
        c = int(case.split("_")[2])
        mod = tm.ComplexConvTranspose2d(2 * c, c // 2, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0),
                                        output_padding=(1, 0))
        inputs = (act(2, 33, 16, c), act(2, 33, 16, c))
    elif case.startswith("abf3x3F"):
        F_ = int(case.split("_")[0][7:])
        cin, cout = (int(v) for v in case.split("_")[1:])
        mod = fw.RealConv2d(cin, cout, 3, padding=1, bias=False)
        inputs = (act(2, 9, F_, cin),)
    elif case.startswith("abf3x3_"):
        cin, cout = (int(v) for v in case.split("_")[1:])
        mod = fw.RealConv2d(cin, cout, 3, padding=1, bias=False)
        inputs = (act(2, 41, 32, cin),)
    else:
        cin, cout = (int(v) for v in case.split("_")[1:])
        mod = fw.RealConv2d(cin, cout, 1, bias=False)
        inputs = (act(2, 41, 64, cin),)
    _round_params(mod)
    mod = mod.to(cuda_dev)
    from clskd_b200 import _lib
    _lib.load().clskd_set_tuning(6, wmode)
    try:
        res = _wgrad_both_policies(cuda_dev, mod, inputs, ops)
    finally:
        _lib.load().clskd_set_tuning(6, 0)
    assert res["bf16"][1] >= 2 and res["fp32"][1] == 0, "tcgen05 forward + wgrad launches expected"
    for k, gref in res["fp32"][0].items():
        gu = res["bf16"][0][k]
        scale = gref.abs().max().item()
        assert (gu - gref).abs().max().item() < 1e-4 * max(scale, 1.0), (case, k)


@pytest.mark.parametrize("B,K", [(64, 128 * 643), (5, 40008), (128, 8192 + 64), (16, 4096), (256, 16384 + 64), (192, 8192),
                                 (130, 4096 + 128)])
def test_umma_gram_fwd_bwd_vs_torch(cuda_dev, B, K):
    import clskd_b200
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(B)
    z = torch.randn(B, K, generator=g).bfloat16()
    zt = torch.randn(B, K // 2, generator=g).bfloat16()
    zd = z.to(cuda_dev).requires_grad_(True)
    clskd_b200.set_precision("bf16")
    n0 = ops.umma_launches
    G = ops.gram(zd.detach())
    assert ops.umma_launches == n0 + 1, "tcgen05 Gram path was not taken"
    ref = z.double() @ z.double().t()
    assert (G.cpu().double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()
    # SPKD loss + gradient through the tensor-core Gram / Gram-gradient kernels vs torch autograd (fp64)
    from oracle import losses_oracle as LO
    loss = ops.SPKDFn.apply(zd, zt.to(cuda_dev), 1.0 / (B * B))
    loss.backward()
    z64 = z.double().requires_grad_(True)
    lref = LO.spkd(z64, zt.double())
    lref.backward()
    assert abs(float(loss) - float(lref)) < 1e-4 * abs(float(lref))
    gref = z64.grad
    err = (zd.grad.float().cpu().double() - gref).abs().max().item()
    assert err < 1.5e-2 * gref.abs().max().item(), err          # dz is rounded to bf16 (2^-8 relative)


@pytest.mark.parametrize("cin,cout,ks,F,T", [(128, 2, 3, 32, 45), (64, 1, 3, 32, 45), (32, 2, 3, 32, 45), (32, 2, 3, 128, 301),
                                             (64, 2, 3, 256, 131)])
def test_tap_in_channel_narrow_conv_vs_torch(cuda_dev, cin, cout, ks, F, T):
    """bf16 policy: k x k conv onto <= 2 channels = pointwise tcgen05 GEMM + tap gather-sum (the last two sizes take
    the shared-memory tiled gather-sum kernels, >= 65536 positions)"""
    import clskd_b200
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(cin + ks)
    conv = fw.RealConv2d(cin, cout, ks, padding=ks // 2, bias=False)
    _round_params(conv)
    x = torch.randn(2, cin, F, T, generator=g).bfloat16().float()
    up = torch.randn(2, cout, F, T, generator=g).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    wr = conv.weight.detach().clone().requires_grad_(True)
    ref = torch.nn.functional.conv2d(xr, wr, padding=ks // 2)
    (ref * up).sum().backward()
    conv = conv.to(cuda_dev)
    clskd_b200.set_precision("bf16")
    xp = x.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16().requires_grad_(True)
    assert conv._use_narrow(xp, None)
    n0 = ops.umma_launches
    y = conv.forward_phys(xp, out_dtype=torch.float32)
    (y * up.permute(0, 3, 2, 1).to(cuda_dev)).sum().backward()
    assert ops.umma_launches >= n0 + 3                      # pointwise fwd, dgrad and wgrad on tensor cores
    yl = y.detach().permute(0, 3, 2, 1).cpu()
    s = ref.abs().max().item()
    assert (yl - ref.detach()).abs().max().item() < 2e-2 * s       # z is rounded to bf16 per tap
    gx = xp.grad.float().permute(0, 3, 2, 1).cpu()
    assert (gx - xr.grad).abs().max().item() < 2e-2 * xr.grad.abs().max().item()
    gw = conv.weight.grad.cpu()
    assert (gw - wr.grad).abs().max().item() < 2e-2 * wr.grad.abs().max().item()


@pytest.mark.parametrize("B,L", [(8, 16000), (2, 64000)])
def test_split_bf16_stft_istft_vs_oracle(cuda_dev, B, L):
    """STFT / iSTFT DFT GEMMs of the bf16 policy run on tcgen05 with split-bf16 (hi+lo) operands:
    spectra, round trip and the waveform gradient must stay at fp32-GEMM accuracy (1e-5 relative to
    the largest value), far inside the bf16 policy's 1e-3 waveform budget."""
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(L)
    x = 0.1 * torch.randn(B, L, generator=g)
    stft, istft = tm.ConvSTFT(400, 100, 512, 'hamming', 'complex').to(cuda_dev), \
        tm.ConviSTFT(400, 100, 512, 'hamming', 'complex').to(cuda_dev)
    w_a, _ = D.init_kernels(400, 100, 512, 'hamming')
    w_s, win = D.init_kernels(400, 100, 512, 'hamming', invers=True)
    ref_spec = D.conv_stft(x, w_a, 400, 100)
    gw = torch.randn(ref_spec.shape, generator=g)
    xr = x.clone().requires_grad_(True)
    spec_r = D.conv_stft(xr, w_a, 400, 100)
    wav_r = D.conv_istft(spec_r * 0.5, w_s, win, 400, 100)
    (wav_r * wav_r).sum().backward()
    clskd_b200.set_precision("bf16")
    n0 = ops.umma_launches
    xd = x.to(cuda_dev).requires_grad_(True)
    spec = stft(xd)
    assert ops.umma_launches == n0 + 1, "split-bf16 tcgen05 route was not taken for the STFT GEMM"
    wav = istft(spec * 0.5)
    assert ops.umma_launches == n0 + 2, "split-bf16 tcgen05 route was not taken for the iSTFT GEMM"
    (wav * wav).sum().backward()
    assert ops.umma_launches == n0 + 4, "data gradients of the DFT GEMMs must take the split route too"
    s = ref_spec.abs().max().item()
    assert (spec.detach().cpu() - ref_spec).abs().max().item() < 1e-5 * s
    assert (wav.detach().cpu() - wav_r.detach()).abs().max().item() < 1e-5 * max(wav_r.abs().max().item(), 1e-3) + 2e-7
    gs = xr.grad.abs().max().item()
    assert (xd.grad.cpu() - xr.grad).abs().max().item() < 2e-5 * gs


def test_split_bf16_gemm_bias_and_padding(cuda_dev):
    """the split route pads K to 64 and N to the tcgen05 tile (scratch + strided copy) and adds the bias"""
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200.clstm import _LinearParams
    g = torch.Generator().manual_seed(5)
    lin = _LinearParams(72, 200)
    x = torch.randn(5000, 72, generator=g)
    ref = x.double() @ lin.weight.detach().double().t() + lin.bias.detach().double()
    lin = lin.to(cuda_dev)
    clskd_b200.set_precision("bf16")
    n0 = ops.umma_launches
    with torch.no_grad():
        y = lin.forward_rows(x.to(cuda_dev))
    assert ops.umma_launches == n0 + 1
    assert (y.cpu().double() - ref).abs().max().item() < 2e-5 * ref.abs().max().item()


@pytest.mark.parametrize("kind,cin,cout,F,T,B", [("conv", 32, 64, 64, 37, 3), ("conv", 64, 256, 16, 130, 2),
                                                  ("deconv", 64, 32, 8, 33, 2), ("deconv", 32, 16, 64, 70, 1)])
def test_fused_conv_epilogue_batch_statistics_and_folded_eval_bn(cuda_dev, kind, cin, cout, F, T, B):
    """tcgen05 conv epilogue: (a) train-mode BatchNorm statistics accumulated from the stored outputs
    (sub-pixel phases of the transposed conv accumulate into one buffer, ragged last time tile masked),
    (b) eval-mode BatchNorm folded to scale/shift + PReLU applied before the single store.  Both against
    the separate statistics / normalise kernels on the same operands, and (a) against the fp32 oracle."""
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    g = torch.Generator().manual_seed(cin + F + T)
    if kind == "conv":
        conv = tm.ComplexConv2d(cin, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1))
        x1 = None
    else:
        conv = tm.ComplexConvTranspose2d(2 * cin, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 0),
                                         output_padding=(1, 0))
        x1 = torch.randn(B, T, F, cin, generator=g).bfloat16().to(cuda_dev)
    conv.real_conv.bias.data.normal_(generator=g)
    conv.imag_conv.bias.data.normal_(generator=g)
    blk = tm.ConvBNAct(conv, tm.BatchNorm2d(cout), tm.PReLU()).to(cuda_dev)
    blk[1].weight.data.uniform_(0.5, 1.5, generator=None)
    blk[1].bias.data.normal_()
    blk[1].running_mean.normal_(0, 0.1)
    blk[1].running_var.uniform_(0.5, 1.5)
    x0 = torch.randn(B, T, F, cin, generator=g).bfloat16().to(cuda_dev)
    clskd_b200.set_precision("bf16")
    outs = {}
    for fuse in (True, False):
        ops.policy.fuse_epilogue = fuse
        for mode in ("train", "eval"):
            blk.train(mode == "train")
            rm0, rv0 = blk[1].running_mean.clone(), blk[1].running_var.clone()
            n0 = ops.fused_epilogues
            with torch.no_grad():
                y = blk.forward_phys(x0, x1)
            assert (ops.fused_epilogues - n0) == (1 if fuse else 0), (fuse, mode)
            outs[(fuse, mode)] = (y.float().cpu(), blk[1].running_mean.clone().cpu(), blk[1].running_var.clone().cpu())
            blk[1].running_mean.copy_(rm0)
            blk[1].running_var.copy_(rv0)
    ops.policy.fuse_epilogue = True
    yf, rmf, rvf = outs[(True, "train")]
    yu, rmu, rvu = outs[(False, "train")]
    scale = yu.abs().max().item()
    # same stored bf16 conv output, statistics differ only by fp32-vs-fp64 partial sums
    assert (rmf - rmu).abs().max().item() < 1e-5 * max(1.0, rmu.abs().max().item())
    assert (rvf - rvu).abs().max().item() < 1e-4 * max(1.0, rvu.abs().max().item())
    assert (yf - yu).abs().max().item() < 1e-2 * scale          # at most a bf16 ulp here and there
    assert (yf - yu).abs().mean().item() < 1e-4 * scale
    ye, yeu = outs[(True, "eval")][0], outs[(False, "eval")][0]
    se = yeu.abs().max().item()
    # folded: one rounding of the normalised value instead of two (conv output, then BN output)
    assert (ye - yeu).abs().max().item() < 2e-2 * se
    assert (ye - yeu).abs().mean().item() < 2e-3 * se


@pytest.mark.parametrize("mid,F,T,B,up", [(64, 32, 21, 2, True), (128, 16, 40, 3, False)])
def test_abf_rank2_mid_stage_vs_stored_z1_path(cuda_dev, mid, F, T, B, up):
    """ABF level whose 1x1 conv has 2 input channels (mask-level decoder map): the kernels that recompute
    z1 = W1 x per row (and return dx, dW1 directly) against the path that stores z1 and runs the conv's own
    forward / data-gradient / weight-gradient launches.  Differences are bf16 roundings of z1 only."""
    import clskd_b200
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(mid + F)
    torch.manual_seed(mid)
    abf = fw.ABF(2, mid, 2, True).to(cuda_dev).train()
    abf.conv1[1].weight.data.uniform_(0.5, 1.5)
    abf.conv1[1].bias.data.normal_(0, 0.2)
    abf.att_conv[0].bias.data.normal_(0, 0.2)
    Fy = F // 2 if up else F
    x0 = torch.randn(B, 2, F, T, generator=g)
    y0 = torch.randn(B, mid, Fy, T, generator=g)
    gw = torch.randn(B, mid, F, T, generator=g).to(cuda_dev)
    clskd_b200.set_precision("bf16")
    res = {}
    for rank2 in (True, False):
        ops.policy.abf_rank2 = rank2
        for p in abf.parameters():
            p.grad = None
        rm = abf.conv1[1].running_mean.clone()
        x = x0.to(cuda_dev).bfloat16().requires_grad_(True)
        y = y0.to(cuda_dev).bfloat16().requires_grad_(True)
        out, fused = abf(x, y, F, None, "decoder")
        (fused.float() * gw).sum().backward()
        res[rank2] = dict(fused=fused.detach().float().cpu(), dx=x.grad.float().cpu(), dy=y.grad.float().cpu(),
                          dw1=abf.conv1[0].weight.grad.float().cpu().reshape(-1),
                          dgamma=abf.conv1[1].weight.grad.float().cpu(), dbeta=abf.conv1[1].bias.grad.float().cpu(),
                          dwatt=abf.att_conv[0].weight.grad.float().cpu().reshape(-1),
                          rmean=abf.conv1[1].running_mean.clone().cpu())
        abf.conv1[1].running_mean.copy_(rm)
    ops.policy.abf_rank2 = True
    a, b = res[True], res[False]
    for k in ("fused", "dy", "dx"):
        s = b[k].abs().max().item()
        assert (a[k] - b[k]).abs().max().item() < 4e-2 * s, k
        assert (a[k] - b[k]).abs().mean().item() < 4e-3 * s, k
    for k in ("dw1", "dgamma", "dbeta", "dwatt", "rmean"):
        s = max(b[k].abs().max().item(), 1e-6)
        assert (a[k] - b[k]).abs().max().item() < 2e-2 * s, k


@pytest.mark.parametrize("cout,F,T,B", [(16, 256, 45, 3), (32, 256, 33, 2), (24, 64, 130, 2), (64, 32, 300, 1)])
def test_first_layer_mma_kernel_vs_oracle(cuda_dev, cout, F, T, B):
    """First encoder layer under the bf16 policy (tapconv_c2_mma.cu: fp32 two-channel spectrogram -> bf16 maps on
    mma.sync with split-bf16 inputs AND weights) against the oracle's complex conv in float64: only the bf16 rounding
    of the stored output may differ (2^-9 relative), the contraction itself keeps ~2^-16."""
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200 import tools_for_model as tm
    from oracle import dccrn_oracle as D
    g = torch.Generator().manual_seed(cout + F)
    conv = tm.ComplexConv2d(2, cout, kernel_size=(5, 2), stride=(2, 1), padding=(2, 1))
    conv.real_conv.bias.data.normal_(generator=g)
    conv.imag_conv.bias.data.normal_(generator=g)
    x0 = torch.randn(B, 2, F, T, generator=g) * torch.rand(B, 1, F, 1, generator=g) * 3
    ref = D.complex_conv2d(x0.double(), conv.real_conv.weight.double(), conv.real_conv.bias.double(),
                           conv.imag_conv.weight.double(), conv.imag_conv.bias.double())
    clskd_b200.set_precision("bf16")
    conv = conv.to(cuda_dev)
    xp = ops.to_phys(x0.to(cuda_dev))                      # fp32 [B, T, F, 2] view
    n0 = ops.core_launches
    y = conv.forward_phys(ops.dense(xp), torch.bfloat16)
    assert ops.core_launches == n0 + 1                     # the CUDA-core entry point (routes to the mma.sync kernel)
    out = ops.to_logical(y).float().cpu().double()
    assert out.shape == ref.shape
    s = ref.abs().max().item()
    err = (out - ref).abs()
    assert err.max().item() < 6e-3 * s and err.mean().item() < 1.5e-3 * ref.abs().mean().item()
    # against the same values rounded to bf16: the contraction error itself
    assert (out - ref.float().bfloat16().double()).abs().max().item() < 1e-2 * s
    exact = (out == ref.float().bfloat16().double()).double().mean().item()
    assert exact > 0.97, exact
    # weight / bias gradients (tapconv_wgrad_c2_mma_kernel: x split into bf16 hi + lo, dY is bf16) against float64 autograd
    # on the same bf16-representable output gradient
    gw = torch.randn(ref.shape, generator=g).bfloat16()
    for p in conv.parameters():
        p.grad = None
    n1 = ops.core_launches
    y2 = conv.forward_phys(ops.dense(xp), torch.bfloat16)
    (ops.to_logical(y2).float() * gw.to(cuda_dev).float()).sum().backward()
    assert ops.core_launches >= n1 + 2                     # forward + weight gradient on the CUDA-core entry points
    ws = [conv.real_conv.weight, conv.real_conv.bias, conv.imag_conv.weight, conv.imag_conv.bias]
    wd = [w.detach().double().cpu().requires_grad_(True) for w in ws]
    refo = D.complex_conv2d(x0.double(), wd[0], wd[1], wd[2], wd[3])
    (refo * gw.double()).sum().backward()
    for w, r in zip(ws, wd):
        err = float((w.grad.double().cpu() - r.grad).norm() / r.grad.norm())
        assert err < 1e-4, err


@pytest.mark.parametrize("sd,dd", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16),
                                   (torch.bfloat16, torch.float32), (torch.float32, torch.float32)])
@pytest.mark.parametrize("B,T,F,Cc", [(3, 37, 4, 64), (2, 21, 4, 128), (2, 9, 4, 20), (2, 5, 3, 16)])
def test_lstm_layout_transposes_vs_torch(cuda_dev, sd, dd, B, T, F, Cc):
    """The (channel, frequency) transposes either side of the LSTM (DCCRN.py:178-199) through clskd_strided_copy4d: with
    F = 4 they run on the 4 x 4 block transpose kernel (both directions, every dtype pair); F = 3 keeps the generic kernel."""
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(B * T + Cc)
    # direction 0: physical [B,T,F,2Cc] (one complex part) -> [T,B,Cc,F]
    x = torch.randn(B, T, F, 2 * Cc, generator=g).to(cuda_dev).to(sd)
    for part in range(2):
        out = torch.full((T, B, Cc * F), 7.0, dtype=dd, device=cuda_dev)
        ops.strided_copy_into(x[..., part * Cc:(part + 1) * Cc].permute(1, 0, 3, 2), out.view(T, B, Cc, F))
        ref = x[..., part * Cc:(part + 1) * Cc].permute(1, 0, 3, 2).to(dd)
        assert torch.equal(out.view(T, B, Cc, F), ref)
    # direction 1: LSTM output [T,B,Cc*F] -> physical [B,T,F,2Cc] (one complex part)
    y = torch.randn(T, B, Cc * F, generator=g).to(cuda_dev).to(sd)
    phys = torch.full((B, T, F, 2 * Cc), -3.0, dtype=dd, device=cuda_dev)
    for part in range(2):
        ops.strided_copy_into(y.reshape(T, B, Cc, F).permute(1, 0, 3, 2), phys[..., part * Cc:(part + 1) * Cc])
        assert torch.equal(phys[..., part * Cc:(part + 1) * Cc], y.reshape(T, B, Cc, F).permute(1, 0, 3, 2).to(dd))


@pytest.mark.parametrize("C,M", [(16, 70001), (32, 12345), (64, 4099), (128, 130), (128, 50000), (16, 7)])
def test_colgram_vs_float64(cuda_dev, C, M):
    """clskd_colgram (x^T x and column sums of a bf16 map in one mma.sync pass) against float64 torch on the same bf16
    values; rows with a non-zero mean, M not a multiple of the 128-row tile."""
    from clskd_b200 import _lib
    g = torch.Generator().manual_seed(C + M)
    x = (torch.randn(M, C, generator=g) + torch.randn(1, C, generator=g)).bfloat16().to(cuda_dev)
    out = torch.full((C * C + C,), 7.0, dtype=torch.float64, device=cuda_dev)      # the call zeroes its outputs
    assert _lib.load().clskd_colgram_supported(_lib.BF16, C) == 1
    _lib.call("clskd_colgram", x.data_ptr(), _lib.BF16, M, C, out.data_ptr(), out[C * C:].data_ptr(),
              torch.cuda.current_stream().cuda_stream)
    xd = x.double()
    G, sx = xd.t() @ xd, xd.sum(0)
    assert (out[:C * C].view(C, C) - G).abs().max().item() < 2e-6 * G.abs().max().item()      # fp32 partial sums
    assert (out[C * C:] - sx).abs().max().item() < 2e-6 * max(sx.abs().max().item(), 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,mid,F,T,B,up,mode", [(16, 128, 128, 23, 2, True, "train"), (32, 128, 64, 31, 3, True, "train"),
                                                   (64, 128, 32, 40, 2, False, "train"), (128, 128, 16, 33, 3, True, "train"),
                                                   (16, 64, 32, 21, 2, True, "eval")])
def test_abf_fold_vs_two_pass_backward(cuda_dev, cin, mid, F, T, B, up, mode):
    """ops.AbfFoldFn (one-pass middle-stage backward, BatchNorm backward folded into conv1's two-source data gradient
    and into dW1 through x^T dxp, x^T x and sum x) against the two-pass AbfMidFn path + conv1's own gradient launches.
    x has a non-zero mean per channel (post-PReLU maps do), which is what the cancellation in the folded dW1 sees.
    Both paths round dz1-sized intermediates to bf16 once; differences are those roundings."""
    import clskd_b200
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(mid + F + cin)
    torch.manual_seed(mid + cin)
    abf = fw.ABF(cin, mid, 32, True).to(cuda_dev)
    abf = abf.train() if mode == "train" else abf.eval()
    abf.conv1[1].weight.data.uniform_(0.5, 1.5)
    abf.conv1[1].bias.data.normal_(0, 0.2)
    abf.conv1[1].running_mean.normal_(0, 0.3)
    abf.conv1[1].running_var.uniform_(0.5, 1.5)
    abf.att_conv[0].bias.data.normal_(0, 0.2)
    Fy = F // 2 if up else F
    x0 = torch.randn(B, cin, F, T, generator=g) + torch.randn(1, cin, 1, 1, generator=g) * 1.5
    y0 = torch.randn(B, mid, Fy, T, generator=g)
    gw = torch.randn(B, mid, F, T, generator=g).to(cuda_dev)
    clskd_b200.set_precision("bf16")
    res = {}
    calls = {}
    for fold in (True, False):
        ops.policy.abf_fold = fold
        for p in abf.parameters():
            p.grad = None
        rm = abf.conv1[1].running_mean.clone()
        x = x0.to(cuda_dev).bfloat16().requires_grad_(True)
        y = y0.to(cuda_dev).bfloat16().requires_grad_(True)
        out, fused = abf(x, y, F, None, "decoder")
        names, frontier = set(), [fused.grad_fn]
        for _ in range(4):                       # autograd nodes within four hops of the fused map
            nxt = []
            for fn in frontier:
                if fn is not None:
                    names.add(type(fn).__name__)
                    nxt += [f for f, _ in fn.next_functions]
            frontier = nxt
        calls[fold] = " ".join(sorted(names))
        (fused.float() * gw).sum().backward()
        res[fold] = dict(fused=fused.detach().float().cpu(), dx=x.grad.float().cpu(), dy=y.grad.float().cpu(),
                         dw1=abf.conv1[0].weight.grad.float().cpu().reshape(-1),
                         dgamma=abf.conv1[1].weight.grad.float().cpu(), dbeta=abf.conv1[1].bias.grad.float().cpu(),
                         dwatt=abf.att_conv[0].weight.grad.float().cpu().reshape(-1),
                         rmean=abf.conv1[1].running_mean.clone().cpu())
        abf.conv1[1].running_mean.copy_(rm)
    ops.policy.abf_fold = True
    assert "AbfFoldFn" in calls[True] and "AbfFoldFn" not in calls[False], calls
    a, b = res[True], res[False]
    rl2 = lambda u, v: float((u.double() - v.double()).norm() / max(float(v.double().norm()), 1e-30))
    # forward: the same launches except for the batch statistics of z1 - from the moments of x (clskd_colgram ->
    # clskd_abf_fold_stats: statistics of the un-rounded W1 x) instead of the conv epilogue (of the stored bf16 z1)
    errs = {k: rl2(a[k], b[k]) for k in ("fused", "rmean", "dx", "dy", "dw1", "dgamma", "dbeta", "dwatt")}
    assert errs["fused"] < 2e-3 and errs["rmean"] < 2e-4, errs
    assert errs["dy"] < 2e-3 and errs["dwatt"] < 2e-3 and errs["dgamma"] < 2e-3 and errs["dbeta"] < 2e-3, errs
    assert errs["dx"] < 1e-2 and errs["dw1"] < 1e-2, errs


@pytest.mark.parametrize("cin,cout,F,T,B,ks", [(32, 64, 128, 9, 2, 3), (128, 32, 256, 5, 1, 3), (16, 128, 128, 7, 2, 1),
                                              (64, 48, 128, 6, 1, 3), (64, 128, 32, 21, 2, 3), (128, 256, 8, 70, 2, 3)])
@pytest.mark.parametrize("tune", [(0, 0), (1, 1), (2, 1), (3, 1), (3, 0)])
def test_umma_operand_reuse_modes_vs_torch(cuda_dev, cin, cout, F, T, B, ks, tune):
    """halo-patch / time-grouped / per-tap operand loading and resident weights of the tcgen05 forward kernel
    (clskd_set_tuning) against torch's conv2d on bf16-representable operands, fp32 output (fp32 accumulation only)"""
    import clskd_b200
    from clskd_b200 import _lib
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(cin + F + ks)
    conv = fw.RealConv2d(cin, cout, ks, padding=ks // 2, bias=True)
    _round_params(conv)
    x = torch.randn(B, cin, F, T, generator=g).bfloat16().float()
    ref = torch.nn.functional.conv2d(x, conv.weight, conv.bias, padding=ks // 2)
    conv = conv.to(cuda_dev)
    xp = x.permute(0, 3, 2, 1).contiguous().to(cuda_dev).bfloat16()
    lib = _lib.load()
    try:
        lib.clskd_set_tuning(0, tune[0])
        lib.clskd_set_tuning(1, tune[1])
        with torch.no_grad():
            clskd_b200.set_precision("bf16")
            n0 = ops.umma_launches
            y = conv.forward_phys(xp, out_dtype=torch.float32)
            assert ops.umma_launches == n0 + 1
    finally:
        lib.clskd_set_tuning(0, 0)
        lib.clskd_set_tuning(1, 0)
    y = y.permute(0, 3, 2, 1).cpu()
    assert (y - ref).abs().max().item() < 2e-5 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize("C,F,T,B,up,training", [(128, 32, 21, 2, True, True), (64, 16, 40, 3, False, True), (128, 64, 9, 1, True, False),
                                                 (256, 8, 30, 2, True, True)])
def test_abf_xs2_kernels_vs_round1_xs_kernels_fp32(cuda_dev, C, F, T, B, up, training):
    """rank-2 folded 2-channel ABF middle stage (clskd_abf_xs2_fwd / _bwd: one backward pass + finalisation + dx pass)
    against the round-1 kernels that recompute z1 per element in two passes, on fp32 tensors (same math, different
    association): outputs and every gradient to fp32 rounding."""
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(C + F)
    Fy = F // 2 if up else F
    xs = torch.randn(B, T, F, 2, generator=g)
    yp = torch.randn(B, T, Fy, C, generator=g)
    w1 = 0.5 * torch.randn(C, 2, 1, 1, generator=g)
    gamma, beta = 0.5 + torch.rand(C, generator=g), 0.2 * torch.randn(C, generator=g)
    watt, batt = 0.1 * torch.randn(2, 2 * C, 1, 1, generator=g), 0.2 * torch.randn(2, generator=g)
    rm, rv = 0.1 * torch.randn(C, generator=g), 0.5 + torch.rand(C, generator=g)
    up_g = torch.randn(B, T, F, C, generator=g)
    res = {}
    for xs2 in (True, False):
        ops.policy.abf_xs2 = xs2
        leaves = [t.clone().to(cuda_dev).requires_grad_(True) for t in (xs, yp, w1, gamma, beta, watt, batt)]
        out = ops.AbfMidXsFn.apply(leaves[0], leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], leaves[6],
                                   rm.clone().to(cuda_dev), rv.clone().to(cuda_dev), training, 0.1, 1e-5)
        (out * up_g.to(cuda_dev)).sum().backward()
        res[xs2] = [out.detach().cpu()] + [t.grad.detach().cpu() for t in leaves]
    ops.policy.abf_xs2 = True
    names = ["xb", "dx", "dy", "dw1", "dgamma", "dbeta", "dwatt", "dbatt"]
    for n, a, b in zip(names, res[True], res[False]):
        s = max(b.abs().max().item(), 1e-6)
        assert (a - b).abs().max().item() < 2e-4 * s, n


def _lstm_reference(pre, whh_t, h0, c0):
    """float64 recurrence on pre [nsets, P, T, Bp, 4H] with W_hh^T [nsets, H, 4H] (tools_for_model.py:138-178's
    nn.LSTM cell: gates i, f, g, o)."""
    nsets, P, T, Bp, G = pre.shape
    H = G // 4
    h = torch.zeros(nsets, P, T, Bp, H, dtype=torch.float64)
    gates = torch.zeros(nsets, P, T, Bp, G, dtype=torch.float64)
    c = torch.zeros(nsets, P, T, Bp, H, dtype=torch.float64)
    hh, cc = h0.double().clone(), c0.double().clone()          # [nsets, P, Bp, H]
    W = whh_t.double()
    for t in range(T):
        g_ = pre[:, :, t].double() + torch.einsum("spbk,skg->spbg", hh, W)
        i_, f_, gg, o_ = (torch.sigmoid(g_[..., :H]), torch.sigmoid(g_[..., H:2 * H]), torch.tanh(g_[..., 2 * H:3 * H]),
                          torch.sigmoid(g_[..., 3 * H:]))
        cc = f_ * cc + i_ * gg
        hh = o_ * torch.tanh(cc)
        h[:, :, t], c[:, :, t] = hh, cc
        gates[:, :, t] = torch.cat([i_, f_, gg, o_], -1)
    return h, gates, c, hh, cc


@pytest.mark.gpu
@pytest.mark.parametrize("H,P,Bp,T,nsets,state", [(128, 2, 12, 40, 2, False), (64, 2, 5, 33, 2, True),
                                                  (32, 1, 7, 25, 1, True), (64, 2, 300, 6, 2, False),
                                                  (128, 2, 301, 5, 2, True)])
def test_lstm_tensor_core_recurrence_vs_float64_and_cuda_core_kernel(cuda_dev, H, P, Bp, T, nsets, state):
    """bf16 policy: the mma.sync recurrence (W_hh fragments in registers, h as a bf16 hi+lo pair) against a float64
    recurrence with the same bf16-rounded W_hh, and against the CUDA-core kernel (tuning key 7) on the same operands.
    Ragged row counts (rows not a multiple of 8), carried state, and the 16-rows-per-CTA variant (large R)."""
    from clskd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(H + Bp + T)
    G, R = 4 * H, P * Bp
    pre = torch.randn(P, T, Bp, nsets * G, generator=g)
    whh_t = (torch.randn(nsets, H, G, generator=g) / H ** 0.5)
    h0 = 0.5 * torch.randn(nsets, R, H, generator=g) if state else torch.zeros(nsets, R, H)
    c0 = 0.5 * torch.randn(nsets, R, H, generator=g) if state else torch.zeros(nsets, R, H)
    pre_sets = torch.stack([pre[..., s * G:(s + 1) * G] for s in range(nsets)])        # [nsets, P, T, Bp, G]
    ref = _lstm_reference(pre_sets, whh_t.bfloat16().float(), h0.view(nsets, P, Bp, H), c0.view(nsets, P, Bp, H))
    st = torch.cuda.current_stream().cuda_stream
    pre_d, w_d = pre.to(cuda_dev), whh_t.to(cuda_dev).contiguous()
    outs = {}
    for legacy in (0, 1):
        lib.clskd_set_tuning(7, legacy)
        try:
            h = torch.full((nsets, P, T, Bp, H), float("nan"), device=cuda_dev)
            gates = torch.full((nsets, P, T, Bp, G), float("nan"), device=cuda_dev)
            c = torch.full((nsets, P, T, Bp, H), float("nan"), device=cuda_dev)
            hs, cs = h0.to(cuda_dev).clone(), c0.to(cuda_dev).clone()
            _lib.call("clskd_lstm_fwd_state", pre_d.data_ptr(), w_d.data_ptr(), T, R, Bp, H, nsets,
                      T * Bp * nsets * G, Bp * nsets * G, nsets * G, G, H * G, 1, h.data_ptr(), gates.data_ptr(),
                      c.data_ptr(), hs.data_ptr() if state else None, cs.data_ptr() if state else None,
                      hs.data_ptr(), cs.data_ptr(), st)
            torch.cuda.synchronize()
        finally:
            lib.clskd_set_tuning(7, 0)
        outs[legacy] = [x.cpu().double() for x in (h, gates, c, hs.view(nsets, P, Bp, H), cs.view(nsets, P, Bp, H))]
    for name, a, b, r in zip(("h", "gates", "c", "hN", "cN"), outs[0], outs[1], ref):
        assert torch.isfinite(a).all(), name
        assert (a - r).abs().max().item() < 2e-5 * max(1.0, r.abs().max().item()), (name, "vs float64")
        assert (a - b).abs().max().item() < 2e-5 * max(1.0, r.abs().max().item()), (name, "vs CUDA-core kernel")


@pytest.mark.gpu
@pytest.mark.parametrize("H,P,Bp,T,nsets", [(128, 2, 12, 30, 2), (64, 2, 5, 33, 2), (32, 1, 7, 25, 1),
                                            (64, 2, 300, 6, 2), (128, 2, 301, 4, 2)])
def test_lstm_tensor_core_bptt_vs_float64_and_cuda_core_kernel(cuda_dev, H, P, Bp, T, nsets):
    """bf16 policy BPTT on mma.sync (W_hh^T fragments in registers, gate gradients as a bf16 hi+lo pair) against a
    float64 BPTT with the same bf16-rounded W_hh, and against the fp32 CUDA-core kernel (fp32 W_hh: the difference is
    the weight rounding, bounded by 2^-8 of the gradient scale)."""
    from clskd_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3 * H + Bp + T)
    G, R = 4 * H, P * Bp
    pre = torch.randn(P, T, Bp, nsets * G, generator=g)
    whh = torch.randn(nsets, G, H, generator=g) / H ** 0.5                 # W_hh [4H][H]
    pre_sets = torch.stack([pre[..., s * G:(s + 1) * G] for s in range(nsets)])
    z = torch.zeros(nsets, P, Bp, H)
    h, gates, c, _, _ = _lstm_reference(pre_sets, whh.bfloat16().float().transpose(1, 2).contiguous(), z, z)
    dh = torch.randn(nsets, P, T, Bp, H, generator=g)
    # float64 BPTT (cabi_emu.clskd_lstm_bwd's recurrence) with the bf16-rounded weight
    Wb = whh.bfloat16().double()
    ref = torch.zeros(nsets, P, T, Bp, G, dtype=torch.float64)
    dh_rec = torch.zeros(nsets, P, Bp, H, dtype=torch.float64)
    dc_next = torch.zeros_like(dh_rec)
    for t in range(T - 1, -1, -1):
        gi, gf, gg, go = (gates[:, :, t][..., k * H:(k + 1) * H] for k in range(4))
        ct = c[:, :, t]
        cprev = c[:, :, t - 1] if t > 0 else torch.zeros_like(ct)
        d = dh[:, :, t].double() + dh_rec
        tc = torch.tanh(ct)
        do = d * tc * go * (1 - go)
        dc = d * go * (1 - tc * tc) + dc_next
        di, df, dg = dc * gg * gi * (1 - gi), dc * cprev * gf * (1 - gf), dc * gi * (1 - gg * gg)
        dc_next = dc * gf
        dp = torch.cat([di, df, dg, do], -1)
        ref[:, :, t] = dp
        dh_rec = torch.einsum("spbg,sgk->spbk", dp, Wb)
    st = torch.cuda.current_stream().cuda_stream
    dev = cuda_dev
    args = [x.float().contiguous().to(dev) for x in (dh, whh, gates, c)]
    outs = {}
    for w_bf16 in (1, 0):
        dpre = torch.full((P, T, Bp, nsets * G), float("nan"), device=dev)
        _lib.call("clskd_lstm_bwd_policy", args[0].data_ptr(), args[1].data_ptr(), args[2].data_ptr(), args[3].data_ptr(),
                  T, R, Bp, H, nsets, G * H, T * Bp * nsets * G, Bp * nsets * G, nsets * G, G, dpre.data_ptr(), w_bf16, st)
        torch.cuda.synchronize()
        outs[w_bf16] = torch.stack([dpre[..., s * G:(s + 1) * G] for s in range(nsets)]).cpu().double()
    scale = ref.abs().max().item()
    assert torch.isfinite(outs[1]).all()
    assert (outs[1] - ref).abs().max().item() < 3e-5 * scale
    assert (outs[1] - outs[0]).abs().max().item() < 2e-2 * scale


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,F,T,B", [(128, 32, 128, 70, 2), (64, 64, 32, 33, 3), (256, 128, 16, 600, 8)])
def test_conv_with_residual_alias_accumulates_data_gradient(cuda_dev, cin, cout, F, T, B):
    """RealConv2d.forward_phys_res (ABF conv2 + residual of the next level, framework.py:221-224): the convolution's
    data gradient is added to the residual's gradient by the tcgen05 epilogue (TMA reduce-add) - against the plain
    fan-out (separate data gradient + clskd_sum_n) on the same operands, and against torch autograd in fp32."""
    import clskd_b200
    from clskd_b200 import framework as fw
    from clskd_b200 import ops
    g = torch.Generator().manual_seed(cin + F + T)
    conv = fw.RealConv2d(cin, cout, 3, padding=1, bias=False)
    _round_params(conv)
    x = (0.5 * torch.randn(B, T, F, cin, generator=g)).bfloat16()
    gy = torch.randn(B, T, F, cout, generator=g).bfloat16()
    gres = torch.randn(B, T, F, cin, generator=g).bfloat16()
    # fp32 reference
    xr = x.float().permute(0, 3, 2, 1).contiguous().requires_grad_(True)
    yr = torch.nn.functional.conv2d(xr, conv.weight.detach(), None, padding=1)
    (yr * gy.float().permute(0, 3, 2, 1)).sum().backward()
    ref_dx = xr.grad.permute(0, 3, 2, 1) + gres.float()
    conv = conv.to(cuda_dev)
    clskd_b200.set_precision("bf16")
    outs = []
    for mode in ("res", "fanout"):
        xd = x.to(cuda_dev).requires_grad_(True)
        conv.weight.grad = None
        n0 = ops.umma_launches
        if mode == "res":
            y, res = conv.forward_phys_res(xd)
        else:
            x0, res = ops.fanout(xd, 2)
            y = conv.forward_phys(x0)
        torch.autograd.backward([y, res], [gy.to(cuda_dev), gres.to(cuda_dev).clone()])
        assert ops.umma_launches > n0
        outs.append((xd.grad.float().cpu(), conv.weight.grad.float().cpu(), y.detach().float().cpu()))
    s = ref_dx.abs().max().item()
    assert (outs[0][2] - outs[1][2]).abs().max().item() == 0.0
    assert (outs[0][1] - outs[1][1]).abs().max().item() <= 1e-6 * outs[1][1].abs().max().item()
    # both round the sum to bf16 once or twice: within two bf16 ulps of the fp32 result
    assert (outs[0][0] - ref_dx).abs().max().item() < 2.0 ** -7 * s
    assert (outs[1][0] - ref_dx).abs().max().item() < 2.0 ** -7 * s
    assert (outs[0][0] - outs[1][0]).abs().max().item() < 2.0 ** -7 * s


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["clskd", "spkd_all"])
def test_batched_repack_after_optimizer_step_equals_lazy_pack(cuda_dev, mode):
    """FlatAdam rewrites the parameters behind autograd's back and then rebuilds every packed kernel-side weight of
    the step in one launch (ops.repack_registered -> clskd_multi_pack_gather).  After two training steps every
    registered packed weight must be bit-identical to a fresh clskd_pack_gather of the current parameters, the cache
    must serve it without a re-pack, and the losses must follow a trainer that packs every weight lazily."""
    import copy
    import clskd_b200
    from clskd_b200 import ops
    from clskd_b200.distill import DistillTrainer
    from oracle import dccrn_oracle as D
    cfg_t = dict(kernel_num=[32, 64, 64, 64, 64, 64], rnn_units=64)
    cfg_s = dict(kernel_num=[16, 32, 32, 32, 32, 32], rnn_units=32)
    t_sd = D.make_state_dict(cfg_t["kernel_num"], cfg_t["rnn_units"], seed=1)
    s_sd = D.make_state_dict(cfg_s["kernel_num"], cfg_s["rnn_units"], seed=2)

    def mk(cfg, sd):
        m = clskd_b200.DCCRN(rnn_units=cfg["rnn_units"], masking_mode="E", use_clstm=True, kernel_num=cfg["kernel_num"])
        m.load_state_dict(sd)
        return m.to(cuda_dev)
    g = torch.Generator().manual_seed(0)
    X = (0.1 * torch.randn(4, 8000, generator=g)).to(cuda_dev)
    y = (0.1 * torch.randn(4, 8000, generator=g)).to(cuda_dev)
    clskd_b200.set_precision("bf16")
    losses = {}
    for lazy in (False, True):
        torch.manual_seed(11)                        # the ABF blocks are created (randomly initialised) by the trainer
        tr = DistillTrainer(mk(cfg_t, t_sd), mk(cfg_s, s_sd), mode=mode, lr=1e-3, example_input=X)
        out = []
        for it in range(3):
            if lazy:
                ops._pack_registry.clear()           # every weight is packed lazily at its first use
            out.append(float(tr.train_step(X, y)))
            if not lazy and it == 1:
                torch.cuda.synchronize()
                live = [e for e in ops._pack_registry.values() if e.a() is not None]
                assert len(live) > 20
                for e in live:
                    a, b = e.a(), (e.b() if e.b is not None else None)
                    key, w = e.cache[e.ckey]
                    assert w is e.out and key == ops._wkey(a, b)
                    fresh = ops.pack_weights(e.table, a, b, e.dtype)
                    assert torch.equal(fresh.view(torch.int16 if e.dtype == torch.bfloat16 else torch.int32),
                                       w.view(torch.int16 if e.dtype == torch.bfloat16 else torch.int32))
        losses[lazy] = out
    # (atomic accumulation orders differ run to run: the two trainers agree to rounding, not bitwise - and Adam's first
    # updates are sign-like, so that rounding noise grows step by step: the bitwise check above is the test of the re-pack,
    # this one only catches a trainer that runs on stale weights)
    assert abs(losses[False][0] - losses[True][0]) < 1e-5 * abs(losses[True][0]), losses
    for la, lb in zip(losses[False], losses[True]):
        assert abs(la - lb) < 3e-3 * abs(lb), losses


def _shift_tf(x, dt, df):
    """x[b, t + dt, f + df, :] with zeros outside"""
    B, T, F, C = x.shape
    out = torch.zeros_like(x)
    t0, t1 = max(0, -dt), min(T, T - dt)
    f0, f1 = max(0, -df), min(F, F - df)
    if t1 > t0 and f1 > f0:
        out[:, t0:t1, f0:f1] = x[:, t0 + dt:t1 + dt, f0 + df:f1 + df]
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("c0,c1,N,F,T,B,dts,dfs", [
    (128, 0, 32, 128, 70, 8, (-1, 0, 1), (-1, 0, 1)), (128, 0, 64, 64, 131, 8, (-1, 0, 1), (-1, 0, 1)),
    (64, 0, 16, 32, 261, 8, (-1, 0, 1), (-1, 0, 1)), (192, 0, 24, 128, 35, 16, (-1, 0, 1), (-1, 0, 1)),
    (128, 0, 32, 256, 33, 8, (-1, 0, 1), (-1, 0, 1)),
    (32, 32, 16, 64, 130, 8, (-1, 0), (-1, 0, 1)), (64, 64, 32, 32, 261, 8, (-1, 0), (0, 1)),
    (128, 128, 64, 16, 520, 8, (-1, 0), (-1, 0, 1)), (16, 16, 32, 128, 66, 8, (-1, 0), (-1, 0)), (24, 8, 16, 64, 130, 8, (0, 1), (-1, 0, 1)),
    (16, 0, 32, -64, 130, 8, (-1, 0), (-2, -1, 0, 1, 2)), (32, 0, 64, -32, 261, 8, (-1, 0), (-2, -1, 0, 1, 2)),
    (64, 0, 16, -128, 66, 8, (-1, 0), (-2, -1, 0, 1, 2)), (8, 0, 24, -64, 131, 8, (-1, 0), (-2, -1, 0, 1, 2)),
    (32, 32, 16, 64, 1130, 8, (-1, 0), (-1, 0, 1)), (64, 64, 32, 32, 1261, 8, (-1, 0), (0, 1))])
def test_tap_stacked_weight_gradient_vs_torch(cuda_dev, c0, c1, N, F, T, B, dts, dfs):
    """clskd_tapconv_wgrad_umma_stacked (frequency taps as sub-blocks of the MMA's N dimension, one patch row apart;
    one or two sources, 64 / 32 / 16-channel swizzle groups) against shifted einsums on the same bf16-representable
    operands; also checks that clskd_tapconv_wgrad_umma routes such a launch there and that tuning key 6 = 5 gives the
    per-tap kernel's result."""
    import ctypes
    from clskd_b200 import _lib
    lib = _lib.load()
    sf = 1
    if F < 0:                  # negative F: stride 2 along f (the encoder layers), -F output frequencies
        sf, F = 2, -F
    Fi = F * sf
    To = T
    if T >= 1000:              # T + 1000: one more output row than input rows (decoder sub-pixel phases)
        T -= 1000
        To = T + 1
    g = torch.Generator().manual_seed(c0 + N + F)
    taps = [(dt, df) for df in dfs for dt in dts]
    C = c0 + c1
    x = (0.5 * torch.randn(B, T, Fi, C, generator=g)).bfloat16()
    dy = (0.5 * torch.randn(B, To, F, N, generator=g)).bfloat16()
    xd_, dyd_ = x.double(), dy.double()
    if To != T:                # zero rows appended to x so that both have To rows
        xd_ = torch.cat([xd_, torch.zeros(B, To - T, Fi, C, dtype=torch.float64)], 1)
    ref = torch.stack([torch.einsum("btfc,btfn->cn", _shift_tf(xd_, dt, df)[:, :, ::sf], dyd_) for dt, df in taps])     # [taps][C][N]
    x0d = x[..., :c0].contiguous().to(cuda_dev)
    x1d = x[..., c0:].contiguous().to(cuda_dev) if c1 else None
    dyd = dy.to(cuda_dev)
    st = torch.cuda.current_stream().cuda_stream
    d = _lib.TapConv()
    d.x0 = x0d.data_ptr()
    d.x0_sB, d.x0_sT, d.x0_sF = T * Fi * c0, Fi * c0, c0
    d.x1 = x1d.data_ptr() if c1 else None
    if c1:
        d.x1_sB, d.x1_sT, d.x1_sF = T * Fi * c1, Fi * c1, c1
    d.c0, d.c1 = c0, c1
    d.B, d.To, d.Fo, d.Ti, d.Fi = B, To, F, T, Fi
    d.sf, d.ntaps = sf, len(taps)
    for j, (dt, df) in enumerate(taps):
        d.dt[j], d.df[j] = dt, df
    d.bias, d.N = None, N
    d.y = dyd.data_ptr()
    d.y_sB, d.y_sT, d.y_sF = To * F * N, F * N, N
    d.x_dtype, d.y_dtype, d.accumulate = _lib.BF16, _lib.BF16, 0
    assert lib.clskd_tapconv_wgrad_umma_stacked_supported(ctypes.byref(d)) == 1
    outs = {}
    for name, entry, mode in (("stacked", "clskd_tapconv_wgrad_umma_stacked", 0), ("routed", "clskd_tapconv_wgrad_umma", 0),
                              ("per_tap", "clskd_tapconv_wgrad_umma", 5)):
        dw = torch.full((len(taps), C, N), float("nan"), dtype=torch.float32, device=cuda_dev)
        d.w = dw.data_ptr()
        lib.clskd_set_tuning(6, mode)
        try:
            _lib.call(entry, ctypes.byref(d), st)
            torch.cuda.synchronize()
        finally:
            lib.clskd_set_tuning(6, 0)
        outs[name] = dw.cpu().double()
    scale = ref.abs().max().item()
    for name, o in outs.items():
        assert torch.isfinite(o).all(), name
        assert (o - ref).abs().max().item() < 2e-5 * scale, name          # fp32 accumulation of exact bf16 products
