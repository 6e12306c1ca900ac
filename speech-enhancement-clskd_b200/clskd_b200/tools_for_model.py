"""B200-native building blocks with the reference's names, constructor signatures and
parameter/buffer names (reference: tools_for_model.py:15-524), so `state_dict`s interchange and
`feature_extraction` hooks keep working.  Every forward runs hand-written sm_100a kernels through
the C ABI (see ops.py); nothing here falls back to torch's conv / RNN / FFT library calls.

Tensors follow the reference's logical shapes ([B, C, F, T], [T, B, D], [B, 514, T] ...); internally
activations are channels-last [B, T, F, C] and the logical tensors are permuted views of them.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from scipy.signal import get_window

from . import ops
from .ops import (BNActFn, ConvPlan, FramedGemmFn, OlaFn, Pad1dFn, TapConvFn, _codes, _neg, call,
                  dense, to_logical, to_phys)


# ---------------------------------------------------------------------------------------------
# STFT / iSTFT filterbanks (reference: tools_for_model.py:15-109)
# ---------------------------------------------------------------------------------------------

def init_kernels(win_len, win_inc, fft_len, win_type=None, invers=False):
    """Windowed DFT analysis basis [fft_len+2, 1, win_len] (cos rows, then -sin rows) or, with
    invers=True, the pseudo-inverse synthesis basis; plus the window [1, win_len, 1]."""
    if win_type is None or win_type == 'None':
        window = np.ones(win_len)
    else:
        window = get_window(win_type, win_len, fftbins=True)
    n = np.arange(win_len, dtype=np.float64)[None, :]
    k = np.arange(fft_len // 2 + 1, dtype=np.float64)[:, None]
    ang = 2.0 * np.pi * k * n / fft_len
    basis = np.concatenate([np.cos(ang), -np.sin(ang)], axis=0)      # [fft_len+2, win_len]
    if invers:
        basis = np.linalg.pinv(basis).T
    basis = basis * window
    return (torch.from_numpy(np.ascontiguousarray(basis[:, None, :], dtype=np.float32)),
            torch.from_numpy(np.ascontiguousarray(window[None, :, None], dtype=np.float32)))


def _default_fft_len(win_len):
    return int(2 ** math.ceil(math.log2(win_len)))


class ConvSTFT(nn.Module):
    """STFT as a framed-window DFT GEMM.  forward: [B, L] or [B, 1, L] -> [B, fft_len+2, T]
    (rows 0..N/2 real, then imag), T = (L + 2*(win_len-win_inc) - win_len)//win_inc + 1."""

    def __init__(self, win_len, win_inc, fft_len=None, win_type='hamming', feature_type='real', fix=True):
        super().__init__()
        self.fft_len = _default_fft_len(win_len) if fft_len is None else fft_len
        kernel, _ = init_kernels(win_len, win_inc, self.fft_len, win_type)
        self.register_buffer('weight', kernel)
        self.feature_type = feature_type
        self.stride = win_inc
        self.win_len = win_len
        self.dim = self.fft_len
        self._plans = {}

    def _plan(self, interleaved):
        key = (interleaved, self.weight.device)
        if key not in self._plans:
            nrow = self.fft_len + 2
            code = _codes((nrow, self.win_len), 0).T                  # [k, n] -> weight[n, 0, k]
            if interleaved:                                           # columns (bin, part)
                nb = nrow // 2
                order = np.stack([np.arange(nb), nb + np.arange(nb)], 1).reshape(-1)
                code = code[:, order]
            self._plans[key] = ConvPlan("conv", code[None, None], 1, 0, 0, self.win_len, 0, None,
                                        self.weight.numel(), 0, self.weight.device)
        return self._plans[key]

    def spectrum(self, inputs, interleaved=True):
        """[B, L] -> fp32 [B, T, nbins, 2] (interleaved) or [B, T, fft_len+2]."""
        if inputs.dim() == 3:
            inputs = inputs.squeeze(1)
        pad = self.win_len - self.stride
        xpad = Pad1dFn.apply(inputs, pad, pad, 0)
        spec = FramedGemmFn.apply(self._plan(interleaved), self.weight.view(-1), xpad, self.win_len, self.stride)
        if interleaved:
            return spec.view(spec.shape[0], spec.shape[1], -1, 2)
        return spec

    def spectrum_padded(self, xpad, interleaved=True):
        """spectrum of an ALREADY padded signal [B, n]: frames start at 0, win_inc, ... (streaming inference cuts
        frame ranges out of one padded utterance instead of padding every chunk)"""
        spec = FramedGemmFn.apply(self._plan(interleaved), self.weight.view(-1), xpad, self.win_len, self.stride)
        if interleaved:
            return spec.view(spec.shape[0], spec.shape[1], -1, 2)
        return spec

    def forward(self, inputs):
        outputs = self.spectrum(inputs, interleaved=False).permute(0, 2, 1)
        if self.feature_type == 'complex':
            return outputs
        dim = self.dim // 2 + 1
        real, imag = outputs[:, :dim, :], outputs[:, dim:, :]
        return torch.sqrt(real ** 2 + imag ** 2), torch.atan2(imag, real)


class ConviSTFT(nn.Module):
    """iSTFT: synthesis-basis GEMM + overlap-add with window^2 normalisation.
    forward: [B, fft_len+2, T] (or mags [B, N/2+1, T] + phase) -> [B, 1, L]."""

    def __init__(self, win_len, win_inc, fft_len=None, win_type='hamming', feature_type='real', fix=True):
        super().__init__()
        self.fft_len = _default_fft_len(win_len) if fft_len is None else fft_len
        kernel, window = init_kernels(win_len, win_inc, self.fft_len, win_type, invers=True)
        self.register_buffer('weight', kernel)
        self.feature_type = feature_type
        self.win_type = win_type
        self.win_len = win_len
        self.stride = win_inc
        self.dim = self.fft_len
        self.register_buffer('window', window)
        self.register_buffer('enframe', torch.eye(win_len)[:, None, :])   # kept for state_dict parity
        self._plans = {}

    def _plan(self, interleaved):
        key = (interleaved, self.weight.device)
        if key not in self._plans:
            nrow = self.fft_len + 2
            code = _codes((nrow, self.win_len), 0)                    # [k=row, n=sample]
            if interleaved:
                nb = nrow // 2
                order = np.stack([np.arange(nb), nb + np.arange(nb)], 1).reshape(-1)
                code = code[order, :]
            self._plans[key] = ConvPlan("conv", code[None, None], 1, 0, 0, nrow, 0, None,
                                        self.weight.numel(), 0, self.weight.device)
        return self._plans[key]

    def synthesize(self, spec_rows, interleaved, clamp=False):
        """spec_rows: fp32 dense [B, T, fft_len+2] -> wav [B, L]"""
        B, T, K = spec_rows.shape
        frames = TapConvFn.apply(self._plan(interleaved), spec_rows.view(B, T, 1, K), None,
                                 self.weight.view(-1), None, None, None, torch.float32)
        return OlaFn.apply(frames.view(B, T, self.win_len), self.window.view(-1), self.stride,
                           self.win_len - self.stride, clamp)

    def forward(self, inputs, phase=None):
        if phase is not None:
            inputs = torch.cat([inputs * torch.cos(phase), inputs * torch.sin(phase)], 1)
        rows = dense(inputs.permute(0, 2, 1), torch.float32)
        return self.synthesize(rows, interleaved=False).unsqueeze(1)


# ---------------------------------------------------------------------------------------------
# complex conv / deconv (reference: tools_for_model.py:193-330)
# ---------------------------------------------------------------------------------------------

class _ConvParams(nn.Module):
    """Parameter holder with nn.Conv2d's names (`weight`, `bias`)."""

    def __init__(self, shape, nbias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(shape))
        self.bias = nn.Parameter(torch.zeros(nbias))
        nn.init.normal_(self.weight, std=0.05)


def _complex_block(code_r, code_i):
    """[KF,KT,Cin,Cout] codes of Wr / Wi -> block [KF,KT,2Cin,2Cout] = [[Wr, Wi], [-Wi, Wr]]
    (rows: real inputs then imag inputs; columns: real outputs then imag outputs)."""
    top = np.concatenate([code_r, code_i], axis=3)
    bot = np.concatenate([_neg(code_i), code_r], axis=3)
    return np.concatenate([top, bot], axis=2)


def _complex_bias_table(cout):
    """bias = [br - bi, br + bi]  (a = real bias, b = imag bias)"""
    j = np.arange(cout)
    t = np.empty((2 * cout, 2), dtype=np.int32)
    t[:cout, 0] = j * 4
    t[:cout, 1] = j * 4 + 1 + 2
    t[cout:, 0] = j * 4
    t[cout:, 1] = j * 4 + 1
    return t


class ComplexConv2d(nn.Module):
    """Complex convolution as ONE 2x-wide real contraction with block weight [[Wr,-Wi],[Wi,Wr]].
    in_channels / out_channels count real+imag.  Input [B, in_channels, F, T] (real half of the
    channels first), output [B, out_channels, F', T']."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 dilation=1, groups=1, causal=True, complex_axis=1):
        super().__init__()
        if dilation != 1 or groups != 1 or complex_axis != 1 or stride[1] != 1:
            raise NotImplementedError("ComplexConv2d: only dilation=1, groups=1, complex_axis=1, "
                                      "time stride 1 are implemented")
        self.in_channels = in_channels // 2
        self.out_channels = out_channels // 2
        self.kernel_size = tuple(kernel_size)
        self.stride = tuple(stride)
        self.padding = tuple(padding)
        self.causal = causal
        self.groups = groups
        self.dilation = dilation
        self.complex_axis = complex_axis
        shape = (self.out_channels, self.in_channels) + self.kernel_size
        self.real_conv = _ConvParams(shape, self.out_channels)
        self.imag_conv = _ConvParams(shape, self.out_channels)
        self._plans = {}

    def plan(self):
        dev = self.real_conv.weight.device
        if dev not in self._plans:
            shape = tuple(self.real_conv.weight.shape)
            cr = _codes(shape, 0).transpose(2, 3, 1, 0)
            ci = _codes(shape, 1).transpose(2, 3, 1, 0)
            n = self.real_conv.weight.numel()
            pt = self.padding[1]
            self._plans[dev] = ConvPlan("conv", _complex_block(cr, ci), self.stride[0], self.padding[0], pt,
                                        2 * self.in_channels, 0, _complex_bias_table(self.out_channels),
                                        n, n, dev, t_extra=pt if (pt != 0 and self.causal) else 2 * pt)
        return self._plans[dev]

    def forward_phys(self, x_phys, out_dtype=None):
        return TapConvFn.apply(self.plan(), x_phys, None, self.real_conv.weight, self.imag_conv.weight,
                               self.real_conv.bias, self.imag_conv.bias, out_dtype or ops.policy.act_dtype)

    def forward(self, inputs):
        x = to_phys(inputs)
        return to_logical(self.forward_phys(x, x.dtype))


class ComplexConvTranspose2d(nn.Module):
    """Complex transposed convolution, sub-pixel decomposed (one dense contraction per output
    frequency phase).  `forward_phys(x0, x1)` consumes the decoder's skip connection as a second
    K segment instead of a materialised complex_cat."""

    def __init__(self, in_channels, out_channels, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 output_padding=(0, 0), causal=False, complex_axis=1, groups=1):
        super().__init__()
        if groups != 1 or complex_axis != 1 or stride[1] != 1:
            raise NotImplementedError("ComplexConvTranspose2d: only groups=1, complex_axis=1, time stride 1")
        self.in_channels = in_channels // 2
        self.out_channels = out_channels // 2
        self.kernel_size = tuple(kernel_size)
        self.stride = tuple(stride)
        self.padding = tuple(padding)
        self.output_padding = tuple(output_padding)
        self.groups = groups
        self.complex_axis = complex_axis
        shape = (self.in_channels, self.out_channels) + self.kernel_size
        self.real_conv = _ConvParams(shape, self.out_channels)
        self.imag_conv = _ConvParams(shape, self.out_channels)
        self._plans = {}

    def plan(self, ca=None):
        """ca: complex channels of source 0 when the input is (x0, x1); None = single input."""
        dev = self.real_conv.weight.device
        key = (dev, ca)
        if key not in self._plans:
            shape = tuple(self.real_conv.weight.shape)
            cin = self.in_channels
            cr = _codes(shape, 0).transpose(2, 3, 0, 1)
            ci = _codes(shape, 1).transpose(2, 3, 0, 1)
            block = _complex_block(cr, ci)                     # rows [real 0..cin), [imag 0..cin)
            c0, c1 = 2 * cin, 0
            if ca is not None:
                # reference cat order [r_a, r_b, i_a, i_b] (complex_cat) -> internal [r_a, i_a, r_b, i_b]
                perm = np.concatenate([np.arange(0, ca), cin + np.arange(0, ca),
                                       np.arange(ca, cin), cin + np.arange(ca, cin)])
                block = block[:, :, perm, :]
                c0, c1 = 2 * ca, 2 * (cin - ca)
            n = self.real_conv.weight.numel()
            self._plans[key] = ConvPlan("deconv", block, self.stride[0], self.padding[0], self.padding[1],
                                        c0, c1, _complex_bias_table(self.out_channels), n, n, dev,
                                        t_extra=self.output_padding[1], f_extra=self.output_padding[0])
        return self._plans[key]

    def _narrow_plan(self, ca):
        """Tap-in-channel decomposition for a layer with ONE complex output channel (the mask-producing
        last decoder layer): a pointwise GEMM onto (tap, real/imag) channels at every input position,
        then the sub-pixel gather-sum of clskd_tapsum_fwd (sf = frequency stride)."""
        dev = self.real_conv.weight.device
        key = (dev, ca, "narrow")
        if key not in self._plans:
            full = self.plan(ca)                               # block codes incl. the skip permutation
            KF, KT = self.kernel_size
            N = 2 * self.out_channels
            Zc = (KF * KT * N + 15) // 16 * 16
            # recover the [KF,KT,Ctot,N] code block of the full plan from its forward launches
            shape = tuple(self.real_conv.weight.shape)
            cin = self.in_channels
            cr = _codes(shape, 0).transpose(2, 3, 0, 1)
            ci = _codes(shape, 1).transpose(2, 3, 0, 1)
            block = _complex_block(cr, ci)
            c0, c1 = 2 * cin, 0
            if ca is not None:
                perm = np.concatenate([np.arange(0, ca), cin + np.arange(0, ca),
                                       np.arange(ca, cin), cin + np.arange(ca, cin)])
                block = block[:, :, perm, :]
                c0, c1 = 2 * ca, 2 * (cin - ca)
            blk = np.full((2 * cin, Zc), -1, dtype=np.int64)
            dts, dfs = [], []
            j = 0
            for kf in range(KF):
                for kt in range(KT):
                    blk[:, j * N:(j + 1) * N] = block[kf, kt]
                    dts.append(self.padding[1] - kt)           # ti = to + pt - kt
                    dfs.append(self.padding[0] - kf)           # fi = (fo + pf - kf) / sf
                    j += 1
            n = self.real_conv.weight.numel()
            pw = ConvPlan("conv", blk[None, None], 1, 0, 0, c0, c1, None, n, n, dev)
            self._plans[key] = (pw, dts, dfs, full)
        return self._plans[key]

    def _use_narrow(self, x0, x1):
        ok = self.out_channels == 1 and self.kernel_size[0] * self.kernel_size[1] <= 16
        if ops.policy.narrow == "always":
            return ok
        return (ok and ops.policy.use_umma and x0.dtype == torch.bfloat16 and x0.shape[-1] % 8 == 0
                and (x1 is None or (x1.dtype == torch.bfloat16 and x1.shape[-1] % 8 == 0)))

    def forward_phys(self, x0, x1=None, out_dtype=None):
        ca = None if x1 is None else x0.shape[-1] // 2
        if self._use_narrow(x0, x1):
            ops.request_epilogue(None)      # the pointwise stage below is not the conv output
            pw, dts, dfs, full = self._narrow_plan(ca)
            z = TapConvFn.apply(pw, x0, x1, self.real_conv.weight, self.imag_conv.weight, None, None, x0.dtype)
            To, Fo = full.out_size(x0.shape[1], x0.shape[2])
            return ops.TapSumFn.apply(z, dts, dfs, 2 * self.out_channels, (To, Fo), self.stride[0],
                                      out_dtype or ops.policy.act_dtype, full, self.real_conv.bias,
                                      self.imag_conv.bias)
        plan = self.plan(ca)
        return TapConvFn.apply(plan, x0, x1, self.real_conv.weight, self.imag_conv.weight,
                               self.real_conv.bias, self.imag_conv.bias, out_dtype or ops.policy.act_dtype)

    def forward(self, inputs):
        if isinstance(inputs, (tuple, list)):
            inputs = torch.cat([inputs[0], inputs[1]], self.complex_axis)
        x = to_phys(inputs)
        return to_logical(self.forward_phys(x, None, x.dtype))


def complex_cat(inputs, axis):
    """[r_a, i_a], [r_b, i_b] -> [r_a, r_b, i_a, i_b] along `axis` (reference: tools_for_model.py:181-190).
    Inside DCCRN the concat is never materialised; this is the API-parity version."""
    parts = [torch.chunk(x, 2, axis) for x in inputs]
    return torch.cat([p[0] for p in parts] + [p[1] for p in parts], axis)


# ---------------------------------------------------------------------------------------------
# BatchNorm / PReLU on channels-last activations (reference: DCCRN.py:80-82)
# ---------------------------------------------------------------------------------------------

class BatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters/buffers) whose forward is the fused statistics + normalise
    kernel pair; `forward_phys(x, slope)` additionally fuses the following PReLU."""

    def forward_phys(self, x_phys, slope=None, pre_stats=None):
        """pre_stats: fp64 [2, C] column sums of x_phys already produced by the conv epilogue"""
        if self.training and self.track_running_stats and self.num_batches_tracked is not None:
            self.num_batches_tracked.add_(1)
        use_running = (not self.training) and self.track_running_stats
        return BNActFn.apply(x_phys, self.weight, self.bias, slope,
                             self.running_mean if self.track_running_stats else None,
                             self.running_var if self.track_running_stats else None,
                             not use_running, self.momentum, self.eps, pre_stats)

    def forward(self, inputs):
        return to_logical(self.forward_phys(to_phys(inputs, need_dense=True)))


class PReLU(nn.PReLU):
    """Single-slope PReLU; standalone forward = identity-statistics BNAct kernel."""

    def forward(self, inputs):
        if self.weight.numel() != 1:
            raise NotImplementedError("PReLU: only num_parameters=1")
        x = to_phys(inputs, need_dense=True) if inputs.dim() == 4 else dense(inputs)
        C = x.shape[-1]
        zero = torch.zeros(C, dtype=torch.float32, device=x.device)
        one = torch.ones(C, dtype=torch.float32, device=x.device)
        y = BNActFn.apply(x, None, None, self.weight, zero, one, False, 0.0, 0.0, None)
        return to_logical(y) if inputs.dim() == 4 else y


class cPReLU(nn.Module):
    def __init__(self, complex_axis=1):
        super().__init__()
        self.r_prelu = PReLU()
        self.i_prelu = PReLU()
        self.complex_axis = complex_axis

    def forward(self, inputs):
        real, imag = torch.chunk(inputs, 2, self.complex_axis)
        return torch.cat([self.r_prelu(real), self.i_prelu(imag)], self.complex_axis)


class _CBNFn(torch.autograd.Function):
    """y = Z (x - mean) + B with Z = W V^{-1/2} per complex channel (tools_for_model.py:398-508) on the
    moments / whitening kernels; backward = batch sums -> closed-form per-channel whitening derivative
    -> one apply pass.  x: physical dense [..., 2*Cc] (real half then imag half)."""

    @staticmethod
    def forward(ctx, x, Wrr, Wri, Wii, Br, Bi, mod, training):
        Cc = mod.num_features
        M = x.numel() // (2 * Cc)
        dev = x.device
        coef = torch.empty(6, Cc, dtype=torch.float32, device=dev)
        s = torch.empty(5, Cc, dtype=torch.float64, device=dev)
        st = ops._stream()
        if training:
            call("clskd_cbn_moments", x.data_ptr(), ops._tag(x.dtype), M, Cc, s.data_ptr(), st)
        rs = mod.track_running_stats
        P = lambda t: t.data_ptr() if t is not None else None
        call("clskd_cbn_finalize", s.data_ptr(), M, Cc, float(mod.eps), float(mod.momentum or 0.0),
             1 if training else 0, P(Wrr), P(Wri), P(Wii),
             P(mod.RMr) if rs else None, P(mod.RMi) if rs else None, P(mod.RVrr) if rs else None,
             P(mod.RVri) if rs else None, P(mod.RVii) if rs else None, coef.data_ptr(), st)
        y = torch.empty_like(x)
        call("clskd_cbn_apply", x.data_ptr(), ops._tag(x.dtype), M, Cc, coef.data_ptr(), P(Br), P(Bi),
             y.data_ptr(), ops._tag(y.dtype), st)
        ctx.mod, ctx.training, ctx.dims = mod, training, (M, Cc)
        ctx.save_for_backward(x, Wrr, Wri, Wii, s)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, Wrr, Wri, Wii, s = ctx.saved_tensors
        mod, training = ctx.mod, ctx.training
        M, Cc = ctx.dims
        dev = x.device
        st = ops._stream()
        dy = dense(dy, x.dtype)
        P = lambda t: t.data_ptr() if t is not None else None
        s6 = torch.empty(6, Cc, dtype=torch.float64, device=dev)
        call("clskd_cbn_bwd_moments", x.data_ptr(), dy.data_ptr(), ops._tag(x.dtype), M, Cc, s6.data_ptr(), st)
        coefb = torch.empty(11, Cc, dtype=torch.float32, device=dev)
        pg = torch.empty(5, Cc, dtype=torch.float32, device=dev)
        aff = Wrr is not None
        call("clskd_cbn_bwd_finalize", s.data_ptr(), s6.data_ptr(), M, Cc, float(mod.eps), 1 if training else 0,
             P(Wrr), P(Wri), P(Wii), P(mod.RMr), P(mod.RMi), P(mod.RVrr), P(mod.RVri), P(mod.RVii),
             coefb.data_ptr(), pg[0].data_ptr() if aff else None, pg[1].data_ptr() if aff else None,
             pg[2].data_ptr() if aff else None, pg[3].data_ptr() if aff else None,
             pg[4].data_ptr() if aff else None, st)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("clskd_cbn_bwd_apply", x.data_ptr(), dy.data_ptr(), ops._tag(x.dtype), M, Cc, coefb.data_ptr(),
                 dx.data_ptr(), st)
        g = [pg[i] if aff else None for i in range(5)]
        return dx, g[0], g[1], g[2], g[3], g[4], None, None


class ComplexBatchNorm(nn.Module):
    """Trabelsi-style complex batch norm (reference: tools_for_model.py:335-512): 2x2 whitening of
    (real, imag) per channel + complex affine, with batch (training) or running statistics; forward
    and backward (through the batch statistics, like the reference) run on the C-ABI kernels."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 complex_axis=1):
        super().__init__()
        self.num_features = num_features // 2
        self.eps, self.momentum, self.affine = eps, momentum, affine
        self.track_running_stats = track_running_stats
        self.complex_axis = complex_axis
        nf = self.num_features
        if affine:
            self.Wrr = nn.Parameter(torch.ones(nf))
            self.Wri = nn.Parameter(torch.empty(nf).uniform_(-.9, .9))
            self.Wii = nn.Parameter(torch.ones(nf))
            self.Br = nn.Parameter(torch.zeros(nf))
            self.Bi = nn.Parameter(torch.zeros(nf))
        else:
            for n in ('Wrr', 'Wri', 'Wii', 'Br', 'Bi'):
                self.register_parameter(n, None)
        if track_running_stats:
            self.register_buffer('RMr', torch.zeros(nf))
            self.register_buffer('RMi', torch.zeros(nf))
            self.register_buffer('RVrr', torch.ones(nf))
            self.register_buffer('RVri', torch.zeros(nf))
            self.register_buffer('RVii', torch.ones(nf))
            self.register_buffer('num_batches_tracked', torch.tensor(0, dtype=torch.long))
        else:
            for n in ('RMr', 'RMi', 'RVrr', 'RVri', 'RVii', 'num_batches_tracked'):
                self.register_parameter(n, None)

    def forward_phys(self, x, slope=None):
        Cc = self.num_features
        training = self.training or not self.track_running_stats
        if self.training and self.track_running_stats:
            self.num_batches_tracked.add_(1)
        dev = x.device
        x = dense(x)
        y = _CBNFn.apply(x, self.Wrr, self.Wri, self.Wii, self.Br, self.Bi, self, training)
        if slope is not None:
            C = 2 * Cc
            zero = torch.zeros(C, dtype=torch.float32, device=dev)
            one = torch.ones(C, dtype=torch.float32, device=dev)
            y = BNActFn.apply(y, None, None, slope, zero, one, False, 0.0, 0.0, None)
        return y

    def forward(self, inputs):
        return to_logical(self.forward_phys(to_phys(inputs, need_dense=True)))


class ConvBNAct(nn.Sequential):
    """Sequential(conv, [norm, PReLU]) with the reference's child indices (`.0`, `.1`, `.2`) whose
    forward runs the fused pipeline: contraction -> BN statistics -> normalise+PReLU in one pass.
    Forward hooks on this module (feature_extraction) see the same output the reference's
    nn.Sequential would produce."""

    _res_holder = None

    def _conv(self, x0, x1):
        conv = self[0]
        if self._res_holder is not None and x1 is None:
            y, res = conv.forward_phys_res(x0)
            self._res_holder.append(res)
            return y
        return conv.forward_phys(x0, x1) if x1 is not None else conv.forward_phys(x0)

    def forward_phys_res(self, x0):
        """forward_phys(x0) plus an alias of x0 for a second consumer (see RealConv2d.forward_phys_res)"""
        if not hasattr(self[0], "forward_phys_res"):
            x0, res = ops.fanout(x0, 2)
            return self.forward_phys(x0), res
        self._res_holder = []
        try:
            out = self.forward_phys(x0)
            res = self._res_holder[0]
        finally:
            self._res_holder = None
        return out, res

    def forward_phys(self, x0, x1=None):
        if len(self) == 1:
            return self._conv(x0, x1)
        bn = self[1]
        slope = self[2].weight if len(self) > 2 else None
        if ops.policy.use_umma and ops.policy.fuse_epilogue and type(bn) is BatchNorm2d and x0.dtype == torch.bfloat16:
            C = bn.num_features
            use_running = (not bn.training) and bn.track_running_stats
            if use_running and not torch.is_grad_enabled():
                # frozen / inference: BatchNorm folded to a per-channel affine and PReLU applied in the conv
                # epilogue - the activation is written once, already normalised
                fold = torch.empty(2, C, dtype=torch.float32, device=x0.device)
                call("clskd_bn_fold", bn.running_mean.data_ptr(), bn.running_var.data_ptr(),
                     ops._ptr(ops._f32c(bn.weight)) if bn.weight is not None else None,
                     ops._ptr(ops._f32c(bn.bias)) if bn.bias is not None else None, C, float(bn.eps),
                     fold[0].data_ptr(), fold[1].data_ptr(), ops._stream())
                ep = ops.Epilogue(scale=fold[0], shift=fold[1],
                                  slope=ops._f32c(slope) if slope is not None else None)
                ops.request_epilogue(ep)
                try:
                    z = self._conv(x0, x1)
                finally:
                    ops.request_epilogue(None)     # never left dangling for an unrelated conv after an exception
                if ep.fused:
                    return z
                return bn.forward_phys(z, slope)
            if not use_running:
                # training: the conv epilogue accumulates the batch statistics of what it stores
                ep = ops.Epilogue(stats=torch.zeros(2, C, dtype=torch.float64, device=x0.device))
                ops.request_epilogue(ep)
                try:
                    z = self._conv(x0, x1)
                finally:
                    ops.request_epilogue(None)
                return bn.forward_phys(z, slope, ep.stats if ep.fused else None)
        z = self._conv(x0, x1)
        return bn.forward_phys(z, slope)

    def forward(self, inputs, skip=None):
        x0 = to_phys(inputs)
        x1 = to_phys(skip) if skip is not None else None
        return to_logical(self.forward_phys(x0, x1))
