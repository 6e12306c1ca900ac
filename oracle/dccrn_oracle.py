"""ORACLE (test infrastructure only) - CPU restatement of the reference's DCCRN forward path.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
may import this package.  It restates, function by function, what the reference's PyTorch modules
compute (citations are file:line relative to the reference tree), as plain functional torch on a
`state_dict` keyed exactly like the reference's local DCCRN, so identical weights can be fed to
the reference, to this oracle and to the CUDA implementation.

Pinning: tests/golden/*.pt were produced by running the UNMODIFIED reference modules
(tests/golden/make_golden.py imports /root/reference through oracle/ref_shim.py); test_oracle.py
checks this restatement against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F
from scipy.signal import get_window


# ------------------------------------------------------------------ tools_for_model.py:15-32
def init_kernels(win_len, win_inc, fft_len, win_type=None, invers=False):
    window = np.ones(win_len) if (win_type is None or win_type == 'None') else \
        get_window(win_type, win_len, fftbins=True)
    fb = np.fft.rfft(np.eye(fft_len))[:win_len]                       # :22
    kernel = np.concatenate([fb.real, fb.imag], 1).T                  # :23-25
    if invers:
        kernel = np.linalg.pinv(kernel).T                             # :28
    kernel = (kernel * window)[:, None, :]                            # :30-31
    return torch.from_numpy(kernel.astype(np.float32)), torch.from_numpy(window[None, :, None].astype(np.float32))


# ------------------------------------------------------------------ tools_for_model.py:53-60
def conv_stft(x, weight, win_len, hop):
    if x.dim() == 2:
        x = x.unsqueeze(1)
    x = F.pad(x, [win_len - hop, win_len - hop])
    return F.conv1d(x, weight, stride=hop)


# ------------------------------------------------------------------ tools_for_model.py:90-109
def conv_istft(spec, weight, window, win_len, hop):
    out = F.conv_transpose1d(spec, weight, stride=hop)
    t = window.repeat(1, 1, spec.size(-1)) ** 2
    enframe = torch.eye(win_len, dtype=spec.dtype)[:, None, :]
    coff = F.conv_transpose1d(t, enframe, stride=hop)
    out = out / (coff + 1e-8)
    return out[..., win_len - hop:-(win_len - hop)]


# ------------------------------------------------------------------ tools_for_model.py:236-262
def complex_conv2d(x, wr, br, wi, bi, stride=(2, 1), padding=(2, 1), causal=True):
    if padding[1] != 0 and causal:
        x = F.pad(x, [padding[1], 0, 0, 0])
    else:
        x = F.pad(x, [padding[1], padding[1], 0, 0])
    real, imag = torch.chunk(x, 2, 1)
    pad = [padding[0], 0]
    r2r = F.conv2d(real, wr, br, stride, pad)
    i2i = F.conv2d(imag, wi, bi, stride, pad)
    r2i = F.conv2d(real, wi, bi, stride, pad)
    i2r = F.conv2d(imag, wr, br, stride, pad)
    return torch.cat([r2r - i2i, r2i + i2r], 1)


# ------------------------------------------------------------------ tools_for_model.py:303-330
def complex_deconv2d(x, wr, br, wi, bi, stride=(2, 1), padding=(2, 0), output_padding=(1, 0)):
    real, imag = torch.chunk(x, 2, 1)
    ct = lambda t, w, b: F.conv_transpose2d(t, w, b, stride, padding, output_padding)
    return torch.cat([ct(real, wr, br) - ct(imag, wi, bi), ct(real, wi, bi) + ct(imag, wr, br)], 1)


# ------------------------------------------------------------------ tools_for_model.py:181-190
def complex_cat(inputs, axis=1):
    parts = [torch.chunk(t, 2, axis) for t in inputs]
    return torch.cat([p[0] for p in parts] + [p[1] for p in parts], axis)


def _lstm(x, w_ih, w_hh, b_ih, b_hh):
    """single-layer unidirectional nn.LSTM, zero initial state, gate order i,f,g,o; x [T,B,D]"""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    pre = x @ w_ih.t() + b_ih + b_hh
    outs = []
    for t in range(T):
        g = pre[t] + h @ w_hh.t()
        i, f, gg, o = g.chunk(4, 1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 0)


# ------------------------------------------------------------------ tools_for_model.py:159-174
def complex_lstm(real, imag, sd, prefix, project):
    def p(n):
        return sd[prefix + n]
    rl = lambda x: _lstm(x, p('real_lstm.weight_ih_l0'), p('real_lstm.weight_hh_l0'),
                         p('real_lstm.bias_ih_l0'), p('real_lstm.bias_hh_l0'))
    il = lambda x: _lstm(x, p('imag_lstm.weight_ih_l0'), p('imag_lstm.weight_hh_l0'),
                         p('imag_lstm.bias_ih_l0'), p('imag_lstm.bias_hh_l0'))
    real_out = rl(real) - il(imag)
    imag_out = rl(imag) + il(real)
    if project:
        real_out = F.linear(real_out, p('r_trans.weight'), p('r_trans.bias'))
        imag_out = F.linear(imag_out, p('i_trans.weight'), p('i_trans.bias'))
    return real_out, imag_out


def _bn_prelu(x, sd, prefix, training, eps=1e-5, momentum=0.1, update=None):
    """nn.BatchNorm2d + nn.PReLU of DCCRN.py:80-82; running stats are updated in `update` (a dict)
    when training."""
    if (prefix + '1.Wrr') in sd:                 # use_cbn=True: ComplexBatchNorm in slot .1 (DCCRN.py:80-81)
        keys = ('Wrr', 'Wri', 'Wii', 'Br', 'Bi', 'RMr', 'RMi', 'RVrr', 'RVri', 'RVii')
        upd = {} if (training and update is not None) else None
        y = complex_batch_norm(x, {k: sd[prefix + '1.' + k] for k in keys}, training, eps, momentum, update=upd)
        if upd is not None:
            for k, v in upd.items():
                update[prefix + '1.' + k] = v
        return F.prelu(y, sd[prefix + '2.weight'])
    w, b = sd[prefix + '1.weight'], sd[prefix + '1.bias']
    rm, rv = sd[prefix + '1.running_mean'], sd[prefix + '1.running_var']
    if training:
        rm, rv = rm.clone(), rv.clone()
        y = F.batch_norm(x, rm, rv, w, b, True, momentum, eps)
        if update is not None:
            update[prefix + '1.running_mean'], update[prefix + '1.running_var'] = rm, rv
    else:
        y = F.batch_norm(x, rm, rv, w, b, False, momentum, eps)
    return F.prelu(y, sd[prefix + '2.weight'])


# ------------------------------------------------------------------ tools_for_model.py:398-508
def complex_batch_norm(x, p, training, eps=1e-5, momentum=0.1, update=None):
    """ComplexBatchNorm.forward (Trabelsi 2x2 whitening).  p: dict with Wrr, Wri, Wii, Br, Bi and the
    running buffers RMr, RMi, RVrr, RVri, RVii; x: [B, 2C, ...] (real half then imag half of dim 1).
    When training, the lerp-updated running buffers are returned in `update`."""
    xr, xi = torch.chunk(x, 2, 1)
    red = [d for d in range(xr.dim()) if d != 1]
    vdim = [1] * xr.dim()
    vdim[1] = xr.size(1)
    if training:
        Mr, Mi = xr.mean(red, keepdim=True), xi.mean(red, keepdim=True)          # :424-428
    else:
        Mr, Mi = p['RMr'].view(vdim), p['RMi'].view(vdim)
    xr, xi = xr - Mr, xi - Mi                                                   # :435
    if training:
        Vrr, Vri, Vii = (xr * xr).mean(red, keepdim=True), (xr * xi).mean(red, keepdim=True), \
            (xi * xi).mean(red, keepdim=True)                                   # :444-450
        if update is not None:                                                  # lerp_ :433-434,455-457
            for k, v in (('RMr', Mr), ('RMi', Mi), ('RVrr', Vrr), ('RVri', Vri), ('RVii', Vii)):
                update[k] = p[k] + momentum * (v.reshape(-1) - p[k])
    else:
        Vrr, Vri, Vii = p['RVrr'].view(vdim), p['RVri'].view(vdim), p['RVii'].view(vdim)
    Vrr, Vii = Vrr + eps, Vii + eps                                             # :462-464
    tau = Vrr + Vii
    delta = Vrr * Vii - Vri * Vri                                               # :472
    s_ = delta.sqrt()
    t = (tau + 2 * s_).sqrt()
    rst = (s_ * t).reciprocal()                                                 # :477
    Urr, Uii, Uri = (s_ + Vii) * rst, (s_ + Vrr) * rst, -Vri * rst
    Wrr, Wri, Wii = p['Wrr'].view(vdim), p['Wri'].view(vdim), p['Wii'].view(vdim)
    Zrr, Zri = Wrr * Urr + Wri * Uri, Wrr * Uri + Wri * Uii                     # :493-496
    Zir, Zii = Wri * Urr + Wii * Uri, Wri * Uri + Wii * Uii
    yr = Zrr * xr + Zri * xi + p['Br'].view(vdim)
    yi = Zir * xr + Zii * xi + p['Bi'].view(vdim)
    return torch.cat([yr, yi], 1)


def mask_apply(real, imag, mask_real, mask_imag, mode):
    """DCCRN.py:153,159,212-230"""
    if mode == 'E':
        spec_mags = torch.sqrt(real ** 2 + imag ** 2 + 1e-8)
        spec_phase = torch.atan2(imag, real)
        mask_mags = (mask_real ** 2 + mask_imag ** 2) ** 0.5
        real_phase = mask_real / (mask_mags + 1e-8)
        imag_phase = mask_imag / (mask_mags + 1e-8)
        mask_phase = torch.atan2(imag_phase, real_phase)
        est_mags = torch.tanh(mask_mags) * spec_mags
        est_phase = spec_phase + mask_phase
        return est_mags * torch.cos(est_phase), est_mags * torch.sin(est_phase)
    if mode == 'C':
        return real * mask_real - imag * mask_imag, real * mask_imag + imag * mask_real
    if mode == 'R':
        return real * mask_real, imag * mask_imag
    raise ValueError(mode)


def dccrn_forward(sd, x, masking_mode='E', win_len=400, hop=100, fft_len=512, training=False, taps=None,
                  update=None):
    """DCCRN.forward (DCCRN.py:149-240) for use_clstm=True; use_cbn is inferred from the state_dict keys
    (`*.1.Wrr` present = ComplexBatchNorm in the norm slot).
    sd: state_dict with the reference's keys; returns (mask_real, mask_imag, real, imag, wav).
    taps (dict) collects what feature_extraction.DCCRN's hooks see (feature_extraction.py:11-13)."""
    n_layers = len([k for k in sd if k.startswith('encoder.') and k.endswith('.0.real_conv.weight')])
    nb = fft_len // 2 + 1
    specs = conv_stft(x, sd['stft.weight'], win_len, hop)                       # :150
    real, imag = specs[:, :nb], specs[:, nb:]
    out = torch.stack([real, imag], 1)[:, :, 1:]                                  # :161-162
    enc = []
    for i in range(n_layers):                                                     # :173-176
        pre = 'encoder.%d.' % i
        out = complex_conv2d(out, sd[pre + '0.real_conv.weight'], sd[pre + '0.real_conv.bias'],
                             sd[pre + '0.imag_conv.weight'], sd[pre + '0.imag_conv.bias'])
        out = _bn_prelu(out, sd, pre, training, update=update)
        enc.append(out)
    B, C, D, T = out.shape                                                        # :178-184
    o = out.permute(3, 0, 1, 2)
    r = o[:, :, :C // 2].reshape(T, B, C // 2 * D)
    i_ = o[:, :, C // 2:].reshape(T, B, C // 2 * D)
    if 'enhance.weight_ih_l0' in sd:        # use_clstm=False: 2-layer nn.LSTM + Linear "tranform" (:100-110, :193-199)
        y = o.reshape(T, B, C * D)
        l = 0
        while ('enhance.weight_ih_l%d' % l) in sd:
            y = _lstm(y, sd['enhance.weight_ih_l%d' % l], sd['enhance.weight_hh_l%d' % l],
                      sd['enhance.bias_ih_l%d' % l], sd['enhance.bias_hh_l%d' % l])
            l += 1
        lstm_tap = [y]
        y = F.linear(y, sd['tranform.weight'], sd['tranform.bias'])
        out = y.reshape(T, B, C, D).permute(1, 2, 3, 0)
    else:
        n_rnn = len([k for k in sd if k.startswith('enhance.') and k.endswith('.real_lstm.weight_ih_l0')])
        for l in range(n_rnn):                                                    # :186
            r, i_ = complex_lstm(r, i_, sd, 'enhance.%d.' % l, ('enhance.%d.r_trans.weight' % l) in sd)
        lstm_tap = [r, i_]
        r = r.reshape(T, B, C // 2, D)
        i_ = i_.reshape(T, B, C // 2, D)
        out = torch.cat([r, i_], 2).permute(1, 2, 3, 0)                           # :188-199
    dec = []
    for idx in range(n_layers):                                                   # :201-205
        pre = 'decoder.%d.' % idx
        out = complex_cat([out, enc[-1 - idx]], 1)
        out = complex_deconv2d(out, sd[pre + '0.real_conv.weight'], sd[pre + '0.real_conv.bias'],
                               sd[pre + '0.imag_conv.weight'], sd[pre + '0.imag_conv.bias'])
        if (pre + '1.weight') in sd or (pre + '1.Wrr') in sd:
            out = _bn_prelu(out, sd, pre, training, update=update)
        dec.append(out)
        out = out[..., 1:]
    mask_real = F.pad(out[:, 0], [0, 0, 1, 0])                                    # :207-210
    mask_imag = F.pad(out[:, 1], [0, 0, 1, 0])
    est_r, est_i = mask_apply(real, imag, mask_real, mask_imag, masking_mode)
    wav = conv_istft(torch.cat([est_r, est_i], 1), sd['istft.weight'], sd['istft.window'], win_len, hop)
    wav = torch.clamp(wav.squeeze(1), -1, 1)                                      # :235-237
    if taps is not None:
        taps['encoder'], taps['decoder'], taps['clstm'] = enc, dec, lstm_tap
    return mask_real, mask_imag, est_r, est_i, wav


def make_state_dict(kernel_num, rnn_units, seed=0, rnn_layers=2, win_len=400, hop=100, fft_len=512,
                    win_type='hamming', kernel_size=5, randomize_bn=True):
    """Random weights with the reference constructors' distributions (tools_for_model.py:231-234,
    298-301; torch defaults elsewhere) and the reference's state_dict keys / shapes.
    BN affine parameters / running stats are randomised so that parity tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    ru = lambda *s: torch.rand(*s, generator=g)
    kn = [2] + list(kernel_num)
    sd = {}
    sd['stft.weight'], _ = init_kernels(win_len, hop, fft_len, win_type)
    sd['istft.weight'], sd['istft.window'] = init_kernels(win_len, hop, fft_len, win_type, invers=True)
    sd['istft.enframe'] = torch.eye(win_len)[:, None, :]

    def bn(prefix, ch):
        sd[prefix + '1.weight'] = 1 + 0.2 * rn(ch) if randomize_bn else torch.ones(ch)
        sd[prefix + '1.bias'] = 0.1 * rn(ch) if randomize_bn else torch.zeros(ch)
        sd[prefix + '1.running_mean'] = 0.1 * rn(ch) if randomize_bn else torch.zeros(ch)
        sd[prefix + '1.running_var'] = 0.5 + ru(ch) if randomize_bn else torch.ones(ch)
        sd[prefix + '1.num_batches_tracked'] = torch.tensor(0)
        sd[prefix + '2.weight'] = torch.tensor([0.25])

    for i in range(len(kn) - 1):
        cin, cout = kn[i] // 2, kn[i + 1] // 2
        p = 'encoder.%d.' % i
        for part in ('real_conv', 'imag_conv'):
            sd[p + '0.%s.weight' % part] = 0.05 * rn(cout, cin, kernel_size, 2)
            sd[p + '0.%s.bias' % part] = 0.02 * rn(cout) if randomize_bn else torch.zeros(cout)
        bn(p, kn[i + 1])
    hidden = fft_len // (2 ** len(kn))
    D0 = hidden * kn[-1] // 2
    H = rnn_units // 2
    for l in range(rnn_layers):
        D = D0 if l == 0 else H
        k = 1.0 / math.sqrt(H)
        for part in ('real_lstm', 'imag_lstm'):
            p = 'enhance.%d.%s.' % (l, part)
            sd[p + 'weight_ih_l0'] = (ru(4 * H, D) * 2 - 1) * k
            sd[p + 'weight_hh_l0'] = (ru(4 * H, H) * 2 - 1) * k
            sd[p + 'bias_ih_l0'] = (ru(4 * H) * 2 - 1) * k
            sd[p + 'bias_hh_l0'] = (ru(4 * H) * 2 - 1) * k
        if l == rnn_layers - 1:
            for part in ('r_trans', 'i_trans'):
                sd['enhance.%d.%s.weight' % (l, part)] = (ru(D0, H) * 2 - 1) * k
                sd['enhance.%d.%s.bias' % (l, part)] = (ru(D0) * 2 - 1) * k
    for j, idx in enumerate(range(len(kn) - 1, 0, -1)):
        cin, cout = kn[idx], kn[idx - 1] // 2          # ComplexConvTranspose2d(kn[idx]*2, kn[idx-1])
        p = 'decoder.%d.' % j
        for part in ('real_conv', 'imag_conv'):
            sd[p + '0.%s.weight' % part] = 0.05 * rn(cin, cout, kernel_size, 2)
            sd[p + '0.%s.bias' % part] = 0.02 * rn(cout) if randomize_bn else torch.zeros(cout)
        if idx != 1:
            bn(p, kn[idx - 1])
    return sd
