"""ORACLE (test infrastructure only) - CPU restatement of the reference's objectives and
distillation losses.  Citations are file:line relative to the reference tree."""
import torch
import torch.nn.functional as F

from .dccrn_oracle import dccrn_forward


# ------------------------------------------------------------------ tools_for_loss.py:22-47
def _dot(a, b):
    return torch.sum(a * b, -1, keepdim=True)


def sdr(s1, s2, eps=1e-8):
    sn, d = _dot(s1, s1), _dot(s1 - s2, s1 - s2)
    return torch.mean(10 * torch.log10(sn ** 2 / (d ** 2 + eps)))


def si_snr(s1, s2, eps=1e-8):
    s_target = _dot(s1, s2) / (_dot(s2, s2) + eps) * s2
    e_noise = s1 - s_target
    return torch.mean(10 * torch.log10(_dot(s_target, s_target) / (_dot(e_noise, e_noise) + eps) + eps))


# ------------------------------------------------------------------ tools_for_loss.py:83-97
def si_sdr(reference, estimation, eps=1e-8):
    ref_e = torch.sum(reference ** 2, -1, keepdim=True)
    a = torch.sum(reference * estimation, -1, keepdim=True) / ref_e + eps
    proj = a * reference
    noise = estimation - proj
    ratio = torch.sum(proj ** 2, -1) / torch.sum(noise ** 2, -1) + eps
    return 10 * torch.log10(torch.mean(ratio) + eps)


# ------------------------------------------------------------------ framework.py:16-32 (restated
# with return_complex=True because torch>=2 rejects the reference's call; otherwise identical)
def stft_mag(x, fft_size, hop, win_length, window):
    s = torch.stft(x, fft_size, hop, win_length, window, return_complex=True)
    return torch.sqrt(torch.clamp(s.real ** 2 + s.imag ** 2, min=1e-7)).transpose(2, 1)


# ------------------------------------------------------------------ framework.py:35-99
def stft_loss(x, y, fft_size, hop, win_length, window=None):
    window = torch.hann_window(win_length) if window is None else window
    xm, ym = stft_mag(x, fft_size, hop, win_length, window), stft_mag(y, fft_size, hop, win_length, window)
    sc = torch.norm(ym - xm, p="fro") / torch.norm(ym, p="fro")
    mag = F.l1_loss(torch.log(ym), torch.log(xm))
    return sc, mag


# ------------------------------------------------------------------ framework.py:128-146
def mr_stft_loss(x, y, fft_sizes, hop_sizes, win_lengths, factor_sc=0.1, factor_mag=0.1):
    sc = mag = 0.0
    for fs, hs, wl in zip(fft_sizes, hop_sizes, win_lengths):
        s, m = stft_loss(x, y, fs, hs, wl)
        sc, mag = sc + s, mag + m
    n = len(fft_sizes)
    return factor_sc * sc / n, factor_mag * mag / n


# ------------------------------------------------------------------ framework.py:157-172
def spkd(student, teacher, reduction='batchmean'):
    def g(z):
        z = torch.flatten(z, 1)
        return F.normalize(z @ z.t(), 1)          # NB: positional 1 is p -> L1 row normalisation (:159)
    loss = torch.norm(g(teacher) - g(student)) ** 2
    return loss / (teacher.shape[0] ** 2) if reduction == 'batchmean' else loss


# ------------------------------------------------------------------ framework.py:206-224
def abf_forward(x, y, sd, prefix, shape, out_shape, training=True, eps=1e-5):
    """sd keys: conv1.0.weight, conv1.1.{weight,bias,running_mean,running_var}, conv2.*, att_conv.0.*"""
    def bn(t, p):
        return F.batch_norm(t, sd[p + 'running_mean'].clone(), sd[p + 'running_var'].clone(), sd[p + 'weight'],
                            sd[p + 'bias'], training, 0.1, eps)
    n, _, h, w = x.shape
    x = bn(F.conv2d(x, sd[prefix + 'conv1.0.weight']), prefix + 'conv1.1.')
    if (prefix + 'att_conv.0.weight') in sd:
        y = F.interpolate(y, (shape, w), mode="nearest")
        z = torch.sigmoid(F.conv2d(torch.cat([x, y], 1), sd[prefix + 'att_conv.0.weight'],
                                   sd[prefix + 'att_conv.0.bias']))
        x = x * z[:, 0].view(n, 1, h, w) + y * z[:, 1].view(n, 1, h, w)
    if x.shape[-1] != out_shape:                  # compares the TIME width to a freq size (:221): a no-op resize
        x = F.interpolate(x, (out_shape, w), mode="nearest")
    y = bn(F.conv2d(x, sd[prefix + 'conv2.0.weight'], padding=1), prefix + 'conv2.1.')
    return y, x


# ------------------------------------------------------------------ framework.py:244-263
def review_kd_forward(feature_maps, sd, shapes, out_shapes, ft_type, training=True):
    """sd holds ReviewKD's state_dict ('abfs.{i}.…', i = 0 deepest)."""
    maps = feature_maps[::-1] if ft_type == 'encoder' else list(feature_maps)
    out, res = abf_forward(maps[0], None, sd, 'abfs.0.', None, out_shapes[0], training)
    results = [out]
    for i in range(1, len(maps)):
        out, res = abf_forward(maps[i], res, sd, 'abfs.%d.' % i, shapes[i], out_shapes[i], training)
        if ft_type == 'encoder':
            results.insert(0, out)
        else:
            results.append(out)
    return results


# ------------------------------------------------------------------ framework.py:287-306 (on 4-D maps)
def hcl(fstudent, fteacher):
    total = 0.0
    for fs, ft in zip(fstudent, fteacher):
        h = fs.shape[2]
        loss = F.mse_loss(fs, ft)
        cnt, tot = 1.0, 1.0
        for l in (4, 2, 1):
            if l >= h:
                continue
            cnt /= 2.0
            loss = loss + F.mse_loss(F.adaptive_avg_pool2d(fs, (l, l)), F.adaptive_avg_pool2d(ft, (l, l))) * cnt
            tot += cnt
        total = total + loss / tot
    return total


# ------------------------------------------------------------------ distill.py:72-148 on the local DCCRN
def clskd_step_loss(teacher_sd, student_sd, X, y, abf_enc_sd=None, abf_dec_sd=None, mode='clskd',
                    teacher_training=False, student_training=True):
    """Loss of one distillation step (SURVEY.md 3.1 adaptations: local DCCRN taps, batch-first LSTM
    taps, ReviewKD channel/shape lists read off the maps).  Returns (loss, dict of terms)."""
    tt, st = {}, {}
    with torch.no_grad():
        t_wav = dccrn_forward(teacher_sd, X, training=teacher_training, taps=tt)[-1]
    s_wav = dccrn_forward(student_sd, X, training=student_training, taps=st)[-1]
    if mode == 'reviewkd':          # distill_ReviewKD.py:56 base-loss resolution
        terms = {'base': mr_stft_loss(s_wav, y, [512], [16], [32])[1]}
    else:
        terms = {'base': mr_stft_loss(s_wav, y, [512], [100], [400])[1]}
    if mode == 'reviewkd':
        # distill_ReviewKD.py:92-125 on the local DCCRN: hcl between fused student maps and teacher maps;
        # the LSTM taps go through hcl when their widths agree and SPKD otherwise (hcl needs equal shapes)
        e_shapes = [m.shape[2] for m in st['encoder']][::-1]
        d_shapes = [m.shape[2] for m in st['decoder']]
        f_enc = review_kd_forward(st['encoder'], abf_enc_sd, e_shapes, e_shapes, 'encoder')
        f_dec = review_kd_forward(st['decoder'], abf_dec_sd, d_shapes, d_shapes, 'decoder')
        terms['encoder'] = hcl(f_enc, tt['encoder'])
        terms['decoder'] = hcl(f_dec, tt['decoder'])
        for key, i in (('clstm_real', 0), ('clstm_img', 1)):
            a, b = st['clstm'][i].transpose(0, 1), tt['clstm'][i].transpose(0, 1)
            terms[key] = hcl([a.unsqueeze(1)], [b.unsqueeze(1)]) if a.shape == b.shape else spkd(a, b)
    elif mode in ('clskd', 'spkd_all'):
        if mode == 'clskd':
            e_shapes = [m.shape[2] for m in st['encoder']][::-1]
            d_shapes = [m.shape[2] for m in st['decoder']]
            f_enc = review_kd_forward(st['encoder'], abf_enc_sd, e_shapes, e_shapes, 'encoder')
            f_dec = review_kd_forward(st['decoder'], abf_dec_sd, d_shapes, d_shapes, 'decoder')
        else:
            f_enc, f_dec = st['encoder'], st['decoder']
        terms['encoder'] = sum(spkd(a, b) for a, b in zip(f_enc, tt['encoder']))
        terms['decoder'] = sum(spkd(a, b) for a, b in zip(f_dec, tt['decoder']))
        terms['clstm_real'] = spkd(st['clstm'][0].transpose(0, 1), tt['clstm'][0].transpose(0, 1))
        terms['clstm_img'] = spkd(st['clstm'][1].transpose(0, 1), tt['clstm'][1].transpose(0, 1))
    elif mode == 'spkd':
        terms['kd'] = spkd(s_wav.unsqueeze(1), t_wav.unsqueeze(1))
    elif mode == 'mse':
        terms['kd'] = F.mse_loss(s_wav, t_wav)
    elif mode == 'stft':
        terms['kd'] = mr_stft_loss(s_wav, t_wav, [512], [100], [400])[1]
    else:
        raise ValueError(mode)
    return sum(terms.values()), terms


# ------------------------------------------------------------------ framework.py:176-202, 226-242
def make_abf_state_dict(in_channels, out_channels, seed=0, randomize_bn=True):
    """Deterministic ReviewKD weights ('abfs.{i}.…', i = 0 deepest) with the reference constructor's
    distributions: conv1 / conv2 kaiming_uniform(a=1) (framework.py:194-195), att_conv torch default
    (kaiming_uniform(a=sqrt(5)) + uniform bias).  in_channels / out_channels are listed shallow -> deep
    like ReviewKD's arguments; mid = min(512, in_channels[-1]) (framework.py:238).  Lets a parity test
    inject identical fusion weights into the reference, this oracle and the CUDA implementation without
    shipping megabytes of weights in a fixture."""
    import math
    g = torch.Generator().manual_seed(seed)
    mid = min(512, in_channels[-1])
    last = len(in_channels) - 1
    sd = {}

    def ku(shape, a):
        fan_in = shape[1] * shape[2] * shape[3]
        bound = math.sqrt(2.0 / (1 + a * a)) * math.sqrt(3.0 / fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    def bn(p, ch):
        sd[p + 'weight'] = 1 + 0.2 * torch.randn(ch, generator=g) if randomize_bn else torch.ones(ch)
        sd[p + 'bias'] = 0.1 * torch.randn(ch, generator=g) if randomize_bn else torch.zeros(ch)
        sd[p + 'running_mean'] = torch.zeros(ch)
        sd[p + 'running_var'] = torch.ones(ch)
        sd[p + 'num_batches_tracked'] = torch.tensor(0)

    for idx in range(last, -1, -1):                       # abfs = blocks[::-1]: deepest first
        p = 'abfs.%d.' % (last - idx)
        sd[p + 'conv1.0.weight'] = ku((mid, in_channels[idx], 1, 1), 1.0)
        bn(p + 'conv1.1.', mid)
        sd[p + 'conv2.0.weight'] = ku((out_channels[idx], mid, 3, 3), 1.0)
        bn(p + 'conv2.1.', out_channels[idx])
        if idx < last:
            sd[p + 'att_conv.0.weight'] = ku((2, 2 * mid, 1, 1), math.sqrt(5))
            b = 1.0 / math.sqrt(2 * mid)
            sd[p + 'att_conv.0.bias'] = (torch.rand(2, generator=g) * 2 - 1) * b
    return sd
