"""DCCRN: deep complex convolution recurrent network on B200-native kernels.

Same class name, constructor signature, submodule names (`stft`, `istft`, `encoder`, `enhance`,
`decoder` - the hook targets of feature_extraction) and `state_dict` keys as the reference
(DCCRN.py:14-257), but the forward is a chain of fused sm_100a kernels on channels-last
activations: framed-window DFT GEMM -> 6x (complex conv as one 2x-wide implicit GEMM -> BatchNorm
statistics -> normalise+PReLU) -> complex LSTM (batched projections + persistent recurrence) ->
6x (skip connection as a second K segment -> sub-pixel complex transposed conv -> BN+PReLU) ->
trig-free polar mask -> synthesis GEMM + overlap-add + clamp.
"""
import torch
import torch.nn as nn

from . import config as cfg
from . import ops
from .clstm import LSTM, NavieComplexLSTM, _LinearParams
from .ops import MaskFn, dense, strided_copy, strided_copy_into, to_logical, to_phys
from .tools_for_loss import mse, sdr, si_sdr, si_snr
from .tools_for_model import (BatchNorm2d, ComplexBatchNorm, ComplexConv2d, ComplexConvTranspose2d, ConvBNAct,
                              ConviSTFT, ConvSTFT, PReLU)


class _ToLstmFn(torch.autograd.Function):
    """encoder output (physical [B,T,F,2Cc]) -> X [2,T,B,Cc*F] with feature index c*F+f
    (the permute/reshape of DCCRN.py:178-184)."""

    @staticmethod
    def forward(ctx, x, dtype):
        B, T, F, C = x.shape
        Cc = C // 2
        X = torch.empty((2, T, B, Cc * F), dtype=dtype, device=x.device)
        for p in range(2):
            strided_copy_into(x[..., p * Cc:(p + 1) * Cc].permute(1, 0, 3, 2), X[p].view(T, B, Cc, F))
        ctx.meta = (x.shape, x.dtype)
        return X

    @staticmethod
    def backward(ctx, g):
        shape, dt = ctx.meta
        B, T, F, C = shape
        Cc = C // 2
        g = dense(g)
        dx = torch.empty(shape, dtype=dt, device=g.device)
        for p in range(2):
            strided_copy_into(g[p].view(T, B, Cc, F).permute(1, 0, 3, 2), dx[..., p * Cc:(p + 1) * Cc])
        return dx, None


class _FromLstmFn(torch.autograd.Function):
    """LSTM outputs (real, imag), each [T,B,Cc*F] -> physical [B,T,F,2Cc] (DCCRN.py:188-199)."""

    @staticmethod
    def forward(ctx, r, i, F, dtype):
        T, B, D = r.shape
        Cc = D // F
        out = torch.empty((B, T, F, 2 * Cc), dtype=dtype, device=r.device)
        for p, y in enumerate((r, i)):
            strided_copy_into(y.reshape(T, B, Cc, F).permute(1, 0, 3, 2), out[..., p * Cc:(p + 1) * Cc])
        ctx.meta = (r.shape, r.dtype, i.dtype, F)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, dtr, dti, F = ctx.meta
        T, B, D = shape
        Cc = D // F
        g = dense(g)
        outs = []
        for p, dt in enumerate((dtr, dti)):
            d = torch.empty(shape, dtype=dt, device=g.device)
            strided_copy_into(g[..., p * Cc:(p + 1) * Cc].permute(1, 0, 3, 2), d.view(T, B, Cc, F))
            outs.append(d)
        return outs[0], outs[1], None, None


class _Cat2Fn(torch.autograd.Function):
    """X [2, T, B, D] -> [T, B, 2D] (part 0 features, then part 1 features)."""

    @staticmethod
    def forward(ctx, X):
        _, T, B, D = X.shape
        out = torch.empty((T, B, 2 * D), dtype=X.dtype, device=X.device)
        strided_copy_into(X[0], out[..., :D])
        strided_copy_into(X[1], out[..., D:])
        return out

    @staticmethod
    def backward(ctx, g):
        D = g.shape[-1] // 2
        dX = torch.empty((2,) + tuple(g.shape[:2]) + (D,), dtype=g.dtype, device=g.device)
        strided_copy_into(g[..., :D], dX[0])
        strided_copy_into(g[..., D:], dX[1])
        return dX


class _TrimFirstFn(torch.autograd.Function):
    """physical [B,T+1,F,C] -> view [B,T,F,C] without column 0 (DCCRN.py:205)."""

    @staticmethod
    def forward(ctx, x):
        ctx.meta = (x.shape, x.dtype)
        return x[:, 1:]

    @staticmethod
    def backward(ctx, g):
        shape, dt = ctx.meta
        dx = torch.empty(shape, dtype=dt, device=g.device)
        dx[:, 0].zero_()
        strided_copy_into(g, dx[:, 1:])
        return dx


class DCCRN(nn.Module):

    def __init__(self, rnn_layers=cfg.rnn_layers, rnn_units=cfg.rnn_units, win_len=cfg.win_len,
                 win_inc=cfg.win_inc, fft_len=cfg.fft_len, win_type=cfg.window_type, masking_mode='E',
                 use_clstm=False, use_cbn=False, kernel_size=5, kernel_num=[16, 32, 64, 128, 256, 256]):
        """rnn_layers: number of LSTM layers; rnn_units / kernel_num count real+imag channels."""
        super().__init__()
        self.win_len, self.win_inc, self.fft_len, self.win_type = win_len, win_inc, fft_len, win_type
        self.rnn_units = rnn_units
        self.input_dim = self.output_dim = win_len
        self.hidden_layers = rnn_layers
        self.kernel_size = kernel_size
        self.kernel_num = [2] + list(kernel_num)
        self.masking_mode = masking_mode
        self.use_clstm = use_clstm
        self.fix = True
        self.stft = ConvSTFT(win_len, win_inc, fft_len, win_type, 'complex', fix=True)
        self.istft = ConviSTFT(win_len, win_inc, fft_len, win_type, 'complex', fix=True)

        def norm(ch):
            return ComplexBatchNorm(ch) if use_cbn else BatchNorm2d(ch)

        kn = self.kernel_num
        self.encoder = nn.ModuleList()
        self.decoder = nn.ModuleList()
        for idx in range(len(kn) - 1):
            self.encoder.append(ConvBNAct(
                ComplexConv2d(kn[idx], kn[idx + 1], kernel_size=(kernel_size, 2), stride=(2, 1), padding=(2, 1)),
                norm(kn[idx + 1]), PReLU()))
        hidden_dim = fft_len // (2 ** len(kn))
        if use_clstm:
            rnns = []
            for idx in range(rnn_layers):
                rnns.append(NavieComplexLSTM(
                    input_size=hidden_dim * kn[-1] if idx == 0 else rnn_units, hidden_size=rnn_units,
                    bidirectional=False, batch_first=False,
                    projection_dim=hidden_dim * kn[-1] if idx == rnn_layers - 1 else None))
            self.enhance = nn.Sequential(*rnns)
        else:       # plain 2-layer nn.LSTM + Linear "tranform" (DCCRN.py:100-110; the typo is the reference's key)
            self.enhance = LSTM(input_size=hidden_dim * kn[-1], hidden_size=rnn_units, num_layers=2)
            self.tranform = _LinearParams(rnn_units, hidden_dim * kn[-1])
        for idx in range(len(kn) - 1, 0, -1):
            conv = ComplexConvTranspose2d(kn[idx] * 2, kn[idx - 1], kernel_size=(kernel_size, 2), stride=(2, 1),
                                          padding=(2, 0), output_padding=(1, 0))
            self.decoder.append(ConvBNAct(conv, norm(kn[idx - 1]), PReLU()) if idx != 1 else ConvBNAct(conv))
        self.flatten_parameters()

    def flatten_parameters(self):
        pass

    def forward(self, inputs, lens=None, is_feat=None):
        ops._require_cuda(inputs)
        act = ops.policy.act_dtype
        spec = self.stft.spectrum(inputs, interleaved=True)            # fp32 [B, T, 257, 2]
        # DC bin dropped (DCCRN.py:162); logical [B, 2, 256, T]
        out = to_logical(dense(spec[:, :, 1:, :]))
        encoder_out = []
        for layer in self.encoder:
            out = layer(out)
            skip, out = ops.fanout(out, 2)       # skip connection + next layer: gradients summed by one library kernel
            encoder_out.append(skip)
        # ---- complex LSTM bottleneck
        x = to_phys(out, need_dense=True)
        F = x.shape[2]
        if self.use_clstm:
            X = _ToLstmFn.apply(x, act)
            r, i = self.enhance([X[0], X[1]])
            out = to_logical(_FromLstmFn.apply(r, i, F, act))
        else:       # DCCRN.py:193-199: [T, B, C*D] -> LSTM -> Linear -> [T, B, C, D]
            Bb, Tt, _, C = x.shape
            X = _ToLstmFn.apply(x, act)                               # [2, T, B, Cc*F]: feature c*F+f per half
            seq = _Cat2Fn.apply(X)                                    # [T, B, C*F] (real-half features, then imag-half)
            y, _ = self.enhance(seq)
            y = self.tranform(y)                                      # [T, B, C*F] fp32
            half = y.shape[-1] // 2
            out = to_logical(_FromLstmFn.apply(y[..., :half], y[..., half:], F, act))
        # ---- decoder: skip connections are the second K segment of each transposed conv
        for idx in range(len(self.decoder)):
            out = self.decoder[idx](out, encoder_out[-1 - idx])
            out = to_logical(_TrimFirstFn.apply(to_phys(out)))
        mask = to_phys(out)                                            # [B, T, 256, 2] (batch-strided view)
        if mask.stride(3) != 1 or mask.stride(2) != 2:
            mask = dense(mask)
        want = is_feat is not True
        out_spec, mp = MaskFn.apply(spec, mask, self.masking_mode, want)
        B, T = out_spec.shape[0], out_spec.shape[1]
        out_wav = self.istft.synthesize(out_spec.view(B, T, -1), interleaved=True, clamp=True)
        if is_feat is True:
            return out_wav
        mask_real, mask_imag = mp[..., 0].permute(0, 2, 1), mp[..., 1].permute(0, 2, 1)
        real, imag = out_spec[..., 0].permute(0, 2, 1), out_spec[..., 1].permute(0, 2, 1)
        return mask_real, mask_imag, real, imag, out_wav

    # ------------------------------------------------------------------------------------------
    # time-chunked streaming inference (the reference's eval path eval.py:42-60 runs whole
    # utterances one at a time).  Exact, not approximate: apart from the LSTM the model has a bounded temporal
    # footprint - every encoder conv looks back ONE frame (time kernel 2, causal left pad 1,
    # tools_for_model.py:237-238), the LSTM is unidirectional (state carried in (h, c)), every transposed conv
    # followed by the drop of column 0 (DCCRN.py:205) looks AHEAD one frame (6 frames = 37.5 ms in total), eval
    # BatchNorm / PReLU / mask are pointwise, and the iSTFT overlap-add spans 4 frames.  So activation memory is
    # bounded by the chunk length instead of the utterance length.
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _tcat(parts):
        """concatenate logical [B, C, F, t] maps along time (physical [B, t, F, C]) with the library's copy kernel"""
        ps = [to_phys(p) for p in parts]
        if len(ps) == 1 and ps[0].is_contiguous():
            return to_logical(strided_copy(ps[0]))
        B, _, F, C = ps[0].shape
        T = sum(p.shape[1] for p in ps)
        out = torch.empty((B, T, F, C), dtype=ps[0].dtype, device=ps[0].device)
        o = 0
        for p in ps:
            strided_copy_into(p, out[:, o:o + p.shape[1]])
            o += p.shape[1]
        return to_logical(out)

    def enhance_streaming(self, inputs, chunk_frames=400):
        """Enhanced waveform [B, L'] of `inputs` [B, L], identical (to fp32 rounding) to
        `forward(inputs, is_feat=True)` in eval mode, computed in time chunks of `chunk_frames` STFT frames."""
        if self.training:
            raise RuntimeError("enhance_streaming: eval mode only (train-mode BatchNorm couples all frames)")
        if not self.use_clstm:
            raise NotImplementedError("enhance_streaming: implemented for the complex-LSTM bottleneck (DCCRN-CL)")
        ops._require_cuda(inputs)
        if inputs.dim() == 3:
            inputs = inputs.squeeze(1)
        act = ops.policy.act_dtype
        B, L = inputs.shape
        hop, win = self.win_inc, self.win_len
        pad = win - hop
        nlay = len(self.decoder)
        T = (L + 2 * pad - win) // hop + 1
        if T <= chunk_frames + 2 * nlay + 8:
            return self.forward(inputs, is_feat=True)
        from .ops import Pad1dFn
        with torch.no_grad():
            xpad = Pad1dFn.apply(inputs, pad, pad, 0)                       # [B, L + 600]
            bounds = list(range(0, T, chunk_frames)) + [T]
            if bounds[-1] - bounds[-2] < 2 * nlay + 8:                      # fold a short tail into the last chunk
                bounds.pop(-2)
            enc_prev = [None] * len(self.encoder)       # last input frame of every encoder layer
            states = [m.new_state(B, inputs.device) for m in self.enhance]
            tail_in, tail_skip, tail_spec = None, [None] * len(self.encoder), None   # last `nlay` frames
            tail_est = None                                                  # last 3 estimated-spectrum frames
            pieces = []
            for a, b in zip(bounds[:-1], bounds[1:]):
                last = b == T
                seg = dense(xpad[:, a * hop:(b - 1) * hop + win])
                spec = self.stft.spectrum_padded(seg)                        # fp32 [B, b-a, 257, 2]
                out = to_logical(dense(spec[:, :, 1:, :]))
                enc_out = []
                for li, layer in enumerate(self.encoder):
                    if enc_prev[li] is None:
                        nxt = layer(out)
                    else:
                        nxt = layer(self._tcat([enc_prev[li], out]))[..., 1:]
                    enc_prev[li] = self._tcat([out[..., -1:]])
                    out = nxt
                    enc_out.append(out)
                x = to_phys(out, need_dense=True)
                Fd = x.shape[2]
                X = _ToLstmFn.apply(x, act)
                for m, stt in zip(self.enhance, states):
                    X = m.forward_stacked_state(dense(X, act), stt)
                out = to_logical(_FromLstmFn.apply(X[0], X[1], Fd, act))
                # decoder on the window [a - nlay, b): its outputs are final for frames [a - nlay, b - nlay)
                if tail_in is not None:
                    out = self._tcat([tail_in, out])
                    enc_out = [self._tcat([t_, e_]) for t_, e_ in zip(tail_skip, enc_out)]
                    spec = self._cat_rows(tail_spec, spec)
                f0 = a - nlay if tail_in is not None else a
                if not last:
                    tail_in = self._tcat([out[..., -nlay:]])
                    tail_skip = [self._tcat([e_[..., -nlay:]]) for e_ in enc_out]
                    tail_spec = strided_copy(spec[:, -nlay:])
                for idx in range(nlay):
                    out = self.decoder[idx](out, enc_out[-1 - idx])
                    out = to_logical(_TrimFirstFn.apply(to_phys(out)))
                nvalid = out.shape[-1] if last else out.shape[-1] - nlay
                mask = to_phys(out)[:, :nvalid]
                if mask.stride(3) != 1 or mask.stride(2) != 2 or mask.stride(1) != mask.shape[2] * 2:
                    mask = dense(mask)
                est, _ = MaskFn.apply(dense(spec[:, :nvalid]), mask, self.masking_mode, False)
                f1 = f0 + nvalid
                rows = est.view(B, nvalid, -1)
                if tail_est is not None:
                    rows = self._cat_rows(tail_est, rows)
                tail_est = strided_copy(rows[:, -3:])
                pieces.append(self.istft.synthesize(rows, interleaved=True, clamp=True))
                del out, enc_out, spec, est, rows, X, x
            wav = torch.empty((B, sum(p_.shape[1] for p_ in pieces)), dtype=torch.float32, device=inputs.device)
            o = 0
            for p_ in pieces:
                strided_copy_into(p_.view(B, 1, 1, -1), wav[:, o:o + p_.shape[1]].unsqueeze(1).unsqueeze(1))
                o += p_.shape[1]
        return wav

    @staticmethod
    def _cat_rows(a, b):
        """concatenate [B, ta, ...] and [B, tb, ...] along dim 1 with the library's copy kernel"""
        B = a.shape[0]
        out = torch.empty((B, a.shape[1] + b.shape[1]) + tuple(a.shape[2:]), dtype=a.dtype, device=a.device)
        n = a[0, 0].numel()
        strided_copy_into(a.reshape(B, 1, a.shape[1], n), out[:, :a.shape[1]].reshape(B, 1, a.shape[1], n))
        strided_copy_into(b.reshape(B, 1, b.shape[1], n), out[:, a.shape[1]:].reshape(B, 1, b.shape[1], n))
        return out

    def get_params(self, weight_decay=0.0):
        weights = [p for n, p in self.named_parameters() if 'bias' not in n]
        biases = [p for n, p in self.named_parameters() if 'bias' in n]
        return [{'params': weights, 'weight_decay': weight_decay}, {'params': biases, 'weight_decay': 0.0}]

    def loss(self, inputs, labels, real_spec=None, img_spec=None, loss_mode=cfg.loss_mode):
        """Loss dispatcher (reference DCCRN.py:259-411).  The waveform objectives are implemented;
        the mel (LMS) and PMSQE modes depend on asteroid code outside this path."""
        if loss_mode == 'MSE':
            return mse(inputs, labels)
        if loss_mode == 'SDR':
            return -sdr(labels, inputs)
        if loss_mode == 'SI-SNR':
            return -si_snr(inputs, labels)
        if loss_mode == 'SI-SDR':
            return -si_sdr(labels, inputs)
        if loss_mode == 'MSE+SI-SNR':
            return (1 * -si_snr(inputs, labels) + 100 * mse(inputs, labels)) / 101
        if loss_mode == 'SI-SNR+SI-SDR':
            return (-si_snr(inputs, labels) + -si_sdr(inputs, labels)) / 2
        raise NotImplementedError("loss_mode %r needs the mel/PMSQE code of asteroid, which is outside the "
                                  "distillation path implemented here" % (loss_mode,))
