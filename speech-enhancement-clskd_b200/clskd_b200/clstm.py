"""Complex LSTM (reference: NavieComplexLSTM, tools_for_model.py:138-178).

The four LSTM passes of the reference (real_lstm / imag_lstm applied to the real and the imaginary
input) are two weight sets applied to a doubled batch: one input-projection GEMM for both sets,
one persistent-recurrence launch covering all four passes, and `real = rr - ii`, `imag = ir + ri`
as an axpby over contiguous blocks.  Internal layout is part-major: X [2, T, B, D] with X[0] the
real and X[1] the imaginary input, so the reference's lists of [T, B, D] tensors are views.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ops import ConvPlan, Launch, _codes, call, colstats, dense, f64_to_f32, pack_weights, run_tapconv, \
    run_wgrad, strided_copy_into, unpack_grads


class _LSTMParams(nn.Module):
    """Parameter holder with nn.LSTM's single-layer names and default init."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        k = 1.0 / math.sqrt(hidden_size)
        self.weight_ih_l0 = nn.Parameter(torch.empty(4 * hidden_size, input_size).uniform_(-k, k))
        self.weight_hh_l0 = nn.Parameter(torch.empty(4 * hidden_size, hidden_size).uniform_(-k, k))
        self.bias_ih_l0 = nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k))
        self.bias_hh_l0 = nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k))

    def flatten_parameters(self):
        pass


class _ClstmPlans:
    """Static tables of one complex-LSTM layer (two weight sets: sel 0 = real_lstm, 1 = imag_lstm)."""

    def __init__(self, D, H, device):
        G = 4 * H
        self.D, self.H = D, H
        # input projection: X[.., D] @ [Wr_ih^T | Wi_ih^T] -> [.., 8H]
        cr = _codes((G, D), 0).T
        ci = _codes((G, D), 1).T
        self.ih = ConvPlan("conv", np.concatenate([cr, ci], 1)[None, None], 1, 0, 0, D, 0, None, G * D, G * D,
                           device)
        # W_hh^T for the forward recurrence [2][H][4H], W_hh [2][4H][H] for BPTT
        hr, hi = _codes((G, H), 0), _codes((G, H), 1)
        self.whh_t = torch.from_numpy(ops._pairs(np.stack([hr.T, hi.T]).reshape(-1))).to(device)
        self.whh = torch.from_numpy(ops._pairs(np.stack([hr, hi]).reshape(-1))).to(device)
        # bias pairs (b_ih + b_hh) per set
        j = np.arange(G)
        self.bias = torch.from_numpy(np.stack([j * 4, j * 4 + 1], 1).astype(np.int32)).to(device)
        # dW_hh: wgrad gives [H][4H] per set -> parameter [4H][H]
        t = np.full((G * H, 2), -1, dtype=np.int32)
        gg, kk = np.meshgrid(np.arange(G), np.arange(H), indexing="ij")
        t[:, 0] = ((kk * G + gg) * 2).reshape(-1)
        self.hh_unpack = torch.from_numpy(t).to(device)
        self.cache = {}


def _clstm_forward(plans, X, wr_ih, wr_hh, br_ih, br_hh, wi_ih, wi_hh, bi_ih, bi_hh, w_bf16, train, state):
    """One complex LSTM layer on X [2, T, B, D] (fp32 contiguous weights) -> (Y [2,T,B,H], h, gates, c).
    state: None (zero initial state, the reference's default) or a dict with 'h' / 'c' tensors
    [2 sets][2*B rows][H] that initialise the recurrence and are UPDATED IN PLACE with the state after the last
    step (time-chunked streaming inference)."""
    P, T, B, D = X.shape
    H, G = plans.H, 4 * plans.H
    dev = X.device
    st = ops._stream()
    # biases of both sets -> [8H]
    bkey = (ops._wkey(br_ih, br_hh), ops._wkey(bi_ih, bi_hh))
    ent = plans.cache.get("bias8")
    if ent is not None and ent[0] == bkey:
        bias8 = ent[1]
    else:
        bias8 = torch.empty(2 * G, dtype=torch.float32, device=dev)
        call("clskd_pack_gather", br_ih.data_ptr(), br_hh.data_ptr(), plans.bias.data_ptr(), G,
             bias8.data_ptr(), 0, st)
        call("clskd_pack_gather", bi_ih.data_ptr(), bi_hh.data_ptr(), plans.bias.data_ptr(), G,
             bias8.data_ptr() + 4 * G, 0, st)
        plans.cache["bias8"] = (bkey, bias8)
    # input projections for both sets: rows (p,t,b) -> pre [2, T, B, 8H]
    pre = torch.empty((P, T, B, 2 * G), dtype=torch.float32, device=dev)
    run_tapconv(X.view(1, P * T, B, D), None, D, 0, 1, P * T, B, P * T, B, plans.ih.fwd[0], wr_ih, wi_ih,
                bias8, pre.view(1, P * T, B, 2 * G))
    whh_t = ops.packed_weights(plans.cache, "whh_t", lambda: plans.whh_t, wr_hh, wi_hh, torch.float32)
    h = torch.empty((2, P, T, B, H), dtype=torch.float32, device=dev)
    gates = torch.empty((2, P, T, B, G), dtype=torch.float32, device=dev) if train else None
    c = torch.empty((2, P, T, B, H), dtype=torch.float32, device=dev) if train else None
    if state is None:
        call("clskd_lstm_fwd", pre.data_ptr(), whh_t.data_ptr(), T, P * B, B, H, 2,
             T * B * 2 * G, B * 2 * G, 2 * G, G, H * G, 1 if w_bf16 else 0,
             h.data_ptr(), ops._ptr(gates), ops._ptr(c), st)
    else:
        hs, cs = state["h"], state["c"]
        assert hs.shape == (2, P * B, H) and cs.shape == (2, P * B, H) and hs.is_contiguous() and cs.is_contiguous()
        call("clskd_lstm_fwd_state", pre.data_ptr(), whh_t.data_ptr(), T, P * B, B, H, 2,
             T * B * 2 * G, B * 2 * G, 2 * G, G, H * G, 1 if w_bf16 else 0,
             h.data_ptr(), ops._ptr(gates), ops._ptr(c), hs.data_ptr(), cs.data_ptr(), hs.data_ptr(), cs.data_ptr(), st)
    # combine: real = rr - ii = h[set0, part0] - h[set1, part1]; imag = ir + ri = h[0,1] + h[1,0]
    Y = torch.empty((2, T, B, H), dtype=torch.float32, device=dev)
    n = T * B * H
    call("clskd_axpby_f32", h[0, 0].data_ptr(), h[1, 1].data_ptr(), 1.0, -1.0, Y[0].data_ptr(), n, st)
    call("clskd_axpby_f32", h[0, 1].data_ptr(), h[1, 0].data_ptr(), 1.0, 1.0, Y[1].data_ptr(), n, st)
    return Y, h, gates, c


class ComplexLSTMLayerFn(torch.autograd.Function):
    """X [2, T, B, D] -> Y [2, T, B, H] (Y[0] = rr - ii, Y[1] = ir + ri)."""

    @staticmethod
    def forward(ctx, plans: _ClstmPlans, X, wr_ih, wr_hh, br_ih, br_hh, wi_ih, wi_hh, bi_ih, bi_hh, w_bf16):
        f = ops._f32c
        wr_ih, wr_hh, br_ih, br_hh = f(wr_ih), f(wr_hh), f(br_ih), f(br_hh)
        wi_ih, wi_hh, bi_ih, bi_hh = f(wi_ih), f(wi_hh), f(bi_ih), f(bi_hh)
        train = any(ctx.needs_input_grad)   # False under torch.no_grad()
        Y, h, gates, c = _clstm_forward(plans, X, wr_ih, wr_hh, br_ih, br_hh, wi_ih, wi_hh, bi_ih, bi_hh, w_bf16,
                                        train, None)
        P, T, B, D = X.shape
        ctx.plans = plans
        ctx.save_for_backward(X, wr_ih, wi_ih, wr_hh, wi_hh, h, gates, c)
        ctx.dims = (P, T, B, D)
        ctx.w_bf16 = bool(w_bf16)
        return Y

    @staticmethod
    def backward(ctx, dY):
        plans = ctx.plans
        X, wr_ih, wi_ih, wr_hh, wi_hh, h, gates, c = ctx.saved_tensors
        if gates is None:
            raise RuntimeError("complex LSTM: backward requested but the forward did not save its gates")
        P, T, B, D = ctx.dims
        H, G = plans.H, 4 * plans.H
        dev = dY.device
        st = ops._stream()
        dY = dense(dY, torch.float32)
        n = T * B * H
        dh = torch.empty((2, P, T, B, H), dtype=torch.float32, device=dev)
        call("clskd_axpby_f32", dY[0].data_ptr(), None, 1.0, 0.0, dh[0, 0].data_ptr(), n, st)
        call("clskd_axpby_f32", dY[0].data_ptr(), None, -1.0, 0.0, dh[1, 1].data_ptr(), n, st)
        call("clskd_axpby_f32", dY[1].data_ptr(), None, 1.0, 0.0, dh[0, 1].data_ptr(), n, st)
        call("clskd_axpby_f32", dY[1].data_ptr(), None, 1.0, 0.0, dh[1, 0].data_ptr(), n, st)
        whh = pack_weights(plans.whh, wr_hh, wi_hh, torch.float32)
        dpre = torch.empty((P, T, B, 2 * G), dtype=torch.float32, device=dev)
        call("clskd_lstm_bwd_policy", dh.data_ptr(), whh.data_ptr(), gates.data_ptr(), c.data_ptr(), T, P * B, B, H, 2,
             G * H, T * B * 2 * G, B * 2 * G, 2 * G, G, dpre.data_ptr(), 1 if ctx.w_bf16 else 0, st)
        need = ctx.needs_input_grad
        dX = None
        if need[1]:
            dX = torch.empty(X.shape, dtype=X.dtype, device=dev)
            run_tapconv(dpre.view(1, P * T, B, 2 * G), None, 2 * G, 0, 1, P * T, B, P * T, B,
                        plans.ih.dgrad[0][0], wr_ih, wi_ih, None, dX.view(1, P * T, B, D))
        # tensor-core policy: the weight gradients contract bf16 copies of dpre / X / h on tcgen05 (fp32
        # accumulation over the T*B rows), like every other weight gradient of that policy
        dpre_w, X_w, h_w = dpre, X, h
        if ops.policy.use_umma and D % 16 == 0 and H % 16 == 0 and P * T * B >= 4096:
            dpre_w = dense(dpre, torch.bfloat16)
            X_w = dense(X, torch.bfloat16)
            h_w = dense(h, torch.bfloat16)
        # input-projection weight grads (both sets at once)
        dwcat = torch.empty(plans.ih.wcat, dtype=torch.float32, device=dev)
        run_wgrad(X_w.view(1, P * T, B, D), None, D, 0, 1, P * T, B, P * T, B, plans.ih.fwd[0],
                  dpre_w.view(1, P * T, B, 2 * G), dwcat)
        dwr_ih = unpack_grads(dwcat, plans.ih.unpack_a, G * D).view(G, D)
        dwi_ih = unpack_grads(dwcat, plans.ih.unpack_b, G * D).view(G, D)
        # recurrent weight grads: dW_hh[s][g][k] = sum_{p,t>=1,b} dpre[p,t,b,s*4H+g] * h[s,p,t-1,b,k]
        # the part axis is folded into "B" of the wgrad so the t-1 tap never crosses parts
        l = Launch(dt=[-1], df=[0], K=H, N=G)
        dw_hh = []
        for s in range(2):
            tmp = torch.empty(H * G, dtype=torch.float32, device=dev)
            run_wgrad(h_w[s], None, H, 0, P, T, B, T, B, l, dpre_w, tmp,
                      dy_view=(s * G, (T * B * 2 * G, B * 2 * G, 2 * G)))
            dw_hh.append(unpack_grads(tmp, plans.hh_unpack, G * H).view(G, H))
        # biases: column sums of dpre
        sb, _ = colstats(dpre.view(-1, 2 * G))
        sb = f64_to_f32(sb)
        dbr, dbi = sb[:G], sb[G:]
        return None, dX, dwr_ih, dw_hh[0], dbr, dbr, dwi_ih, dw_hh[1], dbi, dbi, None


class _LinearParams(nn.Module):
    """Parameter holder with nn.Linear's names and default init."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features))
        self.bias = nn.Parameter(torch.empty(out_features))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(in_features)
        nn.init.uniform_(self.bias, -bound, bound)
        self._plans = {}

    def plan(self):
        dev = self.weight.device
        if dev not in self._plans:
            code = _codes((self.out_features, self.in_features), 0).T[None, None]   # [1,1,K,N]
            j = np.arange(self.out_features)
            bt = np.stack([j * 4, np.full_like(j, -1)], 1).astype(np.int32)
            self._plans[dev] = ConvPlan("conv", code, 1, 0, 0, self.in_features, 0, bt,
                                        self.weight.numel(), 0, dev)
        return self._plans[dev]

    def forward_rows(self, x_rows, out_dtype=torch.float32):
        """x_rows dense [..., in] -> [..., out]"""
        shp = x_rows.shape
        M = x_rows.numel() // shp[-1]
        y = ops.TapConvFn.apply(self.plan(), x_rows.view(1, M, 1, shp[-1]), None, self.weight, None,
                                self.bias, None, out_dtype)
        return y.view(*shp[:-1], self.out_features)

    def forward(self, x):
        return self.forward_rows(dense(x))


class Stack2Fn(torch.autograd.Function):
    """[real, imag] (each [T, B, D], any strides) -> dense X [2, T, B, D] in `dtype`."""

    @staticmethod
    def forward(ctx, real, imag, dtype):
        X = torch.empty((2,) + tuple(real.shape), dtype=dtype, device=real.device)
        strided_copy_into(real, X[0])
        strided_copy_into(imag, X[1])
        ctx.dts = (real.dtype, imag.dtype)
        return X

    @staticmethod
    def backward(ctx, g):
        g = dense(g)
        gr, gi = g[0], g[1]
        if gr.dtype != ctx.dts[0]:
            gr = dense(gr, ctx.dts[0])
        if gi.dtype != ctx.dts[1]:
            gi = dense(gi, ctx.dts[1])
        return gr, gi, None


class NavieComplexLSTM(nn.Module):
    """Same constructor / parameter names as the reference; forward takes and returns
    [real, imag] lists of [T, B, D] tensors."""

    def __init__(self, input_size, hidden_size, projection_dim=None, bidirectional=False, batch_first=False):
        super().__init__()
        if bidirectional or batch_first:
            raise NotImplementedError("NavieComplexLSTM: bidirectional / batch_first are not implemented")
        self.input_dim = input_size // 2
        self.rnn_units = hidden_size // 2
        self.real_lstm = _LSTMParams(self.input_dim, self.rnn_units)
        self.imag_lstm = _LSTMParams(self.input_dim, self.rnn_units)
        if projection_dim is not None:
            self.projection_dim = projection_dim // 2
            self.r_trans = _LinearParams(self.rnn_units, self.projection_dim)
            self.i_trans = _LinearParams(self.rnn_units, self.projection_dim)
        else:
            self.projection_dim = None
        self._plans = {}

    def flatten_parameters(self):
        pass

    def _get_plans(self, device):
        if device not in self._plans:
            self._plans[device] = _ClstmPlans(self.input_dim, self.rnn_units, device)
        return self._plans[device]

    def forward_stacked(self, X):
        """X dense [2, T, B, D] -> Y dense [2, T, B, H or projection_dim] (fp32)"""
        r, i = self.real_lstm, self.imag_lstm
        Y = ComplexLSTMLayerFn.apply(self._get_plans(X.device), X, r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0,
                                     r.bias_hh_l0, i.weight_ih_l0, i.weight_hh_l0, i.bias_ih_l0, i.bias_hh_l0,
                                     ops.policy.name == "bf16")
        if self.projection_dim is not None:
            Y = _project2(self, Y)
        return Y

    def new_state(self, batch, device):
        """zero (h, c) of the four LSTM passes for `batch` utterances: [2 weight sets][2 parts * batch][H]"""
        z = lambda: torch.zeros((2, 2 * batch, self.rnn_units), dtype=torch.float32, device=device)
        return {"h": z(), "c": z()}

    def forward_stacked_state(self, X, state):
        """Inference with a carried recurrent state (no autograd): X dense [2, T, B, D] -> Y like forward_stacked;
        `state` (new_state) is advanced in place, so consecutive time chunks reproduce the whole-sequence output."""
        f = ops._f32c
        r, i = self.real_lstm, self.imag_lstm
        with torch.no_grad():
            Y, _, _, _ = _clstm_forward(self._get_plans(X.device), X, f(r.weight_ih_l0), f(r.weight_hh_l0),
                                        f(r.bias_ih_l0), f(r.bias_hh_l0), f(i.weight_ih_l0), f(i.weight_hh_l0),
                                        f(i.bias_ih_l0), f(i.bias_hh_l0), ops.policy.name == "bf16", False, state)
            if self.projection_dim is not None:
                Y = _project2(self, Y)
        return Y

    def forward(self, inputs):
        if isinstance(inputs, (list, tuple)):
            real, imag = inputs
        else:
            real, imag = torch.chunk(inputs, 2, -1)
        ops._require_cuda(real, imag)
        X = Stack2Fn.apply(real, imag, ops.policy.act_dtype)
        Y = self.forward_stacked(X)
        return [Y[0], Y[1]]


class _Project2Fn(torch.autograd.Function):
    """Y [2,T,B,H] -> [2,T,B,P]: r_trans on part 0, i_trans on part 1, written into one buffer."""

    @staticmethod
    def forward(ctx, Y, wr, br, wi, bi, plan_r, plan_i):
        _, T, B, H = Y.shape
        P = wr.shape[0]
        out = torch.empty((2, T, B, P), dtype=torch.float32, device=Y.device)
        f = ops._f32c
        for p, (w, b_, plan) in enumerate(((wr, br, plan_r), (wi, bi, plan_i))):
            bias = pack_weights(plan.bias_table, f(b_), None, torch.float32)
            run_tapconv(Y[p].view(1, T, B, H), None, H, 0, 1, T, B, T, B, plan.fwd[0], f(w), None, bias,
                        out[p].view(1, T, B, P))
        ctx.save_for_backward(Y, wr, wi)
        ctx.plans = (plan_r, plan_i)
        return out

    @staticmethod
    def backward(ctx, g):
        Y, wr, wi = ctx.saved_tensors
        plan_r, plan_i = ctx.plans
        _, T, B, H = Y.shape
        P = wr.shape[0]
        g = dense(g, torch.float32)
        f = ops._f32c
        dY = torch.empty_like(Y)
        grads = []
        Y_w, g_w = Y, g
        if ops.policy.use_umma and H % 16 == 0 and P % 16 == 0 and T * B >= 4096:
            Y_w, g_w = dense(Y, torch.bfloat16), dense(g, torch.bfloat16)     # tcgen05 weight gradient
        for p, (w, plan) in enumerate(((wr, plan_r), (wi, plan_i))):
            run_tapconv(g[p].view(1, T, B, P), None, P, 0, 1, T, B, T, B, plan.dgrad[0][0], f(w), None, None,
                        dY[p].view(1, T, B, H))
            dw = torch.empty(plan.wcat, dtype=torch.float32, device=g.device)
            run_wgrad(Y_w[p].view(1, T, B, H), None, H, 0, 1, T, B, T, B, plan.fwd[0], g_w[p].view(1, T, B, P), dw)
            s, _ = colstats(g[p].view(-1, P))
            grads.append((unpack_grads(dw, plan.unpack_a, plan.na).view(P, H), f64_to_f32(s)))
        return dY, grads[0][0], grads[0][1], grads[1][0], grads[1][1], None, None


def _project2(mod, Y):
    return _Project2Fn.apply(Y, mod.r_trans.weight, mod.r_trans.bias, mod.i_trans.weight, mod.i_trans.bias,
                             mod.r_trans.plan(), mod.i_trans.plan())


# ---------------------------------------------------------------------------------------------
# plain (real) LSTM bottleneck of DCCRN(use_clstm=False)  (reference: DCCRN.py:100-110, 193-199)
# ---------------------------------------------------------------------------------------------

class _LstmPlans:
    """Static tables of one real LSTM layer (one weight set, one 'part')."""

    def __init__(self, D, H, device):
        G = 4 * H
        self.D, self.H = D, H
        self.ih = ConvPlan("conv", _codes((G, D), 0).T[None, None], 1, 0, 0, D, 0, None, G * D, 0, device)
        hr = _codes((G, H), 0)
        self.whh_t = torch.from_numpy(ops._pairs(hr.T.reshape(-1))).to(device)     # [H][4H]
        self.whh = torch.from_numpy(ops._pairs(hr.reshape(-1))).to(device)         # [4H][H]
        j = np.arange(G)
        self.bias = torch.from_numpy(np.stack([j * 4, j * 4 + 1], 1).astype(np.int32)).to(device)
        t = np.full((G * H, 2), -1, dtype=np.int32)
        gg, kk = np.meshgrid(np.arange(G), np.arange(H), indexing="ij")
        t[:, 0] = ((kk * G + gg) * 2).reshape(-1)
        self.hh_unpack = torch.from_numpy(t).to(device)
        self.cache = {}


class LSTMLayerFn(torch.autograd.Function):
    """One unidirectional nn.LSTM layer, zero initial state: X [T, B, D] -> Y [T, B, H] (fp32).
    Same kernels as the complex LSTM with one weight set: batched input projection (tapconv / split-bf16
    tcgen05), persistent recurrence (clskd_lstm_fwd), BPTT (clskd_lstm_bwd), tapconv weight gradients."""

    @staticmethod
    def forward(ctx, plans: _LstmPlans, X, w_ih, w_hh, b_ih, b_hh, w_bf16):
        T, B, D = X.shape
        H, G = plans.H, 4 * plans.H
        dev = X.device
        st = ops._stream()
        f = ops._f32c
        w_ih, w_hh, b_ih, b_hh = f(w_ih), f(w_hh), f(b_ih), f(b_hh)
        bias = torch.empty(G, dtype=torch.float32, device=dev)
        call("clskd_pack_gather", b_ih.data_ptr(), b_hh.data_ptr(), plans.bias.data_ptr(), G, bias.data_ptr(), 0, st)
        pre = torch.empty((T, B, G), dtype=torch.float32, device=dev)
        run_tapconv(X.view(1, T, B, D), None, D, 0, 1, T, B, T, B, plans.ih.fwd[0], w_ih, None, bias,
                    pre.view(1, T, B, G))
        whh_t = ops.packed_weights(plans.cache, "whh_t", lambda: plans.whh_t, w_hh, None, torch.float32)
        train = any(ctx.needs_input_grad)
        h = torch.empty((T, B, H), dtype=torch.float32, device=dev)
        gates = torch.empty((T, B, G), dtype=torch.float32, device=dev) if train else None
        c = torch.empty((T, B, H), dtype=torch.float32, device=dev) if train else None
        call("clskd_lstm_fwd", pre.data_ptr(), whh_t.data_ptr(), T, B, B, H, 1, T * B * G, B * G, G, 0, H * G,
             1 if w_bf16 else 0, h.data_ptr(), ops._ptr(gates), ops._ptr(c), st)
        ctx.plans = plans
        ctx.w_bf16 = bool(w_bf16)
        ctx.save_for_backward(X, w_ih, w_hh, h, gates, c)
        return h

    @staticmethod
    def backward(ctx, dY):
        plans = ctx.plans
        X, w_ih, w_hh, h, gates, c = ctx.saved_tensors
        if gates is None:
            raise RuntimeError("LSTM: backward requested but the forward did not save its gates")
        T, B, D = X.shape
        H, G = plans.H, 4 * plans.H
        dev = dY.device
        st = ops._stream()
        dh = dense(dY, torch.float32)
        whh = pack_weights(plans.whh, w_hh, None, torch.float32)
        dpre = torch.empty((T, B, G), dtype=torch.float32, device=dev)
        call("clskd_lstm_bwd_policy", dh.data_ptr(), whh.data_ptr(), gates.data_ptr(), c.data_ptr(), T, B, B, H, 1,
             G * H, T * B * G, B * G, G, 0, dpre.data_ptr(), 1 if ctx.w_bf16 else 0, st)
        dX = None
        if ctx.needs_input_grad[1]:
            dX = torch.empty(X.shape, dtype=X.dtype, device=dev)
            run_tapconv(dpre.view(1, T, B, G), None, G, 0, 1, T, B, T, B, plans.ih.dgrad[0][0], w_ih, None, None,
                        dX.view(1, T, B, D))
        dpre_w, X_w, h_w = dpre, X, h
        if ops.policy.use_umma and D % 16 == 0 and H % 16 == 0 and T * B >= 4096:
            dpre_w, X_w, h_w = dense(dpre, torch.bfloat16), dense(X, torch.bfloat16), dense(h, torch.bfloat16)
        dwcat = torch.empty(plans.ih.wcat, dtype=torch.float32, device=dev)
        run_wgrad(X_w.view(1, T, B, D), None, D, 0, 1, T, B, T, B, plans.ih.fwd[0], dpre_w.view(1, T, B, G), dwcat)
        dw_ih = unpack_grads(dwcat, plans.ih.unpack_a, G * D).view(G, D)
        # dW_hh[g][k] = sum_{t>=1,b} dpre[t,b,g] * h[t-1,b,k]
        l = Launch(dt=[-1], df=[0], K=H, N=G)
        tmp = torch.empty(H * G, dtype=torch.float32, device=dev)
        run_wgrad(h_w.view(1, T, B, H), None, H, 0, 1, T, B, T, B, l, dpre_w.view(1, T, B, G), tmp)
        dw_hh = unpack_grads(tmp, plans.hh_unpack, G * H).view(G, H)
        sb, _ = colstats(dpre.view(-1, G))
        db = f64_to_f32(sb)
        return None, dX, dw_ih, dw_hh, db, db, None


class LSTM(nn.Module):
    """nn.LSTM(input_size, hidden_size, num_layers, bidirectional=False, batch_first=False) with the same
    parameter names (`weight_ih_l{k}`, `weight_hh_l{k}`, `bias_ih_l{k}`, `bias_hh_l{k}`); forward(x [T,B,D])
    returns (output [T,B,H], (h_n, c_n)): h_n [num_layers,B,H] like nn.LSTM; c_n = None (the recurrence kernel
    keeps the cell state in registers and only stores it for training; no caller of the path consumes it)."""

    def __init__(self, input_size, hidden_size, num_layers=1, dropout=0.0, bidirectional=False, batch_first=False):
        super().__init__()
        if bidirectional or batch_first or dropout:
            raise NotImplementedError("LSTM: bidirectional / batch_first / dropout are not implemented")
        self.input_size, self.hidden_size, self.num_layers = input_size, hidden_size, num_layers
        k = 1.0 / math.sqrt(hidden_size)
        for l in range(num_layers):
            d = input_size if l == 0 else hidden_size
            setattr(self, "weight_ih_l%d" % l, nn.Parameter(torch.empty(4 * hidden_size, d).uniform_(-k, k)))
            setattr(self, "weight_hh_l%d" % l, nn.Parameter(torch.empty(4 * hidden_size, hidden_size).uniform_(-k, k)))
            setattr(self, "bias_ih_l%d" % l, nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k)))
            setattr(self, "bias_hh_l%d" % l, nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k)))
        self._plans = {}

    def flatten_parameters(self):
        pass

    def forward(self, x, hx=None):
        if hx is not None:
            raise NotImplementedError("LSTM: only the zero initial state is implemented")
        ops._require_cuda(x)
        y = dense(x, ops.policy.act_dtype)
        last = []
        for l in range(self.num_layers):
            key = (l, y.device)
            if key not in self._plans:
                self._plans[key] = _LstmPlans(y.shape[-1], self.hidden_size, y.device)
            y = LSTMLayerFn.apply(self._plans[key], y, getattr(self, "weight_ih_l%d" % l),
                                  getattr(self, "weight_hh_l%d" % l), getattr(self, "bias_ih_l%d" % l),
                                  getattr(self, "bias_hh_l%d" % l), ops.policy.name == "bf16")
            last.append(y[-1])
        return y, (torch.stack(last, 0), None)
