"""Host-side plumbing between torch tensors and the C ABI: precision policy, tensor-layout
helpers, the tap-list convolution plans and every autograd Function of the path.

Internal ("physical") activation layout is channels-last [B, T, F, C]; the reference's logical
NCHW tensors [B, C, F, T] are exposed as `.permute(0, 3, 2, 1)` views of it (no copies).
"""
import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import TapConv, call

# --------------------------------------------------------------------------------------------
# precision policy
# --------------------------------------------------------------------------------------------


class _Policy:
    """fp32: fp32 activations, fp32 FMA contractions (exact-parity policy, <=1e-5 waveform error).
    bf16: bf16 activations/weights on the tcgen05 tensor cores with fp32 accumulation; STFT/iSTFT,
    BatchNorm statistics, LSTM state, mask math and all loss reductions stay fp32/fp64."""

    def __init__(self):
        self.name = "fp32"
        self.act_dtype = torch.float32
        self.use_umma = False
        self.abf_rank2 = True       # ABF level whose 1x1 conv has 2 input channels: z1 recomputed in the mid kernels
        self.fuse_epilogue = True   # BatchNorm statistics / folded eval BatchNorm + PReLU in the tcgen05 conv epilogue
        self.abf_xs2 = True      # rank-2 folded kernels for the 2-channel ABF level (clskd_abf_xs2_*)
        # ABF conv1 + middle stage as one node, BatchNorm backward folded into conv1's gradients (CLSKD_ABF_FOLD=0: A/B)
        self.abf_fold = os.environ.get("CLSKD_ABF_FOLD", "1") != "0"
        self.split_gemm = True   # fp32-input GEMMs (STFT/iSTFT/LSTM projections) as split-bf16 tcgen05 contractions
        self.narrow = "auto"     # tap-in-channel decomposition of narrow convs: "auto" (tensor-core policy) / "always"


policy = _Policy()


def set_precision(name: str):
    if name not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    policy.name = name
    policy.act_dtype = torch.float32 if name == "fp32" else torch.bfloat16
    policy.use_umma = name == "bf16"


def get_precision() -> str:
    return policy.name


def _tag(dtype):
    if dtype == torch.float32:
        return _lib.F32
    if dtype == torch.bfloat16:
        return _lib.BF16
    raise RuntimeError("clskd_b200: unsupported dtype %s (fp32 / bf16 only)" % dtype)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("clskd_b200 runs on CUDA tensors only (no CPU fallback); got a %s tensor"
                               % t.device)


def _ptr(t, offset=0):
    if t is None:
        return None
    return t.data_ptr() + offset * t.element_size()


def _f32c(t):
    """fp32 contiguous version of a (small) parameter tensor without launching kernels if possible."""
    if t.dtype == torch.float32 and t.is_contiguous():
        return t
    return t.detach().float().contiguous()


# --------------------------------------------------------------------------------------------
# layout helpers
# --------------------------------------------------------------------------------------------

def _pad4(v, fill):
    v = list(v)
    while len(v) < 4:
        v.insert(0, fill)
    if len(v) > 4:
        raise RuntimeError("strided copy supports up to 4 dims")
    return v


def strided_copy_into(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """dst[...] = src[...] for two same-shape (<=4-D) strided views; dtype conversion allowed."""
    _require_cuda(src, dst)
    if tuple(src.shape) != tuple(dst.shape):
        raise RuntimeError("strided_copy_into: shape mismatch %s vs %s" % (tuple(src.shape), tuple(dst.shape)))
    shape = _pad4(src.shape, 1)
    call("clskd_strided_copy4d", src.data_ptr(), _tag(src.dtype), (ctypes.c_int64 * 4)(*_pad4(src.stride(), 0)),
         dst.data_ptr(), _tag(dst.dtype), (ctypes.c_int64 * 4)(*_pad4(dst.stride(), 0)),
         (ctypes.c_int64 * 4)(*shape), _stream())
    return dst


def strided_copy(src: torch.Tensor, dtype=None) -> torch.Tensor:
    """Dense copy of an arbitrary <=4-D strided view, with optional dtype conversion."""
    dst = torch.empty(src.shape, dtype=dtype or src.dtype, device=src.device)
    if src.dim() > 4 and src.is_contiguous():          # pure dtype conversion of a dense tensor
        strided_copy_into(src.view(-1, src.shape[-1]), dst.view(-1, src.shape[-1]))
        return dst
    return strided_copy_into(src, dst)


class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.in_dtype = x.dtype
        return strided_copy(x, dtype)

    @staticmethod
    def backward(ctx, g):
        g = dense(g, ctx.in_dtype) if (g.dtype != ctx.in_dtype or not g.is_contiguous()) else g
        return g, None


def dense(x: torch.Tensor, dtype=None) -> torch.Tensor:
    """x as a contiguous tensor of `dtype` (own copy kernel; identity when already so)."""
    dtype = dtype or x.dtype
    if x.dtype == dtype and x.is_contiguous():
        return x
    if x.requires_grad and torch.is_grad_enabled():
        return _DenseFn.apply(x, dtype)
    return strided_copy(x, dtype)


class FanoutFn(torch.autograd.Function):
    """k aliases of a tensor that has k consumers (skip connection + next layer + feature tap ...): the
    gradients of the aliases are summed by ONE library kernel (clskd_sum_n, fp32 accumulation) instead of
    autograd's chain of add kernels."""

    @staticmethod
    def forward(ctx, x, k):
        ctx.set_materialize_grads(False)
        return tuple(x.view_as(x) for _ in range(k))

    @staticmethod
    def backward(ctx, *gs):
        gs = [g for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        dt = gs[0].dtype
        base = gs[0]
        # sum in the memory order of the first gradient (they are all gradients of the same tensor, so they
        # normally share it); anything else is made dense in that order first
        ds = []
        for g in gs:
            if g.dtype != dt or g.stride() != base.stride() or not _is_dense(g):
                g = _dense_like(g, base, dt)
            ds.append(g)
        if not _is_dense(base):
            ds = [_dense_like(g, None, dt) for g in gs]
        out = torch.empty_like(ds[0])
        while len(ds) > 4:                     # more than four consumers: fold four at a time
            part = torch.empty_like(ds[0])
            call("clskd_sum_n", ds[0].data_ptr(), ds[1].data_ptr(), ds[2].data_ptr(), ds[3].data_ptr(), 4, _tag(dt),
                 ds[0].numel(), part.data_ptr(), _stream())
            ds = [part] + ds[4:]
        p = [d_.data_ptr() for d_ in ds] + [None] * (4 - len(ds))
        call("clskd_sum_n", p[0], p[1], p[2], p[3], len(ds), _tag(dt), ds[0].numel(), out.data_ptr(), _stream())
        return out, None


def _is_dense(t):
    """occupies numel contiguous elements in SOME dimension order (e.g. a permuted view of a contiguous tensor)"""
    if t.is_contiguous():
        return True
    if t.numel() == 0:
        return True
    order = sorted(range(t.dim()), key=lambda i: (-t.stride(i), -t.shape[i]))
    exp = 1
    for i in reversed(order):
        if t.shape[i] != 1 and t.stride(i) != exp:
            return False
        exp *= t.shape[i]
    return True


def _dense_like(g, base, dtype):
    """g re-laid out with base's strides (or contiguous when base is None / not dense) in `dtype`"""
    if base is None or not _is_dense(base):
        return dense(g, dtype)
    out = torch.empty_strided(base.shape, base.stride(), dtype=dtype, device=g.device)
    if g.dim() <= 4:
        strided_copy_into(g, out)
    else:
        out.copy_(g)
    return out


def fanout(x, k=2):
    """k aliases of x whose gradients are summed by the library (no-op outside autograd)"""
    if k < 2 or not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * k
    return FanoutFn.apply(x, k)


def to_phys(x_logical: torch.Tensor, dtype=None, need_dense=False) -> torch.Tensor:
    """logical [B, C, F, T] -> physical [B, T, F, C] with unit channel stride (a view whenever the
    logical tensor is itself a view of a physical one; copied by the layout kernel otherwise)."""
    p = x_logical.permute(0, 3, 2, 1)
    dtype = dtype or p.dtype
    if p.dtype == dtype and not need_dense and (p.shape[3] == 1 or p.stride(3) == 1) and \
            (p.shape[2] == 1 or p.stride(2) == p.shape[3]):
        return p
    return dense(p, dtype)


def to_logical(x_phys: torch.Tensor) -> torch.Tensor:
    return x_phys.permute(0, 3, 2, 1)


# --------------------------------------------------------------------------------------------
# tap-list convolution plans
# --------------------------------------------------------------------------------------------

@dataclass
class Launch:
    """One tapconv launch: taps, input stride, and where its outputs go (sub-pixel phase)."""
    dt: List[int]
    df: List[int]
    sf: int = 1          # input-f step per output row
    osf: int = 1         # output f = osf*m + ooff
    ooff: int = 0
    blk: Optional[np.ndarray] = None      # int64 codes [ntaps][K][N] (host)
    device: Optional[torch.device] = None
    _t_cn: Optional[torch.Tensor] = None
    _t_nc: Optional[torch.Tensor] = None
    K: int = 0           # contraction channels (c)
    N: int = 0           # output channels (n)
    c_lo: int = 0        # for dgrad launches: slice [c_lo, c_lo+N) of the input channels
    woff: int = 0        # offset of this launch's dW block in the concatenated wgrad buffer
    _cache: dict = field(default_factory=dict)   # packed kernel-side weights keyed by parameter version

    # pack tables (int32 pairs on the device), built on first use: [tap][c][n] and [tap][n][c]
    @property
    def t_cn(self):
        if self._t_cn is None:
            self._t_cn = torch.from_numpy(_pairs(self.blk.reshape(-1))).to(self.device)
        return self._t_cn

    @property
    def t_nc(self):
        if self._t_nc is None:
            self._t_nc = torch.from_numpy(_pairs(self.blk.transpose(0, 2, 1).reshape(-1))).to(self.device)
        return self._t_nc

    def t_nc_padded(self, c0, c1, c1p=None):
        """[tap][Np][c0p + c1p] table for the tcgen05 kernel when a channel count is a multiple of 8 but not of 16
        (quarter-width student): each source's K range and N are padded to 16 with skipped (= zero) entries; the
        kernel's TMA boxes are 16 channels wide over the 8 that exist (clskd_tapconv_fwd_umma)."""
        p16 = lambda v: (v + 15) // 16 * 16
        c1p = p16(c1) if c1p is None else c1p      # clskd_tapconv_umma_c1p: a narrow second source fills one K chunk
        key = ("t_nc_pad", c0, c1, c1p)
        if key not in self._cache:
            nt, K, N = self.blk.shape
            assert K == c0 + c1
            out = np.full((nt, p16(c0) + c1p, p16(N)), -1, dtype=np.int64)
            out[:, :c0, :N] = self.blk[:, :c0, :]
            if c1:
                out[:, p16(c0):p16(c0) + c1, :N] = self.blk[:, c0:, :]
            self._cache[key] = torch.from_numpy(_pairs(out.transpose(0, 2, 1).reshape(-1))).to(self.device)
        return self._cache[key]


def _codes(shape, sel):
    n = int(np.prod(shape))
    return (np.arange(n, dtype=np.int64) * 4 + sel).reshape(shape)


def _neg(code):
    return code + 2


def _pairs(code_flat):
    """int32 pair table from a flat array of single codes."""
    t = np.full((code_flat.size, 2), -1, dtype=np.int32)
    t[:, 0] = code_flat.astype(np.int32)
    return t


class ConvPlan:
    """Everything static about one convolution-like layer: forward / data-gradient / weight-gradient
    launches and the gather tables that build packed weights from (and fold gradients back onto)
    the reference's parameter tensors.

    `block` is an int64 code array [KF, KT, Ctot, N] (see clskd_pack_gather) describing the dense
    real-valued weight of the layer in terms of the parameter tensors a (sel 0) and b (sel 1).
    kind 'conv': Y[t,f] = sum X[t - pt + kt, f*sf - pf + kf]      (strided convolution)
    kind 'deconv': Y[ti - pt + kt, fi*sf - pf + kf] += X[ti, fi]  (transposed convolution)
    """

    def __init__(self, kind, block, sf, pf, pt, c0, c1, bias_table, na, nb, device,
                 t_extra=0, f_extra=0):
        self.kind, self.sf, self.pf, self.pt = kind, sf, pf, pt
        KF, KT, Ctot, N = block.shape
        self.KF, self.KT, self.Ctot, self.N = KF, KT, Ctot, N
        self.c0, self.c1 = c0, c1
        assert c0 + c1 == Ctot
        self.t_extra, self.f_extra = t_extra, f_extra   # output-size corrections (see out_size)
        self.na, self.nb = na, nb                       # numel of parameter a / b
        self.device = device
        self.fwd: List[Launch] = []
        self.dgrad: List[List[Launch]] = [[], []]       # per source
        taps = [(kf, kt) for kf in range(KF) for kt in range(KT)]

        def mk(sel_taps, dts, dfs, sf_, osf, ooff, transpose=False, c_lo=0, c_n=None):
            blk = np.stack([block[kf, kt] for kf, kt in sel_taps], 0)      # [ntaps, Ctot, N]
            if c_n is not None:
                blk = blk[:, c_lo:c_lo + c_n, :]
            if transpose:                                                   # dgrad: contract n
                blk = blk.transpose(0, 2, 1)
            return Launch(dt=dts, df=dfs, sf=sf_, osf=osf, ooff=ooff, K=blk.shape[1], N=blk.shape[2],
                          c_lo=c_lo, blk=blk, device=device)

        if kind == "conv":
            self.fwd.append(mk(taps, [kt - pt for _, kt in taps], [kf - pf for kf, _ in taps], sf, 1, 0))
            for src, (lo, cn) in enumerate(((0, c0), (c0, c1))):
                if cn == 0:
                    continue
                for p in range(sf):
                    sel = [(kf, kt) for kf, kt in taps if (kf - pf - p) % sf == 0]
                    if not sel:
                        continue
                    self.dgrad[src].append(mk(sel, [-(kt - pt) for _, kt in sel],
                                              [(p - (kf - pf)) // sf for kf, _ in sel], 1, sf, p,
                                              transpose=True, c_lo=lo, c_n=cn))
        elif kind == "deconv":
            for p in range(sf):
                sel = [(kf, kt) for kf, kt in taps if (p + pf - kf) % sf == 0]
                if not sel:
                    continue
                self.fwd.append(mk(sel, [pt - kt for _, kt in sel], [(p + pf - kf) // sf for kf, _ in sel],
                                   1, sf, p))
            for src, (lo, cn) in enumerate(((0, c0), (c0, c1))):
                if cn == 0:
                    continue
                self.dgrad[src].append(mk(taps, [kt - pt for _, kt in taps], [kf - pf for kf, _ in taps],
                                          sf, 1, 0, transpose=True, c_lo=lo, c_n=cn))
        else:
            raise ValueError(kind)

        # ---- wgrad: concatenated dW buffers of the forward launches -> parameters a / b
        off = 0
        for l in self.fwd:
            l.woff = off
            off += l.blk.size
        self.wcat = off
        self._unpack = [None, None]
        self._cache = {}
        # ---- bias: bias_table is an int32 pair table [N,2] over (bias_a, bias_b) or None
        self.bias_table = None
        self.bias_unpack_a = self.bias_unpack_b = None
        if bias_table is not None:
            self.bias_table = torch.from_numpy(bias_table.astype(np.int32)).to(device)
            nba = nbb = 0
            ents = {0: {}, 1: {}}
            for i in range(bias_table.shape[0]):
                for e in bias_table[i]:
                    if e >= 0:
                        ents[e & 1].setdefault(e >> 2, []).append(i * 2 + ((e >> 1) & 1))
            for sel in (0, 1):
                if ents[sel]:
                    n = max(ents[sel]) + 1
                    t = np.full((n, 2), -1, dtype=np.int32)
                    for j, es in ents[sel].items():
                        for k, e in enumerate(es[:2]):
                            t[j, k] = e
                    tt = torch.from_numpy(t).to(device)
                    if sel == 0:
                        self.bias_unpack_a = tt
                    else:
                        self.bias_unpack_b = tt

    def _unpack_table(self, sel):
        """[n_param, 2] int32 table: where each parameter element sits in the concatenated dW buffer."""
        if self._unpack[sel] is None:
            n = self.na if sel == 0 else self.nb
            idx_all, val_all = [], []
            for l in self.fwd:
                codes = l.blk.reshape(-1)
                pos = np.nonzero((codes >= 0) & ((codes & 1) == sel))[0]
                e = codes[pos]
                idx_all.append(e >> 2)
                val_all.append((l.woff + pos) * 2 + ((e >> 1) & 1))
            idx = np.concatenate(idx_all)
            val = np.concatenate(val_all)
            order = np.argsort(idx, kind="stable")
            idx, val = idx[order], val[order]
            rank = np.arange(idx.size) - np.searchsorted(idx, idx, side="left")
            assert rank.size == 0 or rank.max() <= 1, "a parameter element may appear at most twice"
            t = np.full((n, 2), -1, dtype=np.int32)
            t[idx, rank] = val.astype(np.int32)
            self._unpack[sel] = torch.from_numpy(t).to(self.device)
        return self._unpack[sel]

    @property
    def unpack_a(self):
        return self._unpack_table(0)

    @property
    def unpack_b(self):
        return self._unpack_table(1) if self.nb else None

    # output extents for an input of (Ti, Fi)
    def out_size(self, Ti, Fi):
        if self.kind == "conv":
            To = Ti + self.t_extra - (self.KT - 1)
            Fo = (Fi + 2 * self.pf - self.KF) // self.sf + 1
        else:
            To = (Ti - 1) - 2 * self.pt + self.KT + self.t_extra
            Fo = (Fi - 1) * self.sf - 2 * self.pf + self.KF + self.f_extra
        return To, Fo


weights_epoch = 0   # bumped by anything that rewrites parameters behind autograd's back (FlatAdam's kernel)


def invalidate_weight_cache():
    """Packed (kernel-layout) weights are cached per parameter version; FlatAdam calls this after its
    kernel rewrote the flat parameter bucket (a registered volatile range) behind autograd's back."""
    global weights_epoch
    weights_epoch += 1


_volatile_ranges = []   # [lo, hi) device-address ranges rewritten by kernels outside autograd (FlatAdam buckets)


def register_volatile_range(lo, hi):
    _volatile_ranges.append((lo, hi))


def _wkey(a, b):
    pa = a.data_ptr()
    pb = b.data_ptr() if b is not None else 0
    vol = any(lo <= pa < hi or lo <= pb < hi for lo, hi in _volatile_ranges) if _volatile_ranges else False
    return (pa, a._version, pb, b._version if b is not None else 0, weights_epoch if vol else -1)


# Packed weights of parameters that live in a volatile range (FlatAdam's bucket) are registered here, so that after
# the optimizer step ONE launch (clskd_multi_pack_gather) rebuilds all of those the step actually used, in place,
# instead of one clskd_pack_gather launch per weight at its first use in the next step.
_pack_registry = {}          # (id(cache), name, dtype) -> _PackEntry
_pack_desc = None            # (signature, device descriptor tensor, total)


class _PackEntry:
    __slots__ = ("cache", "ckey", "table", "a", "b", "dtype", "out", "used")

    def __init__(self, cache, ckey, table, a, b, dtype, out):
        import weakref
        self.cache, self.ckey, self.table, self.dtype, self.out = cache, ckey, table, dtype, out
        self.a = weakref.ref(a)
        self.b = weakref.ref(b) if b is not None else None
        self.used = True


def packed_weights(cache, name, table_fn, a, b, dtype):
    """pack_weights(table, a, b, dtype), memoised in `cache` until a / b change."""
    key = _wkey(a, b)
    ent = cache.get((name, dtype))
    if ent is not None and ent[0] == key:
        reg = _pack_registry.get((id(cache), name, dtype))
        if reg is not None:
            reg.used = True
        return ent[1]
    table = table_fn()
    w = pack_weights(table, a, b, dtype)
    cache[(name, dtype)] = (key, w)
    if key[4] >= 0 and a.is_cuda:          # volatile parameters: candidate for the batched re-pack
        _pack_registry[(id(cache), name, dtype)] = _PackEntry(cache, (name, dtype), table, a, b, dtype, w)
    return w


def repack_registered():
    """Rebuild, in ONE launch, every registered packed weight that was used since the last call (FlatAdam calls this
    right after its kernel rewrote the parameters).  Entries whose parameters died, moved or were re-created drop out."""
    global _pack_desc
    live = []
    for rk in list(_pack_registry.keys()):
        e = _pack_registry[rk]
        a = e.a()
        b = e.b() if e.b is not None else None
        cur = e.cache.get(e.ckey)
        if a is None or (e.b is not None and b is None) or cur is None or cur[1] is not e.out or not e.used:
            del _pack_registry[rk]
            continue
        live.append((e, a, b))
    if not live:
        _pack_desc = None
        return 0
    sig = tuple((a.data_ptr(), b.data_ptr() if b is not None else 0, e.table.data_ptr(), e.out.data_ptr())
                for e, a, b in live)
    if _pack_desc is None or _pack_desc[0] != sig:
        rows, start = [], 0
        for e, a, b in live:
            n = e.table.shape[0]
            rows.append([a.data_ptr(), b.data_ptr() if b is not None else a.data_ptr(), e.table.data_ptr(),
                         e.out.data_ptr(), start, 2 * n + (1 if e.dtype == torch.bfloat16 else 0)])
            start += n
        _pack_desc = (sig, torch.tensor(rows, dtype=torch.int64).to(live[0][0].out.device), start)
    call("clskd_multi_pack_gather", _pack_desc[1].data_ptr(), len(live), _pack_desc[2], _stream())
    for e, a, b in live:
        e.cache[e.ckey] = (_wkey(a, b), e.out)
        e.used = False
    return len(live)


def pack_weights(table, a, b, dtype):
    n = table.shape[0]
    out = torch.empty(n, dtype=dtype, device=table.device)
    call("clskd_pack_gather", a.data_ptr(), b.data_ptr() if b is not None else None,
         table.data_ptr(), n, out.data_ptr(), _tag(dtype), _stream())
    return out


def unpack_grads(src, table2, n):
    dst = torch.empty(n, dtype=torch.float32, device=src.device)
    call("clskd_unpack_gather2", src.data_ptr(), table2.data_ptr(), n, dst.data_ptr(), 0, _stream())
    return dst


def _fill_desc(d: TapConv, x0, x0_off, x0_str, x1, x1_off, x1_str, c0, c1, B, To, Fo, Ti, Fi, l: Launch,
               w, bias, N, y, y_off, y_str, accumulate=False):
    d.x0 = _ptr(x0, x0_off)
    d.x1 = _ptr(x1, x1_off) if x1 is not None else None
    d.x0_sB, d.x0_sT, d.x0_sF = x0_str
    if x1 is not None:
        d.x1_sB, d.x1_sT, d.x1_sF = x1_str
    d.c0, d.c1 = c0, c1
    d.B, d.To, d.Fo, d.Ti, d.Fi = B, To, Fo, Ti, Fi
    d.sf = l.sf
    d.ntaps = len(l.dt)
    for i, (a, b_) in enumerate(zip(l.dt, l.df)):
        d.dt[i] = a
        d.df[i] = b_
    d.w = w.data_ptr()
    d.bias = bias.data_ptr() if bias is not None else None
    d.N = N
    d.y = _ptr(y, y_off)
    d.y_sB, d.y_sT, d.y_sF = y_str
    d.x_dtype = _tag(x0.dtype)
    d.y_dtype = _tag(y.dtype)
    d.accumulate = 1 if accumulate else 0


def _umma_ok(d: TapConv) -> bool:
    if not policy.use_umma:
        return False
    return bool(_lib.load().clskd_tapconv_umma_supported(ctypes.byref(d)))


umma_launches = 0
core_launches = 0


class Epilogue:
    """Request for the fused epilogue of the tcgen05 forward kernel (ClskdTapConv.ep_* / stats_*):
    `scale`/`shift` [N] fp32 (eval BatchNorm folded), `slope` (PReLU weight, 1 element) and/or
    `stats` fp64 [2, N] (zeroed; receives column sums / sums of squares of the stored output = the
    batch statistics of the BatchNorm that follows).  `fused` says whether the launch honoured it."""

    def __init__(self, scale=None, shift=None, slope=None, stats=None):
        self.scale, self.shift, self.slope, self.stats = scale, shift, slope, stats
        self.fused = False


_pending_ep = None
fused_epilogues = 0      # number of convolutions whose BatchNorm work ran in the tcgen05 epilogue


def request_epilogue(ep):
    """The next TapConvFn.forward consumes this request (set by ConvBNAct right before the conv)."""
    global _pending_ep
    _pending_ep = ep


def _take_epilogue():
    global _pending_ep
    ep, _pending_ep = _pending_ep, None
    return ep


def _view_of(t):
    """(elem_offset, (sB, sT, sF)) of a 4-D [B,T,F,C] tensor whose channel stride is 1."""
    if t.dim() != 4 or (t.shape[3] > 1 and t.stride(3) != 1):
        raise RuntimeError("tapconv operand must be a 4-D [B,T,F,C] tensor with unit channel stride")
    return 0, (t.stride(0), t.stride(1), t.stride(2))


def _round_up(v, m):
    return (v + m - 1) // m * m


def _split_gemm(d: TapConv, l: Launch, a, b, bias, x0, y) -> bool:
    """fp32-input pointwise contraction (STFT / iSTFT DFT GEMMs, fp32 LSTM and Linear projections) on
    the tcgen05 kernel with split-bf16 operands: x = hi + lo and w = hi + lo are staged as
    [x_hi | x_lo | x_hi] x [w_hi ; w_hi ; w_lo] so that ONE bf16 contraction over 3K with fp32 TMEM
    accumulation gives x_hi w_hi + x_lo w_hi + x_hi w_lo (~1e-5 relative; the fp32 policy keeps the
    fp32 FMA kernel).  Returns False when the launch is not such a GEMM (caller falls through)."""
    global umma_launches
    if d.x_dtype != _lib.F32 or d.ntaps != 1 or d.dt[0] or d.df[0] or d.sf != 1 \
            or d.c1 or d.accumulate or d.Ti != d.To or d.Fi != d.Fo:
        return False
    K, N = d.c0, d.N
    M = d.B * d.To * d.Fo
    if M < 1024 or K < 32 or N < 32:          # tiny / skinny problems stay on the FMA / GEMV kernels
        return False
    Kp = _round_up(K, 64)
    y_bf16 = d.y_dtype == _lib.BF16
    ye = 2 if y_bf16 else 4
    Np = _round_up(N, 16) if N <= (256 if y_bf16 else 128) else _round_up(N, 128)
    dev = x0.device
    # weights: fp32 [K][N] (per-parameter-version cache) -> bf16 [Np][3Kp] = [hi | hi | lo]
    w32 = packed_weights(l._cache, "cn", lambda: l.t_cn, a, b, torch.float32)
    ent = l._cache.get("split_nc")
    if ent is not None and ent[0] is w32:
        ws = ent[1]
    else:
        ws = torch.empty((Np, 3 * Kp), dtype=torch.bfloat16, device=dev)
        call("clskd_split_bf16x3", w32.data_ptr(), 0, 0, 1, N, 1, 1, N, K, Kp, Np, 1, 3, ws.data_ptr(), _stream())
        l._cache["split_nc"] = (w32, ws)
    # activations: rows (b, t, f) gathered by strides -> bf16 [M][2Kp] = [hi | lo]; the third K segment
    # re-reads the hi half through the second tensor map
    xs = torch.empty((M, 2 * Kp), dtype=torch.bfloat16, device=dev)
    call("clskd_split_bf16x3", d.x0, d.x0_sB, d.x0_sT, d.x0_sF, 1, d.B, d.To, d.Fo, K, Kp, M, 0, 2,
         xs.data_ptr(), _stream())
    # output: straight into y when its rows are uniformly strided and TMA-storable, else via scratch
    dims = [(n, st) for n, st in ((d.B, d.y_sB), (d.To, d.y_sT), (d.Fo, d.y_sF)) if n > 1]
    rows_uniform = all(o[1] == i[0] * i[1] for o, i in zip(dims[:-1], dims[1:]))
    ld = dims[-1][1] if dims else N
    direct = Np == N and rows_uniform and (ld * ye) % 16 == 0 and d.y % 16 == 0
    bias_p = bias
    if bias is not None and Np != N:
        bias_p = torch.zeros(Np, dtype=torch.float32, device=dev)
        strided_copy_into(bias.view(-1)[:N], bias_p[:N])
    if direct:
        yt, y_ptr, y_ld = None, d.y, ld
    else:
        yt = torch.empty((M, Np), dtype=y.dtype, device=dev)
        y_ptr, y_ld = yt.data_ptr(), Np
    g = TapConv()
    g.x0, g.x1 = xs.data_ptr(), xs.data_ptr()
    g.x0_sB = g.x1_sB = M * 2 * Kp
    g.x0_sT = g.x1_sT = 2 * Kp
    g.x0_sF = g.x1_sF = 2 * Kp
    g.c0, g.c1 = 2 * Kp, Kp
    g.B, g.To, g.Fo, g.Ti, g.Fi = 1, M, 1, M, 1
    g.sf, g.ntaps = 1, 1
    g.dt[0] = g.df[0] = 0
    g.w = ws.data_ptr()
    g.bias = bias_p.data_ptr() if bias_p is not None else None
    g.N = Np
    g.y = y_ptr
    g.y_sB, g.y_sT, g.y_sF = M * y_ld, y_ld, y_ld
    g.x_dtype, g.y_dtype, g.accumulate = _lib.BF16, d.y_dtype, 0
    if not _lib.load().clskd_tapconv_umma_supported(ctypes.byref(g)):
        return False
    call("clskd_tapconv_fwd_umma", ctypes.byref(g), _stream())
    umma_launches += 1
    if yt is not None:
        shape = (d.B, d.To, d.Fo, N)
        src = yt.as_strided(shape, (d.To * d.Fo * Np, d.Fo * Np, Np, 1))
        dst = y.as_strided(shape, (d.y_sB, d.y_sT, d.y_sF, 1),
                           y.storage_offset() + (d.y - y.data_ptr()) // ye)
        strided_copy_into(src, dst)
    return True


def run_tapconv(x0, x1, c0, c1, B, To, Fo, Ti, Fi, l: Launch, a, b, bias, y, x0_view=None, x1_view=None,
                y_view=None, accumulate=False, allow_umma=True, ep=None):
    """Run one launch.  *_view = (elem_offset, (sB, sT, sF)) override the tensors' own strides.
    ep: Epilogue request; returns None WITHOUT launching when the request cannot be fused (the caller
    then runs the plain contraction and the separate BatchNorm kernels)."""
    global umma_launches, core_launches
    x0_off, x0_str = x0_view if x0_view is not None else _view_of(x0)
    if x1 is not None:
        x1_off, x1_str = x1_view if x1_view is not None else _view_of(x1)
    else:
        x1_off, x1_str = 0, (0, 0, 0)
    y_off, y_str = y_view if y_view is not None else _view_of(y)
    d = TapConv()
    _fill_desc(d, x0, x0_off, x0_str, x1, x1_off, x1_str, c0, c1, B, To, Fo, Ti, Fi, l, x0, bias,
               l.N, y, y_off, y_str, accumulate)
    if ep is not None:
        d.ep_scale, d.ep_shift = _ptr(ep.scale), _ptr(ep.shift)
        d.ep_slope = _ptr(ep.slope)
        if ep.stats is not None:
            d.stats_sum, d.stats_sumsq = ep.stats[0].data_ptr(), ep.stats[1].data_ptr()
        if not (allow_umma and _umma_ok(d)):
            return None
    if ep is None and policy.use_umma and policy.split_gemm and _split_gemm(d, l, a, b, bias, x0, y):
        return y
    if allow_umma and _umma_ok(d):
        c1p = int(_lib.load().clskd_tapconv_umma_c1p(c0, c1)) if c1 else 0
        if c0 % 16 or c1p != c1 or l.N % 16:
            w = packed_weights(l._cache, "nc_pad%d_%d" % (c0, c1), lambda: l.t_nc_padded(c0, c1, c1p), a, b, torch.bfloat16)
        else:
            w = packed_weights(l._cache, "nc", lambda: l.t_nc, a, b, torch.bfloat16)
        d.w = w.data_ptr()
        call("clskd_tapconv_fwd_umma", ctypes.byref(d), _stream())
        umma_launches += 1
    else:
        w = packed_weights(l._cache, "cn", lambda: l.t_cn, a, b, torch.float32)
        d.w = w.data_ptr()
        call("clskd_tapconv_fwd", ctypes.byref(d), _stream())
        core_launches += 1
    return y


def run_wgrad(x0, x1, c0, c1, B, To, Fo, Ti, Fi, l: Launch, dy, dw_out, x0_view=None, x1_view=None,
              dy_view=None):
    x0_off, x0_str = x0_view if x0_view is not None else _view_of(x0)
    if x1 is not None:
        x1_off, x1_str = x1_view if x1_view is not None else _view_of(x1)
    else:
        x1_off, x1_str = 0, (0, 0, 0)
    y_off, y_str = dy_view if dy_view is not None else _view_of(dy)
    d = TapConv()
    _fill_desc(d, x0, x0_off, x0_str, x1, x1_off, x1_str, c0, c1, B, To, Fo, Ti, Fi, l, dw_out, None,
               l.N, dy, y_off, y_str, False)
    global umma_launches, core_launches
    if policy.use_umma and _lib.load().clskd_tapconv_wgrad_umma_supported(ctypes.byref(d)):
        call("clskd_tapconv_wgrad_umma", ctypes.byref(d), _stream())
        umma_launches += 1
    else:
        call("clskd_tapconv_wgrad", ctypes.byref(d), _stream())
        core_launches += 1


def _chan_ok(t):
    """usable as a tapconv operand without a copy: 4-D, unit channel stride"""
    return t.dim() == 4 and (t.shape[3] == 1 or t.stride(3) == 1)


class TapConvFn(torch.autograd.Function):
    """Generic convolution-like operator on physical tensors.

    forward(plan, x0, x1, a, b, bias_a, bias_b) -> y [B, To, Fo, N]
    x0/x1: dense [B, Ti, Fi, c0/c1]; a/b: parameter tensors referenced by the plan's tables.
    """

    @staticmethod
    def forward(ctx, plan: ConvPlan, x0, x1, a, b, bias_a, bias_b, out_dtype):
        _require_cuda(x0, x1, a)
        if not _chan_ok(x0):
            x0 = strided_copy(x0)
        if x1 is not None and (not _chan_ok(x1) or x1.dtype != x0.dtype):
            x1 = strided_copy(x1, x0.dtype)
        B, Ti, Fi, _ = x0.shape
        To, Fo = plan.out_size(Ti, Fi)
        a32 = _f32c(a)
        b32 = _f32c(b) if b is not None else None
        bias = None
        if plan.bias_table is not None and bias_a is not None:
            bias = packed_weights(plan._cache, "bias", lambda: plan.bias_table, _f32c(bias_a),
                                  _f32c(bias_b) if bias_b is not None else None, torch.float32)
        y = torch.empty((B, To, Fo, plan.N), dtype=out_dtype, device=x0.device)
        ep = _take_epilogue()
        if ep is not None and ((ep.stats is not None and ep.stats.shape[1] != plan.N) or
                               (ep.scale is not None and ep.scale.numel() != plan.N)):
            ep = None                # not the conv this request was made for
        for i, l in enumerate(plan.fwd):
            Fo_l = Fo // l.osf
            yv = (l.ooff * plan.N, (To * Fo * plan.N, Fo * plan.N, l.osf * plan.N))
            r = run_tapconv(x0, x1, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, a32, b32, bias, y, y_view=yv, ep=ep)
            if r is None:            # epilogue not fusable: only ever refused on the first launch of a plan
                if i:
                    raise RuntimeError("tapconv: fused epilogue refused after the first launch of a plan")
                ep = None
                run_tapconv(x0, x1, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, a32, b32, bias, y, y_view=yv)
        if ep is not None:
            global fused_epilogues
            ep.fused = True
            fused_epilogues += 1
        ctx.plan = plan
        ctx.save_for_backward(x0, x1, a, b)
        ctx.has_bias = (bias_a is not None, bias_b is not None)
        ctx.dims = (B, Ti, Fi, To, Fo)
        return y

    @staticmethod
    def backward(ctx, dy):
        plan: ConvPlan = ctx.plan
        x0, x1, a, b = ctx.saved_tensors
        B, Ti, Fi, To, Fo = ctx.dims
        dy = dense(dy, dy.dtype)
        a32 = _f32c(a)
        b32 = _f32c(b) if b is not None else None
        need = ctx.needs_input_grad
        dx0 = dx1 = da = db = dba = dbb = None
        # ---- data gradients
        for src, (xs, cn) in enumerate(((x0, plan.c0), (x1, plan.c1))):
            if xs is None or not need[1 + src]:
                continue
            dx = torch.empty(xs.shape, dtype=xs.dtype, device=xs.device)
            for l in plan.dgrad[src]:
                Fi_l = Fi // l.osf
                yv = (l.ooff * cn, (Ti * Fi * cn, Fi * cn, l.osf * cn))
                run_tapconv(dy, None, plan.N, 0, B, Ti, Fi_l, To, Fo, l, a32, b32, None, dx, y_view=yv)
            if src == 0:
                dx0 = dx
            else:
                dx1 = dx
        # ---- weight gradients
        if need[3] or (b is not None and need[4]):
            dwcat = torch.empty(plan.wcat, dtype=torch.float32, device=dy.device)
            for l in plan.fwd:
                Fo_l = Fo // l.osf
                dyv = (l.ooff * plan.N, (To * Fo * plan.N, Fo * plan.N, l.osf * plan.N))
                dw_view = dwcat[l.woff:l.woff + len(l.dt) * plan.Ctot * plan.N]
                run_wgrad(x0, x1, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, dy, dw_view, dy_view=dyv)
            if need[3]:
                da = unpack_grads(dwcat, plan.unpack_a, plan.na).view_as(a)
            if b is not None and need[4]:
                db = unpack_grads(dwcat, plan.unpack_b, plan.nb).view_as(b)
        # ---- bias gradients
        if (ctx.has_bias[0] and need[5]) or (ctx.has_bias[1] and need[6]):
            s, _ = colstats(dy.view(-1, plan.N))
            s32 = f64_to_f32(s)
            if ctx.has_bias[0] and need[5]:
                dba = unpack_grads(s32, plan.bias_unpack_a, plan.bias_unpack_a.shape[0])
            if ctx.has_bias[1] and need[6]:
                dbb = unpack_grads(s32, plan.bias_unpack_b, plan.bias_unpack_b.shape[0])
        return None, dx0, dx1, da, db, dba, dbb, None


class TapConvResFn(torch.autograd.Function):
    """conv(x) together with an alias of x for a second consumer (the ABF residual, framework.py:221-224):
    forward(plan, x0, a, out_dtype) -> (y, x0 alias).  The backward receives both gradients at once, so the data
    gradient of the convolution is ACCUMULATED into the residual's gradient by the tcgen05 epilogue (TMA reduce-add)
    instead of being written out and summed by a separate pass over three tensors (clskd_sum_n)."""

    @staticmethod
    def forward(ctx, plan: ConvPlan, x0, a, out_dtype):
        _require_cuda(x0, None, a)
        ctx.set_materialize_grads(False)
        if not _chan_ok(x0):
            x0 = strided_copy(x0)
        B, Ti, Fi, _ = x0.shape
        To, Fo = plan.out_size(Ti, Fi)
        a32 = _f32c(a)
        y = torch.empty((B, To, Fo, plan.N), dtype=out_dtype, device=x0.device)
        ep = _take_epilogue()
        if ep is not None and ((ep.stats is not None and ep.stats.shape[1] != plan.N) or
                               (ep.scale is not None and ep.scale.numel() != plan.N)):
            ep = None
        for i, l in enumerate(plan.fwd):
            Fo_l = Fo // l.osf
            yv = (l.ooff * plan.N, (To * Fo * plan.N, Fo * plan.N, l.osf * plan.N))
            r = run_tapconv(x0, None, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, a32, None, None, y, y_view=yv, ep=ep)
            if r is None:
                if i:
                    raise RuntimeError("tapconv: fused epilogue refused after the first launch of a plan")
                ep = None
                run_tapconv(x0, None, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, a32, None, None, y, y_view=yv)
        if ep is not None:
            global fused_epilogues
            ep.fused = True
            fused_epilogues += 1
        ctx.plan = plan
        ctx.save_for_backward(x0, a)
        ctx.dims = (B, Ti, Fi, To, Fo)
        return y, x0.detach().view(x0.shape)

    @staticmethod
    def backward(ctx, dy, dres):
        plan: ConvPlan = ctx.plan
        x0, a = ctx.saved_tensors
        B, Ti, Fi, To, Fo = ctx.dims
        a32 = _f32c(a)
        need = ctx.needs_input_grad
        dx0 = da = None
        cn = plan.c0
        if dy is None:
            return None, dres, None, None
        dy = dense(dy, dy.dtype)
        if need[1]:
            acc = (policy.use_umma and dres is not None and dres.dtype == x0.dtype and dres.shape == x0.shape
                   and dres.is_contiguous())
            if acc and not getattr(plan, "_res_tuned", False):
                # an accumulating launch cannot be autotuned on its own output: tune the shape once with a plain launch
                tmp = torch.empty(x0.shape, dtype=x0.dtype, device=x0.device)
                for l in plan.dgrad[0]:
                    yv = (l.ooff * cn, (Ti * Fi * cn, Fi * cn, l.osf * cn))
                    run_tapconv(dy, None, plan.N, 0, B, Ti, Fi // l.osf, To, Fo, l, a32, None, None, tmp, y_view=yv)
                plan._res_tuned = True
                del tmp
            dx = dres if acc else torch.empty(x0.shape, dtype=x0.dtype, device=x0.device)
            for l in plan.dgrad[0]:
                yv = (l.ooff * cn, (Ti * Fi * cn, Fi * cn, l.osf * cn))
                run_tapconv(dy, None, plan.N, 0, B, Ti, Fi // l.osf, To, Fo, l, a32, None, None, dx, y_view=yv,
                            accumulate=acc)
            if not acc and dres is not None:
                dx = FanoutFn.backward(None, dx, dres)[0]
            dx0 = dx
        elif dres is not None:
            dx0 = dres
        if need[2]:
            dwcat = torch.empty(plan.wcat, dtype=torch.float32, device=dy.device)
            for l in plan.fwd:
                Fo_l = Fo // l.osf
                dyv = (l.ooff * plan.N, (To * Fo * plan.N, Fo * plan.N, l.osf * plan.N))
                dw_view = dwcat[l.woff:l.woff + len(l.dt) * plan.Ctot * plan.N]
                run_wgrad(x0, None, plan.c0, plan.c1, B, To, Fo_l, Ti, Fi, l, dy, dw_view, dy_view=dyv)
            da = unpack_grads(dwcat, plan.unpack_a, plan.na).view_as(a)
        return None, dx0, da, None


# --------------------------------------------------------------------------------------------
# small wrappers
# --------------------------------------------------------------------------------------------

def colstats(x2d):
    """x2d dense [M, C] -> (sum fp64 [C], sumsq fp64 [C])"""
    M, C = x2d.shape
    s = torch.empty(2, C, dtype=torch.float64, device=x2d.device)
    call("clskd_colstats", x2d.data_ptr(), _tag(x2d.dtype), M, C, s[0].data_ptr(), s[1].data_ptr(), _stream())
    return s[0], s[1]


def f64_to_f32(t, scale=1.0):
    out = torch.empty(t.shape, dtype=torch.float32, device=t.device)
    call("clskd_f64_to_f32", t.data_ptr(), t.numel(), float(scale), out.data_ptr(), _stream())
    return out


class BNActFn(torch.autograd.Function):
    """y = prelu(batchnorm(x)) over the channel (last) axis of a dense physical tensor.
    Replaces nn.BatchNorm2d + nn.PReLU (DCCRN.py:80-82); slope=None -> no activation."""

    @staticmethod
    def forward(ctx, x, gamma, beta, slope, running_mean, running_var, training, momentum, eps, pre_stats=None):
        _require_cuda(x)
        C = x.shape[-1]
        M = x.numel() // C
        dev = x.device
        stats = torch.empty(2, C, dtype=torch.float32, device=dev)
        mean, invstd = stats[0], stats[1]
        use_batch = training or running_mean is None
        if use_batch:
            # pre_stats: fp64 [2, C] column sums already accumulated by the producing conv's epilogue
            s, ss = (pre_stats[0], pre_stats[1]) if pre_stats is not None else colstats(x.view(M, C))
            call("clskd_bn_finalize", s.data_ptr(), ss.data_ptr(), M, C, float(eps),
                 float(momentum if momentum is not None else 0.0), mean.data_ptr(), invstd.data_ptr(),
                 running_mean.data_ptr() if (running_mean is not None and training) else None,
                 running_var.data_ptr() if (running_var is not None and training) else None, _stream())
        else:
            call("clskd_bn_eval_stats", running_mean.data_ptr(), running_var.data_ptr(), C, float(eps),
                 mean.data_ptr(), invstd.data_ptr(), _stream())
        y = torch.empty_like(x)
        g32 = _f32c(gamma) if gamma is not None else None
        b32 = _f32c(beta) if beta is not None else None
        s32 = _f32c(slope) if slope is not None else None
        call("clskd_bn_act_fwd", x.data_ptr(), _tag(x.dtype), M, C, mean.data_ptr(), invstd.data_ptr(),
             _ptr(g32), _ptr(b32), _ptr(s32), y.data_ptr(), _tag(y.dtype), _stream())
        ctx.save_for_backward(x, stats, gamma, beta, slope)
        ctx.use_batch = use_batch
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma, beta, slope = ctx.saved_tensors
        C = x.shape[-1]
        M = x.numel() // C
        dy = dense(dy, x.dtype)
        mean, invstd = stats[0], stats[1]
        g32 = _f32c(gamma) if gamma is not None else None
        b32 = _f32c(beta) if beta is not None else None
        s32 = _f32c(slope) if slope is not None else None
        sums = torch.empty(2 * C + 1, dtype=torch.float64, device=x.device)
        p_dz, p_dzx, p_ds = sums.data_ptr(), sums.data_ptr() + 8 * C, sums.data_ptr() + 16 * C
        call("clskd_bn_act_bwd_stats", x.data_ptr(), _tag(x.dtype), dy.data_ptr(), _tag(dy.dtype), M, C,
             mean.data_ptr(), invstd.data_ptr(), _ptr(g32), _ptr(b32), _ptr(s32), p_dz, p_dzx, p_ds, _stream())
        dx = torch.empty_like(x)
        pg = torch.empty(2 * C + 1, dtype=torch.float32, device=x.device)
        call("clskd_bn_act_bwd_apply", x.data_ptr(), _tag(x.dtype), dy.data_ptr(), _tag(dy.dtype), M, C,
             mean.data_ptr(), invstd.data_ptr(), _ptr(g32), _ptr(b32), _ptr(s32), p_dz, p_dzx, p_ds,
             1 if ctx.use_batch else 0, dx.data_ptr(), _tag(dx.dtype), pg.data_ptr(),
             pg.data_ptr() + 4 * C, pg.data_ptr() + 8 * C, _stream())
        dgamma = pg[:C].view_as(gamma) if gamma is not None else None
        dbeta = pg[C:2 * C].view_as(beta) if beta is not None else None
        dslope = pg[2 * C:].view_as(slope) if slope is not None else None
        return dx, dgamma, dbeta, dslope, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------
# STFT / iSTFT (framed-window DFT as a GEMM)
# --------------------------------------------------------------------------------------------

def _gemm_plan(weight_kn: np.ndarray, device):
    """ConvPlan of a plain GEMM Y[m, n] = sum_k X[m, k] W[k, n] whose weight is a fixed buffer;
    the 'parameter' a is the flat weight itself."""
    K, N = weight_kn.shape
    block = _codes((1, 1, K, N), 0)
    return ConvPlan("conv", block, 1, 0, 0, K, 0, None, K * N, 0, device)


class FramedGemmFn(torch.autograd.Function):
    """spec[b, t, :] = frames(xpad)[b, t, :] @ W   with frames(xpad)[b,t,k] = xpad[b, t*hop + k].
    xpad is a dense fp32 [B, Lp] signal (already padded); W = plan weight [win, N]."""

    @staticmethod
    def forward(ctx, plan: ConvPlan, w_flat, xpad, win, hop):
        B, Lp = xpad.shape
        T = (Lp - win) // hop + 1
        y = torch.empty((B, T, 1, plan.N), dtype=torch.float32, device=xpad.device)
        run_tapconv(xpad, None, win, 0, B, T, 1, T, 1, plan.fwd[0], w_flat, None, None, y,
                    x0_view=(0, (Lp, hop, 0)), allow_umma=False)
        ctx.plan, ctx.win, ctx.hop, ctx.dims = plan, win, hop, (B, Lp, T)
        ctx.save_for_backward(w_flat)
        return y.view(B, T, plan.N)

    @staticmethod
    def backward(ctx, dy):
        (w_flat,) = ctx.saved_tensors
        plan, win, hop = ctx.plan, ctx.win, ctx.hop
        B, Lp, T = ctx.dims
        dy = dense(dy, torch.float32)
        dframes = torch.empty((B, T, 1, win), dtype=torch.float32, device=dy.device)
        run_tapconv(dy.view(B, T, 1, plan.N), None, plan.N, 0, B, T, 1, T, 1, plan.dgrad[0][0], w_flat, None,
                    None, dframes, allow_umma=False)
        # overlap-add of the frame gradients (plain conv_transpose1d): length (T-1)*hop+win <= Lp
        Lo = (T - 1) * hop + win
        dx = torch.zeros((B, Lp), dtype=torch.float32, device=dy.device) if Lo != Lp else \
            torch.empty((B, Lp), dtype=torch.float32, device=dy.device)
        tmp = dx if Lo == Lp else torch.empty((B, Lo), dtype=torch.float32, device=dy.device)
        call("clskd_ola_fwd", dframes.data_ptr(), None, B, T, win, hop, 0, 0, tmp.data_ptr(), _stream())
        if Lo != Lp:
            dx[:, :Lo] = tmp
        return None, None, dx, None, None


class Pad1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, left, right, mode):
        _require_cuda(x)
        B, L = x.shape
        if x.stride(1) != 1:
            x = dense(x)
        out = torch.empty((B, L + left + right), dtype=torch.float32, device=x.device)
        call("clskd_pad1d", x.data_ptr(), _tag(x.dtype), x.stride(0), B, L, left, right, mode,
             out.data_ptr(), _stream())
        ctx.args = (B, L, left, right, mode, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        B, L, left, right, mode, dt = ctx.args
        g = dense(g, torch.float32)
        dx = torch.empty((B, L), dtype=torch.float32, device=g.device)
        call("clskd_pad1d_bwd", g.data_ptr(), B, L, left, right, mode, dx.data_ptr(), 0, _stream())
        if dt != torch.float32:
            dx = dense(dx, dt)
        return dx, None, None, None


class OlaFn(torch.autograd.Function):
    """wav = clamp(overlap_add(frames) / window^2-overlap)[trim:-trim]  (tools_for_model.py:100-107,
    DCCRN.py:235-237)."""

    @staticmethod
    def forward(ctx, frames, window, hop, trim, do_clamp):
        B, T, win = frames.shape
        L = (T - 1) * hop + win - 2 * trim
        wav = torch.empty((B, L), dtype=torch.float32, device=frames.device)
        call("clskd_ola_fwd", frames.data_ptr(), _ptr(window), B, T, win, hop, trim, 1 if do_clamp else 0,
             wav.data_ptr(), _stream())
        ctx.args = (B, T, win, hop, trim, do_clamp)
        ctx.save_for_backward(wav, window)
        return wav

    @staticmethod
    def backward(ctx, g):
        wav, window = ctx.saved_tensors
        B, T, win, hop, trim, do_clamp = ctx.args
        g = dense(g, torch.float32)
        dframes = torch.empty((B, T, win), dtype=torch.float32, device=g.device)
        call("clskd_ola_bwd", g.data_ptr(), wav.data_ptr(), _ptr(window), B, T, win, hop, trim,
             1 if do_clamp else 0, dframes.data_ptr(), _stream())
        return dframes, None, None, None, None


# --------------------------------------------------------------------------------------------
# mask
# --------------------------------------------------------------------------------------------
_MODES = {"E": 0, "C": 1, "R": 2}


class MaskFn(torch.autograd.Function):
    """Applies the decoder mask to the noisy spectrum (DCCRN.py:207-232).
    spec: fp32 [B, T, nb, 2]; mask: dense physical [B, T, nb-1, 2] (any act dtype).
    Returns (out_spec [B,T,nb,2], mask_padded [B,T,nb,2] or None)."""

    @staticmethod
    def forward(ctx, spec, mask, mode, want_masks):
        _require_cuda(spec, mask)
        B, T, nb, _ = spec.shape
        out = torch.empty_like(spec)
        mp = torch.empty_like(spec) if want_masks else None
        call("clskd_mask_fwd", spec.data_ptr(), mask.data_ptr(), _tag(mask.dtype), mask.stride(0),
             mask.stride(1), B, T, nb, _MODES[mode], out.data_ptr(), _ptr(mp), _stream())
        ctx.mode = mode
        ctx.save_for_backward(spec, mask)
        if mp is None:
            return out, None
        ctx.mark_non_differentiable(mp)
        return out, mp

    @staticmethod
    def backward(ctx, g, _gmp):
        spec, mask = ctx.saved_tensors
        B, T, nb, _ = spec.shape
        g = dense(g, torch.float32)
        dmask = torch.empty(mask.shape, dtype=mask.dtype, device=mask.device)
        call("clskd_mask_bwd", spec.data_ptr(), mask.data_ptr(), _tag(mask.dtype), mask.stride(0),
             mask.stride(1), B, T, nb, _MODES[ctx.mode], g.data_ptr(), dmask.data_ptr(), _tag(dmask.dtype),
             dmask.stride(0), dmask.stride(1), _stream())
        return None, dmask, None, None


# --------------------------------------------------------------------------------------------
# losses
# --------------------------------------------------------------------------------------------
_WAVE_KINDS = {"si_snr": 0, "sdr": 1, "si_sdr": 2, "mse": 3}


class WaveLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, s1, s2, kind, eps):
        _require_cuda(s1, s2)
        s1 = dense(s1.reshape(-1, s1.shape[-1]), torch.float32)
        s2 = dense(s2.reshape(-1, s2.shape[-1]), torch.float32)
        if s1.shape != s2.shape:
            raise RuntimeError("wave loss: shape mismatch %s vs %s" % (tuple(s1.shape), tuple(s2.shape)))
        B, L = s1.shape
        part = torch.empty(B, 4, dtype=torch.float64, device=s1.device)
        out = torch.empty((), dtype=torch.float32, device=s1.device)
        call("clskd_wave_loss_fwd", s1.data_ptr(), s2.data_ptr(), B, L, _WAVE_KINDS[kind], float(eps),
             part.data_ptr(), out.data_ptr(), _stream())
        ctx.kind, ctx.eps = kind, eps
        ctx.save_for_backward(s1, s2, part)
        return out

    @staticmethod
    def backward(ctx, g):
        s1, s2, part = ctx.saved_tensors
        B, L = s1.shape
        kind = ctx.kind
        g = dense(g, torch.float32)
        n1, n2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if kind == "si_snr" and n2:
            raise RuntimeError("si_snr: gradient w.r.t. the reference signal is not implemented")
        if kind == "si_sdr" and n1:
            raise RuntimeError("si_sdr: gradient w.r.t. the reference signal is not implemented")
        d1 = torch.empty_like(s1) if n1 else None
        d2 = torch.empty_like(s2) if n2 else None
        call("clskd_wave_loss_bwd", s1.data_ptr(), s2.data_ptr(), B, L, _WAVE_KINDS[kind], float(ctx.eps),
             part.data_ptr(), g.data_ptr(), _ptr(d1), _ptr(d2), _stream())
        return d1, d2, None, None


class StftMagLossFn(torch.autograd.Function):
    """(sc_loss, mag_loss) of framework.py:35-68 from interleaved spectra [M, nb, 2] (fp32)."""

    @staticmethod
    def forward(ctx, xs, ys):
        xs, ys = dense(xs, torch.float32), dense(ys, torch.float32)
        n = xs.numel() // 2
        part = torch.empty(3, dtype=torch.float64, device=xs.device)
        out = torch.empty(2, dtype=torch.float32, device=xs.device)
        call("clskd_stftmag_loss_fwd", xs.data_ptr(), ys.data_ptr(), n, part.data_ptr(), out.data_ptr(),
             _stream())
        ctx.save_for_backward(xs, ys, part)
        ctx.n = n
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_sc, g_mag):
        xs, ys, part = ctx.saved_tensors
        n = ctx.n
        dxs = torch.empty_like(xs)
        g_sc = dense(g_sc, torch.float32) if g_sc is not None else None
        g_mag = dense(g_mag, torch.float32) if g_mag is not None else None
        call("clskd_stftmag_loss_bwd", xs.data_ptr(), ys.data_ptr(), n, part.data_ptr(), _ptr(g_mag),
             _ptr(g_sc), 1.0 / n, 1.0, dxs.data_ptr(), _stream())
        return dxs, None


def gram(z2d: torch.Tensor) -> torch.Tensor:
    """G = Z Z^T (fp32 [B,B]) of a dense [B, K] matrix."""
    B, K = z2d.shape
    G = torch.empty(B, B, dtype=torch.float32, device=z2d.device)
    if _gram_umma_ok(z2d):
        call("clskd_gram_fwd_umma", z2d.data_ptr(), _tag(z2d.dtype), B, K, z2d.stride(0), G.data_ptr(), 0, _stream())
    else:
        call("clskd_gram_fwd", z2d.data_ptr(), _tag(z2d.dtype), B, K, z2d.stride(0), G.data_ptr(), 0, _stream())
    return G


def _gram_umma_ok(z2d):
    global umma_launches, core_launches
    ok = policy.use_umma and z2d.dtype == torch.bfloat16 and bool(_lib.load().clskd_gram_umma_supported(
        z2d.data_ptr(), _tag(z2d.dtype), z2d.shape[0], z2d.shape[1], z2d.stride(0)))
    if ok:
        umma_launches += 1
    else:
        core_launches += 1
    return ok


class SPKDFn(torch.autograd.Function):
    """SPKD loss (framework.py:157-172) between a student feature (needs grad) and a teacher
    feature (constant): || rownorm1(Zt Zt^T) - rownorm1(Zs Zs^T) ||_F^2 * scale."""

    @staticmethod
    def forward(ctx, zs, zt, scale):
        _require_cuda(zs, zt)
        B = zs.shape[0]
        zs2 = dense(zs).view(B, -1)
        zt2 = dense(zt).view(zt.shape[0], -1)
        if zt2.shape[0] != B:
            raise RuntimeError("SPKD: batch mismatch %d vs %d" % (B, zt2.shape[0]))
        Gs, Gt = gram(zs2), gram(zt2)
        loss = torch.empty((), dtype=torch.float32, device=zs.device)
        dGs = torch.empty_like(Gs) if ctx.needs_input_grad[0] else None
        call("clskd_spkd_loss", Gt.data_ptr(), Gs.data_ptr(), B, float(scale), loss.data_ptr(), _ptr(dGs),
             _stream())
        ctx.save_for_backward(zs2, dGs)
        ctx.shape = zs.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        zs2, dGs = ctx.saved_tensors
        B, K = zs2.shape
        g = dense(g, torch.float32)
        dz = torch.empty_like(zs2)
        if _gram_umma_ok(zs2):
            call("clskd_gram_bwd_umma", zs2.data_ptr(), _tag(zs2.dtype), B, K, zs2.stride(0), dGs.data_ptr(),
                 g.data_ptr(), dz.data_ptr(), _tag(dz.dtype), dz.stride(0), _stream())
        else:
            call("clskd_gram_bwd", zs2.data_ptr(), _tag(zs2.dtype), B, K, zs2.stride(0), dGs.data_ptr(),
                 g.data_ptr(), dz.data_ptr(), _tag(dz.dtype), dz.stride(0), 0, _stream())
        return dz.view(ctx.shape), None, None


# --------------------------------------------------------------------------------------------
# ABF helpers
# --------------------------------------------------------------------------------------------

class ResizeFFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Fo):
        B, T, Fi, C = x.shape
        y = torch.empty((B, T, Fo, C), dtype=x.dtype, device=x.device)
        call("clskd_resize_f_fwd", x.data_ptr(), _tag(x.dtype), B * T, Fi, Fo, C, y.data_ptr(), _stream())
        ctx.dims = (B, T, Fi, Fo, C)
        return y

    @staticmethod
    def backward(ctx, g):
        B, T, Fi, Fo, C = ctx.dims
        g = dense(g)
        dx = torch.empty((B, T, Fi, C), dtype=g.dtype, device=g.device)
        call("clskd_resize_f_bwd", g.data_ptr(), _tag(g.dtype), B * T, Fi, Fo, C, dx.data_ptr(), _stream())
        return dx, None


class AttBlendFn(torch.autograd.Function):
    """x*sigmoid(z0) + y*sigmoid(z1)  (framework.py:218-219); z fp32 [.., 2] logits."""

    @staticmethod
    def forward(ctx, x, y, z):
        C = x.shape[-1]
        M = x.numel() // C
        out = torch.empty_like(x)
        call("clskd_att_blend_fwd", x.data_ptr(), y.data_ptr(), _tag(x.dtype), z.data_ptr(), M, C,
             out.data_ptr(), _stream())
        ctx.save_for_backward(x, y, z)
        return out

    @staticmethod
    def backward(ctx, g):
        x, y, z = ctx.saved_tensors
        C = x.shape[-1]
        M = x.numel() // C
        g = dense(g, x.dtype)
        dx, dy, dz = torch.empty_like(x), torch.empty_like(y), torch.empty_like(z)
        call("clskd_att_blend_bwd", x.data_ptr(), y.data_ptr(), _tag(x.dtype), z.data_ptr(), g.data_ptr(), M, C,
             dx.data_ptr(), dy.data_ptr(), dz.data_ptr(), _stream())
        return dx, dy, dz


def _abf_mid_forward(z1, y_prev, gamma, beta, watt, batt, running_mean, running_var, training, momentum, eps, pre_stats):
    """BatchNorm statistics of z1 (batch or running) + clskd_abf_mid_fwd -> (xb, stats [2,C] = mean / invstd, logits,
    use_batch)"""
    B, T, F, C = z1.shape
    Fy = y_prev.shape[2]
    M = B * T * F
    dev = z1.device
    stats = torch.empty(2, C, dtype=torch.float32, device=dev)
    mean, invstd = stats[0], stats[1]
    use_batch = training or running_mean is None
    if use_batch:
        s, ss = (pre_stats[0], pre_stats[1]) if pre_stats is not None else colstats(z1.view(M, C))
        call("clskd_bn_finalize", s.data_ptr(), ss.data_ptr(), M, C, float(eps),
             float(momentum if momentum is not None else 0.0), mean.data_ptr(), invstd.data_ptr(),
             running_mean.data_ptr() if (running_mean is not None and training) else None,
             running_var.data_ptr() if (running_var is not None and training) else None, _stream())
    else:
        call("clskd_bn_eval_stats", running_mean.data_ptr(), running_var.data_ptr(), C, float(eps),
             mean.data_ptr(), invstd.data_ptr(), _stream())
    g32, b32 = _f32c(gamma), _f32c(beta)
    w32 = _f32c(watt).view(2, 2 * C)
    ba32 = _f32c(batt) if batt is not None else None
    xb = torch.empty_like(z1)
    logits = torch.empty((B, T, F, 2), dtype=torch.float32, device=dev)
    call("clskd_abf_mid_fwd", z1.data_ptr(), y_prev.data_ptr(), _tag(z1.dtype), B, T, F, Fy, C, mean.data_ptr(),
         invstd.data_ptr(), g32.data_ptr(), b32.data_ptr(), w32.data_ptr(), _ptr(ba32), xb.data_ptr(),
         logits.data_ptr(), _stream())
    return xb, stats, logits, use_batch


class AbfMidFn(torch.autograd.Function):
    """Fused ABF middle stage: BatchNorm(z1) -> nearest-resize(y_prev) -> 2-logit attention conv ->
    sigmoid blend, one pass forward and two passes backward (clskd_abf_mid_* in clskd.h).
    z1: dense [B,T,F,C] (1x1-conv output), y_prev: dense [B,T,Fy,C]; returns xb [B,T,F,C]."""

    @staticmethod
    def forward(ctx, z1, y_prev, gamma, beta, watt, batt, running_mean, running_var, training, momentum, eps,
                pre_stats=None):
        xb, stats, logits, use_batch = _abf_mid_forward(z1, y_prev, gamma, beta, watt, batt, running_mean, running_var,
                                                        training, momentum, eps, pre_stats)
        ctx.save_for_backward(z1, y_prev, stats, gamma, beta, watt, logits)
        ctx.use_batch = use_batch
        ctx.has_bias = batt is not None
        return xb

    @staticmethod
    def backward(ctx, g):
        z1, y_prev, stats, gamma, beta, watt, logits = ctx.saved_tensors
        B, T, F, C = z1.shape
        Fy = y_prev.shape[2]
        dev = z1.device
        g = dense(g, z1.dtype)
        acc = torch.empty(2 * C + 4 * C + 2, dtype=torch.float64, device=dev)
        sums, dwatt, dbatt = acc[:2 * C], acc[2 * C:6 * C], acc[6 * C:]
        dz1 = torch.empty_like(z1)
        dy = torch.empty_like(y_prev)
        g32, b32 = _f32c(gamma), _f32c(beta)
        w32 = _f32c(watt).view(2, 2 * C)
        call("clskd_abf_mid_bwd", g.data_ptr(), z1.data_ptr(), y_prev.data_ptr(), _tag(z1.dtype), B, T, F, Fy, C,
             stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), w32.data_ptr(),
             logits.data_ptr(), 1 if ctx.use_batch else 0, sums.data_ptr(), dwatt.data_ptr(), dbatt.data_ptr(),
             dz1.data_ptr(), dy.data_ptr(), _stream())
        a32 = f64_to_f32(acc)
        dbeta, dgamma = a32[:C].view_as(beta), a32[C:2 * C].view_as(gamma)
        dw = a32[2 * C:6 * C].view_as(watt)
        db = a32[6 * C:] if ctx.has_bias else None
        return dz1, dy, dgamma, dbeta, dw, db, None, None, None, None, None, None


_gram_launches = {}


def _pointwise_launch(K, N, device):
    """descriptor-only Launch of a 1x1 contraction [K -> N] (weight-gradient launches on ad-hoc operands)"""
    key = (K, N, device)
    if key not in _gram_launches:
        _gram_launches[key] = Launch(dt=[0], df=[0], K=K, N=N, device=device)
    return _gram_launches[key]


def abf_fold_supported(x, y_prev, w1):
    """conv1 + middle stage as one autograd node with the BatchNorm backward folded into conv1's gradients
    (AbfFoldFn): tensor-core policy, bf16 maps, 16 | Cin, 16 | C"""
    if not (policy.use_umma and policy.abf_fold and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        return False
    C, Cin = w1.shape[0], w1.shape[1]
    if Cin % 16 or C % 16 or Cin > 256 or C > 256 or x.shape[3] != Cin:
        return False
    B, T, F, _ = x.shape
    if y_prev.shape[0] != B or y_prev.shape[1] != T or y_prev.shape[3] != C:
        return False
    lib = _lib.load()
    return bool(lib.clskd_colgram_supported(_tag(x.dtype), Cin)) and bool(lib.clskd_abf_mid_supported(B, T, F, y_prev.shape[2], C))


class AbfFoldFn(torch.autograd.Function):
    """ABF conv1 (1x1, no bias) + BatchNorm + fused middle stage as ONE autograd node (framework.py:209-219).
    Forward: the usual launches (tcgen05 1x1 conv with the batch statistics in its epilogue, clskd_abf_mid_fwd).
    Backward: ONE pass over gout / z1 / y_prev (clskd_abf_mid_bwd_fold: batch sums, dW_att, dy_prev and dxp = the
    gradient of the BatchNorm output); the BatchNorm-backward affine is folded into conv1's gradients
    (clskd_abf_fold_dgrad / _dw1 in clskd.h): dx is one two-source 1x1 contraction over [dxp | x], dW1 follows from
    x^T dxp, x^T x and the column sums of x (clskd_colgram: one pass over x).  dz1 is never written and the second pass over the three big maps of
    AbfMidFn.backward is gone.
    x: [B,T,F,Cin] bf16; y_prev: dense [B,T,Fy,C]; w1: conv1 weight [C,Cin,1,1]; returns xb [B,T,F,C]."""

    @staticmethod
    def forward(ctx, plan, x, y_prev, w1, gamma, beta, watt, batt, running_mean, running_var, training, momentum, eps):
        if not (_chan_ok(x) and x.is_contiguous()):
            x = strided_copy(x)
        C, Cin = w1.shape[0], w1.shape[1]
        use_batch = training or running_mean is None
        pre = gs = None
        if use_batch:
            # batch statistics of z1 = W1 x from the Cin x Cin moments of the narrow input (which the backward needs
            # anyway) instead of the conv's statistics epilogue: the 1x1 conv then runs at its HBM floor
            # (0.48 -> 0.36 ms at F = 128)
            M = x.shape[0] * x.shape[1] * x.shape[2]
            gs = torch.empty(Cin * Cin + Cin, dtype=torch.float64, device=x.device)
            call("clskd_colgram", x.data_ptr(), _tag(x.dtype), M, Cin, gs.data_ptr(), gs[Cin * Cin:].data_ptr(), _stream())
            pre = torch.empty(2, C, dtype=torch.float64, device=x.device)
            call("clskd_abf_fold_stats", gs.data_ptr(), gs[Cin * Cin:].data_ptr(), _f32c(w1).data_ptr(), C, Cin,
                 pre[0].data_ptr(), pre[1].data_ptr(), _stream())
        request_epilogue(None)
        z1 = TapConvFn.apply(plan, x, None, w1, None, None, None, x.dtype)      # (no graph inside forward)
        xb, stats, logits, use_batch = _abf_mid_forward(z1, y_prev, gamma, beta, watt, batt, running_mean, running_var,
                                                        training, momentum, eps, pre)
        ctx.save_for_backward(x, z1, y_prev, stats, gamma, beta, watt, logits, w1, gs)
        ctx.use_batch = use_batch
        ctx.has_bias = batt is not None
        ctx.plan = plan
        return xb

    @staticmethod
    def backward(ctx, g):
        global umma_launches
        x, z1, y_prev, stats, gamma, beta, watt, logits, w1, gs = ctx.saved_tensors
        B, T, F, C = z1.shape
        Cin = x.shape[3]
        Fy = y_prev.shape[2]
        M = B * T * F
        dev = z1.device
        st = _stream()
        g = dense(g, z1.dtype)
        acc = torch.empty(2 * C + 4 * C + 2, dtype=torch.float64, device=dev)
        sums, dwatt, dbatt = acc[:2 * C], acc[2 * C:6 * C], acc[6 * C:]
        dxp = torch.empty_like(z1)
        dy = torch.empty_like(y_prev)
        g32, b32 = _f32c(gamma), _f32c(beta)
        w32 = _f32c(watt).view(2, 2 * C)
        w1f = _f32c(w1).view(C, Cin)
        mean, invstd = stats[0], stats[1]
        call("clskd_abf_mid_bwd_fold", g.data_ptr(), z1.data_ptr(), y_prev.data_ptr(), _tag(z1.dtype), B, T, F, Fy, C,
             mean.data_ptr(), invstd.data_ptr(), g32.data_ptr(), b32.data_ptr(), w32.data_ptr(), logits.data_ptr(),
             sums.data_ptr(), dwatt.data_ptr(), dbatt.data_ptr(), dxp.data_ptr(), dy.data_ptr(), st)
        training = 1 if ctx.use_batch else 0
        need = ctx.needs_input_grad
        dx = dw1 = None
        if need[1]:
            c1p = int(_lib.load().clskd_tapconv_umma_c1p(C, Cin))
            weff = torch.empty((Cin, C + c1p), dtype=torch.bfloat16, device=dev)
            bias = torch.empty(Cin, dtype=torch.float32, device=dev)
            call("clskd_abf_fold_dgrad", w1f.data_ptr(), g32.data_ptr(), mean.data_ptr(), invstd.data_ptr(),
                 sums.data_ptr(), M, training, C, Cin, c1p, weff.data_ptr(), bias.data_ptr(), st)
            dx = torch.empty_like(x)
            d = TapConv()
            l = _pointwise_launch(C + Cin, Cin, dev)
            _fill_desc(d, dxp, 0, _view_of(dxp)[1], x, 0, _view_of(x)[1], C, Cin, B, T, F, T, F, l, weff, bias, Cin,
                       dx, 0, _view_of(dx)[1])
            call("clskd_tapconv_fwd_umma", ctypes.byref(d), st)
            umma_launches += 1
        if need[3]:
            P = torch.empty(Cin * C, dtype=torch.float32, device=dev)
            run_wgrad(x, None, Cin, 0, B, T, F, T, F, _pointwise_launch(Cin, C, dev), dxp, P)
            Gm = sx = None
            if training:
                Gm, sx = gs[:Cin * Cin], gs[Cin * Cin:]          # moments of x from the forward pass
            dw1 = torch.empty((C, Cin), dtype=torch.float32, device=dev)
            call("clskd_abf_fold_dw1", P.data_ptr(), _ptr(Gm), _ptr(sx), w1f.data_ptr(), g32.data_ptr(), mean.data_ptr(),
                 invstd.data_ptr(), sums.data_ptr(), M, training, C, Cin, dw1.data_ptr(), st)
            dw1 = dw1.view_as(w1)
        a32 = f64_to_f32(acc)
        dbeta, dgamma = a32[:C].view_as(beta), a32[C:2 * C].view_as(gamma)
        dw = a32[2 * C:6 * C].view_as(watt)
        db = a32[6 * C:] if ctx.has_bias else None
        return None, dx, dy, dw1, dgamma, dbeta, dw, db, None, None, None, None, None


class AbfMidXsFn(torch.autograd.Function):
    """AbfMidFn for a 1x1 conv with TWO input channels in front of the block (the mask-level map of the
    decoder side): z1 = W1 x is recomputed per row inside the kernels, the BatchNorm statistics of z1
    come from the 2x2 moments of x, and the backward returns dx and dW1 directly (clskd_abf_mid_xs_*).
    xs: dense [B,T,F,2]; w1: conv1 weight [C,2,1,1]; y_prev: dense [B,T,Fy,C]; returns xb [B,T,F,C]."""

    @staticmethod
    def forward(ctx, xs, y_prev, w1, gamma, beta, watt, batt, running_mean, running_var, training, momentum, eps):
        B, T, F, _ = xs.shape
        C = w1.shape[0]
        Fy = y_prev.shape[2]
        M = B * T * F
        dev = xs.device
        st = _stream()
        w1f = _f32c(w1).view(C, 2)
        stats = torch.empty(2, C, dtype=torch.float32, device=dev)
        mean, invstd = stats[0], stats[1]
        use_batch = training or running_mean is None
        if use_batch:
            s5 = torch.empty(5, dtype=torch.float64, device=dev)
            call("clskd_cbn_moments", xs.data_ptr(), _tag(xs.dtype), M, 1, s5.data_ptr(), st)
            ss = torch.empty(2, C, dtype=torch.float64, device=dev)
            call("clskd_rank2_colstats", s5.data_ptr(), w1f.data_ptr(), C, ss[0].data_ptr(), ss[1].data_ptr(), st)
            call("clskd_bn_finalize", ss[0].data_ptr(), ss[1].data_ptr(), M, C, float(eps),
                 float(momentum if momentum is not None else 0.0), mean.data_ptr(), invstd.data_ptr(),
                 running_mean.data_ptr() if (running_mean is not None and training) else None,
                 running_var.data_ptr() if (running_var is not None and training) else None, st)
        else:
            call("clskd_bn_eval_stats", running_mean.data_ptr(), running_var.data_ptr(), C, float(eps),
                 mean.data_ptr(), invstd.data_ptr(), st)
        g32, b32 = _f32c(gamma), _f32c(beta)
        w32 = _f32c(watt).view(2, 2 * C)
        ba32 = _f32c(batt) if batt is not None else None
        xb = torch.empty((B, T, F, C), dtype=xs.dtype, device=dev)
        logits = torch.empty((B, T, F, 2), dtype=torch.float32, device=dev)
        # rank-2 folded kernels (clskd_abf_xs2_*); policy.abf_xs2 = False keeps the round-1 kernels (A/B, tests)
        ctx.xs2 = bool(policy.abf_xs2) and C >= 16
        call("clskd_abf_xs2_fwd" if ctx.xs2 else "clskd_abf_mid_xs_fwd", xs.data_ptr(), w1f.data_ptr(), y_prev.data_ptr(),
             _tag(xs.dtype), B, T, F, Fy, C, mean.data_ptr(), invstd.data_ptr(), g32.data_ptr(), b32.data_ptr(),
             w32.data_ptr(), _ptr(ba32), xb.data_ptr(), logits.data_ptr(), st)
        ctx.save_for_backward(xs, y_prev, stats, gamma, beta, watt, logits, w1)
        ctx.use_batch = use_batch
        ctx.has_bias = batt is not None
        return xb

    @staticmethod
    def backward(ctx, g):
        xs, y_prev, stats, gamma, beta, watt, logits, w1 = ctx.saved_tensors
        B, T, F, _ = xs.shape
        C = w1.shape[0]
        Fy = y_prev.shape[2]
        dev = xs.device
        g = dense(g, xs.dtype)
        acc = torch.empty(8 * C + 2, dtype=torch.float64, device=dev)
        sums, dwatt, dbatt, dw1 = acc[:2 * C], acc[2 * C:6 * C], acc[6 * C:6 * C + 2], acc[6 * C + 2:]
        dxs = torch.empty_like(xs)
        dy = torch.empty_like(y_prev)
        g32, b32 = _f32c(gamma), _f32c(beta)
        w32 = _f32c(watt).view(2, 2 * C)
        w1f = _f32c(w1).view(C, 2)
        if ctx.xs2:
            nb = ctypes.c_int64(0)
            call("clskd_abf_xs2_bwd_workspace", B, T, F, C, ctypes.byref(nb))
            ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
            call("clskd_abf_xs2_bwd", g.data_ptr(), xs.data_ptr(), w1f.data_ptr(), y_prev.data_ptr(), _tag(xs.dtype),
                 B, T, F, Fy, C, stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), w32.data_ptr(),
                 logits.data_ptr(), 1 if ctx.use_batch else 0, sums.data_ptr(), dwatt.data_ptr(), dbatt.data_ptr(),
                 dw1.data_ptr(), dxs.data_ptr(), dy.data_ptr(), ws.data_ptr(), nb.value, _stream())
        else:
            call("clskd_abf_mid_xs_bwd", g.data_ptr(), xs.data_ptr(), w1f.data_ptr(), y_prev.data_ptr(), _tag(xs.dtype),
                 B, T, F, Fy, C, stats[0].data_ptr(), stats[1].data_ptr(), g32.data_ptr(), b32.data_ptr(), w32.data_ptr(),
                 logits.data_ptr(), 1 if ctx.use_batch else 0, sums.data_ptr(), dwatt.data_ptr(), dbatt.data_ptr(),
                 dw1.data_ptr(), dxs.data_ptr(), dy.data_ptr(), _stream())
        a32 = f64_to_f32(acc)
        dbeta, dgamma = a32[:C].view_as(beta), a32[C:2 * C].view_as(gamma)
        dw = a32[2 * C:6 * C].view_as(watt)
        db = a32[6 * C:6 * C + 2] if ctx.has_bias else None
        dw1_ = a32[6 * C + 2:].view_as(w1)
        return dxs, dy, dw1_, dgamma, dbeta, dw, db, None, None, None, None, None


def abf_mid_supported(z1, y_prev):
    if z1.dim() != 4 or y_prev.dim() != 4 or z1.dtype != y_prev.dtype:
        return False
    if not (z1.is_contiguous() and y_prev.is_contiguous()):
        return False
    B, T, F, C = z1.shape
    if y_prev.shape[0] != B or y_prev.shape[1] != T or y_prev.shape[3] != C:
        return False
    return bool(_lib.load().clskd_abf_mid_supported(B, T, F, y_prev.shape[2], C))


class TapSumFn(torch.autograd.Function):
    """y[b,t,f,n] = bias[n] + sum_j z[b, t+dt[j], (f+df[j])/sf, j*N+n]: the gather half of the
    tap-in-channel decomposition of a narrow convolution (see clskd_tapsum_fwd in clskd.h).
    `bias_plan` (a ConvPlan with a bias table) + bias_a / bias_b give the optional per-channel bias."""

    @staticmethod
    def forward(ctx, z, dts, dfs, N, out_size, sf, out_dtype, bias_plan, bias_a, bias_b):
        B, Ti, Fi, Zc = z.shape
        To, Fo = out_size
        y = torch.empty((B, To, Fo, N), dtype=out_dtype, device=z.device)
        n = len(dts)
        dt = (ctypes.c_int32 * n)(*dts)
        df = (ctypes.c_int32 * n)(*dfs)
        bias = None
        if bias_plan is not None and bias_a is not None:
            bias = packed_weights(bias_plan._cache, "bias", lambda: bias_plan.bias_table, _f32c(bias_a),
                                  _f32c(bias_b) if bias_b is not None else None, torch.float32)
        call("clskd_tapsum_fwd", z.data_ptr(), _tag(z.dtype), B, Ti, Fi, To, Fo, sf, Zc, n, dt, df, N, _ptr(bias),
             y.data_ptr(), _tag(y.dtype), _stream())
        ctx.meta = (B, Ti, Fi, To, Fo, sf, Zc, dts, dfs, N, z.dtype)
        ctx.bias_plan = bias_plan
        ctx.has_bias = (bias_a is not None, bias_b is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        B, Ti, Fi, To, Fo, sf, Zc, dts, dfs, N, zdt = ctx.meta
        g = dense(g)
        dz = torch.empty((B, Ti, Fi, Zc), dtype=zdt, device=g.device)
        n = len(dts)
        dt = (ctypes.c_int32 * n)(*dts)
        df = (ctypes.c_int32 * n)(*dfs)
        call("clskd_tapsum_bwd", g.data_ptr(), _tag(g.dtype), B, Ti, Fi, To, Fo, sf, Zc, n, dt, df, N, dz.data_ptr(),
             _tag(dz.dtype), _stream())
        dba = dbb = None
        plan = ctx.bias_plan
        if plan is not None and (ctx.has_bias[0] or ctx.has_bias[1]):
            s_, _ = colstats(g.view(-1, N))
            s32 = f64_to_f32(s_)
            if ctx.has_bias[0] and ctx.needs_input_grad[8]:
                dba = unpack_grads(s32, plan.bias_unpack_a, plan.bias_unpack_a.shape[0])
            if ctx.has_bias[1] and ctx.needs_input_grad[9]:
                dbb = unpack_grads(s32, plan.bias_unpack_b, plan.bias_unpack_b.shape[0])
        return dz, None, None, None, None, None, None, None, dba, dbb


class SqDiffMeanFn(torch.autograd.Function):
    """mean((a-b)^2): F.mse_loss (DCCRN.py:261, distill_MSE.py, hcl framework.py:293)."""

    @staticmethod
    def forward(ctx, a, b):
        a, b = dense(a), dense(b)
        n = a.numel()
        acc = torch.zeros(1, dtype=torch.float64, device=a.device)
        call("clskd_sqdiff_sum", a.data_ptr(), _tag(a.dtype), b.data_ptr(), _tag(b.dtype), n, acc.data_ptr(),
             _stream())
        ctx.save_for_backward(a, b)
        return f64_to_f32(acc, 1.0 / n).view(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        n = a.numel()
        g = dense(g, torch.float32)
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty_like(a)
            call("clskd_sqdiff_bwd", a.data_ptr(), _tag(a.dtype), b.data_ptr(), _tag(b.dtype), n, g.data_ptr(),
                 1.0 / n, da.data_ptr(), _tag(da.dtype), 0, _stream())
        if ctx.needs_input_grad[1]:
            db = torch.empty_like(b)
            call("clskd_sqdiff_bwd", b.data_ptr(), _tag(b.dtype), a.data_ptr(), _tag(a.dtype), n, g.data_ptr(),
                 1.0 / n, db.data_ptr(), _tag(db.dtype), 0, _stream())
        return da, db


class AdaptivePoolFn(torch.autograd.Function):
    """adaptive_avg_pool2d over the logical (F, T) plane of a physical [B,T,F,C] map -> [B,C,l,l]."""

    @staticmethod
    def forward(ctx, x, l):
        B, T, F, C = x.shape
        out = torch.empty((B, C, l, l), dtype=torch.float32, device=x.device)
        call("clskd_adaptive_pool_fwd", x.data_ptr(), _tag(x.dtype), B, T, F, C, l, out.data_ptr(), _stream())
        ctx.dims = (B, T, F, C, l, x.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        B, T, F, C, l, dt = ctx.dims
        g = dense(g, torch.float32)
        dx = torch.empty((B, T, F, C), dtype=dt, device=g.device)
        call("clskd_adaptive_pool_bwd", g.data_ptr(), B, T, F, C, l, dx.data_ptr(), _tag(dt), 0, _stream())
        return dx, None
