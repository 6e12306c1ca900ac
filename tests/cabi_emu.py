"""TEST INFRASTRUCTURE ONLY - a numpy model of the C ABI declared in include/clskd.h, operating on
HOST memory through the raw pointers the Python host code passes.

It lets the CPU test-suite exercise the host-side logic of the product (convolution plans, weight
packing tables, stride/view arithmetic, the autograd wiring, the flat-bucket optimizer and the
world_size-2 gloo data-parallel path) without a GPU.  The product never imports this module; on a
machine without the CUDA library the product raises (see clskd_b200/_lib.py).  Each function
restates the contract written in include/clskd.h, not the CUDA implementation.
"""
import ctypes
import math

import numpy as np

F32, BF16 = 0, 1


def _arr(ptr, n, dtype=np.float32):
    n = int(n)
    if n == 0:
        return np.zeros(0, dtype)
    if ptr is None or ptr == 0:
        return None
    if hasattr(ptr, "value"):
        ptr = ptr.value
    ct = {np.float32: ctypes.c_float, np.float64: ctypes.c_double, np.int32: ctypes.c_int32,
          np.int64: ctypes.c_int64, np.uint64: ctypes.c_uint64}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(int(ptr)))


def _need_f32(tag):
    if tag != F32:
        raise RuntimeError("cabi_emu models the fp32 policy only")


def _strided(ptr, shape, strides):
    """float32 view of host memory at ptr with element strides (may be 0 / overlapping)."""
    ext = 1 + sum((s - 1) * abs(st) for s, st in zip(shape, strides) if s > 0)
    base = _arr(ptr, ext)
    return np.lib.stride_tricks.as_strided(base, shape=tuple(int(s) for s in shape),
                                           strides=tuple(int(st) * 4 for st in strides), writeable=True)


def _ints(a, n):
    return [int(a[i]) for i in range(n)]


class EmuLib:
    """Attribute access returns the emulated entry point (mirrors ctypes.CDLL usage in _lib.py)."""

    def __init__(self):
        self.err = b""
        self.calls = []

    # ------------------------------------------------------------------ misc
    def clskd_last_error(self):
        return self.err

    def clskd_abi_version(self):
        return 1

    def clskd_has_tcgen05(self):
        return 0

    def clskd_tapconv_umma_c1p(self, c0, c1):
        c0p, c1p, bk0 = (c0 + 15) // 16 * 16, (c1 + 15) // 16 * 16, 64
        while bk0 > 16 and c0p % bk0:
            bk0 >>= 1
        return bk0 if (c1 and c1p < bk0) else c1p

    def clskd_tapconv_umma_supported(self, d):
        return 0

    def clskd_tapconv_wgrad_umma_supported(self, d):
        return 0

    def clskd_gram_umma_supported(self, z, dt, B, K, ldz):
        return 0

    def clskd_gram_fwd_umma(self, *a):
        raise RuntimeError("cabi_emu: the tcgen05 path has no CPU model")

    def clskd_gram_bwd_umma(self, *a):
        raise RuntimeError("cabi_emu: the tcgen05 path has no CPU model")

    def clskd_tapconv_wgrad_umma(self, dref, stream):
        raise RuntimeError("cabi_emu: the tcgen05 path has no CPU model")

    # ------------------------------------------------------------------ tapconv
    @staticmethod
    def _desc(dref):
        return dref._obj if hasattr(dref, "_obj") else dref

    @staticmethod
    def _gather(d, j):
        """im2col slab of tap j: [B, To, Fo, Ctot] (zero outside the valid input)."""
        B, To, Fo, Ti, Fi = d.B, d.To, d.Fo, d.Ti, d.Fi
        ti = np.arange(To) + d.dt[j]
        fi = np.arange(Fo) * d.sf + d.df[j]
        tv, fv = (ti >= 0) & (ti < Ti), (fi >= 0) & (fi < Fi)
        tic, fic = np.clip(ti, 0, max(Ti - 1, 0)), np.clip(fi, 0, max(Fi - 1, 0))
        parts = []
        for ptr, sB, sT, sF, c in ((d.x0, d.x0_sB, d.x0_sT, d.x0_sF, d.c0), (d.x1, d.x1_sB, d.x1_sT, d.x1_sF, d.c1)):
            if c == 0:
                continue
            X = _strided(ptr, (B, Ti, Fi, c), (sB, sT, sF, 1))
            g = X[:, tic][:, :, fic]
            g = g * (tv[None, :, None, None] & fv[None, None, :, None])
            parts.append(g)
        return np.concatenate(parts, axis=3) if len(parts) > 1 else parts[0]

    def clskd_tapconv_fwd(self, dref, stream):
        d = self._desc(dref)
        _need_f32(d.x_dtype), _need_f32(d.y_dtype)
        Ctot, N = d.c0 + d.c1, d.N
        if d.B * d.To * d.Fo == 0:
            return 0
        W = _arr(d.w, d.ntaps * Ctot * N).reshape(d.ntaps, Ctot, N)
        acc = np.zeros((d.B, d.To, d.Fo, N), np.float64)
        for j in range(d.ntaps):
            acc += self._gather(d, j).astype(np.float64) @ W[j].astype(np.float64)
        if d.bias:
            acc += _arr(d.bias, N).astype(np.float64)
        Y = _strided(d.y, (d.B, d.To, d.Fo, N), (d.y_sB, d.y_sT, d.y_sF, 1))
        if d.accumulate:
            Y += acc.astype(np.float32)
        else:
            Y[...] = acc.astype(np.float32)
        return 0

    def clskd_tapconv_fwd_umma(self, dref, stream):
        raise RuntimeError("cabi_emu: the tcgen05 path has no CPU model")

    def clskd_tapconv_wgrad(self, dref, stream):
        d = self._desc(dref)
        _need_f32(d.x_dtype), _need_f32(d.y_dtype)
        Ctot, N = d.c0 + d.c1, d.N
        dW = _arr(d.w, d.ntaps * Ctot * N).reshape(d.ntaps, Ctot, N)
        if not d.accumulate:
            dW[...] = 0
        if d.B * d.To * d.Fo == 0:
            return 0
        dY = _strided(d.y, (d.B, d.To, d.Fo, N), (d.y_sB, d.y_sT, d.y_sF, 1)).reshape(-1, N).astype(np.float64)
        for j in range(d.ntaps):
            A = self._gather(d, j).reshape(-1, Ctot).astype(np.float64)
            dW[j] += (A.T @ dY).astype(np.float32)
        return 0

    # ------------------------------------------------------------------ layout / packing
    def clskd_strided_copy4d(self, src, sdt, ss, dst, ddt, ds, shape, stream):
        _need_f32(sdt), _need_f32(ddt)
        shp, ss, ds = _ints(shape, 4), _ints(ss, 4), _ints(ds, 4)
        if 0 in shp:
            return 0
        S = _strided(src, shp, ss)
        D = _strided(dst, shp, ds)
        D[...] = S.copy()
        return 0

    def clskd_multi_pack_gather(self, desc, n_entries, total, stream):
        d = _arr(desc, 6 * int(n_entries), np.int64).reshape(int(n_entries), 6)
        for a, b, table, out, start, nn in d:
            if nn & 1:
                raise NotImplementedError("cabi_emu: bf16 packed weights")
            self.clskd_pack_gather(int(a), int(b), int(table), int(nn >> 1), int(out), 0, stream)
        return 0

    def clskd_pack_gather(self, a, b, table, n, out, out_dtype, stream):
        _need_f32(out_dtype)
        n = int(n)
        t = _arr(table, 2 * n, np.int32).reshape(n, 2)
        o = _arr(out, n)
        v = np.zeros(n, np.float32)
        for k in range(2):
            e = t[:, k]
            ok = e >= 0
            if not ok.any():
                continue
            idx = e[ok] >> 2
            selb = (e[ok] & 1).astype(bool)
            vals = np.zeros(idx.shape, np.float32)
            if (~selb).any():
                A = _arr(a, int(idx[~selb].max()) + 1)
                vals[~selb] = A[idx[~selb]]
            if selb.any():
                Bm = _arr(b, int(idx[selb].max()) + 1)
                vals[selb] = Bm[idx[selb]]
            vals = np.where(e[ok] & 2, -vals, vals)
            v[ok] += vals
        o[...] = v
        return 0

    def clskd_unpack_gather2(self, src, table2, n, dst, accumulate, stream):
        n = int(n)
        t = _arr(table2, 2 * n, np.int32).reshape(n, 2)
        o = _arr(dst, n)
        v = o.copy() if accumulate else np.zeros(n, np.float32)
        mx = int((t.max() >> 1)) + 1 if (t >= 0).any() else 0
        S = _arr(src, mx)
        for k in range(2):
            e = t[:, k]
            ok = e >= 0
            vals = S[e[ok] >> 1]
            v[ok] += np.where(e[ok] & 1, -vals, vals)
        o[...] = v
        return 0

    @staticmethod
    def _padmap(B, L, left, right, mode):
        Lp = L + left + right
        j = np.arange(Lp) - left
        if mode == 1:
            j = np.abs(j)
            j = np.where(j >= L, 2 * (L - 1) - j, j)
        valid = (j >= 0) & (j < L)
        return Lp, np.clip(j, 0, L - 1), valid

    def clskd_pad1d(self, src, sdt, sB, B, L, left, right, mode, dst, stream):
        _need_f32(sdt)
        Lp, j, valid = self._padmap(B, L, left, right, mode)
        S = _strided(src, (B, L), (sB, 1))
        D = _arr(dst, B * Lp).reshape(B, Lp)
        D[...] = S[:, j] * valid
        return 0

    def clskd_pad1d_bwd(self, ddst, B, L, left, right, mode, dsrc, accumulate, stream):
        Lp, j, valid = self._padmap(B, L, left, right, mode)
        G = _arr(ddst, B * Lp).reshape(B, Lp)
        D = _arr(dsrc, B * L).reshape(B, L)
        acc = np.zeros((B, L), np.float32)
        for p in range(Lp):
            if valid[p]:
                acc[:, j[p]] += G[:, p]
        D[...] = D + acc if accumulate else acc
        return 0

    # ------------------------------------------------------------------ BatchNorm
    def clskd_colstats(self, x, dt, M, C, s, ss, stream):
        _need_f32(dt)
        X = _arr(x, M * C).reshape(M, C).astype(np.float64)
        _arr(s, C, np.float64)[...] = X.sum(0)
        _arr(ss, C, np.float64)[...] = (X * X).sum(0)
        return 0

    def clskd_bn_finalize(self, s, ss, M, C, eps, mom, mean, invstd, rm, rv, stream):
        S, SS = _arr(s, C, np.float64), _arr(ss, C, np.float64)
        mu = S / M
        var = np.maximum(SS / M - mu * mu, 0)
        _arr(mean, C)[...] = mu
        _arr(invstd, C)[...] = 1.0 / np.sqrt(var + eps)
        if rm:
            r = _arr(rm, C)
            r[...] = (1 - mom) * r + mom * mu.astype(np.float32)
        if rv:
            r = _arr(rv, C)
            unb = var * M / (M - 1) if M > 1 else var
            r[...] = (1 - mom) * r + mom * unb.astype(np.float32)
        return 0

    def clskd_bn_eval_stats(self, rm, rv, C, eps, mean, invstd, stream):
        _arr(mean, C)[...] = _arr(rm, C)
        _arr(invstd, C)[...] = 1.0 / np.sqrt(_arr(rv, C) + np.float32(eps))
        return 0

    @staticmethod
    def _bn_terms(x, M, C, mean, invstd, gamma, beta, slope):
        X = _arr(x, M * C).reshape(M, C)
        mu, isd = _arr(mean, C), _arr(invstd, C)
        g = _arr(gamma, C) if gamma else np.ones(C, np.float32)
        b = _arr(beta, C) if beta else np.zeros(C, np.float32)
        sl = float(_arr(slope, 1)[0]) if slope else 1.0
        xh = (X - mu) * isd
        u = xh * g + b
        return X, xh, u, g, isd, sl

    def clskd_bn_act_fwd(self, x, xdt, M, C, mean, invstd, gamma, beta, slope, y, ydt, stream):
        _need_f32(xdt), _need_f32(ydt)
        _, _, u, _, _, sl = self._bn_terms(x, M, C, mean, invstd, gamma, beta, slope)
        _arr(y, M * C).reshape(M, C)[...] = np.where(u > 0, u, u * sl)
        return 0

    def clskd_bn_act_bwd_stats(self, x, xdt, dy, ddt, M, C, mean, invstd, gamma, beta, slope, s_dz, s_dzx, dsl, stream):
        _need_f32(xdt), _need_f32(ddt)
        _, xh, u, _, _, sl = self._bn_terms(x, M, C, mean, invstd, gamma, beta, slope)
        D = _arr(dy, M * C).reshape(M, C)
        dz = np.where(u > 0, D, D * sl).astype(np.float64)
        _arr(s_dz, C, np.float64)[...] = dz.sum(0)
        _arr(s_dzx, C, np.float64)[...] = (dz * xh).sum(0)
        if dsl:
            _arr(dsl, 1, np.float64)[0] = np.where(u > 0, 0, D * u).astype(np.float64).sum()
        return 0

    def clskd_bn_act_bwd_apply(self, x, xdt, dy, ddt, M, C, mean, invstd, gamma, beta, slope, s_dz, s_dzx, dsl,
                               training, dx, dxdt, dgamma, dbeta, dslope_out, stream):
        _need_f32(xdt), _need_f32(ddt), _need_f32(dxdt)
        _, xh, u, g, isd, sl = self._bn_terms(x, M, C, mean, invstd, gamma, beta, slope)
        D = _arr(dy, M * C).reshape(M, C)
        dz = np.where(u > 0, D, D * sl)
        sdz, sdzx = _arr(s_dz, C, np.float64), _arr(s_dzx, C, np.float64)
        r = dz
        if training:
            r = dz - (sdz / M).astype(np.float32) - xh * (sdzx / M).astype(np.float32)
        _arr(dx, M * C).reshape(M, C)[...] = g * isd * r
        if dgamma:
            _arr(dgamma, C)[...] = sdzx
        if dbeta:
            _arr(dbeta, C)[...] = sdz
        if dslope_out and dsl:
            _arr(dslope_out, 1)[0] = _arr(dsl, 1, np.float64)[0]
        return 0

    # ------------------------------------------------------------------ complex BN (forward)
    def clskd_cbn_moments(self, x, dt, M, Cc, s, stream):
        _need_f32(dt)
        X = _arr(x, M * 2 * Cc).reshape(M, 2 * Cc).astype(np.float64)
        xr, xi = X[:, :Cc], X[:, Cc:]
        S = _arr(s, 5 * Cc, np.float64).reshape(5, Cc)
        S[0], S[1], S[2], S[3], S[4] = xr.sum(0), xi.sum(0), (xr * xr).sum(0), (xr * xi).sum(0), (xi * xi).sum(0)
        return 0

    def clskd_cbn_finalize(self, s, M, Cc, eps, mom, training, Wrr, Wri, Wii, RMr, RMi, RVrr, RVri, RVii, coef, stream):
        if training:
            S = _arr(s, 5 * Cc, np.float64).reshape(5, Cc)
            mr, mi = S[0] / M, S[1] / M
            Vrr, Vri, Vii = S[2] / M - mr * mr, S[3] / M - mr * mi, S[4] / M - mi * mi
            if RMr:
                for ptr, new in ((RMr, mr), (RMi, mi), (RVrr, Vrr), (RVri, Vri), (RVii, Vii)):
                    r = _arr(ptr, Cc)
                    r[...] = r + mom * (new.astype(np.float32) - r)
        else:
            mr, mi, Vrr, Vri, Vii = (_arr(p, Cc).astype(np.float64) for p in (RMr, RMi, RVrr, RVri, RVii))
        Vrr, Vii = Vrr + eps, Vii + eps
        tau, delta = Vrr + Vii, Vrr * Vii - Vri * Vri
        sq = np.sqrt(delta)
        t = np.sqrt(tau + 2 * sq)
        rst = 1.0 / (sq * t)
        Urr, Uii, Uri = (sq + Vii) * rst, (sq + Vrr) * rst, -Vri * rst
        Zrr, Zri, Zir, Zii = Urr, Uri, Uri, Uii
        if Wrr:
            wrr, wri, wii = _arr(Wrr, Cc), _arr(Wri, Cc), _arr(Wii, Cc)
            Zrr, Zri = wrr * Urr + wri * Uri, wrr * Uri + wri * Uii
            Zir, Zii = wri * Urr + wii * Uri, wri * Uri + wii * Uii
        C = _arr(coef, 6 * Cc).reshape(6, Cc)
        C[0], C[1], C[2], C[3], C[4], C[5] = mr, mi, Zrr, Zri, Zir, Zii
        return 0

    def clskd_cbn_apply(self, x, xdt, M, Cc, coef, Br, Bi, y, ydt, stream):
        _need_f32(xdt), _need_f32(ydt)
        X = _arr(x, M * 2 * Cc).reshape(M, 2 * Cc)
        Y = _arr(y, M * 2 * Cc).reshape(M, 2 * Cc)
        C = _arr(coef, 6 * Cc).reshape(6, Cc)
        xr, xi = X[:, :Cc] - C[0], X[:, Cc:] - C[1]
        Y[:, :Cc] = C[2] * xr + C[3] * xi + (_arr(Br, Cc) if Br else 0)
        Y[:, Cc:] = C[4] * xr + C[5] * xi + (_arr(Bi, Cc) if Bi else 0)
        return 0

    def clskd_cbn_bwd_moments(self, x, dy, dt, M, Cc, s6, stream):
        _need_f32(dt)
        X = _arr(x, M * 2 * Cc).reshape(M, 2 * Cc).astype(np.float64)
        G = _arr(dy, M * 2 * Cc).reshape(M, 2 * Cc).astype(np.float64)
        xr, xi, gr, gi = X[:, :Cc], X[:, Cc:], G[:, :Cc], G[:, Cc:]
        S = _arr(s6, 6 * Cc, np.float64).reshape(6, Cc)
        S[0], S[1] = gr.sum(0), gi.sum(0)
        S[2], S[3], S[4], S[5] = (gr * xr).sum(0), (gr * xi).sum(0), (gi * xr).sum(0), (gi * xi).sum(0)
        return 0

    def clskd_cbn_bwd_finalize(self, s, s6, M, Cc, eps, training, Wrr, Wri, Wii, RMr, RMi, RVrr, RVri, RVii,
                               coefb, dWrr, dWri, dWii, dBr, dBi, stream):
        if training:
            S = _arr(s, 5 * Cc, np.float64).reshape(5, Cc)
            mr, mi = S[0] / M, S[1] / M
            a, b, d = S[2] / M - mr * mr, S[3] / M - mr * mi, S[4] / M - mi * mi
        else:
            mr, mi, a, b, d = (_arr(p, Cc).astype(np.float64) for p in (RMr, RMi, RVrr, RVri, RVii))
        a, d = a + np.float64(np.float32(eps)), d + np.float64(np.float32(eps))
        sq = np.sqrt(a * d - b * b)
        t = np.sqrt(a + d + 2 * sq)
        r = 1.0 / (sq * t)
        Urr, Uii, Uri = (sq + d) * r, (sq + a) * r, -b * r
        one, zero = np.ones(Cc), np.zeros(Cc)
        wrr = _arr(Wrr, Cc).astype(np.float64) if Wrr else one
        wri = _arr(Wri, Cc).astype(np.float64) if Wri else zero
        wii = _arr(Wii, Cc).astype(np.float64) if Wii else one
        Zrr, Zri = wrr * Urr + wri * Uri, wrr * Uri + wri * Uii
        Zir, Zii = wri * Urr + wii * Uri, wri * Uri + wii * Uii
        S6 = _arr(s6, 6 * Cc, np.float64).reshape(6, Cc)
        gr, gi = S6[0], S6[1]
        Arr, Ari, Air, Aii = S6[2] - mr * gr, S6[3] - mi * gr, S6[4] - mr * gi, S6[5] - mi * gi
        for ptr, val in ((dWrr, Arr * Urr + Ari * Uri), (dWri, Arr * Uri + Ari * Uii + Air * Urr + Aii * Uri),
                         (dWii, Air * Uri + Aii * Uii), (dBr, gr), (dBi, gi)):
            if ptr:
                _arr(ptr, Cc)[...] = val
        Grr = Gri = Gii = zero
        mgr = mgi = zero
        if training:
            dUrr = wrr * Arr + wri * Air
            dUri = wrr * Ari + wri * Aii + wri * Arr + wii * Air
            dUii = wri * Ari + wii * Aii
            ds = [d / (2 * sq), -b / sq, a / (2 * sq)]
            dv = []
            for k in range(3):
                dt_ = ((0.0 if k == 1 else 1.0) + 2 * ds[k]) / (2 * t)
                dr = -r * (ds[k] / sq + dt_ / t)
                dUrr_k = (ds[k] + (1.0 if k == 2 else 0.0)) * r + (sq + d) * dr
                dUii_k = (ds[k] + (1.0 if k == 0 else 0.0)) * r + (sq + a) * dr
                dUri_k = -(1.0 if k == 1 else 0.0) * r - b * dr
                dv.append(dUrr * dUrr_k + dUii * dUii_k + dUri * dUri_k)
            Grr, Gri, Gii = 2 * dv[0] / M, dv[1] / M, 2 * dv[2] / M
            mgr, mgi = gr / M, gi / M
        C = _arr(coefb, 11 * Cc).reshape(11, Cc)
        for i, v in enumerate((mr, mi, Zrr, Zir, Zri, Zii, Grr, Gri, Gii, Zrr * mgr + Zir * mgi, Zri * mgr + Zii * mgi)):
            C[i] = v
        return 0

    def clskd_cbn_bwd_apply(self, x, dy, dt, M, Cc, coefb, dx, stream):
        _need_f32(dt)
        X = _arr(x, M * 2 * Cc).reshape(M, 2 * Cc)
        G = _arr(dy, M * 2 * Cc).reshape(M, 2 * Cc)
        D = _arr(dx, M * 2 * Cc).reshape(M, 2 * Cc)
        C = _arr(coefb, 11 * Cc).reshape(11, Cc)
        xr, xi, gr, gi = X[:, :Cc] - C[0], X[:, Cc:] - C[1], G[:, :Cc], G[:, Cc:]
        D[:, :Cc] = C[2] * gr + C[3] * gi + C[6] * xr + C[7] * xi - C[9]
        D[:, Cc:] = C[4] * gr + C[5] * gi + C[7] * xr + C[8] * xi - C[10]
        return 0

    # ------------------------------------------------------------------ mask
    @staticmethod
    def _mask_inputs(spec, mask, m_sB, m_sT, B, T, nb):
        S = _arr(spec, B * T * nb * 2).reshape(B, T, nb, 2)
        Mk = _strided(mask, (B, T, nb - 1, 2), (m_sB, m_sT, 2, 1))
        return S, Mk

    def clskd_mask_fwd(self, spec, mask, mdt, m_sB, m_sT, B, T, nb, mode, out, mp, stream):
        _need_f32(mdt)
        S, Mk = self._mask_inputs(spec, mask, m_sB, m_sT, B, T, nb)
        mr = np.zeros((B, T, nb), np.float32)
        mi = np.zeros((B, T, nb), np.float32)
        mr[:, :, 1:], mi[:, :, 1:] = Mk[..., 0], Mk[..., 1]
        sr, si = S[..., 0].astype(np.float64), S[..., 1].astype(np.float64)
        O = _arr(out, B * T * nb * 2).reshape(B, T, nb, 2)
        if mode == 0:
            mags = np.sqrt(sr * sr + si * si + 1e-8)
            ph = np.arctan2(si, sr) + np.arctan2(mi.astype(np.float64), mr.astype(np.float64))
            a = np.tanh(np.sqrt(mr.astype(np.float64) ** 2 + mi.astype(np.float64) ** 2)) * mags
            O[..., 0], O[..., 1] = a * np.cos(ph), a * np.sin(ph)
        elif mode == 1:
            O[..., 0], O[..., 1] = sr * mr - si * mi, sr * mi + si * mr
        else:
            O[..., 0], O[..., 1] = sr * mr, si * mi
        if mp:
            P = _arr(mp, B * T * nb * 2).reshape(B, T, nb, 2)
            P[..., 0], P[..., 1] = mr, mi
        return 0

    def clskd_mask_bwd(self, spec, mask, mdt, m_sB, m_sT, B, T, nb, mode, dout, dmask, ddt, dm_sB, dm_sT, stream):
        _need_f32(mdt), _need_f32(ddt)
        S, Mk = self._mask_inputs(spec, mask, m_sB, m_sT, B, T, nb)
        G = _arr(dout, B * T * nb * 2).reshape(B, T, nb, 2)[:, :, 1:].astype(np.float64)
        sr, si = S[:, :, 1:, 0].astype(np.float64), S[:, :, 1:, 1].astype(np.float64)
        mr, mi = Mk[..., 0].astype(np.float64), Mk[..., 1].astype(np.float64)
        D = _strided(dmask, (B, T, nb - 1, 2), (dm_sB, dm_sT, 2, 1))
        if mode == 0:
            mags = np.sqrt(sr * sr + si * si + 1e-8)
            phs = np.arctan2(si, sr)
            mm = np.sqrt(mr * mr + mi * mi)
            safe = np.where(mm > 0, mm, 1.0)
            phm = np.arctan2(mi, mr)
            th = np.tanh(mm)
            ct, st = np.cos(phs + phm), np.sin(phs + phm)
            # out = A(mm) * (cos, sin)(phs + phm);  dmm/dm = m/mm ; dphm/dmr = -mi/mm^2, dphm/dmi = mr/mm^2
            dA = (1 - th * th) * mags
            A = th * mags
            dor_mr = dA * (mr / safe) * ct + A * (-st) * (-mi / safe ** 2)
            dor_mi = dA * (mi / safe) * ct + A * (-st) * (mr / safe ** 2)
            doi_mr = dA * (mr / safe) * st + A * ct * (-mi / safe ** 2)
            doi_mi = dA * (mi / safe) * st + A * ct * (mr / safe ** 2)
            dmr = np.where(mm > 0, G[..., 0] * dor_mr + G[..., 1] * doi_mr, 0)
            dmi = np.where(mm > 0, G[..., 0] * dor_mi + G[..., 1] * doi_mi, 0)
        elif mode == 1:
            dmr = G[..., 0] * sr + G[..., 1] * si
            dmi = -G[..., 0] * si + G[..., 1] * sr
        else:
            dmr, dmi = G[..., 0] * sr, G[..., 1] * si
        D[..., 0], D[..., 1] = dmr, dmi
        return 0

    # ------------------------------------------------------------------ overlap-add
    @staticmethod
    def _coff(window, T, win, hop):
        Lf = (T - 1) * hop + win
        c = np.zeros(Lf, np.float64)
        w2 = _arr(window, win).astype(np.float64) ** 2
        for t in range(T):
            c[t * hop:t * hop + win] += w2
        return c

    def clskd_ola_fwd(self, frames, window, B, T, win, hop, trim, do_clamp, wav, stream):
        Fr = _arr(frames, B * T * win).reshape(B, T, win).astype(np.float64)
        Lf = (T - 1) * hop + win
        acc = np.zeros((B, Lf))
        for t in range(T):
            acc[:, t * hop:t * hop + win] += Fr[:, t]
        if window:
            acc = acc / (self._coff(window, T, win, hop) + 1e-8)
        L = Lf - 2 * trim
        v = acc[:, trim:trim + L]
        if do_clamp:
            v = np.clip(v, -1, 1)
        _arr(wav, B * L).reshape(B, L)[...] = v
        return 0

    def clskd_ola_bwd(self, dwav, wav, window, B, T, win, hop, trim, do_clamp, dframes, stream):
        Lf = (T - 1) * hop + win
        L = Lf - 2 * trim
        G = _arr(dwav, B * L).reshape(B, L).astype(np.float64)
        if do_clamp:
            Wv = _arr(wav, B * L).reshape(B, L)
            G = G * ((Wv > -1) & (Wv < 1))
        full = np.zeros((B, Lf))
        full[:, trim:trim + L] = G
        if window:
            full = full / (self._coff(window, T, win, hop) + 1e-8)
        D = _arr(dframes, B * T * win).reshape(B, T, win)
        for t in range(T):
            D[:, t] = full[:, t * hop:t * hop + win]
        return 0

    # ------------------------------------------------------------------ wave losses
    def clskd_wave_loss_fwd(self, s1, s2, B, L, kind, eps, part, out, stream):
        a = _arr(s1, B * L).reshape(B, L).astype(np.float64)
        b = _arr(s2, B * L).reshape(B, L).astype(np.float64)
        P = _arr(part, 4 * B, np.float64).reshape(B, 4)
        P[...] = 0
        o = _arr(out, 1)
        if kind == 0:
            P[:, 0], P[:, 1] = (a * b).sum(1), (b * b).sum(1)
            al = (P[:, 0] / (P[:, 1] + eps))[:, None]
            tg = al * b
            P[:, 2], P[:, 3] = (tg * tg).sum(1), ((a - tg) ** 2).sum(1)
            o[0] = np.mean(10 * np.log10(P[:, 2] / (P[:, 3] + eps) + eps))
        elif kind == 1:
            P[:, 0], P[:, 1] = (a * a).sum(1), ((a - b) ** 2).sum(1)
            o[0] = np.mean(10 * np.log10(P[:, 0] ** 2 / (P[:, 1] ** 2 + eps)))
        elif kind == 2:
            P[:, 0], P[:, 1] = (a * a).sum(1), (a * b).sum(1)
            al = (P[:, 1] / P[:, 0] + eps)[:, None]
            pr = al * a
            P[:, 2], P[:, 3] = (pr * pr).sum(1), ((b - pr) ** 2).sum(1)
            o[0] = 10 * np.log10(np.mean(P[:, 2] / P[:, 3] + eps) + eps)
        else:
            P[:, 0] = ((a - b) ** 2).sum(1)
            o[0] = P[:, 0].sum() / (B * L)
        return 0

    def clskd_wave_loss_bwd(self, s1, s2, B, L, kind, eps, part, gout, ds1, ds2, stream):
        """numerical-free closed forms (derived independently of the CUDA kernel)"""
        a = _arr(s1, B * L).reshape(B, L).astype(np.float64)
        b = _arr(s2, B * L).reshape(B, L).astype(np.float64)
        g = float(_arr(gout, 1)[0])
        k10 = 10.0 / math.log(10.0)
        d1 = d2 = None
        if kind == 0:        # wrt s1 (estimate)
            ab, bb = (a * b).sum(1, keepdims=True), (b * b).sum(1, keepdims=True)
            al = ab / (bb + eps)
            tg = al * b
            en = a - tg
            tn, nn = (tg * tg).sum(1, keepdims=True), (en * en).sum(1, keepdims=True)
            R = tn / (nn + eps)
            # d tn / d a = 2 al bb/(bb+eps) b ; d nn / d a = 2 en - 2 <en,b>/(bb+eps) b
            dtn = 2 * al * bb / (bb + eps) * b
            dnn = 2 * en - 2 * (en * b).sum(1, keepdims=True) / (bb + eps) * b
            d1 = g * k10 / (R + eps) * (dtn / (nn + eps) - tn / (nn + eps) ** 2 * dnn) / B
        elif kind == 1:
            sn, dd = (a * a).sum(1, keepdims=True), ((a - b) ** 2).sum(1, keepdims=True)
            c = g * k10 / B
            d_dd = -c * 2 * dd / (dd * dd + eps)
            d1 = c * 2 / sn * 2 * a + d_dd * 2 * (a - b)
            d2 = d_dd * -2 * (a - b)
        elif kind == 2:      # wrt s2 (estimation)
            re, ab = (a * a).sum(1, keepdims=True), (a * b).sum(1, keepdims=True)
            al = ab / re + eps
            pr = al * a
            no = b - pr
            P, Nn = (pr * pr).sum(1, keepdims=True), (no * no).sum(1, keepdims=True)
            Rm = np.mean(P / Nn + eps)
            c = g * k10 / (Rm + eps) / B
            dal = a / re                                   # d al / d b
            dP = 2 * al * re * dal
            dN = 2 * no - 2 * (no * a).sum(1, keepdims=True) * dal
            d2 = c * (dP / Nn - P / Nn ** 2 * dN)
        else:
            c = g * 2.0 / (B * L)
            d1, d2 = c * (a - b), -c * (a - b)
        if ds1:
            _arr(ds1, B * L).reshape(B, L)[...] = d1
        if ds2:
            _arr(ds2, B * L).reshape(B, L)[...] = d2
        return 0

    # ------------------------------------------------------------------ STFT magnitude loss
    def clskd_stftmag_loss_fwd(self, xs, ys, n, part, out, stream):
        X = _arr(xs, 2 * n).reshape(n, 2).astype(np.float64)
        Y = _arr(ys, 2 * n).reshape(n, 2).astype(np.float64)
        xm = np.sqrt(np.maximum((X * X).sum(1), 1e-7))
        ym = np.sqrt(np.maximum((Y * Y).sum(1), 1e-7))
        P = _arr(part, 3, np.float64)
        P[0], P[1], P[2] = np.abs(np.log(ym) - np.log(xm)).sum(), ((ym - xm) ** 2).sum(), (ym * ym).sum()
        o = _arr(out, 2)
        o[0], o[1] = math.sqrt(P[1]) / math.sqrt(P[2]), P[0] / n
        return 0

    def clskd_stftmag_loss_bwd(self, xs, ys, n, part, gmag, gsc, scale_mag, scale_sc, dxs, stream):
        X = _arr(xs, 2 * n).reshape(n, 2).astype(np.float64)
        Y = _arr(ys, 2 * n).reshape(n, 2).astype(np.float64)
        P = _arr(part, 3, np.float64)
        px = (X * X).sum(1)
        xm = np.sqrt(np.maximum(px, 1e-7))
        ym = np.sqrt(np.maximum((Y * Y).sum(1), 1e-7))
        cm = float(_arr(gmag, 1)[0]) * scale_mag if gmag else 0.0
        cs = 0.0
        if gsc:
            den = math.sqrt(P[1]) * math.sqrt(P[2])
            cs = float(_arr(gsc, 1)[0]) * scale_sc / den if den > 0 else 0.0
        dm = cm * np.sign(np.log(xm) - np.log(ym)) / xm + cs * (xm - ym)
        D = dm[:, None] * X / xm[:, None]
        D[px < 1e-7] = 0
        _arr(dxs, 2 * n).reshape(n, 2)[...] = D
        return 0

    # ------------------------------------------------------------------ SPKD
    def clskd_gram_fwd(self, z, dt, B, K, ldz, G, accumulate, stream):
        _need_f32(dt)
        Z = _strided(z, (B, K), (ldz, 1)).astype(np.float64)
        Gm = _arr(G, B * B).reshape(B, B)
        r = (Z @ Z.T).astype(np.float32)
        Gm[...] = Gm + r if accumulate else r
        return 0

    def clskd_spkd_loss(self, Gt, Gs, B, scale, loss, dGs, stream):
        T_ = _arr(Gt, B * B).reshape(B, B).astype(np.float64)
        S_ = _arr(Gs, B * B).reshape(B, B).astype(np.float64)
        nt = np.maximum(np.abs(T_).sum(1, keepdims=True), 1e-12)
        ns = np.maximum(np.abs(S_).sum(1, keepdims=True), 1e-12)
        Dm = T_ / nt - S_ / ns
        _arr(loss, 1)[0] = (Dm * Dm).sum() * scale
        if dGs:
            E = -2 * scale * Dm
            rowdot = (E * S_).sum(1, keepdims=True)
            _arr(dGs, B * B).reshape(B, B)[...] = E / ns - np.sign(S_) * rowdot / ns ** 2
        return 0

    def clskd_gram_bwd(self, z, dt, B, K, ldz, dG, gout, dz, dzdt, lddz, accumulate, stream):
        _need_f32(dt), _need_f32(dzdt)
        Z = _strided(z, (B, K), (ldz, 1)).astype(np.float64)
        Dg = _arr(dG, B * B).reshape(B, B).astype(np.float64)
        g = float(_arr(gout, 1)[0]) if gout else 1.0
        r = (g * (Dg + Dg.T) @ Z).astype(np.float32)
        O = _strided(dz, (B, K), (lddz, 1))
        O[...] = O + r if accumulate else r
        return 0

    # ------------------------------------------------------------------ LSTM
    def clskd_lstm_fwd(self, pre, whh_t, T, R, Bp, H, nsets, pps, pts, pld, pss, wss, w_bf16, h, gates, c, stream):
        return self.clskd_lstm_fwd_state(pre, whh_t, T, R, Bp, H, nsets, pps, pts, pld, pss, wss, w_bf16, h, gates, c,
                                         None, None, None, None, stream)

    def clskd_lstm_fwd_state(self, pre, whh_t, T, R, Bp, H, nsets, pps, pts, pld, pss, wss, w_bf16, h, gates, c,
                             h0, c0, hN, cN, stream):
        if T == 0:
            return 0
        G = 4 * H
        P = R // Bp
        sig = lambda v: 1.0 / (1.0 + np.exp(-v))
        Hh = _arr(h, nsets * P * T * Bp * H).reshape(nsets, P, T, Bp, H)
        Gt = _arr(gates, nsets * P * T * Bp * G).reshape(nsets, P, T, Bp, G) if gates else None
        Ct = _arr(c, nsets * P * T * Bp * H).reshape(nsets, P, T, Bp, H) if c else None
        for s in range(nsets):
            W = _arr(whh_t + 4 * s * wss, H * G).reshape(H, G).astype(np.float64)
            for p in range(P):
                Pre = _strided(pre + 4 * (s * pss + p * pps), (T, Bp, G), (pts, pld, 1)).astype(np.float64)
                hh = _arr(h0, nsets * R * H).reshape(nsets, P, Bp, H)[s, p].astype(np.float64) if h0 else np.zeros((Bp, H))
                cc = _arr(c0, nsets * R * H).reshape(nsets, P, Bp, H)[s, p].astype(np.float64) if c0 else np.zeros((Bp, H))
                for t in range(T):
                    g_ = Pre[t] + hh @ W
                    i_, f_, gg, o_ = sig(g_[:, :H]), sig(g_[:, H:2 * H]), np.tanh(g_[:, 2 * H:3 * H]), sig(g_[:, 3 * H:])
                    cc = f_ * cc + i_ * gg
                    hh = o_ * np.tanh(cc)
                    Hh[s, p, t] = hh
                    if Gt is not None:
                        Gt[s, p, t] = np.concatenate([i_, f_, gg, o_], 1)
                    if Ct is not None:
                        Ct[s, p, t] = cc
                if hN:
                    _arr(hN, nsets * R * H).reshape(nsets, P, Bp, H)[s, p] = hh
                if cN:
                    _arr(cN, nsets * R * H).reshape(nsets, P, Bp, H)[s, p] = cc
        return 0

    def clskd_lstm_bwd_policy(self, dh_out, whh, gates, c, T, R, Bp, H, nsets, wss, pps, pts, pld, pss, dpre, w_bf16,
                              stream):
        return self.clskd_lstm_bwd(dh_out, whh, gates, c, T, R, Bp, H, nsets, wss, pps, pts, pld, pss, dpre, stream)

    def clskd_lstm_bwd(self, dh_out, whh, gates, c, T, R, Bp, H, nsets, wss, pps, pts, pld, pss, dpre, stream):
        G = 4 * H
        P = R // Bp
        DH = _arr(dh_out, nsets * P * T * Bp * H).reshape(nsets, P, T, Bp, H).astype(np.float64)
        Gt = _arr(gates, nsets * P * T * Bp * G).reshape(nsets, P, T, Bp, G).astype(np.float64)
        Ct = _arr(c, nsets * P * T * Bp * H).reshape(nsets, P, T, Bp, H).astype(np.float64)
        for s in range(nsets):
            W = _arr(whh + 4 * s * wss, G * H).reshape(G, H).astype(np.float64)
            for p in range(P):
                Dp = _strided(dpre + 4 * (s * pss + p * pps), (T, Bp, G), (pts, pld, 1))
                dh_rec = np.zeros((Bp, H))
                dc_next = np.zeros((Bp, H))
                for t in range(T - 1, -1, -1):
                    gi, gf, gg, go = (Gt[s, p, t][:, k * H:(k + 1) * H] for k in range(4))
                    ct = Ct[s, p, t]
                    cprev = Ct[s, p, t - 1] if t > 0 else np.zeros_like(ct)
                    dh = DH[s, p, t] + dh_rec
                    tc = np.tanh(ct)
                    do = dh * tc * go * (1 - go)
                    dc = dh * go * (1 - tc * tc) + dc_next
                    di = dc * gg * gi * (1 - gi)
                    df = dc * cprev * gf * (1 - gf)
                    dg = dc * gi * (1 - gg * gg)
                    dc_next = dc * gf
                    dp = np.concatenate([di, df, dg, do], 1)
                    Dp[t] = dp
                    dh_rec = dp @ W
        return 0

    # ------------------------------------------------------------------ ABF helpers / hcl
    def clskd_resize_f_fwd(self, x, dt, BT, Fi, Fo, C, y, stream):
        _need_f32(dt)
        X = _arr(x, BT * Fi * C).reshape(BT, Fi, C)
        idx = (np.arange(Fo) * Fi) // Fo
        _arr(y, BT * Fo * C).reshape(BT, Fo, C)[...] = X[:, idx]
        return 0

    def clskd_resize_f_bwd(self, dy, dt, BT, Fi, Fo, C, dx, stream):
        _need_f32(dt)
        D = _arr(dy, BT * Fo * C).reshape(BT, Fo, C)
        O = _arr(dx, BT * Fi * C).reshape(BT, Fi, C)
        O[...] = 0
        idx = (np.arange(Fo) * Fi) // Fo
        for fo in range(Fo):
            O[:, idx[fo]] += D[:, fo]
        return 0

    def clskd_att_blend_fwd(self, x, y, dt, z, M, C, out, stream):
        _need_f32(dt)
        X, Y = _arr(x, M * C).reshape(M, C), _arr(y, M * C).reshape(M, C)
        Z = 1.0 / (1.0 + np.exp(-_arr(z, 2 * M).reshape(M, 2).astype(np.float64)))
        _arr(out, M * C).reshape(M, C)[...] = X * Z[:, :1] + Y * Z[:, 1:]
        return 0

    def clskd_att_blend_bwd(self, x, y, dt, z, dout, M, C, dx, dy, dz, stream):
        _need_f32(dt)
        X, Y = _arr(x, M * C).reshape(M, C), _arr(y, M * C).reshape(M, C)
        G = _arr(dout, M * C).reshape(M, C).astype(np.float64)
        Z = 1.0 / (1.0 + np.exp(-_arr(z, 2 * M).reshape(M, 2).astype(np.float64)))
        _arr(dx, M * C).reshape(M, C)[...] = G * Z[:, :1]
        _arr(dy, M * C).reshape(M, C)[...] = G * Z[:, 1:]
        DZ = _arr(dz, 2 * M).reshape(M, 2)
        DZ[:, 0] = (G * X).sum(1) * Z[:, 0] * (1 - Z[:, 0])
        DZ[:, 1] = (G * Y).sum(1) * Z[:, 1] * (1 - Z[:, 1])
        return 0

    # ------------------------------------------------------------------ fused ABF middle stage
    def clskd_abf_mid_supported(self, B, T, F, Fy, C):
        tpr = C // 8
        return int(C % 8 == 0 and 8 <= C <= 256 and (tpr & (tpr - 1)) == 0 and (Fy == F or 2 * Fy == F) and F % 2 == 0)

    @staticmethod
    def _abf_terms(z1, y, B, T, F, Fy, C, mean, invstd, gamma, beta, watt):
        Z = _arr(z1, B * T * F * C).reshape(B, T, F, C).astype(np.float64)
        Y = _arr(y, B * T * Fy * C).reshape(B, T, Fy, C).astype(np.float64)
        idx = (np.arange(F) * Fy) // F
        yv = Y[:, :, idx]
        mu, isd = _arr(mean, C).astype(np.float64), _arr(invstd, C).astype(np.float64)
        g = _arr(gamma, C).astype(np.float64) if gamma else np.ones(C)
        b = _arr(beta, C).astype(np.float64) if beta else np.zeros(C)
        xh = (Z - mu) * isd
        xp = xh * g + b
        W = _arr(watt, 4 * C).reshape(2, 2 * C).astype(np.float64)
        return xh, xp, yv, g, isd, W, idx

    def clskd_abf_mid_fwd(self, z1, y, dt, B, T, F, Fy, C, mean, invstd, gamma, beta, watt, batt, xb, logits, stream):
        _need_f32(dt)
        xh, xp, yv, g, isd, W, idx = self._abf_terms(z1, y, B, T, F, Fy, C, mean, invstd, gamma, beta, watt)
        lg = xp @ W[:, :C].T + yv @ W[:, C:].T + (_arr(batt, 2).astype(np.float64) if batt else 0)
        s = 1.0 / (1.0 + np.exp(-lg))
        _arr(xb, B * T * F * C).reshape(B, T, F, C)[...] = xp * s[..., :1] + yv * s[..., 1:]
        _arr(logits, B * T * F * 2).reshape(B, T, F, 2)[...] = lg
        return 0

    def clskd_abf_mid_bwd(self, gout, z1, y, dt, B, T, F, Fy, C, mean, invstd, gamma, beta, watt, logits, training,
                          sums, dwatt, dbatt, dz1, dy, stream):
        _need_f32(dt)
        xh, xp, yv, g, isd, W, idx = self._abf_terms(z1, y, B, T, F, Fy, C, mean, invstd, gamma, beta, watt)
        G = _arr(gout, B * T * F * C).reshape(B, T, F, C).astype(np.float64)
        lg = _arr(logits, B * T * F * 2).reshape(B, T, F, 2).astype(np.float64)
        s = 1.0 / (1.0 + np.exp(-lg))
        dl = np.stack([(G * xp).sum(-1), (G * yv).sum(-1)], -1) * s * (1 - s)
        dxp = G * s[..., :1] + dl @ W[:, :C]
        dyv = G * s[..., 1:] + dl @ W[:, C:]
        M = B * T * F
        S = _arr(sums, 2 * C, np.float64).reshape(2, C)
        S[0], S[1] = dxp.reshape(M, C).sum(0), (dxp * xh).reshape(M, C).sum(0)
        DW = _arr(dwatt, 4 * C, np.float64).reshape(2, 2 * C)
        DW[:, :C] = dl.reshape(M, 2).T @ xp.reshape(M, C)
        DW[:, C:] = dl.reshape(M, 2).T @ yv.reshape(M, C)
        _arr(dbatt, 2, np.float64)[...] = dl.reshape(M, 2).sum(0)
        r = dxp - (S[0] / M + xh * (S[1] / M) if training else 0)
        _arr(dz1, M * C).reshape(B, T, F, C)[...] = g * isd * r
        DY = _arr(dy, B * T * Fy * C).reshape(B, T, Fy, C)
        acc = np.zeros((B, T, Fy, C))
        for fo in range(F):
            acc[:, :, idx[fo]] += dyv[:, :, fo]
        DY[...] = acc
        return 0

    def clskd_tapsum_fwd(self, z, zdt, B, Ti, Fi, To, Fo, sf, Zc, ntaps, dt, df, N, bias, y, ydt, stream):
        _need_f32(zdt), _need_f32(ydt)
        Z = _arr(z, B * Ti * Fi * Zc).reshape(B, Ti, Fi, Zc)
        Y = _arr(y, B * To * Fo * N).reshape(B, To, Fo, N)
        acc = np.zeros((B, To, Fo, N), np.float32)
        for j in range(ntaps):
            for t in range(To):
                tt = t + dt[j]
                if tt < 0 or tt >= Ti:
                    continue
                for f in range(Fo):
                    ff = f + df[j]
                    if ff < 0 or ff % sf:
                        continue
                    ff //= sf
                    if ff < Fi:
                        acc[:, t, f] += Z[:, tt, ff, j * N:(j + 1) * N]
        Y[...] = acc + (_arr(bias, N) if bias else 0)
        return 0

    def clskd_tapsum_bwd(self, dy, ddt, B, Ti, Fi, To, Fo, sf, Zc, ntaps, dt, df, N, dz, zdt, stream):
        _need_f32(ddt), _need_f32(zdt)
        G = _arr(dy, B * To * Fo * N).reshape(B, To, Fo, N)
        D = _arr(dz, B * Ti * Fi * Zc).reshape(B, Ti, Fi, Zc)
        D[...] = 0
        for j in range(ntaps):
            for t in range(Ti):
                tt = t - dt[j]
                if tt < 0 or tt >= To:
                    continue
                for f in range(Fi):
                    ff = f * sf - df[j]
                    if 0 <= ff < Fo:
                        D[:, t, f, j * N:(j + 1) * N] = G[:, tt, ff]
        return 0

    @staticmethod
    def _pool_bounds(n, l):
        return [((i * n) // l, ((i + 1) * n + l - 1) // l) for i in range(l)]

    def clskd_adaptive_pool_fwd(self, x, dt, B, T, F, C, l, out, stream):
        _need_f32(dt)
        X = _arr(x, B * T * F * C).reshape(B, T, F, C).astype(np.float64)
        O = _arr(out, B * C * l * l).reshape(B, C, l, l)
        for i, (h0, h1) in enumerate(self._pool_bounds(F, l)):
            for j, (w0, w1) in enumerate(self._pool_bounds(T, l)):
                O[:, :, i, j] = X[:, w0:w1, h0:h1].mean((1, 2))
        return 0

    def clskd_adaptive_pool_bwd(self, dout, B, T, F, C, l, dx, dt, accumulate, stream):
        _need_f32(dt)
        G = _arr(dout, B * C * l * l).reshape(B, C, l, l)
        D = _arr(dx, B * T * F * C).reshape(B, T, F, C)
        acc = np.zeros((B, T, F, C), np.float32)
        for i, (h0, h1) in enumerate(self._pool_bounds(F, l)):
            for j, (w0, w1) in enumerate(self._pool_bounds(T, l)):
                acc[:, w0:w1, h0:h1] += (G[:, :, i, j] / ((h1 - h0) * (w1 - w0)))[:, None, None, :]
        D[...] = D + acc if accumulate else acc
        return 0

    def clskd_sqdiff_sum(self, a, adt, b, bdt, n, out, stream):
        _need_f32(adt), _need_f32(bdt)
        d = _arr(a, n).astype(np.float64) - _arr(b, n)
        _arr(out, 1, np.float64)[0] += (d * d).sum()
        return 0

    def clskd_sqdiff_bwd(self, a, adt, b, bdt, n, gout, scale, da, dadt, accumulate, stream):
        _need_f32(adt), _need_f32(bdt), _need_f32(dadt)
        g = float(_arr(gout, 1)[0]) * scale * 2
        v = g * (_arr(a, n) - _arr(b, n))
        D = _arr(da, n)
        D[...] = D + v if accumulate else v
        return 0

    # ------------------------------------------------------------------ optimizer / utilities
    def clskd_adam_step(self, p, g, m, v, n, lr, b1, b2, eps, wd, step, gscale, stream):
        P, G_, M_, V_ = _arr(p, n), _arr(g, n), _arr(m, n), _arr(v, n)
        gi = G_ * np.float32(gscale)
        if wd != 0:
            gi = gi + np.float32(wd) * P
        M_[...] = np.float32(b1) * M_ + np.float32(1 - b1) * gi
        V_[...] = np.float32(b2) * V_ + np.float32(1 - b2) * gi * gi
        bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
        P[...] = P - np.float32(lr / bc1) * (M_ / (np.sqrt(V_) / np.float32(math.sqrt(bc2)) + np.float32(eps)))
        return 0

    def clskd_multi_pack_f32(self, ptrs, offsets, n, total, flat, stream):
        Pp = _arr(ptrs, n, np.uint64)
        Of = _arr(offsets, n + 1, np.int64)
        Fl = _arr(flat, total)
        for i in range(n):
            lo, hi = int(Of[i]), int(Of[i + 1])
            Fl[lo:hi] = _arr(int(Pp[i]), hi - lo) if int(Pp[i]) else 0
        return 0

    def clskd_fill_f32(self, p, n, v, stream):
        _arr(p, n)[...] = v
        return 0

    def clskd_axpy_f32(self, y, x, n, a, stream):
        Y = _arr(y, n)
        Y[...] = Y + np.float32(a) * _arr(x, n)
        return 0

    def clskd_axpby_f32(self, x, y, a, b, out, n, stream):
        r = np.float32(a) * _arr(x, n)
        if y:
            r = r + np.float32(b) * _arr(y, n)
        _arr(out, n)[...] = r
        return 0

    def clskd_pair_moments(self, a, b, B, L, a_sB, b_sB, out, stream):
        O = _arr(out, B * 5, np.float64).reshape(B, 5)
        for i in range(B):
            x = _arr(a + 4 * i * a_sB, L).astype(np.float64)
            y = _arr(b + 4 * i * b_sB, L).astype(np.float64)
            O[i] = [x.sum(), y.sum(), (x * x).sum(), (x * y).sum(), (y * y).sum()]
        return 0

    def clskd_sum_n(self, in0, in1, in2, in3, k, dtype, n, out, stream):
        _need_f32(dtype)
        r = _arr(in0, n) + _arr(in1, n)
        if k > 2:
            r = r + _arr(in2, n)
        if k > 3:
            r = r + _arr(in3, n)
        _arr(out, n)[...] = r
        return 0

    def clskd_f64_to_f32(self, inp, n, scale, out, stream):
        _arr(out, n)[...] = _arr(inp, n, np.float64) * scale
        return 0


def install(monkeypatch=None):
    """Route clskd_b200's C-ABI calls to the numpy model and lift the CUDA-only guards.
    Returns the EmuLib instance.  With `monkeypatch` (pytest) everything is undone at teardown."""
    import clskd_b200
    from clskd_b200 import _lib, ops
    emu = EmuLib()

    def setattr_(obj, name, val):
        if monkeypatch is not None:
            monkeypatch.setattr(obj, name, val)
        else:
            setattr(obj, name, val)
    setattr_(_lib, "_lib", emu)
    setattr_(_lib, "load", lambda: emu)
    setattr_(ops, "_require_cuda", lambda *ts: None)
    setattr_(ops, "_stream", lambda: 0)
    return emu
