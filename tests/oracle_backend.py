"""TEST-ONLY binding of the host layer's backend table (sass_backend_vtbl) to the CPU oracle.

Lets the multi-rank host logic (factory, partitions, staging decomposition, all-reduce of partials, result
sink) run on CPU with gloo.  "Device" pointers are host pointers here.  The packed partial of this backend is
[atfinal sum (2 NF) | afinal (2) | a2final (2)] for every dsp type (correlation is linear in the timelines' sums).
Never imported by the product."""
import ctypes as C

import numpy as np

from oracle import oracle as o
from sassena_b200 import _host

_DSP = {0: "autocorrelate", 1: "square", 2: "plain"}


class OracleBackend:
    def __init__(self):
        self.ctxs = {}
        self.bufs = {}
        self.next = 1
        self.err = b""
        cbs = dict(
            init=_host.BE_INIT(self._init), destroy=_host.BE_DESTROY(self._destroy),
            last_error=_host.BE_LAST_ERROR(self._last_error), synchronize=_host.BE_SYNC(lambda c: 0),
            stage_frames=_host.BE_STAGE_FRAMES(self._stage_frames), frames_to_spherical=_host.BE_TO_SPH(self._to_sph),
            stage_atoms=_host.BE_STAGE_ATOMS(self._stage_atoms),
            stage_atoms_from_frames=_host.BE_STAGE_ATOMS_FF(self._stage_atoms_ff),
            set_factors=_host.BE_SET_FACTORS(self._set_factors), partial_len=_host.BE_PARTIAL_LEN(self._partial_len),
            compute_all_vectors_partial=_host.BE_COMPUTE_VEC(self._compute_all),
            compute_self_vectors_partial=_host.BE_COMPUTE_VEC(self._compute_self),
            compute_mpsphere_partial=_host.BE_COMPUTE_MP(self._compute_mp), finalize=_host.BE_FINALIZE(self._finalize),
            device_alloc=_host.BE_ALLOC(self._alloc), device_free=_host.BE_FREE(self._free),
            set_factors_batch=_host.BE_SET_FACTORS_BATCH(self._set_factors_batch),
            mpsphere_amplitudes=_host.BE_MP_AMPL(self._mp_amplitudes), mpsphere_dsp_partial=_host.BE_MP_DSP(self._mp_dsp),
            set_frame_window=_host.BE_SET_WINDOW(self._set_window),
            all_vectors_amplitudes=_host.BE_AV_AMPL(self._av_amplitudes),
            all_vectors_dsp_partial=_host.BE_AV_DSP(self._av_dsp),
            compute_all_vectors_scan_partial=_host.BE_AV_SCAN(self._av_scan),
            all_vectors_scan_amplitudes=_host.BE_AV_SCAN_AMPL(self._av_scan_amplitudes),
            stage_atoms_wave=_host.BE_STAGE_WAVE(self._stage_wave), accumulate=_host.BE_ACCUMULATE(self._accumulate),
            frames_to_cylindrical=_host.BE_TO_CYL(self._to_cyl), mpcylinder_amplitudes=_host.BE_CYL_AMPL(self._cyl_amplitudes),
            stage_atoms_prefetch=_host.BE_PREFETCH(self._prefetch), stage_atoms_swap=_host.BE_SWAP(self._swap),
            host_alloc=_host.BE_ALLOC(self._alloc), host_free=_host.BE_FREE(self._free))
        self._cbs = cbs
        self.vtbl = _host.BackendVtbl(**cbs)

    # -- helpers
    def _ctx(self, c):
        return self.ctxs[int(c)]

    def _init(self, dev, out):
        h = self.next
        self.next += 1
        self.ctxs[h] = {}
        out[0] = h
        return 0

    def _destroy(self, c):
        self.ctxs.pop(int(c or 0), None)

    def _last_error(self, c):
        return self.err

    def _stage_frames(self, c, xyz, NF, NA, repr_):
        a = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(NF, NA, 3)).copy()
        self._ctx(c).update(mode=1, xyz=a, NF=NF, NA=NA, repr=repr_, NFt=NF, f_first=0)
        return 0

    def _set_window(self, c, NF_total, f_first):
        ctx = self._ctx(c)
        if ctx.get("mode") != 1 or f_first + ctx["NF"] > NF_total:
            self.err = b"set_frame_window: bad window"
            return 1
        ctx.update(NFt=NF_total, f_first=f_first)
        return 0

    def _av_amplitudes(self, c, q, NM, out):
        ctx = self._ctx(c)
        NFt, f0, NF = ctx["NFt"], ctx["f_first"], ctx["NF"]
        qv = np.ctypeslib.as_array(q, shape=(NM, 3)).copy()
        *_, amp = o.compute_all_vectors(ctx["xyz"], ctx["b"], qv, dsp="plain", return_amplitudes=True)
        A = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(NM, NFt, 2))
        A[:] = 0
        A[:, f0:f0 + NF] = np.ascontiguousarray(amp).view(np.float64).reshape(NM, NF, 2)
        return 0

    def _scan_b(self, ctx, NQ, n):
        return ctx["bq"][n] if "bq" in ctx and len(ctx["bq"]) == NQ else ctx["b"]

    def _av_scan(self, c, v, NM, sv, NQ, dsp, ptr):
        ctx = self._ctx(c)
        plen = 2 * ctx["NFt"] + 4
        for n in range(NQ):
            p = ptr + n * plen * 8
            if NM == 0:
                self._zero(ctx, p)
                continue
            qv = sv[n] * np.ctypeslib.as_array(v, shape=(NM, 3))
            fqt, fq, fq2 = o.compute_all_vectors(ctx["xyz"], self._scan_b(ctx, NQ, n), qv, dsp=_DSP[dsp])
            self._store(ctx, p, fqt, fq, fq2, NM)
        return 0

    def _av_scan_amplitudes(self, c, v, NM, sv, NQ, out):
        ctx = self._ctx(c)
        NFt, f0, NF = ctx["NFt"], ctx["f_first"], ctx["NF"]
        A = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(NQ, NM, NFt, 2))
        A[:] = 0
        for n in range(NQ):
            qv = sv[n] * np.ctypeslib.as_array(v, shape=(NM, 3))
            *_, amp = o.compute_all_vectors(ctx["xyz"], self._scan_b(ctx, NQ, n), qv, dsp="plain", return_amplitudes=True)
            A[n, :, f0:f0 + NF] = np.ascontiguousarray(amp).view(np.float64).reshape(NM, NF, 2)
        return 0

    def _av_dsp(self, c, amp, m0, mc, dsp, ptr):
        ctx = self._ctx(c)
        NFt = ctx["NFt"]
        if mc == 0:
            self._zero(ctx, ptr)
            return 0
        nm_total = m0 + mc  # rows beyond are not touched
        A = np.ctypeslib.as_array(C.cast(amp, C.POINTER(C.c_double)), shape=(nm_total, NFt, 2))
        Ac = np.ascontiguousarray(A[m0:m0 + mc]).view(np.complex128).reshape(mc, NFt)
        fqt, fq, fq2 = o.np_dsp_store(Ac, _DSP[dsp], "fftw", norm=1.0)
        self._store(ctx, ptr, fqt, fq, fq2, 1.0)
        return 0

    def _to_sph(self, c):
        ctx = self._ctx(c)
        ctx["xyz"] = o.cart_to_spherical(ctx["xyz"])
        ctx["repr"] = 1
        return 0

    def _to_cyl(self, c, axis):
        ctx = self._ctx(c)
        ctx["xyz"] = o.cart_to_cylindrical(ctx["xyz"], [axis[0], axis[1], axis[2]])
        ctx["repr"] = 2
        return 0

    def _cyl_amplitudes(self, c, q, axis, lm, NM, a0, na, out):
        ctx = self._ctx(c)
        NF = ctx["NF"]
        mom = np.ctypeslib.as_array(lm, shape=(NM, 2)).copy()
        A = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(NM, NF, 2))
        if na == 0:
            A[:] = 0
            return 0
        *_, amp = o.compute_mpcylinder(ctx["xyz"][:, a0:a0 + na], ctx["b"][a0:a0 + na], [q[0], q[1], q[2]],
                                       [axis[0], axis[1], axis[2]], mom, dsp="plain", return_amplitudes=True)
        A[:] = amp.view(np.float64).reshape(NM, NF, 2)
        return 0

    def _stage_atoms(self, c, xyz, NA, NF):
        a = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(NA, NF, 3)).copy()
        self._ctx(c).update(mode=2, xyz=a, NF=NF, NA=NA)
        return 0

    def _stage_atoms_ff(self, c, xyz, NF, NA, nranks, rank):
        a = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(NF, NA, 3))
        off, size, _ = o.mod_assignment(nranks, rank, NA)
        ids = off + nranks * np.arange(size)
        self._ctx(c).update(mode=2, xyz=np.ascontiguousarray(a[:, ids].transpose(1, 0, 2)), NF=NF, NA=size)
        return 0

    def _stage_wave(self, c, xyz, NF, NA, first, stride, count):
        a = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(NF, NA, 3))
        ids = first + stride * np.arange(count)
        self.waves_staged = getattr(self, "waves_staged", 0) + 1
        self._ctx(c).update(mode=2, xyz=np.ascontiguousarray(a[:, ids].transpose(1, 0, 2)), NF=NF, NA=count)
        return 0

    def _prefetch(self, c, xyz, count, NF):
        ctx = self._ctx(c)
        if ctx.get("pending") is not None:
            self.err = b"a prefetched wave is waiting for the swap"
            return 5
        ctx["pending"] = np.ctypeslib.as_array(C.cast(xyz, C.POINTER(C.c_float)), shape=(count, NF, 3)).copy()
        return 0

    def _swap(self, c):
        ctx = self._ctx(c)
        a = ctx.get("pending")
        if a is None:
            self.err = b"no prefetched wave"
            return 5
        ctx["pending"] = None
        self.waves_staged = getattr(self, "waves_staged", 0) + 1
        ctx.update(mode=2, xyz=a, NF=a.shape[1], NA=a.shape[0])
        return 0

    def _accumulate(self, c, dst, src, n):
        d = np.ctypeslib.as_array(C.cast(dst, C.POINTER(C.c_double)), shape=(n,))
        d += np.ctypeslib.as_array(C.cast(src, C.POINTER(C.c_double)), shape=(n,))
        return 0

    def _set_factors(self, c, b, n):
        self._ctx(c)["b"] = np.ctypeslib.as_array(b, shape=(n,)).copy()
        return 0

    def _set_factors_batch(self, c, b, NQ, n):
        self._ctx(c)["bq"] = np.ctypeslib.as_array(b, shape=(NQ, n)).copy()
        return 0

    def _mp_amplitudes(self, c, qlens, NQ, lm, NM, a0, na, out):
        ctx = self._ctx(c)
        NF = ctx["NF"]
        ql = np.ctypeslib.as_array(qlens, shape=(NQ,)).copy()
        mom = np.ctypeslib.as_array(lm, shape=(NM, 2)).copy()
        A = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_double)), shape=(NQ, NM, NF, 2))
        for i in range(NQ):
            if na == 0:
                A[i] = 0
                continue
            b = ctx["bq"][i] if "bq" in ctx and len(ctx["bq"]) == NQ else ctx["b"]
            *_, amp = o.compute_mpsphere(ctx["xyz"][:, a0:a0 + na], b[a0:a0 + na], ql[i], mom, dsp="plain",
                                         return_amplitudes=True)
            A[i] = amp.view(np.float64).reshape(NM, NF, 2)
        return 0

    def _mp_dsp(self, c, amp, NQ, NM, dsp, ptr):
        ctx = self._ctx(c)
        NF = ctx["NF"]
        A = np.ctypeslib.as_array(C.cast(amp, C.POINTER(C.c_double)), shape=(NQ, NM, NF, 2))
        plen = 2 * NF + 4
        P = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(NQ, plen))
        for i in range(NQ):
            Ac = np.ascontiguousarray(A[i]).view(np.complex128).reshape(NM, NF)
            fqt, fq, fq2 = o.np_dsp_store(Ac, _DSP[dsp], "fftw", norm=1.0)
            P[i, :2 * NF] = np.ascontiguousarray(fqt).view(np.float64)
            P[i, 2 * NF:] = (fq.real, fq.imag, fq2.real, fq2.imag)
        return 0

    def _partial_len(self, c, dsp, out):
        ctx = self._ctx(c)
        out[0] = 2 * ctx.get("NFt", ctx["NF"]) + 4
        return 0

    def _store(self, ctx, ptr, fqt, fq, fq2, unscale):
        NF = ctx.get("NFt", ctx["NF"])
        p = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * NF + 4,))
        p[:2 * NF] = (fqt * unscale).view(np.float64)
        p[2 * NF:2 * NF + 2] = (fq.real * unscale, fq.imag * unscale)
        p[2 * NF + 2:] = (fq2.real * unscale, fq2.imag * unscale)

    def _zero(self, ctx, ptr):
        np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * ctx.get("NFt", ctx["NF"]) + 4,))[:] = 0

    def _compute_all(self, c, q, NM, dsp, ptr):
        ctx = self._ctx(c)
        if NM == 0:
            self._zero(ctx, ptr)
            return 0
        qv = np.ctypeslib.as_array(q, shape=(NM, 3)).copy()
        fqt, fq, fq2 = o.compute_all_vectors(ctx["xyz"], ctx["b"], qv, dsp=_DSP[dsp])
        self._store(ctx, ptr, fqt, fq, fq2, NM)
        return 0

    def _compute_self(self, c, q, NM, dsp, ptr):
        ctx = self._ctx(c)
        if NM == 0:
            self._zero(ctx, ptr)
            return 0
        qv = np.ctypeslib.as_array(q, shape=(NM, 3)).copy()
        fqt, fq, fq2 = o.compute_self_vectors(ctx["xyz"], ctx["b"], qv, dsp=_DSP[dsp])
        self._store(ctx, ptr, fqt, fq, fq2, NM)
        return 0

    def _compute_mp(self, c, ql, lm, NM, dsp, ptr):
        ctx = self._ctx(c)
        if NM == 0:
            self._zero(ctx, ptr)
            return 0
        mom = np.ctypeslib.as_array(lm, shape=(NM, 2)).copy()
        fqt, fq, fq2 = o.compute_mpsphere(ctx["xyz"], ctx["b"], ql, mom, dsp=_DSP[dsp])
        self._store(ctx, ptr, fqt, fq, fq2, 4 * np.pi)
        return 0

    def _finalize(self, c, ptr, dsp, method, scale, at, af, a2f):
        ctx = self._ctx(c)
        NF = ctx.get("NFt", ctx["NF"])
        p = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(2 * NF + 4,))
        out = np.ctypeslib.as_array(at, shape=(2 * NF,))
        out[:] = p[:2 * NF] * scale
        conj = dsp == 0 and method == 1
        if conj:
            out[1::2] *= -1
        af[0], af[1] = p[2 * NF] * scale, (-1 if conj else 1) * p[2 * NF + 1] * scale
        a2f[0], a2f[1] = p[2 * NF + 2] * scale, p[2 * NF + 3] * scale
        return 0

    def _alloc(self, out, nbytes):
        buf = (C.c_char * max(nbytes, 8))()
        addr = C.addressof(buf)
        self.bufs[addr] = buf
        out[0] = addr
        return 0

    def _free(self, p):
        self.bufs.pop(int(p or 0), None)
        return 0
