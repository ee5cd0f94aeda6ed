"""Python face of the C++ host layer (csrc/host): Params, ScatterDeviceFactory::create + IScatterDevice::run,
communicator adapters for torch.distributed, and the pure host logic.  No compute happens in Python."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _host
from ._lib import load_library


class HostError(RuntimeError):
    pass


def _lib():
    return load_library()


def _ck(rc):
    if rc:
        raise HostError(_lib().sass_last_error().decode(errors="replace"))


def _ctxp(ctx):
    """sgpu_ctx* from None, an int, a c_void_p or a ScatterContext."""
    if ctx is None:
        return None
    if hasattr(ctx, "h"):
        ctx = ctx.h
    if isinstance(ctx, C.c_void_p):
        return ctx
    return C.c_void_p(int(ctx))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


# ---------------------------------------------------------------------------------------------------------------
# host logic
# ---------------------------------------------------------------------------------------------------------------
def motion_transforms(kind, n, displace=0.0, frequency=0.001, radius=0.0, seed=0, sampling=1, direction=(1, 0, 0)):
    """4x4 transforms (row vectors: (x, y, z, 1) @ T) of frames 0..n-1 of a sample.motions walker (motion_walker.cpp)."""
    d = np.ascontiguousarray(direction, dtype=np.float64)
    out = np.zeros((n, 4, 4))
    _ck(_lib().sass_motion_transforms(kind.encode(), float(displace), float(frequency), float(radius), int(seed), int(sampling),
                                      _dp(d), n, _dp(out)))
    return out


def div_assignment(NN, rank, NAF):
    o, s, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _ck(_lib().sass_div_assignment(NN, rank, NAF, C.byref(o), C.byref(s), C.byref(m)))
    return o.value, s.value, m.value


def mod_assignment(NN, rank, NAF):
    o, s, m = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _ck(_lib().sass_mod_assignment(NN, rank, NAF, C.byref(o), C.byref(s), C.byref(m)))
    return o.value, s.value, m.value


def decomposition_plan(nn, nq, naf, elbytes, maxbytes, utilization=0.95, automatic=True, manual_size=1):
    p, ps, pen = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _ck(_lib().sass_decomposition_plan(nn, nq, naf, elbytes, maxbytes, utilization, int(automatic), manual_size,
                                       C.byref(p), C.byref(ps), C.byref(pen)))
    return p.value, ps.value, pen.value


def create_from_scans(scans):
    """scans: list of dict(base, from, to, points, exponent) -> q-vectors [N][3] (parameters.cpp:1125-1189)."""
    rows = np.array([[*s.get("base", (1, 0, 0)), s.get("from", 0.0), s.get("to", 1.0), s.get("points", 100),
                      s.get("exponent", 1.0)] for s in scans], dtype=np.float64).reshape(-1, 7)
    n = _lib().sass_create_from_scans(_dp(rows), len(rows), None, 0)
    if n == 0 and _lib().sass_last_error():
        if len(scans) > 3:
            raise HostError(_lib().sass_last_error().decode(errors="replace"))
    out = np.zeros((max(n, 1), 3))
    _lib().sass_create_from_scans(_dp(rows), len(rows), _dp(out), n)
    return out[:n]


class DCDFile:
    """DCDFrameset (frames.cpp:272-436) with the frameset's first/last/stride trimming."""

    def __init__(self, path, first=0, last=None, stride=1):
        h = C.c_void_p()
        _ck(_lib().sass_dcd_open(str(path).encode(), first, 0 if last is None else last, 0 if last is None else 1,
                                 stride, C.byref(h)))
        self.h = h
        nf, na, uc = C.c_size_t(), C.c_size_t(), C.c_int()
        _ck(_lib().sass_dcd_info(self.h, C.byref(nf), C.byref(na), C.byref(uc)))
        self.number_of_frames, self.number_of_atoms, self.has_unitcell = nf.value, na.value, bool(uc.value)

    def read(self, first=0, count=None):
        count = self.number_of_frames - first if count is None else count
        out = np.empty((count, self.number_of_atoms, 3), dtype=np.float32)
        _ck(_lib().sass_dcd_read(self.h, first, count, out.ctypes.data))
        return out

    def close(self):
        if self.h:
            _lib().sass_dcd_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class XdrFile:
    """XTCFrameset / TRRFrameset (frames.cpp:592-858): GROMACS trajectories, coordinates in Angstrom."""

    def __init__(self, path, format=None, first=0, last=None, stride=1):
        if format is None:
            format = "trr" if str(path).lower().endswith(".trr") else "xtc"
        h = C.c_void_p()
        _ck(_lib().sass_xdr_open(str(path).encode(), format.encode(), first, 0 if last is None else last,
                                 0 if last is None else 1, stride, C.byref(h)))
        self.h = h
        nf, na = C.c_size_t(), C.c_size_t()
        _ck(_lib().sass_xdr_info(self.h, C.byref(nf), C.byref(na)))
        self.number_of_frames, self.number_of_atoms = nf.value, na.value

    def read(self, first=0, count=None, with_box=False):
        count = self.number_of_frames - first if count is None else count
        out = np.empty((count, self.number_of_atoms, 3), dtype=np.float32)
        box = np.zeros((count, 3, 3)) if with_box else None
        _ck(_lib().sass_xdr_read(self.h, first, count, out.ctypes.data, box.ctypes.data if with_box else None))
        return (out, box) if with_box else out

    def close(self):
        if self.h:
            _lib().sass_xdr_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def write_dcd(path, xyz):
    """xyz float32 [NF][NA][3] -> CHARMM DCD in the layout of the reference's DCDCoordinateWriter."""
    a = np.ascontiguousarray(xyz, dtype=np.float32)
    _ck(_lib().sass_dcd_write(str(path).encode(), a.ctypes.data, a.shape[0], a.shape[1]))


class Params:
    """The part of the scatter.xml Params the hot path reads; keys are the XML paths."""

    def __init__(self, **kv):
        self.h = C.c_void_p(_lib().sass_params_new())
        for k, v in kv.items():
            self.set(k.replace("__", "."), v)

    def __del__(self):
        try:
            if self.h and not getattr(self, "_borrowed", False):
                _lib().sass_params_free(self.h)
                self.h = None
        except Exception:
            pass

    def set(self, key, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        _ck(_lib().sass_params_set(self.h, key.encode(), str(value).encode()))
        return self

    def set_vectors(self, v):
        v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1, 3)
        _ck(_lib().sass_params_set_vectors(self.h, _dp(v), len(v)))
        return self

    def set_moments(self, lm):
        lm = np.ascontiguousarray(lm, dtype=np.int64).reshape(-1, 2)
        _ck(_lib().sass_params_set_moments(self.h, lm.ctypes.data_as(C.POINTER(C.c_long)), len(lm)))
        return self

    def create(self):
        _ck(_lib().sass_params_create(self.h))
        return self

    @property
    def vectors(self):
        n = _lib().sass_params_num_vectors(self.h)
        out = np.zeros((max(n, 1), 3))
        _ck(_lib().sass_params_get_vectors(self.h, _dp(out)))
        return out[:n]

    @property
    def moments(self):
        n = _lib().sass_params_num_moments(self.h)
        out = np.zeros((max(n, 1), 2), dtype=np.int64)
        _ck(_lib().sass_params_get_moments(self.h, out.ctypes.data_as(C.POINTER(C.c_long))))
        return out[:n]

    def init_subvectors(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        n = _lib().sass_init_subvectors(self.h, _dp(q), None, 0)
        out = np.zeros((max(n, 1), 3))
        _lib().sass_init_subvectors(self.h, _dp(q), _dp(out), n)
        return out[:n]


# ---------------------------------------------------------------------------------------------------------------
# communicators
# ---------------------------------------------------------------------------------------------------------------
class _DevPtr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class TorchDistCommunicator:
    """boost::mpi::communicator stand-in over torch.distributed.  device_memory=True: buffers are CUDA pointers and
    the group is NCCL; False: host pointers (gloo), used by the CPU tests of the host logic.
    The C++ side copies the callback table into the communicators it derives with split(), replacing only `user`,
    so every callback dispatches on the `user` handle through the registry."""

    _registry = {}
    _next = [1]

    def __init__(self, group=None, device_memory=True):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.device_memory = device_memory
        self.handle = self._next[0]
        self._next[0] += 1
        self._registry[self.handle] = self
        cls = TorchDistCommunicator
        self._cbs = (_host.RANK_FN(cls._rank), _host.SIZE_FN(cls._size), _host.ALLREDUCE_FN(cls._allreduce),
                     _host.BARRIER_FN(cls._barrier), _host.SPLIT_FN(cls._split), _host.RELEASE_FN(cls._release))
        self.vtbl = _host.CommVtbl(C.c_void_p(self.handle), *self._cbs)

    @classmethod
    def _get(cls, user):
        return cls._registry[int(user)]

    @staticmethod
    def _rank(user):
        self = TorchDistCommunicator._get(user)
        return self.dist.get_rank(self.group)

    @staticmethod
    def _size(user):
        self = TorchDistCommunicator._get(user)
        return self.dist.get_world_size(self.group)

    @staticmethod
    def _allreduce(user, ptr, n):
        try:
            import torch
            self = TorchDistCommunicator._get(user)
            if self.device_memory:
                t = torch.as_tensor(_DevPtr(ptr, n), device="cuda")
                self.dist.all_reduce(t, group=self.group)
                torch.cuda.synchronize()
            else:
                a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n,))
                t = torch.from_numpy(a)
                self.dist.all_reduce(t, group=self.group)
            return 0
        except Exception as e:  # pragma: no cover
            print("allreduce callback failed:", e)
            return 1

    @staticmethod
    def _barrier(user):
        try:
            self = TorchDistCommunicator._get(user)
            self.dist.barrier(group=self.group)
            return 0
        except Exception as e:  # pragma: no cover
            print("barrier callback failed:", e)
            return 1

    @staticmethod
    def _split(user, color):
        # every rank of this communicator calls split: gather the colors, create one group per color in the same
        # order everywhere.  torch.distributed.new_group must be entered by ALL ranks of the job, so split() is only
        # valid when called collectively by the whole job (true for the factory: scatter_device_factory.cpp:116,131).
        try:
            self = TorchDistCommunicator._get(user)
            world = self.dist.get_world_size(self.group)
            colors = [None] * world
            self.dist.all_gather_object(colors, (int(color), self.dist.get_rank()), group=self.group)
            # all communicators of the same generation must create the same groups: share the colour table job-wide
            job = [None] * self.dist.get_world_size()
            self.dist.all_gather_object(job, sorted(colors))
            mine = None
            seen = []
            for table in job:
                if table in seen:
                    continue
                seen.append(table)
                for c in sorted(set(col for col, _ in table)):
                    ranks = sorted(gr for col, gr in table if col == c)
                    g = self.dist.new_group(ranks=ranks)
                    if table == sorted(colors) and c == int(color):
                        mine = g
            child = TorchDistCommunicator(mine, self.device_memory)
            return child.handle
        except Exception as e:  # pragma: no cover
            print("split callback failed:", e)
            return None

    @staticmethod
    def _release(user):
        TorchDistCommunicator._registry.pop(int(user), None)


class NcclCommunicator:
    """the library's own communicator on NCCL (csrc/host/nccl_comm.cpp): all callbacks are native code, torch.distributed (or
    anything else) is only used to hand the 128-byte unique id from rank 0 to the others.  With it the frame-sharded
    coherent device exchanges amplitudes inside the library (ncclSend/ncclRecv on the device streams)."""

    def __init__(self, unique_id: bytes, nranks: int, rank: int, device: int):
        assert len(unique_id) == 128
        self.vtbl = _host.CommVtbl()
        if _lib().sass_comm_nccl_create(unique_id, nranks, rank, device, C.byref(self.vtbl)) != 0:
            raise RuntimeError("sass_comm_nccl_create failed (libnccl.so.2 not loadable, or ncclCommInitRank failed)")

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if _lib().sass_comm_nccl_unique_id(buf) != 0:
            raise RuntimeError("sass_comm_nccl_unique_id failed")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int):
        """bootstrap over an initialised torch.distributed job: rank 0's id is broadcast, then ncclCommInitRank"""
        import torch.distributed as dist
        box = [cls.unique_id() if dist.get_rank() == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(box[0], dist.get_world_size(), dist.get_rank(), device)

    def close(self):
        if getattr(self, "vtbl", None) is not None and self.vtbl.user:
            self.vtbl.release(self.vtbl.user)
            self.vtbl.user = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------------
# run
# ---------------------------------------------------------------------------------------------------------------
def run_scatter(params: Params, frames, qvectors, b=None, factors_fn=None, comm: TorchDistCommunicator | None = None,
                backend=None, ctx=None):
    """ScatterDeviceFactory::create(...)->run().  frames: float32 [NF][NA][3]; qvectors [NQ][3].
    Returns (records, has_device, timers): records = list of dict(q, fqt, fq0, fq, fq2) written by partition rank 0."""
    frames = np.ascontiguousarray(frames, dtype=np.float32)
    NF, NA, _ = frames.shape
    q = np.ascontiguousarray(qvectors, dtype=np.float64).reshape(-1, 3)
    records = []

    def _write(user, qp, fqt, nf, fq, fq2):
        records.append({"q": np.array([qp[0], qp[1], qp[2]]),
                        "fqt": np.ctypeslib.as_array(fqt, shape=(2 * nf,)).copy().view(np.complex128),
                        "fq": complex(fq[0], fq[1]), "fq2": complex(fq2[0], fq2[1])})
        records[-1]["fq0"] = records[-1]["fqt"][0]

    wcb = _host.WRITE_FN(_write)
    if factors_fn is not None:
        def _factors(user, ql, bp, na):
            np.ctypeslib.as_array(bp, shape=(na,))[:] = factors_fn(ql)
        fcb = _host.FACTORS_FN(_factors)
        bptr = None
    else:
        fcb = C.cast(None, _host.FACTORS_FN)
        barr = np.ascontiguousarray(b, dtype=np.float64)
        bptr = _dp(barr)
    has = C.c_int(0)
    timers = C.create_string_buffer(4096)
    rc = _lib().sass_scatter_run(params.h, C.byref(comm.vtbl) if comm is not None else None,
                                 C.byref(backend) if backend is not None else None,
                                 _ctxp(ctx), NA, NF, frames.ctypes.data, bptr, fcb, None,
                                 _dp(q), len(q), wcb, None, C.byref(has), timers, len(timers))
    _ck(rc)
    tm = {}
    for item in timers.value.decode().split(";"):
        if "=" in item:
            k, v = item.split("=")
            s, c = v.split(":")
            tm[k] = (float(s), int(c))
    return records, bool(has.value), tm


# ---------------------------------------------------------------------------------------------------------------
# control plane: scatter.xml -> signal directory (the `sassena` executable's flow, src/main/sassena.cpp:132-417)
# ---------------------------------------------------------------------------------------------------------------
class Job:
    """scatter.xml + db.xml + PDB + DCD loaded by the native control plane (csrc/host/control.cpp)."""

    def __init__(self, config_file, overwrites=None):
        """overwrites: {key: value} of the reference's command-line overwrite options (parameters.cpp:795-875), e.g.
        {"stager.target": "carbons", "scattering.signal.file": "run2.h5"}; applied after the configuration is read."""
        h = C.c_void_p()
        kv = [(str(k).encode(), (("true" if v else "false") if isinstance(v, bool) else str(v)).encode())
              for k, v in (overwrites or {}).items()]
        keys = (C.c_char_p * max(len(kv), 1))(*[k for k, _ in kv])
        vals = (C.c_char_p * max(len(kv), 1))(*[v for _, v in kv])
        _ck(_lib().sass_job_load_overwrite(str(config_file).encode(), keys, vals, len(kv), C.byref(h)))
        self.h = h
        n = [C.c_size_t() for _ in range(4)]
        _ck(_lib().sass_job_info(self.h, *[C.byref(x) for x in n]))
        self.natoms, self.ntarget, self.nframes, self.nqvectors = [x.value for x in n]

    @property
    def signal_file(self):
        """scattering.signal.file resolved like the reference does (default: signal.h5 next to the configuration)"""
        return _lib().sass_job_signal_file(self.h).decode()

    def option(self, key):
        """value in effect of one of the overwritable options (file names resolved); None for any other key"""
        v = _lib().sass_job_option(self.h, key.encode())
        return None if v is None else v.decode()

    def close(self):
        if getattr(self, "h", None):
            _lib().sass_job_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def qvectors(self):
        q = np.empty((self.nqvectors, 3))
        _ck(_lib().sass_job_qvectors(self.h, _dp(q)))
        return q

    def factors(self, ql):
        b = np.empty(self.ntarget)
        _ck(_lib().sass_job_factors(self.h, float(ql), _dp(b)))
        return b

    def frames(self):
        p = C.POINTER(C.c_float)()
        _ck(_lib().sass_job_frames(self.h, C.byref(p)))
        return np.ctypeslib.as_array(p, shape=(self.nframes, self.ntarget, 3)).copy()

    def selection(self, name):
        n = C.c_size_t()
        _ck(_lib().sass_job_selection(self.h, name.encode(), None, 0, C.byref(n)))
        ids = np.empty(n.value, dtype=np.uintp)
        _ck(_lib().sass_job_selection(self.h, name.encode(), ids.ctypes.data_as(_host.c_size_p), n.value, C.byref(n)))
        return ids.astype(np.int64)

    def params(self):
        """Borrowed Params view (valid while the job lives)."""
        p = Params.__new__(Params)
        p.h = C.c_void_p(_lib().sass_job_params(self.h))
        p._borrowed = True
        return p

    def run(self, signal_dir, comm: TorchDistCommunicator | None = None, backend=None, ctx=None):
        """Runs every q-vector and writes <signal_dir>/{qvectors,fqt,fq0,fq,fq2}.npy.  Returns (written, report)."""
        n = C.c_size_t()
        rep = C.create_string_buffer(1024)
        _ck(_lib().sass_job_run(self.h, str(signal_dir).encode(), C.byref(comm.vtbl) if comm is not None else None,
                                C.byref(backend) if backend is not None else None, _ctxp(ctx),
                                C.byref(n), rep, len(rep)))
        return n.value, rep.value.decode()


    def stage(self, comm: TorchDistCommunicator | None = None, backend=None, ctx=None):
        """The reference's `s_stage` executable (s_stage.cpp:205-232): stages stager.target the way stager.mode says and,
        with stager.dump, writes the staged coordinates to stager.file.  Returns (bytes staged on this rank, report)."""
        n = C.c_size_t()
        rep = C.create_string_buffer(1024)
        _ck(_lib().sass_job_stage(self.h, C.byref(comm.vtbl) if comm is not None else None,
                                  C.byref(backend) if backend is not None else None, _ctxp(ctx), C.byref(n), rep, len(rep)))
        return n.value, rep.value.decode()


H5_DATASET_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_char_p, C.c_int, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                            C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t)


def read_h5(path, with_layout=False):
    """Reads an HDF5 file through the built-in minimal reader (csrc/host/h5mini.cpp): dict path -> float64 array /
    bytes (char arrays) / str (strings).  with_layout=True returns (values, {path: dict(maxdims, chunk)})."""
    out, layout = {}, {}

    def _cb(user, name, kind, rank, dims, maxdims, chunk, data, nbytes):
        shape = tuple(int(dims[i]) for i in range(rank))
        raw = C.string_at(data, nbytes) if nbytes else b""
        key = name.decode(errors="replace")
        if kind == 0:
            out[key] = np.frombuffer(raw, dtype=np.float64).reshape(shape).copy()
        elif kind == 1:
            out[key] = raw
        else:
            out[key] = raw.decode("latin-1")
        layout[key] = {"maxdims": tuple(int(maxdims[i]) for i in range(rank)) if maxdims else None,
                       "chunk": tuple(int(chunk[i]) for i in range(rank)) if chunk else None}

    cb = H5_DATASET_FN(_cb)
    _ck(_lib().sass_h5_read(str(path).encode(), C.cast(cb, C.c_void_p), None))
    return (out, layout) if with_layout else out


def write_signal_h5(path, qvectors, fqt, fq, fq2, chunksize=10000, resume=False, rawconfig="", config="", database="",
                    datasets=("fqt", "fq0", "fq", "fq2")):
    """Writes (or, with resume=True, appends to) a signal file in the reference's HDF5 layout
    (file_writer_service.cpp:44-171): fqt complex [N][NF], fq / fq2 complex [N]; fq0 = fqt[:, 0]."""
    q = np.ascontiguousarray(qvectors, dtype=np.float64).reshape(-1, 3)
    t = np.asarray(fqt, dtype=np.complex128)
    NF = t.shape[1] if t.ndim == 2 else (t.size // max(len(q), 1))
    t = np.ascontiguousarray(t.reshape(len(q), NF)).view(np.float64).reshape(len(q), 2 * NF)
    a = np.ascontiguousarray(np.asarray(fq, dtype=np.complex128).reshape(len(q))).view(np.float64)
    a2 = np.ascontiguousarray(np.asarray(fq2, dtype=np.complex128).reshape(len(q))).view(np.float64)
    flags = sum(1 << i for i, k in enumerate(("fqt", "fq0", "fq", "fq2")) if k in datasets)
    total = C.c_size_t()
    _ck(_lib().sass_h5_write_signal(str(path).encode(), NF, chunksize, flags, int(resume), rawconfig.encode(),
                                    config.encode(), database.encode(), len(q), _dp(q), _dp(t), _dp(a), _dp(a2),
                                    C.byref(total)))
    return total.value


def load_signal_h5(path):
    """signal.h5 -> the dict load_signal returns (complex fqt / fq0 / fq / fq2)."""
    d = read_h5(path)
    out = {"qvectors": d["qvectors"]}
    for k in ("fqt", "fq0", "fq", "fq2"):
        if k in d:
            out[k] = d[k][..., 0] + 1j * d[k][..., 1]
    return out


def load_signal(signal_dir):
    """Reads a signal directory back: dict of qvectors [N,3], fqt [N,NF] complex, fq0/fq/fq2 [N] complex.
    Multi-rank runs store each writer's rows under rank_<r>/; they are concatenated in rank order (rows pair up
    through qvectors, as in the reference's signal.h5)."""
    import glob
    import os
    dirs = [str(signal_dir)]
    if not os.path.exists(os.path.join(dirs[0], "qvectors.npy")):
        dirs = sorted(glob.glob(os.path.join(dirs[0], "rank_*")), key=lambda d: int(d.rsplit("_", 1)[1]))
        dirs = [d for d in dirs if os.path.exists(os.path.join(d, "qvectors.npy"))]
    out = {}
    for k in ("qvectors", "fqt", "fq0", "fq", "fq2"):
        parts = [np.load(os.path.join(d, k + ".npy")) for d in dirs if os.path.exists(os.path.join(d, k + ".npy"))]
        if parts:
            a = np.concatenate(parts, axis=0)
            out[k] = a if k == "qvectors" else a[..., 0] + 1j * a[..., 1]
    return out
