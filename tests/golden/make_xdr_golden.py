"""Writes the XTC / TRR fixtures of tests/golden/ with the REFERENCE's own trajectory codec and records what that
codec decodes from them.

The codec is the reference's vendored xdrfile 1.1.1 (vendor/xdrfile-1.1.1, plain C), built from where it lies under
/root/reference into oracle/_ref/libxdrfile_ref.so by `make -C oracle ref`.  It only exists in the build container, so the
fixtures (small files) and the decoded coordinates (xdr_golden.npz) are committed; tests/test_xdr_traj.py compares the
product's from-scratch readers with them bit for bit, and, where oracle/_ref is present, with xdrfile live on random input.

Run:  make -C oracle ref && python tests/golden/make_xdr_golden.py
"""
import ctypes as C
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def load_ref():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libxdrfile_ref.so"))
    lib.xdrfile_open.restype = C.c_void_p
    lib.xdrfile_open.argtypes = [C.c_char_p, C.c_char_p]
    lib.xdrfile_close.argtypes = [C.c_void_p]
    fp = C.POINTER(C.c_float)
    ip = C.POINTER(C.c_int)
    lib.write_xtc.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, fp, fp, C.c_float]
    lib.read_xtc.argtypes = [C.c_void_p, C.c_int, ip, fp, fp, fp, fp]
    lib.read_xtc_natoms.argtypes = [C.c_char_p, ip]
    lib.write_trr.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, fp, fp, fp, fp]
    lib.read_trr.argtypes = [C.c_void_p, C.c_int, ip, fp, fp, fp, fp, fp, fp]
    lib.read_trr_natoms.argtypes = [C.c_char_p, ip]
    return lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def write_xtc(lib, path, xyz_nm, box_nm, precision):
    xd = lib.xdrfile_open(path.encode(), b"w")
    assert xd
    for f in range(xyz_nm.shape[0]):
        x = np.ascontiguousarray(xyz_nm[f], dtype=np.float32)
        b = np.ascontiguousarray(box_nm[f], dtype=np.float32)
        assert lib.write_xtc(xd, x.shape[0], f * 10, C.c_float(0.5 * f), _fp(b), _fp(x), C.c_float(precision)) == 0
    lib.xdrfile_close(xd)


def read_xtc(lib, path):
    n = C.c_int()
    assert lib.read_xtc_natoms(path.encode(), C.byref(n)) == 0
    xd = lib.xdrfile_open(path.encode(), b"r")
    frames, boxes = [], []
    while True:
        x = np.zeros((n.value, 3), dtype=np.float32)
        b = np.zeros((3, 3), dtype=np.float32)
        step, t, prec = C.c_int(), C.c_float(), C.c_float(1000.0)
        if lib.read_xtc(xd, n.value, C.byref(step), C.byref(t), _fp(b), _fp(x), C.byref(prec)) != 0:
            break
        frames.append(x)
        boxes.append(b)
    lib.xdrfile_close(xd)
    return np.array(frames), np.array(boxes)


def write_trr(lib, path, xyz_nm, box_nm, with_vf=False):
    xd = lib.xdrfile_open(path.encode(), b"w")
    for f in range(xyz_nm.shape[0]):
        x = np.ascontiguousarray(xyz_nm[f], dtype=np.float32)
        b = np.ascontiguousarray(box_nm[f], dtype=np.float32)
        v = np.ascontiguousarray(x * 0.25) if with_vf else None
        assert lib.write_trr(xd, x.shape[0], f, C.c_float(0.1 * f), C.c_float(0.0), _fp(b), _fp(x),
                             _fp(v) if with_vf else None, _fp(v) if with_vf else None) == 0
    lib.xdrfile_close(xd)


def write_trr_double(path, xyz_nm, box_nm):
    """A double-precision TRR written field by field (xdrfile itself only writes single precision)."""
    with open(path, "wb") as f:
        for i in range(xyz_nm.shape[0]):
            na = xyz_nm.shape[1]
            f.write(struct.pack(">ii", 1993, 13) + struct.pack(">i", 12) + b"GMX_trn_file")
            f.write(struct.pack(">13i", 0, 0, 72, 72, 0, 0, 0, na * 24, 0, 0, na, i, 0))
            f.write(struct.pack(">dd", 0.1 * i, 0.0))
            f.write(np.asarray(box_nm[i], dtype=">f8").tobytes())
            f.write(np.zeros(9, dtype=">f8").tobytes())  # virial
            f.write(np.asarray(xyz_nm[i], dtype=">f8").tobytes())


def read_trr(lib, path):
    n = C.c_int()
    assert lib.read_trr_natoms(path.encode(), C.byref(n)) == 0
    xd = lib.xdrfile_open(path.encode(), b"r")
    frames, boxes = [], []
    while True:
        x = np.zeros((n.value, 3), dtype=np.float32)
        b = np.zeros((3, 3), dtype=np.float32)
        step, t, lam = C.c_int(), C.c_float(), C.c_float()
        if lib.read_trr(xd, n.value, C.byref(step), C.byref(t), C.byref(lam), _fp(b), _fp(x), None, None) != 0:
            break
        frames.append(x)
        boxes.append(b)
    lib.xdrfile_close(xd)
    return np.array(frames), np.array(boxes)


def water_box(rng, nmol, nf, box):
    """nmol three-site molecules random-walking in a cubic box (nm): neighbours in the file are close in space, which
    is what makes the encoder emit runs of small offsets."""
    o = rng.uniform(0, box, (nmol, 3))
    out = np.zeros((nf, nmol * 3, 3), dtype=np.float32)
    for f in range(nf):
        o = o + rng.normal(0, 0.02, o.shape)
        h1 = o + rng.normal(0, 0.06, o.shape)
        h2 = o + rng.normal(0, 0.06, o.shape)
        out[f] = np.stack([o, h1, h2], axis=1).reshape(-1, 3)
    return out


def boxes(nf, box):
    b = np.zeros((nf, 3, 3), dtype=np.float32)
    for f in range(nf):
        b[f] = np.diag([box, box * 1.1, box * 0.9]) + 0.001 * f
    return b


def cases(rng):
    """name -> (kind, xyz_nm, box_nm, precision)"""
    c = {}
    c["water"] = ("xtc", water_box(rng, 60, 4, 3.0), boxes(4, 3.0), 1000.0)
    c["gas"] = ("xtc", rng.uniform(-4, 9, (3, 157, 3)).astype(np.float32), boxes(3, 13.0), 1000.0)
    c["tiny"] = ("xtc", rng.uniform(0, 2, (3, 7, 3)).astype(np.float32), boxes(3, 2.0), 1000.0)  # <= 9 atoms: raw floats
    c["wide"] = ("xtc", rng.uniform(-40, 40, (2, 40, 3)).astype(np.float32), boxes(2, 80.0), 1.0e6)  # ranges > 2^24
    c["coarse"] = ("xtc", water_box(rng, 25, 3, 2.0), boxes(3, 2.0), 100.0)
    c["mixed"] = ("xtc", np.concatenate([water_box(rng, 30, 3, 2.5), rng.uniform(0, 2.5, (3, 31, 3)).astype(np.float32)],
                                        axis=1), boxes(3, 2.5), 1000.0)
    c["trr_single"] = ("trr", rng.uniform(0, 3, (4, 33, 3)).astype(np.float32), boxes(4, 3.0), None)
    c["trr_vf"] = ("trr_vf", rng.uniform(0, 3, (3, 12, 3)).astype(np.float32), boxes(3, 3.0), None)
    c["trr_double"] = ("trr_double", rng.uniform(0, 3, (3, 21, 3)), boxes(3, 3.0), None)
    return c


def main():
    lib = load_ref()
    rng = np.random.default_rng(20261017)
    out = {}
    for name, (kind, xyz, box, prec) in cases(rng).items():
        ext = "xtc" if kind == "xtc" else "trr"
        path = os.path.join(HERE, f"xdr_{name}.{ext}")
        if kind == "xtc":
            write_xtc(lib, path, xyz, box, prec)
            dec, bx = read_xtc(lib, path)
        elif kind == "trr_double":
            write_trr_double(path, xyz, box)
            dec, bx = read_trr(lib, path)
        else:
            write_trr(lib, path, xyz, box, with_vf=(kind == "trr_vf"))
            dec, bx = read_trr(lib, path)
        assert dec.shape == xyz.shape, (name, dec.shape, xyz.shape)
        out[name + "_nm"] = dec   # what xdrfile decodes (nm, float32)
        out[name + "_box_nm"] = bx
        print(name, kind, dec.shape, os.path.getsize(path), "bytes; max |decoded - written| =",
              float(np.max(np.abs(dec - xyz.astype(np.float32)))))
    np.savez_compressed(os.path.join(HERE, "xdr_golden.npz"), **out)


if __name__ == "__main__":
    sys.exit(main())
