"""ctypes binding of oracle/liboracle.so -- the CHECKER, for tests / smoke / cpu_baseline only.

The product package (radiative3d_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import struct

import numpy as np

from radiative3d_b200 import abi

ORACLE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        L = C.CDLL(path)
        pd, pu32, pu64 = C.POINTER(C.c_double), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        pdesc = C.POINTER(abi.ModelDesc)
        L.r3d_oracle_run.argtypes = [pdesc, C.c_uint64, C.c_uint64, C.c_uint64, pd, pu64, pu64, C.c_void_p, C.c_int]
        L.r3d_oracle_run.restype = C.c_int
        L.r3d_oracle_draw.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.r3d_oracle_draw.restype = C.c_uint32
        L.r3d_oracle_cdf_search.argtypes = [pd, C.c_uint32, pu32, C.c_uint32, pu32]
        L.r3d_oracle_path_to_boundary.argtypes = [pdesc, pd, C.c_uint32, pd]
        L.r3d_oracle_advance.argtypes = [pdesc, pd, C.c_uint32, pd]
        L.r3d_oracle_transform.argtypes = [pd, C.c_uint32, pd]
        L.r3d_oracle_rtcoef.argtypes = [pd, C.c_uint32, pd]
        L.r3d_oracle_catch.argtypes = [C.c_double, C.c_uint32, pd, C.c_uint32, pd]
        L.r3d_oracle_build_scatterer_tables.argtypes = [C.POINTER(abi.ScatterParams), pd, pd, C.c_uint32, pd, pd, pd, pd]
        L.r3d_oracle_build_scatterer_tables.restype = None
        for f in ("cdf_search", "path_to_boundary", "advance", "transform", "rtcoef", "catch"):
            getattr(L, "r3d_oracle_" + f).restype = None
        _LIB = L
    return _LIB


def _pd(a):
    return abi.as_ptr(a, C.c_double)


def run(model, first, n, seed, finals=False, nthreads=1):
    """Trace phonons [first, first+n) on the CPU oracle.  Returns (energies, counts, counters, finals|None)."""
    d = model.desc()
    e = np.zeros((model.n_seis, model.n_bins, 5))
    c = np.zeros((model.n_seis, model.n_bins, 2), dtype=np.uint64)
    k = np.zeros(abi.R3D_NCOUNTERS, dtype=np.uint64)
    fin = np.zeros(n, dtype=abi.PHONON_FINAL_DTYPE) if finals else None
    rc = lib().r3d_oracle_run(C.byref(d), first, n, seed, _pd(e), abi.as_ptr(c, C.c_uint64), abi.as_ptr(k, C.c_uint64),
                              fin.ctypes.data if finals else None, nthreads)
    if rc != 0:
        raise MemoryError("r3d_oracle_run failed")
    return e, c, k, fin


def draw(seed, idx, ordinal):
    return lib().r3d_oracle_draw(seed, idx, ordinal)


def cdf_search(cdf, k):
    cdf = np.ascontiguousarray(cdf, dtype=np.float64)
    k = np.ascontiguousarray(k, dtype=np.uint32)
    out = np.zeros(k.size, dtype=np.uint32)
    lib().r3d_oracle_cdf_search(_pd(cdf), cdf.size, abi.as_ptr(k, C.c_uint32), k.size, abi.as_ptr(out, C.c_uint32))
    return out


def _rows(fn, width_in, width_out, x, *pre):
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, width_in)
    out = np.zeros((x.shape[0], width_out))
    fn(*pre, _pd(x), x.shape[0], _pd(out))
    return out


def path_to_boundary(model, x):
    d = model.desc()
    return _rows(lib().r3d_oracle_path_to_boundary, 7, 9, x, C.byref(d))


def advance(model, x):
    d = model.desc()
    return _rows(lib().r3d_oracle_advance, 8, 9, x, C.byref(d))


def transform(x):
    return _rows(lib().r3d_oracle_transform, 6, 3, x)


def rtcoef(x):
    return _rows(lib().r3d_oracle_rtcoef, 15, 13, x)


def catch(bin_dt, n_bins, x):
    return _rows(lib().r3d_oracle_catch, 28, 6, x, C.c_double(bin_dt), C.c_uint32(n_bins))


def build_scatterer_tables(params, toa_theta, toa_phi):
    th = np.ascontiguousarray(toa_theta, dtype=np.float64)
    ph = np.ascontiguousarray(toa_phi, dtype=np.float64)
    n = th.size
    cdf, spol, whole, mfp = np.zeros((4, n)), np.zeros(n), np.zeros((2, 4)), np.zeros(2)
    par = abi.ScatterParams(*[float(x) for x in params])
    lib().r3d_oracle_build_scatterer_tables(C.byref(par), _pd(th), _pd(ph), n, _pd(cdf), _pd(spol), _pd(whole), _pd(mfp))
    return cdf, spol, whole, mfp


def load_bins(path):
    """Read an "R3DBINS1" result file (written by oracle/ref_harness.cpp)."""
    with open(path, "rb") as f:
        if f.read(8) != b"R3DBINS1":
            raise ValueError(f"{path}: not an R3DBINS1 file")
        ns, nb = struct.unpack("<II", f.read(8))
        counters = np.fromfile(f, dtype="<u8", count=abi.R3D_NCOUNTERS)
        energies = np.fromfile(f, dtype="<f8", count=ns * nb * 5).reshape(ns, nb, 5)
        counts = np.fromfile(f, dtype="<u8", count=ns * nb * 2).reshape(ns, nb, 2)
    return energies, counts, counters
