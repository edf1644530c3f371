"""Flattened Radiative3D model: the host-side container behind r3d_model_desc.

A `FlatModel` is what the reference's Model constructor produces (model.cpp:220-501)
after it has been walked into plain arrays: take-off-angle set, source
and scatterer CDF tables, cell / face records and seismometer descriptors.  The
on-disk form is the "R3DMODL1" container of include/r3d_modelfile.h.
"""
import ctypes as C
import struct

import numpy as np

from . import abi

_SCALARS = struct.Struct("<13dQ12I")          # struct r3d_modelfile_scalars
_ARRAYS = [                                    # order of include/r3d_modelfile.h
    ("toa_theta", "<f8"), ("toa_phi", "<f8"), ("src_whole_cdf", "<f8"), ("src_cdf", "<f8"),
    ("scat_mfp", "<f8"), ("scat_whole_cdf", "<f8"), ("scat_cdf", "<f8"), ("scat_spol", "<f8"),
    ("cell_params", "<f8"), ("cell_scat", "<u4"), ("face_flags", "u1"), ("face_other_cell", "<u4"),
    ("seis", "<f8"),
]
_SCALAR_NAMES = ["freq_hz", "ttl", "bin_dt", "ec0", "ec1", "ec2", "min_theta", "max_theta", "slow_concern",
                 "sl0", "sl1", "sl2", "cyl_radius2", "loop_concern", "n_bins", "ecs_radial", "no_deflect",
                 "n_toa", "src_cell", "n_scat", "n_cells", "cell_kind", "cell_nparam", "faces_per_cell",
                 "n_seis", "pad"]


class FlatModel:
    def __init__(self, **kw):
        self.freq_hz = 1.0
        self.ttl = 3600.0
        self.bin_dt = 1.0
        self.n_bins = 100
        self.ecs_radial = 0
        self.earth_center = (0.0, 0.0, 0.0)
        self.min_theta = 1e-7
        self.max_theta = np.pi - 1e-7
        self.slow_concern = 0.001
        self.loop_concern = 1 << 20
        self.no_deflect = 0
        self.src_loc = (0.0, 0.0, 0.0)
        self.src_cell = 0
        self.cell_kind = abi.R3D_CELL_CYLINDER
        self.cyl_radius2 = 0.0
        for name, dt in _ARRAYS:
            setattr(self, name, np.zeros(0, dtype=dt))
        for k, v in kw.items():
            setattr(self, k, v)

    # ---- derived sizes -------------------------------------------------
    @property
    def n_toa(self):
        return int(self.toa_theta.size)

    @property
    def n_scat(self):
        return int(self.scat_mfp.size // 2)

    @property
    def cell_nparam(self):
        return {abi.R3D_CELL_CYLINDER: abi.R3D_CYL_NPARAM, abi.R3D_CELL_SHELL: abi.R3D_SHELL_NPARAM,
                abi.R3D_CELL_TETRA: abi.R3D_TETRA_NPARAM}[self.cell_kind]

    @property
    def faces_per_cell(self):
        return {abi.R3D_CELL_CYLINDER: abi.R3D_CYL_NFACES, abi.R3D_CELL_SHELL: abi.R3D_SHELL_NFACES,
                abi.R3D_CELL_TETRA: abi.R3D_TETRA_NFACES}[self.cell_kind]

    @property
    def n_cells(self):
        return int(self.cell_scat.size)

    @property
    def n_seis(self):
        return int(self.seis.size // abi.R3D_SEIS_NPARAM)

    def table_bytes(self):
        return sum(getattr(self, n).nbytes for n, _ in _ARRAYS)

    # ---- validation (mirrors the checks r3d_create makes) ---------------
    def validate(self):
        nt, ns, nc, nf = self.n_toa, self.n_scat, self.n_cells, self.faces_per_cell
        want = {"toa_theta": nt, "toa_phi": nt, "src_whole_cdf": 3, "src_cdf": 3 * nt, "scat_mfp": 2 * ns,
                "scat_whole_cdf": 8 * ns, "scat_cdf": 4 * ns * nt, "scat_spol": ns * nt,
                "cell_params": nc * self.cell_nparam, "cell_scat": nc, "face_flags": nc * nf,
                "face_other_cell": nc * nf, "seis": self.n_seis * abi.R3D_SEIS_NPARAM}
        for (name, dt) in _ARRAYS:
            a = getattr(self, name)
            if a.dtype != np.dtype(dt) or not a.flags.c_contiguous:
                raise ValueError(f"{name}: need contiguous {dt}")
            if a.size != want[name]:
                raise ValueError(f"{name}: {a.size} elements, expected {want[name]}")
        if nt == 0 or nc == 0 or ns == 0:
            raise ValueError("model needs at least one TOA, one cell and one scatterer")
        if self.src_cell >= nc:
            raise ValueError("src_cell out of range")
        if nc and int(self.cell_scat.max()) >= ns:
            raise ValueError("cell_scat out of range")
        return self

    # ---- C view ---------------------------------------------------------
    def desc(self):
        """r3d_model_desc borrowing this object's arrays (keep `self` alive while it is used)."""
        self.validate()
        d = abi.ModelDesc()
        d.freq_hz, d.ttl, d.bin_dt, d.n_bins = self.freq_hz, self.ttl, self.bin_dt, self.n_bins
        d.ecs_radial = self.ecs_radial
        d.earth_center[:] = self.earth_center
        d.min_theta, d.max_theta, d.slow_concern = self.min_theta, self.max_theta, self.slow_concern
        d.loop_concern, d.no_deflect = self.loop_concern, self.no_deflect
        d.n_toa = self.n_toa
        d.src_loc[:] = self.src_loc
        d.src_cell = self.src_cell
        d.n_scat, d.n_cells, d.cell_kind = self.n_scat, self.n_cells, self.cell_kind
        d.cell_nparam, d.faces_per_cell = self.cell_nparam, self.faces_per_cell
        d.cyl_radius2 = self.cyl_radius2
        d.n_seis = self.n_seis
        for name, dt in _ARRAYS:
            ct = {"<f8": C.c_double, "<u4": C.c_uint32, "u1": C.c_uint8}[dt]
            setattr(d, name, abi.as_ptr(getattr(self, name), ct))
        d._keepalive = self
        return d

    # ---- file I/O ---------------------------------------------------------
    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            if f.read(8) != b"R3DMODL1":
                raise ValueError(f"{path}: not an R3DMODL1 file")
            s = dict(zip(_SCALAR_NAMES, _SCALARS.unpack(f.read(_SCALARS.size))))
            m = cls()
            for k in ("freq_hz", "ttl", "bin_dt", "min_theta", "max_theta", "slow_concern", "cyl_radius2",
                      "loop_concern", "n_bins", "ecs_radial", "no_deflect", "src_cell", "cell_kind"):
                setattr(m, k, s[k])
            m.earth_center = (s["ec0"], s["ec1"], s["ec2"])
            m.src_loc = (s["sl0"], s["sl1"], s["sl2"])
            for name, dt in _ARRAYS:
                (nbytes,) = struct.unpack("<Q", f.read(8))
                a = np.fromfile(f, dtype=np.uint8, count=nbytes)
                if a.size != nbytes:
                    raise ValueError(f"{path}: truncated in {name}")
                f.seek((8 - nbytes % 8) % 8, 1)
                setattr(m, name, a.view(dt).copy())
        for k in ("n_toa", "n_scat", "n_cells", "cell_nparam", "faces_per_cell", "n_seis"):
            if getattr(m, k) != s[k]:
                raise ValueError(f"{path}: header {k}={s[k]} disagrees with array sizes ({getattr(m, k)})")
        return m.validate()

    def save(self, path):
        self.validate()
        with open(path, "wb") as f:
            f.write(b"R3DMODL1")
            f.write(_SCALARS.pack(self.freq_hz, self.ttl, self.bin_dt, *self.earth_center, self.min_theta,
                                  self.max_theta, self.slow_concern, *self.src_loc, self.cyl_radius2,
                                  self.loop_concern, self.n_bins, self.ecs_radial, self.no_deflect, self.n_toa,
                                  self.src_cell, self.n_scat, self.n_cells, self.cell_kind, self.cell_nparam,
                                  self.faces_per_cell, self.n_seis, 0))
            for name, _ in _ARRAYS:
                a = getattr(self, name)
                f.write(struct.pack("<Q", a.nbytes))
                f.write(a.tobytes())
                f.write(b"\0" * ((8 - a.nbytes % 8) % 8))
