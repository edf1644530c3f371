"""The C-ABI boundary: the library loads, exports every symbol include/r3d_gpu.h declares, the Python
struct mirrors have the C sizes, and -- with no GPU -- the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden
from radiative3d_b200 import abi, engine
from radiative3d_b200.model import FlatModel

HEADER = os.path.join(ROOT, "include", "r3d_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(r3d_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"libr3dgpu.so does not export {n}"
    assert sorted(engine.EXPORTS) == names


def test_struct_sizes_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "r3d_gpu.h"\n#include "r3d_modelfile.h"\n'
                   'int main(){printf("%zu %zu %zu %zu\\n", sizeof(r3d_model_desc), sizeof(r3d_phonon_final),'
                   ' sizeof(r3d_modelfile_scalars), sizeof(r3d_event));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    a, b, c, d = map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split())
    assert a == C.sizeof(abi.ModelDesc)
    assert b == C.sizeof(abi.PhononFinal) == abi.PHONON_FINAL_DTYPE.itemsize
    assert c == 160
    assert d == abi.EVENT_DTYPE.itemsize


def test_abi_version():
    assert engine.load_library().r3d_abi_version() == abi.R3D_ABI_VERSION


def test_no_cpu_fallback(have_gpu):
    """Without a CUDA device r3d_create must fail with R3D_ENODEV, not run anything on the CPU."""
    if have_gpu:
        pytest.skip("a GPU is present")
    m, _ = load_golden("halfspace")
    with pytest.raises(engine.R3DError) as ei:
        engine.Engine(m)
    assert ei.value.code == 3 and "CUDA" in str(ei.value)
    with pytest.raises(engine.R3DError):
        engine.transform(np.zeros((1, 6)))


def test_bad_descriptor_is_rejected():
    L = engine.load_library()
    m, _ = load_golden("halfspace")
    d = m.desc()
    d.cell_nparam = 5
    h = C.c_void_p()
    rc = L.r3d_create(C.byref(d), None, 1, C.byref(h))
    assert rc == 1 and b"cell_nparam" in L.r3d_last_error()
    assert L.r3d_create(None, None, 1, C.byref(h)) == 1
    assert L.r3d_run(None, 0, 1, 1) == 1


def test_missing_library_message(tmp_path):
    with pytest.raises(FileNotFoundError) as ei:
        engine.load_library(str(tmp_path / "nope.so"))
    assert "no CPU fallback" in str(ei.value)


def test_product_does_not_touch_oracle():
    """Nothing under radiative3d_b200/ may import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "radiative3d_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(base, f), errors="replace").read()
                for needle in ("liboracle", "oracle_binding", "r3d_oracle", "oracle/_ref", "ref_harness"):
                    assert needle not in text, f"{f} mentions {needle}"


def test_modelfile_roundtrip(tmp_path):
    m, _ = load_golden("spherical")
    p = tmp_path / "m.r3dmodel"
    m.save(str(p))
    m2 = FlatModel.load(str(p))
    for k in ("freq_hz", "ttl", "bin_dt", "n_bins", "ecs_radial", "earth_center", "src_loc", "src_cell", "cell_kind",
              "loop_concern", "cyl_radius2"):
        assert getattr(m, k) == getattr(m2, k), k
    for k in ("toa_theta", "src_cdf", "scat_cdf", "cell_params", "face_flags", "face_other_cell", "seis"):
        assert np.array_equal(getattr(m, k), getattr(m2, k)), k
    with open(p, "r+b") as f:
        f.write(b"XXXX")
    with pytest.raises(ValueError):
        FlatModel.load(str(p))
