"""CPU-side checks of the boundary: the library builds, loads and exports every
symbol include/sara_b200.h declares; argument validation that needs no device;
the product never reaches into oracle/."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import sara_b200 as sb
from sara_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "sara_b200.h")).read()
    return sorted(set(re.findall(r"SARA_B200_API[^;(]*?\b(sara_b200_\w+)\s*\(", hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(api.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = sb.load_library()
    for name in _declared_symbols():
        assert hasattr(L, name), name
    out = subprocess.check_output(["nm", "-D", "--defined-only", sb.library_path()], text=True)
    exported = set(re.findall(r" T (sara_b200_\w+)", out))
    assert exported == set(_declared_symbols())
    assert L.sara_b200_version() == 100


def test_header_compiles_as_c_and_cpp(tmp_path):
    for comp, ext in (("gcc", "c"), ("g++", "cpp")):
        src = tmp_path / f"t.{ext}"
        src.write_text('#include "sara_b200.h"\nint main(void){ return sizeof(sara_b200_keypoint) == 52 ? 0 : 1; }\n')
        exe = tmp_path / f"t_{ext}"
        subprocess.check_call([comp, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
        assert subprocess.call([str(exe)]) == 0


def test_cpp_adapter_compiles(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "sara_b200.hpp"\nint main(){ sara_b200::ImagePyramidParams p; return p.scale_count_per_octave() == 6 ? 0 : 1; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), sb.library_path(), f"-Wl,-rpath,{os.path.dirname(sb.library_path())}"])
    assert subprocess.call([str(exe)]) == 0


def test_defaults_match_reference():
    L = sb.load_library()
    a = api._SiftArgs()
    L.sara_b200_default_sift_args(ctypes.byref(a))
    pp = a.pyramid_params
    # ImagePyramid.hpp:36-42, SIFT.hpp:26-32
    assert (pp.first_octave_index, pp.scale_count_per_octave, pp.image_padding_size) == (-1, 6, 1)
    assert pp.scale_geometric_factor == np.float32(2.0) ** np.float32(1.0 / 3.0)
    assert (pp.scale_camera, pp.scale_initial, pp.num_octaves_max) == (0.5, np.float32(1.6), 2**31 - 1)
    assert (a.gauss_truncate, a.extremum_thres, a.edge_ratio_thres, a.extremum_refinement_iter) == (4.0, np.float32(0.01), 10.0, 5)
    d = api._DogArgs()
    L.sara_b200_default_dog_args(ctypes.byref(d))
    assert (d.img_padding_sz, d.extremum_refinement_iter) == (1, 5)  # DoG.hpp:72-78
    assert sb.ImagePyramidParams() == sb.ImagePyramidParams(-1, 6, float(pp.scale_geometric_factor), 1, 0.5, 1.6, 2**31 - 1)


def test_gaussian_kernel_host_side_matches_oracle():
    from oracle import oracle as O

    for sigma in (0.1, 0.5, 1.2262735, 1.5198684, 3.09, 7.7):
        assert api.make_gaussian_kernel(sigma).tobytes() == O.make_gaussian_kernel(sigma).tobytes()
    with pytest.raises(ValueError):
        api.make_gaussian_kernel(100.0)


def test_no_cpu_fallback():
    """Without a device the product refuses to run instead of computing on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(sb.SaraB200Error) as ei:
        sb.SiftContext(64, 64)
    assert ei.value.code == -3
    with pytest.raises(sb.SaraB200Error):
        sb.compute_sift_keypoints(np.zeros((32, 32), np.float32))


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "sara_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|libsara_oracle|sift_oracle", txt, re.M), f
    out = subprocess.check_output(["ldd", sb.library_path()], text=True)
    assert "oracle" not in out


def test_pyramid_kernels_keep_multiply_and_add_apart():
    """The reference's arithmetic is RN(acc + RN(b * k)) per tap.  ptxas contracts
    mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with --fmad=false
    (profiles/microbench), which would change the bits; the kernels issue the add as
    acc * ONE + p instead.  Check in the SASS that every packed multiply kept its own
    packed add, and that the TMA path is really there."""
    import shutil

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.check_output([cuobjdump, "-sass", sb.library_path()], text=True)
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per_kernel.setdefault(m.group(1), {"FMUL2": 0, "FFMA2": 0, "UTMALDG": 0, "FFMA": 0})
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and m.group(1) in cur:
            cur[m.group(1)] += 1
    tma = [k for k in per_kernel if "5stage12stage_kernel" in k or "fused_octave_kernel" in k]
    assert len(tma) >= 7
    for k in tma:
        c = per_kernel[k]
        assert c["UTMALDG"] >= 1, k
        assert c["FMUL2"] > 0 and c["FMUL2"] == c["FFMA2"], (k, c)
        assert c["FFMA"] == 0, (k, c)  # no scalar contraction either
