"""CPU: the C-ABI shared library loads and exports every symbol include/mcshade.h declares (no compute without a GPU)."""
import ctypes
import os
import re

from nvdiffrecmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mcshade.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcs_[a-z0-9_]+)\s*\(", src)))


def test_library_is_built_in_tree_and_loads():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build()"
    assert os.path.commonpath([ROOT, os.path.abspath(_lib.LIB_PATH)]) == ROOT
    l = _lib.lib()
    assert l.mcs_abi_version() == 2


def test_every_declared_symbol_is_exported_and_bound():
    names = _declared()
    assert len(names) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "missing export: " + n
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, "Python binding table and header disagree"


def test_header_is_plain_c():
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write('#include "mcshade.h"\nint main(void){ mcs_tensor t; (void)t; return MCS_ABI_VERSION == 2 ? 0 : 1; }\n')
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", c, "-o", os.path.join(d, "t.o")], check=True)


def test_errors_without_gpu_are_reported_not_swallowed():
    """ctx creation needs a device: on a CPU-only box it must FAIL with a message (the reference drops CUDA errors)."""
    import torch
    if torch.cuda.is_available():
        return
    l = _lib.lib()
    h = ctypes.c_void_p()
    rc = l.mcs_ctx_create(ctypes.byref(h))
    assert rc != 0 and b"cudaGetDevice" in l.mcs_last_error()


def test_argument_validation_returns_status_and_message():
    """Every entry point validates its arguments BEFORE touching the device and reports through the status code + mcs_last_error()
    (the reference's CUDA_CHECK / OPTIX_CHECK format a string and drop it, optixutils/c_src/common.h:37-61).  Null arguments never
    reach a kernel launch, so this runs without a GPU."""
    l = _lib.lib()
    N = None
    T = ctypes.POINTER(_lib.mcs_tensor)()
    calls = [
        ("mcs_bvh_build", lambda: l.mcs_bvh_build(N, N, 0, N, 0, 1, N), b"null context"),
        ("mcs_trace_visibility", lambda: l.mcs_trace_visibility(N, N, N, 4, N, N), b"no acceleration structure"),
        ("mcs_rasterize", lambda: l.mcs_rasterize(N, N, 1, 4, 4, N, N), b"no acceleration structure"),
        ("mcs_interpolate_fwd", lambda: l.mcs_interpolate_fwd(N, 0, 3, 3, N, 1, N, 1, 2, 2, N, N), b"bad arguments"),
        ("mcs_texel_fetch_fwd", lambda: l.mcs_texel_fetch_fwd(N, 4, 3, N, 1, N, N), b"bad arguments"),
        ("mcs_texel_fetch_bwd", lambda: l.mcs_texel_fetch_bwd(4, 3, N, 1, N, N, N), b"bad arguments"),
        ("mcs_update_pdf", lambda: l.mcs_update_pdf(T, N, N, N, N, N), b"null / empty"),
        ("mcs_bilateral_fwd", lambda: l.mcs_bilateral_fwd(T, T, T, ctypes.c_float(1.0), N, N), b"null / empty"),
        ("mcs_shade_combine_fwd", lambda: l.mcs_shade_combine_fwd(T, T, T, T, 1, N, N), b"null / empty"),
        ("mcs_pbr_bsdf_fwd", lambda: l.mcs_pbr_bsdf_fwd(T, T, T, T, T, T, ctypes.c_float(0.08), 0, N, N), b"null / empty"),
        ("mcs_image_loss_fwd", lambda: l.mcs_image_loss_fwd(T, T, 0, 0, N, N), b"null"),
        ("mcs_xfm_fwd", lambda: l.mcs_xfm_fwd(T, T, 1, N, N), b"null / empty"),
    ]
    for name, call, frag in calls:
        rc = call()
        msg = l.mcs_last_error() or b""
        assert rc != 0, name
        assert frag in msg, (name, msg)
