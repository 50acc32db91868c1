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
    assert l.mcs_abi_version() == 1


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
        open(c, "w").write('#include "mcshade.h"\nint main(void){ mcs_tensor t; (void)t; return MCS_ABI_VERSION == 1 ? 0 : 1; }\n')
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
