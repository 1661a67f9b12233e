"""CPU-only: the C-ABI shared library loads and exports every symbol include/parm_b200.h declares
(no compute calls without a GPU), the ctypes table matches the header, and the product path fails
loudly -- never falls back -- when no CUDA device is usable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "parm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(parm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from parm_b200 import capi
    return capi


def test_header_declares_the_expected_surface():
    names = header_functions()
    for must in ("parm_ctx_create", "parm_set_box", "parm_upload_atoms", "parm_download_atoms", "parm_nlist_update",
                 "parm_nlist_download_pairs", "parm_inter_set_forces", "parm_inter_energy", "parm_inter_pressure",
                 "parm_verlet_create", "parm_sol_create", "parm_integ_timestep", "parm_reduce"):
        assert must in names


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built.LIB_PATH)
    missing = [n for n in header_functions() if not hasattr(lib, n)]
    assert not missing, "declared in include/parm_b200.h but not exported: %s" % missing


def test_ctypes_table_matches_header(built):
    assert sorted(built.SIGNATURES) == header_functions()


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only machines")
    h = ctypes.c_void_p()
    with pytest.raises(built.ParmError) as e:
        built.call("parm_ctx_create", 3, 16, 0, ctypes.byref(h))
    assert "no CPU fallback" in str(e.value)
    from parm_b200 import sim
    with pytest.raises(built.ParmError):
        sim.AtomVec(8, 1.0)


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under parm_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "parm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.replace("oracle/", "ORACLEDIR_DOC") or "import oracle" not in txt
                assert "from oracle" not in txt and "import oracle" not in txt and "libparm_oracle" not in txt
