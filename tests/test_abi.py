"""The C-ABI library loads on a CPU-only box, exports every symbol include/mgb200.h declares, fails
loudly without a device, and its host-side setup routines reproduce the oracle's artefacts bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_SMALL, ROOT, have_gpu, load_golden
from multigrid_dolfinx_b200 import _lib, engine as en, problems as pr
from oracle import c_oracle as co
from oracle import restated as rs


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mgb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    lib = C.CDLL(lib_built)
    names = header_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/mgb200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header drifted apart"
    assert _lib.load().mgb_version() == 100


def test_no_cpu_fallback(lib_built):
    if have_gpu():
        pytest.skip("GPU present")
    with pytest.raises(_lib.MGBError) as ei:
        en.MGEngine(0)
    assert ei.value.code == _lib.ERR_CUDA and "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under multigrid_dolfinx_b200/ may import, include, load or link it."""
    pkg = os.path.join(ROOT, "multigrid_dolfinx_b200")
    bad = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|#\s*include\s*[<\"][^>\"]*oracle|libmgoracle|mg_oracle\.c|c_oracle|oracle/_build", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not bad.search(txt), f"{f} reaches into oracle/"
    import subprocess
    needed = subprocess.run(["ldd", os.path.join(pkg, "libmgb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in needed


@pytest.mark.parametrize("name", GOLDEN_SMALL)
def test_host_rj_equals_reference_getJacobiMatrices(lib_built, name):
    """mgb_host_build_rj (what mgb_finalize runs) vs the arrays getJacobiMatrices produced in the reference run."""
    d, kw, K = load_golden(name)
    H = pr.build_hierarchy(dim=2, with_dicts=False, with_rhs=False, **kw)
    A = H.A_sp_dict[H.finest_level][0]
    rip, rix, rax, dinv = en.host_build_rj(A, reversed_order=True)
    assert np.array_equal(rip, d["rj_indptr"]) and np.array_equal(rix, d["rj_indices"])
    assert np.array_equal(rax, d["rj_data"]) and np.array_equal(dinv, d["rj_dinv"])
    ip, ix, ax, di = rs.rj_pattern_from_values(A)       # as-stored order option
    rip2, rix2, rax2, _ = en.host_build_rj(A, reversed_order=False)
    assert np.array_equal(rip2, ip) and np.array_equal(rix2, ix) and np.array_equal(rax2, ax)


@pytest.mark.parametrize("dim,m,seed", [(2, 24, None), (2, 24, 8), (3, 6, None), (3, 6, 2)])
def test_host_gs_artefacts_equal_oracle(lib_built, dim, m, seed):
    A = pr.stencil_p1(m, dim, pr.make_permutation((m + 1) ** dim, seed))
    for a, b in zip(en.host_level_sets(A), co.level_sets(A)):
        assert np.array_equal(a, b)
    for a, b in zip(en.host_colouring(A), co.greedy_colouring(A)):
        assert np.array_equal(a, b)


def test_host_dense_inverse(lib_built):
    A = pr.stencil_p1(16, 2, pr.make_permutation(17 ** 2, 3))
    inv = en.host_dense_inverse(A)
    assert np.abs(inv @ A.toarray() - np.eye(A.shape[0])).max() < 1e-13
    import scipy.sparse as sp
    with pytest.raises(_lib.MGBError):
        en.host_dense_inverse(sp.csr_matrix(np.array([[1.0, 2.0], [2.0, 4.0]])))


def test_singular_diagonal_is_reported(lib_built):
    import scipy.sparse as sp
    with pytest.raises(_lib.MGBError):
        en.host_build_rj(sp.csr_matrix(np.array([[0.0, 1.0], [1.0, 2.0]])))


def test_every_entry_point_refuses_null_arguments_without_crashing(lib_built):
    """Error behaviour of the boundary (every call returns MGB_OK or a negative code, nothing throws or crashes):
    all 50+ entry points called with a null handle / null pointers, in a child process so that a crash is a test failure."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, sys
sys.path.insert(0, %r)
from multigrid_dolfinx_b200 import _lib as L
lib = L.load()
n = 0
for name, (res, args) in L.SYMBOLS.items():
    if name in ("mgb_version", "mgb_last_error"):
        continue
    vals = [None if (a is C.c_void_p or a is C.c_char_p or hasattr(a, "contents")) else 0.0 if a is C.c_double else 5 for a in args]
    rc = getattr(lib, name)(*vals)
    assert rc < 0 or name == "mgb_destroy", (name, rc)          # destroying nothing is not an error
    n += 1
assert lib.mgb_last_error(None)                                 # the text of the last refusal
print("refused", n)
''' % ROOT
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert int(out.stdout.split()[-1]) >= 50


def test_header_is_plain_c_and_a_c_program_links_against_the_library(lib_built, tmp_path):
    """include/mgb200.h is the boundary a maintainer binds: it must be valid C (not only C++), and a C caller must get the
    documented error code -- not a crash, not a silent CPU path -- when there is no device."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "mgb200.h")
    chk = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert chk.returncode == 0, chk.stderr
    src = tmp_path / "caller.c"
    src.write_text('#include <stdio.h>\n#include "mgb200.h"\n'
                   'int main(void) {\n    mgb_handle* h = NULL;\n    int rc = mgb_create(&h, 0);\n'
                   '    printf("version %d rc %d: %s\\n", mgb_version(), rc, rc == MGB_OK ? "created" : mgb_last_error(NULL));\n'
                   '    if (rc == MGB_OK) rc = mgb_destroy(h);\n    return rc == MGB_OK ? 0 : 2;\n}\n')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(lib_built)
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                         "-L" + libdir, "-l:libmgb200.so", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert "version 100" in run.stdout
    if have_gpu():
        assert run.returncode == 0 and "rc 0: created" in run.stdout, run.stdout
    else:
        assert run.returncode == 2 and f"rc {_lib.ERR_CUDA}:" in run.stdout and "no CPU fallback" in run.stdout, run.stdout
