"""The shipped library's SASS, checked on the CPU (cuobjdump reads the cubins; no GPU needed).

Bit-exactness against the reference rests on every row sum being ONE accumulator fed by separately rounded multiplies and adds
(scipy's csr_matvec, multigrid.py:226 / :244 / :260): a fused multiply-add anywhere in a row-sum kernel would change bits.  The
kernels say so in source (`__dmul_rn` / `__dadd_rn`); this test checks what the compiler actually emitted."""
import os
import shutil
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None, reason="needs cuobjdump and c++filt")


@pytest.fixture(scope="module")
def sass():
    import sass_summary as ss
    if not os.path.exists(ss.LIB):
        pytest.skip("libmgb200.so not built")
    return ss, ss.scan()


def test_no_fused_multiply_add_in_any_row_sum_kernel(sass):
    ss, stats = sass
    rowsum = [k for k in stats if ss.ROWSUM.search(k)]
    assert len(rowsum) > 100                                            # the families really were found by name
    for fam in ("k_hotrow<", "k_hotinj<", "k_anchrow<", "k_anchloop<", "k_rowstream<", "k_stream<", "k_tile<"):
        assert any(fam in k for k in rowsum), fam
    assert ss.fused_multiply_add_offenders(stats) == []
    # and they do compute in fp64 with separate multiplies and adds
    k = next(k for k in rowsum if "k_hotrow<6, 128, 2, 8, false, mgb::EpiJacobiRJ>" in k)
    assert stats[k]["DMUL"] > 0 and stats[k]["DADD"] > 0 and stats[k].get("DFMA", 0) == 0


def test_library_is_sm100a_only_with_tma_mbarrier_and_dependent_launch(sass):
    import subprocess
    ss, stats = sass
    elfs = [l for l in subprocess.run(["cuobjdump", "-lelf", ss.LIB], capture_output=True, text=True).stdout.splitlines() if l.strip()]
    assert elfs and all("sm_100a" in l for l in elfs), elfs
    stream = [v for k, v in stats.items() if "k_stream<" in k]
    assert stream and all(v["UBLKCP"] > 0 and v["SYNCS"] > 0 for v in stream)          # 1-D TMA bulk copies behind mbarriers
    hot = [v for k, v in stats.items() if "k_hotrow<" in k]
    assert hot and all(v["ACQBULK"] > 0 for v in hot)                                    # griddepcontrol.wait: programmatic dependent launch
    assert not any(op.startswith(("UTCMMA", "UTCHMMA", "HMMA", "DMMA")) for v in stats.values() for op in v)   # no tensor-core path
