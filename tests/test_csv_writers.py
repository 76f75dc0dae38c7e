"""The three CSV files of the reference's driver (multigrid.py:298-301, :345-356): same file names, same rows, byte for byte.
No GPU needed: the writers only read the module globals initialize_problem sets (multigrid.py:10-45)."""
import csv
import os

import numpy as np
import pytest

from multigrid_dolfinx_b200 import multigrid as mg
from oracle import reference_import as ri


def _set_globals(mod, c, lc, lf):
    mod.coarsest_level_elements_per_dim, mod.coarsest_level, mod.finest_level = c, lc, lf


@pytest.mark.parametrize("c,lc,lf", [(8, 1, 3), (32, 0, 6)])
def test_residual_and_error_csv_round_trip(tmp_path, monkeypatch, c, lc, lf):
    monkeypatch.chdir(tmp_path)
    res = [float(x) for x in np.random.default_rng(0).standard_normal(7) * 1e-3] + [1e-12, 3.0]
    err = [float(x) for x in np.random.default_rng(1).standard_normal(5)]
    saved = (mg.coarsest_level_elements_per_dim, mg.coarsest_level, mg.finest_level)
    try:
        _set_globals(mg, c, lc, lf)
        mg.writing_residual_for_mesh_to_csv(res)
        mg.writing_error_for_mesh_to_csv(err)
    finally:
        mg.coarsest_level_elements_per_dim, mg.coarsest_level, mg.finest_level = saved
    nlev, nel = lf - lc + 1, c * 2 ** lf
    fr, fe = f"residual_for_{nel}_{nlev}_levels.csv", f"error_for_{nel}_{nlev}_levels.csv"      # multigrid.py:346, :353
    assert sorted(os.listdir(tmp_path)) == sorted([fr, fe])
    for name, vals in ((fr, res), (fe, err)):
        rows = list(csv.reader(open(name)))
        assert [int(r[0]) for r in rows] == list(range(len(vals)))
        assert [float(r[1]) for r in rows] == vals               # repr round trip: the values come back exactly
    mine = {n: open(n, "rb").read() for n in (fr, fe)}
    try:
        ref = ri.load_reference()
    except ri.ReferenceUnavailable:
        return
    for n in (fr, fe):
        os.remove(n)
    _set_globals(ref, c, lc, lf)
    ref.writing_residual_for_mesh_to_csv(res)
    ref.writing_error_for_mesh_to_csv(err)
    assert {n: open(n, "rb").read() for n in (fr, fe)} == mine   # byte for byte what the unmodified reference writes


def test_iteration_count_csv_row_format(tmp_path, monkeypatch):
    """multigrid.py:298-301: one appended row [finest cells per dimension, finest-level cycle count]; FullMultiGrid writes it
    through the same csv.writer call (multigrid_dolfinx_b200/multigrid.py, FullMultiGrid) -- checked here on the file format."""
    monkeypatch.chdir(tmp_path)
    import inspect
    src = inspect.getsource(mg.FullMultiGrid)
    assert "iter_count_for_diff_num_elems_{finest_level - coarsest_level + 1}_levels.csv" in src and "mode='a'" in src
    assert "coarsest_level_elements_per_dim * 2 ** finest_level, len(hist)" in src
    with open("iter_count_for_diff_num_elems_3_levels.csv", mode="a") as f1:
        csv.writer(f1, delimiter=",").writerow([64, 17])
    with open("iter_count_for_diff_num_elems_3_levels.csv", mode="a") as f1:
        csv.writer(f1, delimiter=",").writerow([128, 19])
    assert list(csv.reader(open("iter_count_for_diff_num_elems_3_levels.csv"))) == [["64", "17"], ["128", "19"]]
