"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (/root/reference/multigrid.py,
imported unmodified through oracle/reference_import.py) on synthetic hierarchies from
multigrid_dolfinx_b200.problems.  Only runs in the development container (the reference does not
travel to the GPU box); the .npz files are committed.

    python tests/golden/gen_golden.py

Every array is an output of a reference function:
  vcycle_*       V_cycle_scheme (multigrid.py:231-268), finest level, K consecutive cycles, v0 = 0
  dbg_*          its test=True 4-tuple of the LAST cycle (multigrid.py:262-266)
  jac_*          jacobiRelaxation (multigrid.py:223-228)
  interp_*       Interpolation2D (multigrid.py:59-120)
  inj_*          Restriction2D_direct (multigrid.py:123-132)
  fw_*           Restriction2D (multigrid.py:135-198)
  rj_*           getJacobiMatrices (multigrid.py:48-56) CSR arrays + the inverse diagonal
Inputs that are not reproducible from problems.py alone (random vectors) are stored too.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multigrid_dolfinx_b200 import problems as pr   # noqa: E402
from oracle import reference_import as ri            # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (c, coarsest, finest, perm_seed, mu, ncycles)      -- config 1 of BASELINE.json is 9/17/33
    "cfg1_lex_mu2": (8, 0, 2, None, 2, 6),
    "cfg1_lex_mu50": (8, 0, 2, None, 50, 4),
    "cfg1_perm_mu2": (8, 0, 2, 11, 2, 6),
    "cfg1_perm_mu50": (8, 0, 2, 11, 50, 4),
    "proto_perm_mu50": (8, 1, 3, 5, 50, 3),       # the prototype's own setup 17/33/65 (proto:35-46)
    "l4_perm_mu3": (16, 0, 3, 21, 3, 6),          # 17..129, four levels
}


def main():
    ref = ri.load_reference()
    for name, (c, lc, lf, seed, mu, K) in CASES.items():
        H = pr.build_hierarchy(dim=2, c=c, coarsest_level=lc, finest_level=lf, perm_seed=seed, mu1=mu, mu2=mu,
                               with_dicts=True)
        outs, dbg = ri.run_reference_vcycles(H, K, test_tuple=True)
        A = H.A_sp_dict[lf][0]
        f = H.b_dict[lf]
        d = {"meta": np.array([c, lc, lf, -1 if seed is None else seed, mu, K], dtype=np.int64),
             "omega": np.array(H.omega)}
        big = H.n(lf) > 10000            # keep big fixtures small: last cycle only, no per-op vectors
        d["vcycle_v"] = np.stack([o[:, 0] for o in (outs[-1:] if big else outs)])
        d["vcycle_resnorm"] = np.array([np.linalg.norm(f - A.dot(o)) for o in outs])
        d["dbg_f2h"], d["dbg_v2h"], d["dbg_errh"] = (x[:, 0] for x in dbg)
        if big:
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
            print(name, "n_fine", H.n(lf), "resnorms", d["vcycle_resnorm"][:3])
            continue
        # per-op vectors on random input
        rng = np.random.default_rng(1234)
        x = rng.standard_normal((H.n(lf), 1)); g = rng.standard_normal((H.n(lf), 1))
        e = rng.standard_normal((H.n(lf - 1), 1))
        d["in_x"], d["in_g"], d["in_e"] = x[:, 0], g[:, 0], e[:, 0]
        Aj = H.A_jacobi_sp_dict[lf]
        d["jac_1"] = ref.jacobiRelaxation(Aj, x, g, 1)[:, 0]
        d["jac_5"] = ref.jacobiRelaxation(Aj, x, g, 5)[:, 0]
        md = H.mesh_dof_list_dict
        d["interp"] = ref.Interpolation2D(e, md[lf - 1], md[lf], H.element_size[lf - 1], H.element_size[lf], H.n(lf))[:, 0]
        d["inj"] = ref.Restriction2D_direct(x, md[lf - 1], md[lf], H.n(lf - 1))[:, 0]
        d["fw"] = ref.Restriction2D(x, md[lf - 1], md[lf], H.element_size[lf - 1], H.element_size[lf], H.n(lf - 1))[:, 0]
        RO = Aj[0]
        d["rj_indptr"], d["rj_indices"], d["rj_data"] = RO.indptr, RO.indices, RO.data
        d["rj_dinv"] = Aj[1].diagonal()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "n_fine", H.n(lf), "resnorms", d["vcycle_resnorm"][:3])


if __name__ == "__main__":
    main()
