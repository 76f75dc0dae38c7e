"""Generates tests/golden/oracle_*.npz: outputs of the CPU ORACLE (oracle/mg_oracle.c) for the parts of the path that
have no reference text -- 3-D hierarchies, P2, Gauss-Seidel in natural / colour order.  They freeze
the oracle's own definitions so that a later change of the oracle cannot silently move the goalposts; the files produced
by the reference itself are in gen_golden.py.

    python tests/golden/gen_oracle_fixtures.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from multigrid_dolfinx_b200 import problems as pr   # noqa: E402
from oracle import c_oracle as co                    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (builder kwargs, r_mode, smoother, cycles)
    "oracle_3d_p1_inj": (dict(dim=3, c=2, coarsest_level=0, finest_level=3, perm_seed=None), "injection", "jacobi", 4),
    "oracle_3d_p1_perm_transpose": (dict(dim=3, c=2, coarsest_level=0, finest_level=2, perm_seed=3), "transpose", "jacobi", 4),
    "oracle_2d_gs": (dict(dim=2, c=8, coarsest_level=0, finest_level=2, perm_seed=None), "injection", "gs", 3),
    "oracle_2d_gs_color_perm": (dict(dim=2, c=8, coarsest_level=0, finest_level=2, perm_seed=6), "injection", "gs_color", 3),
    "oracle_3d_p2_transpose": ("p2", "transpose", "jacobi", 4),
}


def build(kw):
    if kw == "p2":
        return pr.build_hierarchy_p2(c=1, coarsest_level=0, finest_level=2)
    return pr.build_hierarchy(with_dicts=False, **kw)


def main():
    for name, (kw, r_mode, smoother, K) in CASES.items():
        H = build(kw)
        lf = H.finest_level
        cm = co.from_hierarchy(H, r_mode=r_mode, smoother=smoother)
        f = H.b_dict[lf][:, 0]
        v, hist = cm.vcycle(np.zeros_like(f), f, ncycles=K, history=True)
        d = {"v": v, "resnorm": hist}
        A = H.A_sp_dict[lf][0]
        if smoother.startswith("gs"):
            lev, order, off = co.level_sets(A)
            col, corder, coff = co.greedy_colouring(A)
            d.update(level_of_row=lev, level_offsets=off, colour_of_row=col, colour_offsets=coff)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, "n", H.n(lf), "resnorm", hist)


if __name__ == "__main__":
    main()
