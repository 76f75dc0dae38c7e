import sys, numpy as np
sys.path.insert(0, "/root/repo")
from multigrid_dolfinx_b200 import problems as pr
from multigrid_dolfinx_b200.engine import MGEngine
H = pr.build_hierarchy(dim=2, c=8, coarsest_level=0, finest_level=2, with_dicts=False)
lf = 2
f = H.b_dict[lf][:, 0]
res = {}
for name, opts in [("a", {"hot_inj": 0, "reuse_g": 0, "compress": 2}), ("b", {}), ("c", {"hot_inj": 0}), ("d", {"reuse_g": 0}), ("e", {"compress": 2}),
                   ("f", {"hot_inj": 0, "reuse_g": 0}), ("g", {"compress": 2, "reuse_g": 0}), ("h", {"compress": 2, "hot_inj": 0})]:
    eng = MGEngine.from_hierarchy(H, options=opts)
    v1 = eng.vcycle(lf, np.zeros_like(f), f)
    v1b = eng.vcycle(lf, v1, f)
    v3, hist = eng.vcycle(lf, np.zeros_like(f), f, ncycles=3, history=True)
    res[name] = (v1, v1b, v3)
    if name == "b": print(eng.describe())
    eng.close()
for k in "bcdefgh":
    print(k, [int(np.count_nonzero(res["a"][j] != res[k][j])) for j in range(3)])
