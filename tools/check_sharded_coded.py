"""Row-sharded hierarchy with losslessly coded operators vs the single-GPU CSR kernels: bit identity of the gathered
solution and of the residual history.  Run under torchrun with >= 2 ranks; writes gpurun_out/sharded_coded.json.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_coded.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as td
    from multigrid_dolfinx_b200 import dist as ds
    from multigrid_dolfinx_b200 import problems as pr
    from multigrid_dolfinx_b200.engine import MGEngine
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    td.init_process_group("nccl", device_id=torch.device("cuda", rank))
    out = {}
    cases = [("generated_3d_65", 3, 4, 4, True, {}), ("host_2d_129", 2, 8, 4, False, {}), ("generated_3d_65_overlap2", 3, 4, 4, True, {"overlap_halo": 2})]
    for name, dim, c, lf, gen, opts in cases:
        H = pr.build_hierarchy(dim=dim, c=c, coarsest_level=0, finest_level=lf, with_dicts=False) if (rank == 0 or not gen) else None
        src = ds.StructuredSource(dim, c, 0, lf) if gen else ds.HierarchySource(H)
        mg = ds.DistMG(src, device=rank, gather_level=1, device_gen=gen, options=opts)
        mg.load_rhs()
        hist = mg.cycles(4, history=True)
        v = mg.gather_solution()
        desc = mg.eng.describe()
        mg.close()
        if rank == 0:
            f = src.rhs_rows(lf, 0, H.n(lf)) if gen else H.b_dict[lf][:, 0]
            eng = MGEngine.from_hierarchy(H, options={"compress": 0})
            v1, h1 = eng.vcycle(lf, np.zeros_like(f), f, ncycles=4, history=True)
            eng.close()
            out[name] = {"bit_identical": bool(np.array_equal(v, v1)), "max_abs_diff": float(np.abs(v - v1).max()),
                         "hist_rel": float(np.abs(hist - h1).max() / h1.max()), "coded_operators_rank0": desc.count("coded mode="),
                         "mode3_operators_rank0": desc.count("coded mode=3"), "world": td.get_world_size()}
            print(name, out[name], flush=True)
        td.barrier()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open("gpurun_out/sharded_coded.json", "w"), indent=1)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
