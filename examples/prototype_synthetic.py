#!/usr/bin/env python
"""The reference's driver script (Multigrid_prototype.py) with its dolfinx assembly loop (lines 62-133) replaced by
the synthetic assembler and its `from multigrid import ...` (line 8) replaced by the drop-in module.  Everything from
line 135 on is the reference's own sequence: getJacobiMatrices per level, initialize_problem, FullMultiGrid_test.

    python examples/prototype_synthetic.py            (needs a CUDA device)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_dolfinx_b200 import problems
from multigrid_dolfinx_b200.multigrid import (getJacobiMatrices, FullMultiGrid, initialize_problem, FullMultiGrid_test,
                                              writing_residual_for_mesh_to_csv)

# Multigrid_prototype.py:35-46
finest_level = 3
coarsest_level = finest_level - 2
coarsest_level_elements_per_dim = 8
mu0, mu1, mu2, omega = 2, 50, 50, 2 / 3

H = problems.build_hierarchy(dim=2, c=coarsest_level_elements_per_dim, coarsest_level=coarsest_level, finest_level=finest_level,
                             perm_seed=7, mu0=mu0, mu1=mu1, mu2=mu2, omega=omega, with_dicts=True)     # stands in for :62-133


class Var_initializer:            # Multigrid_prototype.py:15-32 (any object with these 16 attributes works)
    pass


parameter = Var_initializer()
A_jacobi_sp_dict = {key: getJacobiMatrices(value) for key, value in H.A_sp_dict.items()}               # :135-136
for name in ("mesh_dof_list_dict", "element_size", "coarsest_level_elements_per_dim", "coarsest_level", "finest_level",
             "A_sp_dict", "b_dict", "mu0", "mu1", "mu2", "omega", "residual_per_V_cycle_finest", "error_per_V_cycle_finest",
             "u_exact_fine", "V_fine_dolfx"):
    setattr(parameter, name, getattr(H, name))
parameter.A_jacobi_sp_dict = A_jacobi_sp_dict
initialize_problem(parameter)                                                                            # :138-140
u_FMG_test, residual_fine_restricted, error_coarse, error_coarse_to_fine_interp = FullMultiGrid_test(
    A_jacobi_sp_dict[finest_level], H.b_dict[finest_level], True)                                       # :141-143
print(u_FMG_test.shape)                                                                                  # :144-147
print(residual_fine_restricted.shape)
print(error_coarse.shape)
print(error_coarse_to_fine_interp.shape)

# the production path the prototype has commented out (:148-150)
import multigrid_dolfinx_b200.multigrid as mg
mg.max_fmg_cycles = 400
u_FMG = FullMultiGrid(A_jacobi_sp_dict[finest_level], H.b_dict[finest_level])
A = H.A_sp_dict[finest_level][0]
print("FMG cycles on the finest level:", len(H.residual_per_V_cycle_finest), " final ||f - A u||_2 =",
      float(np.linalg.norm(H.b_dict[finest_level] - A.dot(u_FMG))))
