"""Multi-patch test problems (the builders live with the other workloads)."""
from gsstructuralanalysis_b200.workloads import cut, dof_permutation  # noqa: F401
