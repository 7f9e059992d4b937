"""Multi-GPU host layer: one process per GPU.

Two ways the path shards:
  * replicas — every rank owns a whole assembler and assembles independent states (gsAPALM's workers,
    benchmarks/benchmark_Frustrum_APALM.cpp:391-458).  No communication; see bench.py.
  * strips   — ONE matrix split by element rows of the second parametric direction.  Rank g assembles the elements
    of its strip into its (full-size) value array; the contributions that land in columns owned by the next rank
    (the p rows of control points the strips share) are sent to their owner and added there.  This is the one real
    exchange step of the path; it moves only `double` values of the interface columns (about 6 MB per interface at
    1M DOF) with point-to-point sends (NCCL on GPUs, gloo in the CPU tests).

The partition logic below is pure host arithmetic on the DoF map and is shared by the GPU path and the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np


def _ranges(sorted_idx):
    """merge a sorted int array into [begin, end) runs"""
    if len(sorted_idx) == 0:
        return []
    cuts = np.nonzero(np.diff(sorted_idx) != 1)[0]
    starts = np.concatenate([[0], cuts + 1])
    ends = np.concatenate([cuts + 1, [len(sorted_idx)]])
    return [(int(sorted_idx[a]), int(sorted_idx[b - 1]) + 1) for a, b in zip(starts, ends)]


@dataclass
class StripPlan:
    rank: int
    world: int
    e2_begin: int                 # element rows assembled by this rank
    e2_end: int
    owned_cols: list              # [(c0,c1)] column ranges this rank owns after the exchange
    send_cols: list               # column ranges whose partial sums go to rank+1
    recv_cols: list               # column ranges received from rank-1 (== that rank's send_cols)


def plan_strips(n1, n2, p, nel2, dof_map, n_free, world, rank, fhi2=None):
    """Element rows are split into `world` contiguous strips; control-point row i2 is owned by the strip that holds
    its first element, so a strip only ever contributes to its own rows and to the first p rows of the next strip.
    fhi2/flo2-free version for open knot vectors without interior repetitions: function i2 lives on elements
    [i2-p, i2] clipped to [0, nel2)."""
    assert nel2 >= world * p, "each strip needs at least p element rows"
    bounds = [(nel2 * g) // world for g in range(world + 1)]
    ncp = n1 * n2
    dm = np.asarray(dof_map).reshape(3, n2, n1)

    def first_elem(i2):
        return max(i2 - p, 0)

    def cols_of_rows(r0, r1):
        if r1 <= r0:
            return np.zeros(0, dtype=np.int64)
        c = np.unique(dm[:, r0:r1, :].reshape(-1))
        return c[c < n_free]

    # owner of cp row i2: strip containing element min(i2, nel2-1)  (so rows [E_g, E_{g+1}) belong to g, the last strip
    # also owns the trailing p rows)
    def owner_rows(g):
        r0 = bounds[g]
        r1 = bounds[g + 1] if g + 1 < world else n2
        return r0, r1

    plans = []
    for g in range(world):
        r0, r1 = owner_rows(g)
        owned = cols_of_rows(r0, r1)
        if g + 1 < world:
            s0, s1 = bounds[g + 1], min(bounds[g + 1] + p, n2)
            send = cols_of_rows(s0, s1)
        else:
            send = np.zeros(0, dtype=np.int64)
        plans.append((owned, send))
    # matched DoFs can tie rows of different strips together (collapsed sides): a column is owned by the LOWEST rank
    # that lists it, and every other rank that touches it sends it there.  With the slab ordering this only ever
    # involves neighbours for clamped sides; collapsed sides along direction 1 stay inside one strip.
    owned, send = plans[rank]
    recv = plans[rank - 1][1] if rank > 0 else np.zeros(0, dtype=np.int64)
    return StripPlan(rank, world, bounds[rank], bounds[rank + 1], _ranges(owned), _ranges(send), _ranges(recv))


def value_ranges(col_ranges, outer):
    """column ranges -> ranges into the compressed value array"""
    return [(int(outer[c0]), int(outer[c1])) for c0, c1 in col_ranges]


def exchange_halo(plan: StripPlan, outer, values, residual, dist, device_tensor_fn=None):
    """Send the partial sums of the interface columns to rank+1 and add what rank-1 sent (one batched group of
    point-to-point operations).  `values` / `residual` are torch tensors (CPU for gloo, CUDA views for nccl).
    After the call the entries of plan.owned_cols are complete on this rank.  Returns the bytes received."""
    import torch
    ops, bufs = [], []
    if plan.rank + 1 < plan.world:
        for (a, b) in value_ranges(plan.send_cols, outer):
            ops.append(dist.P2POp(dist.isend, values[a:b], plan.rank + 1))
        if residual is not None:
            for (c0, c1) in plan.send_cols:
                ops.append(dist.P2POp(dist.isend, residual[c0:c1], plan.rank + 1))
    if plan.rank > 0:
        for (a, b) in value_ranges(plan.recv_cols, outer):
            t = torch.empty(b - a, dtype=values.dtype, device=values.device)
            ops.append(dist.P2POp(dist.irecv, t, plan.rank - 1))
            bufs.append((values, a, b, t))
        if residual is not None:
            for (c0, c1) in plan.recv_cols:
                t = torch.empty(c1 - c0, dtype=residual.dtype, device=residual.device)
                ops.append(dist.P2POp(dist.irecv, t, plan.rank - 1))
                bufs.append((residual, c0, c1, t))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    for (dst, a, b, t) in bufs:
        dst[a:b] += t
    return sum(t.numel() for (_, _, _, t) in bufs) * 8


class DevicePointerView:
    """Zero-copy torch view of a raw device allocation owned by libkl_shell (via __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda")
