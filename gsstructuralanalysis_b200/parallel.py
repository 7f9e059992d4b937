"""Multi-GPU host layer: one process per GPU.

Two ways the path shards:
  * replicas — every rank owns a whole assembler and assembles independent states (gsAPALM's workers,
    benchmarks/benchmark_Frustrum_APALM.cpp:391-458).  No communication; see bench.py.
  * strips   — ONE matrix split by element rows of the second parametric direction.  Rank g assembles the elements
    of its strip; the contributions that land in columns owned by the next rank (the control-point rows whose support
    reaches into the next strip) are sent to their owner and added there.  This is the one real exchange step of the
    path; it moves only `double` values of the interface columns (about 6 MB per interface at 1M DOF) with
    point-to-point sends (NCCL on GPUs, gloo in the CPU tests).  DoFs that tie control points of several strips
    together (a collapsed west/east side is ONE DoF for the whole side) are completed by a small all-reduce instead.

  * patches  — a multi-patch (kl_mp_*) split by PATCHES: rank g assembles the patches assigned to it (kl_mp_set_active).  Only the
    columns of interface DoFs (functions glued across patches owned by different ranks) receive contributions from more than one
    rank; every such column is owned by the lowest rank that touches it, the other ranks send their partial column (and residual
    entry) to the owner, which adds them: the same grouped point-to-point exchange as for strips (plan_patches / exchange_patches).

The partition logic below is pure host arithmetic on the knot vector and the DoF map and is shared by the GPU path and
the CPU tests.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np


def _ranges(sorted_idx):
    """merge a sorted int array into [begin, end) runs"""
    if len(sorted_idx) == 0:
        return []
    cuts = np.nonzero(np.diff(sorted_idx) != 1)[0]
    starts = np.concatenate([[0], cuts + 1])
    ends = np.concatenate([cuts + 1, [len(sorted_idx)]])
    return [(int(sorted_idx[a]), int(sorted_idx[b - 1]) + 1) for a, b in zip(starts, ends)]


def function_supports(knots, p):
    """(flo, fhi): first / last non-empty element of every 1-D B-spline function of an open knot vector, and the number
    of non-empty elements (what kl_create builds as flo / fhi; repeated interior knots are fine)."""
    U = np.asarray(knots, dtype=np.float64)
    n = len(U) - p - 1
    spans = [k for k in range(p, n) if U[k + 1] > U[k]]
    flo = np.full(n, 1 << 30, dtype=np.int64)
    fhi = np.full(n, -1, dtype=np.int64)
    for e, k in enumerate(spans):
        flo[k - p:k + 1] = np.minimum(flo[k - p:k + 1], e)
        fhi[k - p:k + 1] = np.maximum(fhi[k - p:k + 1], e)
    return flo, fhi, len(spans)


@dataclass
class StripPlan:
    rank: int
    world: int
    e2_begin: int                 # element rows assembled by this rank
    e2_end: int
    owned_cols: list              # [(c0,c1)] column ranges complete on this rank after the exchange
    send_cols: list               # column ranges whose partial sums go to rank+1
    recv_cols: list               # column ranges received from rank-1 (== that rank's send_cols)
    shared_cols: list             # column ranges touched by several strips' OWN rows (matched DoFs): completed by all-reduce
    tail_rows: int = 0            # the last element rows of the strip that reach control-point rows of the next strip
    compute_begin: int = 0        # halo-compute: assembling the element rows [compute_begin, e2_end) completes every owned column on this
                                  # rank WITHOUT any exchange (p redundant element rows per interface); -1 when shared columns forbid it
    _bufs: dict = field(default_factory=dict)     # receive / pack buffers, allocated once per (device, dtype)


def plan_strips(n1, n2, p, nel2, dof_map, n_free, world, rank, knots2=None):
    """Element rows are split into `world` contiguous strips.  Control-point row i2 is owned by the strip that holds the
    LAST element of its support, so a strip contributes to its own rows and to rows owned by the next strip only (every
    strip holds at least p element rows).  knots2: second-direction knot vector (default: uniform open, one element per
    interior span, i.e. function i2 lives on elements [i2-p, i2])."""
    if knots2 is None:
        flo = np.maximum(np.arange(n2) - p, 0)
        fhi = np.minimum(np.arange(n2), nel2 - 1)
    else:
        flo, fhi, ne = function_supports(knots2, p)
        assert ne == nel2 and len(flo) == n2
    if nel2 < world * p:
        raise ValueError("each strip needs at least p element rows")
    bounds = [(nel2 * g) // world for g in range(world + 1)]
    dm = np.asarray(dof_map).reshape(3, n2, n1)
    strip_of_elem = np.zeros(nel2, dtype=np.int64)
    for g in range(world):
        strip_of_elem[bounds[g]:bounds[g + 1]] = g
    owner_row = strip_of_elem[fhi]                          # owner strip of every control-point row

    def cols_of_rows(mask):
        if not mask.any():
            return np.zeros(0, dtype=np.int64)
        c = np.unique(dm[:, mask, :].reshape(-1))
        return c[c < n_free]

    owned_sets = [cols_of_rows(owner_row == g) for g in range(world)]
    # a DoF that appears in the own rows of more than one strip (matched DoFs along direction 2)
    allc = np.concatenate(owned_sets) if world > 1 else owned_sets[0]
    uniq, cnt = np.unique(allc, return_counts=True)
    shared = uniq[cnt > 1]
    send_sets, tails = [], []
    for g in range(world):
        if g + 1 < world:
            # rows owned by a later strip that this strip's elements reach
            mask = (owner_row > g) & (flo < bounds[g + 1])
            if (owner_row[mask] > g + 1).any():
                raise ValueError("a control-point row of strip %d is supported two strips away: strips are too thin" % g)
            send = np.setdiff1d(cols_of_rows(mask), shared)
            # everything sent must be owned by the neighbour
            if len(np.setdiff1d(send, owned_sets[g + 1])):
                raise ValueError("interface column of strip %d is not owned by its neighbour (matched DoFs across strips)" % g)
            tails.append(int(bounds[g + 1] - flo[mask].min()) if mask.any() else 0)
        else:
            send = np.zeros(0, dtype=np.int64)
            tails.append(0)
        send_sets.append(send)
    owned = np.setdiff1d(owned_sets[rank], shared)
    owned = np.union1d(owned, shared)          # shared columns are complete on every rank after the all-reduce
    recv = send_sets[rank - 1] if rank > 0 else np.zeros(0, dtype=np.int64)
    # halo-compute: the owner also integrates the element rows of the previous strip that its own control-point rows
    # reach into, so nothing has to be sent; elements of the overlap are assembled twice, which would double-count shared columns
    mine = owner_row == rank
    compute_begin = int(flo[mine].min()) if mine.any() else bounds[rank]
    if len(shared):
        compute_begin = -1
    return StripPlan(rank, world, bounds[rank], bounds[rank + 1], _ranges(owned), _ranges(send_sets[rank]), _ranges(recv),
                     _ranges(shared), tails[rank], compute_begin)


def value_ranges(col_ranges, outer):
    """column ranges -> ranges into the compressed value array"""
    return [(int(outer[c0]), int(outer[c1])) for c0, c1 in col_ranges]


def _buffers(plan, outer, values, residual):
    """receive buffers (one flat tensor per neighbour message) are allocated once and reused by every exchange"""
    import torch
    key = (str(values.device), values.dtype, None if residual is None else residual.dtype)
    b = plan._bufs.get(key)
    if b is None:
        vr, sr = value_ranges(plan.recv_cols, outer), value_ranges(plan.shared_cols, outer)
        b = {
            "recv_v": [torch.empty(b_ - a_, dtype=values.dtype, device=values.device) for a_, b_ in vr],
            "recv_r": ([torch.empty(c1 - c0, dtype=residual.dtype, device=residual.device) for c0, c1 in plan.recv_cols]
                       if residual is not None else []),
            "shared": torch.empty(sum(b_ - a_ for a_, b_ in sr) + (sum(c1 - c0 for c0, c1 in plan.shared_cols) if residual is not None else 0),
                                  dtype=values.dtype, device=values.device),
            "vr": vr, "sr": sr, "send_vr": value_ranges(plan.send_cols, outer),
        }
        plan._bufs[key] = b
    return b


def assemble_strip_overlapped(asm, plan: StripPlan, outer, values, residual, x_dev_ptr, dist, stream=0):
    """One Jacobian + internal-force assembly of this rank's strip with the halo exchange overlapped: the last element rows
    (plan.tail_rows: the only ones that touch the next strip's columns) are assembled first, their partial sums start travelling, the rest of
    the strip is assembled meanwhile, and the received ranges are added at the end.  Returns the bytes received."""
    tail = min(plan.tail_rows, plan.e2_end - plan.e2_begin)
    asm.strip_begin_device(x_dev_ptr, residual.data_ptr(), 0.0, 1.0, tail, stream)
    pending = exchange_halo_begin(plan, outer, values, residual, dist)
    if plan.e2_end - tail > plan.e2_begin:
        asm.jacobian_rows_device(plan.e2_begin, plan.e2_end - tail, stream)
    return exchange_halo_end(plan, outer, values, residual, dist, pending)


def exchange_halo_begin(plan: StripPlan, outer, values, residual, dist):
    """post the sends of the interface columns and the receives from the previous strip (asynchronous)"""
    b = _buffers(plan, outer, values, residual)
    ops, dst, src = [], [], []
    if plan.rank + 1 < plan.world:
        for (a, e) in b["send_vr"]:
            ops.append(dist.P2POp(dist.isend, values[a:e], plan.rank + 1))
        if residual is not None:
            for (c0, c1) in plan.send_cols:
                ops.append(dist.P2POp(dist.isend, residual[c0:c1], plan.rank + 1))
    if plan.rank > 0:
        for (a, e), t in zip(b["vr"], b["recv_v"]):
            ops.append(dist.P2POp(dist.irecv, t, plan.rank - 1))
            dst.append(values[a:e]); src.append(t)
        if residual is not None:
            for (c0, c1), t in zip(plan.recv_cols, b["recv_r"]):
                ops.append(dist.P2POp(dist.irecv, t, plan.rank - 1))
                dst.append(residual[c0:c1]); src.append(t)
    reqs = dist.batch_isend_irecv(ops) if ops else []
    return reqs, dst, src


def exchange_halo_end(plan: StripPlan, outer, values, residual, dist, pending):
    import torch
    reqs, dst, src = pending
    for r in reqs:
        r.wait()
    if dst:
        torch._foreach_add_(dst, src)
    nbytes = sum(t.numel() for t in src) * 8
    return nbytes + _allreduce_shared(plan, outer, values, residual, dist)


def _allreduce_shared(plan, outer, values, residual, dist):
    if not (plan.shared_cols and plan.world > 1):
        return 0
    b = _buffers(plan, outer, values, residual)
    flat, o = b["shared"], 0
    parts = [values[a:e] for a, e in b["sr"]] + ([residual[c0:c1] for c0, c1 in plan.shared_cols] if residual is not None else [])
    for t in parts:
        flat[o:o + t.numel()].copy_(t); o += t.numel()
    dist.all_reduce(flat)
    o = 0
    for t in parts:
        t.copy_(flat[o:o + t.numel()]); o += t.numel()
    return flat.numel() * 8


def exchange_halo(plan: StripPlan, outer, values, residual, dist):
    """Send the partial sums of the interface columns to rank+1 and add what rank-1 sent (one batched group of
    point-to-point operations, pre-allocated receive buffers, one fused add); columns shared by several strips are
    summed with one all-reduce.  `values` / `residual` are torch tensors (CPU for gloo, CUDA views for nccl).
    After the call the entries of plan.owned_cols are complete on this rank.  Returns the bytes received."""
    return exchange_halo_end(plan, outer, values, residual, dist, exchange_halo_begin(plan, outer, values, residual, dist))


@dataclass
class PatchPlan:
    rank: int
    world: int
    active: list                  # active[q] = 1 for the patches this rank assembles
    owned_cols: list              # [(c0,c1)] column ranges complete on this rank after the exchange
    send: dict                    # owner rank -> column ranges whose partial sums this rank sends there
    recv: dict                    # sender rank -> column ranges this rank receives and adds
    _bufs: dict = field(default_factory=dict)


def plan_patches(dof_maps, n_free, patch_rank, world, rank):
    """dof_maps: the GLOBAL dof map of every patch; patch_rank[q]: the rank that assembles patch q.  A DoF touched by the patches of
    several ranks is owned by the lowest of them."""
    touch = np.zeros((world, n_free), dtype=bool)
    for q, m in enumerate(dof_maps):
        g = np.unique(np.asarray(m))
        touch[patch_rank[q], g[g < n_free]] = True
    owner = np.argmax(touch, axis=0)                     # first (lowest) rank that touches the DoF
    untouched = ~touch.any(axis=0)
    owner[untouched] = 0
    send, recv = {}, {}
    for r in range(world):
        if r == rank:
            continue
        mine_to_r = np.nonzero(touch[rank] & (owner == r))[0]
        if len(mine_to_r):
            send[r] = _ranges(mine_to_r)
        r_to_me = np.nonzero(touch[r] & (owner == rank))[0]
        if len(r_to_me):
            recv[r] = _ranges(r_to_me)
    owned = np.nonzero(owner == rank)[0]
    return PatchPlan(rank, world, [1 if patch_rank[q] == rank else 0 for q in range(len(dof_maps))], _ranges(owned), send, recv)


def exchange_patches(plan: PatchPlan, outer, values, residual, dist):
    """Complete the interface columns on their owners: one batched group of point-to-point operations (NCCL / gloo), pre-allocated
    receive buffers, one fused add.  Returns the bytes received."""
    import torch
    key = (str(values.device), values.dtype)
    b = plan._bufs.get(key)
    if b is None:
        b = {r: ([torch.empty(e - a, dtype=values.dtype, device=values.device) for a, e in value_ranges(cr, outer)],
                 [torch.empty(c1 - c0, dtype=values.dtype, device=values.device) for c0, c1 in cr] if residual is not None else [])
             for r, cr in plan.recv.items()}
        plan._bufs[key] = b
    ops, adds = [], []
    for r in sorted(set(plan.send) | set(plan.recv)):      # the same peer order on both sides of every pair
        if r in plan.send:
            for a, e in value_ranges(plan.send[r], outer):
                ops.append(dist.P2POp(dist.isend, values[a:e], r))
            if residual is not None:
                for c0, c1 in plan.send[r]:
                    ops.append(dist.P2POp(dist.isend, residual[c0:c1], r))
        if r in plan.recv:
            bv, br = b[r]
            dst, src = [], []
            for (a, e), t in zip(value_ranges(plan.recv[r], outer), bv):
                ops.append(dist.P2POp(dist.irecv, t, r))
                dst.append(values[a:e]); src.append(t)
            if residual is not None:
                for (c0, c1), t in zip(plan.recv[r], br):
                    ops.append(dist.P2POp(dist.irecv, t, r))
                    dst.append(residual[c0:c1]); src.append(t)
            adds.append((dst, src))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    # one fused add PER SENDER: the ranges of one sender are disjoint, but several senders may contribute to the same column (a DoF
    # shared by more than two ranks, e.g. a collapsed side) and a fused multi-tensor add must not see the same destination twice
    for dst, src in adds:
        torch._foreach_add_(dst, src)
    return sum(t.numel() for _, src in adds for t in src) * 8


class DevicePointerView:
    """Zero-copy torch view of a raw device allocation owned by libkl_shell (via __cuda_array_interface__)."""

    def __init__(self, ptr, n, typestr="<f8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 3}

    def tensor(self):
        import torch
        return torch.as_tensor(self, device="cuda")
