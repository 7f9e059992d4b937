"""Multi-patch oracle (TEST INFRASTRUCTURE ONLY — never imported by the product package).

CPU restatement of what the reference does with a gsMultiPatch whose conforming interfaces were found by computeTopology() /
addInterface() (benchmarks/benchmark_Wrinkling.cpp:446-522, benchmarks/benchmark_cylinder_DC.cpp:146):

  * gsFeSpace::setupMapper / gsDofMapper (upstream gismo, not in /root/reference; SURVEY Appendix A.6 [UPSTREAM-RECALLED]):
    one mapper over all patches, per component the patches are concatenated (patch offsets), interface functions are glued
    with matchDofs (k-th function of one side to the k-th / (len-1-k)-th of the other), free DoFs are numbered plain first,
    then the coupled groups, eliminated DoFs after all free ones.  `build_dofmap_mp` restates that with dict-based groups,
    independently of the product's union-find in kl_capi.cu.
  * the element loop of gsExprAssembler runs over every patch and pushes into one matrix: K = sum_q K_q, rhs = sum_q rhs_q.
    `MultiPatchOracle` therefore instantiates the single-patch oracle (oracle/kl_oracle.c) once per patch with the patch's
    GLOBAL dof map and adds the results with scipy; the pattern is the union of the patch patterns.

PARITY UNPINNED against upstream like the rest of the oracle; pinned here by an identity that needs no reference: a patch cut
in two along a parameter line and glued C0 spans exactly the space of the uncut patch with a knot of multiplicity p on that
line, so both must give the same matrix and residual up to the DoF permutation (tests/test_oracle_multipatch.py)."""
from __future__ import annotations

import ctypes as C
import numpy as np
import scipy.sparse as sp

from gsstructuralanalysis_b200.problem import MultiPatchProblem, KL_BC_DIRICHLET, KL_BC_CLAMPED, KL_BC_COLLAPSED
from .binding import Oracle


def _side(n1, n2, s, k, layer=0):
    if s == 0:
        return layer + n1 * k
    if s == 1:
        return (n1 - 1 - layer) + n1 * k
    if s == 2:
        return k + n1 * layer
    return k + n1 * (n2 - 1 - layer)


def build_dofmap_mp(npatch, n1p, n2p, bcs, nif, ifs, map_p, nfree_p, nfixed_p):
    """Same signature as the C symbol kl_mp_build_dofmap (called through MultiPatchProblem.number_dofs)."""
    n1 = [int(n1p[q]) for q in range(npatch)]
    n2 = [int(n2p[q]) for q in range(npatch)]
    off = np.concatenate([[0], np.cumsum([a * b for a, b in zip(n1, n2)])]).astype(int)
    N = int(off[-1])
    out = np.zeros(3 * N, dtype=np.int64)
    free_total, elim_total = 0, 0
    elim_ids = []
    for c in range(3):
        group = {}                      # dof -> set object shared by the members of its matched group
        elim = np.zeros(N, dtype=bool)

        def match(a, b):
            ga, gb = group.get(a), group.get(b)
            if ga is None and gb is None:
                g = {a, b}
            elif ga is None:
                g = gb; g.add(a)
            elif gb is None:
                g = ga; g.add(b)
            elif ga is gb:
                g = ga
            else:
                g = ga | gb
            for m in g:
                group[m] = g

        for q in range(npatch):
            for s in range(4):
                kind = int(bcs[q].side[s][c])
                length = n2[q] if s < 2 else n1[q]
                for k in range(length):
                    b0 = off[q] + _side(n1[q], n2[q], s, k)
                    if kind == KL_BC_DIRICHLET:
                        elim[b0] = True
                    elif kind == KL_BC_CLAMPED:
                        match(b0, off[q] + _side(n1[q], n2[q], s, k, 1))
                    elif kind == KL_BC_COLLAPSED and k > 0:
                        match(off[q] + _side(n1[q], n2[q], s, 0), b0)
            corners = [0, n1[q] - 1, n1[q] * (n2[q] - 1), n1[q] * n2[q] - 1]
            for k in range(4):
                if int(bcs[q].corner[k][c]):
                    elim[off[q] + corners[k]] = True
        for k in range(nif):
            qa, qb = int(ifs[k].patch[0]), int(ifs[k].patch[1])
            sa, sb = int(ifs[k].side[0]), int(ifs[k].side[1])
            length = n2[qa] if sa < 2 else n1[qa]
            assert length == (n2[qb] if sb < 2 else n1[qb]), "non-conforming interface"
            for i in range(length):
                j = length - 1 - i if int(ifs[k].reversed) else i
                match(off[qa] + _side(n1[qa], n2[qa], sa, i), off[qb] + _side(n1[qb], n2[qb], sb, j))
        for g in {id(g): g for g in group.values()}.values():
            if any(elim[m] for m in g):
                for m in g:
                    elim[m] = True
        val = np.full(N, -1, dtype=np.int64)
        cnt = 0
        for i in range(N):
            if not elim[i] and i not in group:
                val[i] = free_total + cnt
                cnt += 1
        seen = {}
        for i in range(N):
            if not elim[i] and i in group:
                key = id(group[i])
                if key not in seen:
                    seen[key] = free_total + cnt
                    cnt += 1
                val[i] = seen[key]
        free_total += cnt
        seen = {}
        eid = np.full(N, -1, dtype=np.int64)
        for i in range(N):
            if elim[i]:
                key = id(group[i]) if i in group else ("s", i)
                if key not in seen:
                    seen[key] = elim_total
                    elim_total += 1
                eid[i] = seen[key]
        elim_ids.append(eid)
        for q in range(npatch):
            ncp = n1[q] * n2[q]
            out[3 * off[q] + c * ncp:3 * off[q] + (c + 1) * ncp] = val[off[q]:off[q] + ncp]
    for c in range(3):
        for q in range(npatch):
            ncp = n1[q] * n2[q]
            seg = elim_ids[c][off[q]:off[q] + ncp]
            dst = out[3 * off[q] + c * ncp:3 * off[q] + (c + 1) * ncp]
            dst[seg >= 0] = free_total + seg[seg >= 0]
    for i in range(3 * N):
        map_p[i] = int(out[i])
    nfree_p._obj.value = free_total          # the arguments are ctypes.byref(c_int32) objects
    nfixed_p._obj.value = elim_total
    return 0


class MultiPatchOracle:
    def __init__(self, mprob: MultiPatchProblem, threads=None):
        if any(p.dof_map is None for p in mprob.patches):
            mprob.number_dofs(build_dofmap_mp)
        self.mprob = mprob
        self.parts = [Oracle(p, threads) for p in mprob.patches]
        self.n_dofs = mprob.n_free
        n = self.n_dofs
        pat = None
        for o in self.parts:
            m = sp.csc_matrix((np.ones(o.nnz), o.inner, o.outer), shape=(n, n))
            pat = m if pat is None else pat + m
        pat.sort_indices()
        self.outer, self.inner = pat.indptr.astype(np.int32), pat.indices.astype(np.int32)
        self.nnz = int(self.outer[-1])
        self._pos = [self._positions(o) for o in self.parts]
        self.n_elements = sum(o.n_elements for o in self.parts)
        self.n_qp = sum(o.n_qp for o in self.parts)

    def _positions(self, o):
        """index into the union value array of every stored entry of the patch oracle o (columns are sorted)"""
        key_u = np.repeat(np.arange(self.n_dofs, dtype=np.int64), np.diff(self.outer)) * self.n_dofs + self.inner
        key_q = np.repeat(np.arange(self.n_dofs, dtype=np.int64), np.diff(o.outer)) * self.n_dofs + o.inner
        pos = np.searchsorted(key_u, key_q)
        assert np.array_equal(key_u[pos], key_q)
        return pos

    def _sum_on_pattern(self, per_patch_values, parts=None):
        v = np.zeros(self.nnz)
        for q, o in enumerate(self.parts):
            if parts is not None and q not in parts:
                continue
            v[self._pos[q]] += per_patch_values(o)       # positions are unique within a patch
        return v

    def jacobian(self, x, parts=None):
        return sp.csc_matrix((self.jacobian_values(x, parts), self.inner, self.outer), shape=(self.n_dofs, self.n_dofs))

    def jacobian_values(self, x, parts=None):
        return self._sum_on_pattern(lambda o: o.jacobian_values(x), parts)

    def residual(self, x, parts=None):
        return sum(o.residual(x) for q, o in enumerate(self.parts) if parts is None or q in parts)

    def al_residual(self, x, lam):
        return sum(o.al_residual(x, lam) for o in self.parts)

    def force(self):
        return sum(o.force() for o in self.parts)

    def mass(self, density):
        l = np.zeros(self.n_dofs)
        res = {}

        def one(o):
            v, lq = o.mass(density)
            res[id(o)] = lq
            return v
        v = self._sum_on_pattern(one)
        for lq in res.values():
            l += lq
        return v, l

    def close(self):
        for o in self.parts:
            o.close()
