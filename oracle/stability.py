"""Stability indicator of the arc-length solvers — oracle side (TEST INFRASTRUCTURE ONLY).

Restates gsALMBase<T>::_computeStability / gsStaticBase<T>::_computeStabilityDet, bifurcation method "Determinant"
(src/gsALMSolvers/gsALMBase.hpp:546-611, src/gsStaticSolvers/gsStaticBase.h:161-179): factorise the tangent K = L D L^T
(gsSparseSolver<>::SimplicialLDLT = Eigen 3.4 SimplicialLDLT, third-party, not in /root/reference: a permutation but NO numerical
pivoting), m_stabilityVec = vectorD(), m_negatives = countNegatives(vectorD) (gsALMBase.hpp:621-631 counts entries < 0),
m_indicator = min(vectorD), stability() = (m_indicator < 0) ? 1 : -1.

PARITY UNPINNED against Eigen's AMD ordering (the pivots depend on the ordering); what is ordering-independent — the number of
negative pivots and the sign of the smallest one (Sylvester's law of inertia) — is pinned against numpy's eigenvalues."""
from __future__ import annotations

import numpy as np


def ldlt_pivots(A):
    """D of A = L D L^T without pivoting (dense, right-looking) — plain restatement of the algorithm"""
    A = np.array(A, dtype=np.float64)
    n = A.shape[0]
    D = np.zeros(n)
    for k in range(n):
        d = A[k, k]
        D[k] = d
        l = A[k + 1:, k] / d
        A[k + 1:, k + 1:] -= np.outer(l, l) * d
    return D


def node_major_permutation(prob):
    """perm[g] = position of free DoF g in the node-major ordering along the shorter direction; a matched DoF takes the position
    of its first control point (the ordering the product's band factorisation uses; any ordering gives the same inertia)"""
    n1, n2 = prob.surface.n
    ncp = n1 * n2
    dm = np.asarray(prob.dof_map).reshape(3, ncp)
    i = np.arange(ncp)
    i1, i2 = i % n1, i // n1
    node = i2 * n1 + i1 if n1 <= n2 else i1 * n2 + i2
    key = np.full(prob.n_free, np.iinfo(np.int64).max, dtype=np.int64)
    for c in range(3):
        free = dm[c] < prob.n_free
        np.minimum.at(key, dm[c][free], (node * 3 + c)[free])
    order = np.lexsort((np.arange(prob.n_free), key))
    perm = np.empty(prob.n_free, dtype=np.int64)
    perm[order] = np.arange(prob.n_free)
    return perm


def stability(K_dense, perm=None):
    """(indicator, negatives, D in the original DoF order)"""
    n = K_dense.shape[0]
    if perm is None:
        perm = np.arange(n)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    D = ldlt_pivots(K_dense[np.ix_(inv, inv)])
    return D.min(), int((D < 0).sum()), D[perm]
