"""CPU oracle of the output side — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement (plain loops, numpy ``pinv`` like the reference) of
  ip_state_vals              src/mech/mat/linear-elastic.jl:140-142, von-mises.jl:159-168, drucker-prager.jl:152-163
  stress_strain_dict         src/tools/tensors.jl:162-216   (eigvals :80-93, J2 :24-27)
  reg_terms                  src/fe-model.jl:490-503
  nodal_patch_recovery       src/fe-model.jl:506-692
The reference has no test that pins recovered nodal values numerically (test/mech/*.jl only run the analyses), so for
this piece **parity is unpinned** beyond exactness properties: a field that is a polynomial of the regression basis is
recovered exactly at every node (tests/test_recovery.py).  Only tests/ may import this.
"""
from __future__ import annotations

import numpy as np

SR2 = np.sqrt(2.0)
MAT_LE, MAT_VM, MAT_DP = 1, 2, 3
NCORNER = {1: 4, 2: 4, 3: 8, 4: 8, 5: 4}          # basic_shape.npoints of QUAD4 QUAD8 HEX8 HEX20 TET10


def J2(s):
    t23, t13, t12 = s[3] / SR2, s[4] / SR2, s[5] / SR2
    return 1 / 6 * ((s[0] - s[1]) ** 2 + (s[1] - s[2]) ** 2 + (s[2] - s[0]) ** 2) + t23 * t23 + t13 * t13 + t12 * t12


def eigvals(s):
    t11, t22, t33, t23, t13, t12 = s[0], s[1], s[2], s[3] / SR2, s[4] / SR2, s[5] / SR2
    L = np.linalg.eigvalsh(np.array([[t11, t12, t13], [t12, t22, t23], [t13, t23, t33]]))
    return L[::-1]


def ip_state_vals(kind, planestrain, sig, eps, epa):
    """Ordered (name, value) pairs of one integration point."""
    svm = np.sqrt(3 * J2(sig))
    s1, s2, s3 = eigvals(sig)
    if planestrain:
        d = [("σxx", sig[0]), ("σyy", sig[1]), ("σzz", sig[2]), ("σyz", sig[3] / SR2), ("σxz", sig[4] / SR2),
             ("σxy", sig[5] / SR2), ("σvm", svm), ("σ1", s1), ("σ3", s3), ("εxx", eps[0]), ("εyy", eps[1]), ("εzz", eps[2]),
             ("εxy", eps[5] / SR2)]
    else:
        d = [("σxx", sig[0]), ("σyy", sig[1]), ("σzz", sig[2]), ("σyz", sig[3] / SR2), ("σxz", sig[4] / SR2),
             ("σxy", sig[5] / SR2), ("σvm", svm), ("σ1", s1), ("σ2", s2), ("σ3", s3), ("εxx", eps[0]), ("εyy", eps[1]),
             ("εzz", eps[2]), ("εyz", eps[3] / SR2), ("εxz", eps[4] / SR2), ("εxy", eps[5] / SR2)]
    if kind == MAT_VM:
        d.append(("ep", epa))
    elif kind == MAT_DP:
        d += [("epa", epa), ("j1", sig[0] + sig[1] + sig[2]), ("srj2d", np.sqrt(J2(sig)))]
    return d


def reg_terms(x, y, z, nterms, ndim):
    if ndim == 3:
        return {7: (1.0, x, y, z, x * y, y * z, x * z), 4: (1.0, x, y, z), 1: (1.0,)}[nterms]
    return {6: (1.0, x, y, x * y, x * x, y * y), 4: (1.0, x, y, x * y), 3: (1.0, x, y), 1: (1.0,)}[nterms]


def build_patches(nnodes, conn, ncorner, at_bound):
    """Internal patches + adopted boundary patches (fe-model.jl:529-582) -> list of element lists per node."""
    patches = [[] for _ in range(nnodes)]
    bry = [[] for _ in range(nnodes)]
    for e in range(conn.shape[0]):
        for n in conn[e, :ncorner]:
            (bry if at_bound[n] else patches)[n].append(e)
    haspatch = np.zeros(nnodes, dtype=bool)
    for p in patches:
        for e in p:
            haspatch[conn[e]] = True
    orphans = [n for n in range(nnodes) if not haspatch[n] and at_bound[n]]
    if orphans:
        for k in (3, 2, 1):
            for n in orphans:
                if len(bry[n]) >= k:
                    patches[n] = bry[n]
            for n in orphans:
                for e in patches[n]:
                    haspatch[conn[e]] = True
            orphans = [n for n in orphans if not haspatch[n]]
            if not orphans:
                break
    return patches


def nodal_patch_recovery(ndim, planestrain, coords, conn, shape_id, ipcoords, elem_kind, sig, eps, epa, at_bound,
                         return_deficient=False):
    """-> (V (nnodes, nfields), field names).  ipcoords (nelem*nip, 3), state arrays element-major."""
    nnodes, nelem = coords.shape[0], conn.shape[0]
    nip = ipcoords.shape[0] // nelem
    if not at_bound.any():
        return np.zeros((nnodes, 0)), []
    patches = build_patches(nnodes, conn, NCORNER[shape_id], at_bound)
    vals = [[ip_state_vals(elem_kind[e], planestrain, sig[e * nip + q], eps[e * nip + q], epa[e * nip + q])
             for q in range(nip)] for e in range(nelem)]
    fields = []
    for e in range(nelem):
        for k, _ in vals[e][0]:
            if k not in fields:
                fields.append(k)
    fidx = {k: i for i, k in enumerate(fields)}
    V = np.zeros((nnodes, len(fields)))
    R = np.zeros((nnodes, len(fields)), dtype=np.int64)
    deficient = np.zeros((nnodes, len(fields)), dtype=bool)      # touched by a rank-deficient sub-patch (pinv = minimum norm)
    for patch in patches:
        if not patch:
            continue
        pf = []
        for e in patch:
            for k, _ in vals[e][0]:
                if k not in pf:
                    pf.append(k)
        last, invM, N, nodes = None, None, None, None
        for f in pf:
            sub = [e for e in patch if f in dict(vals[e][0])]
            if sub != last:
                last = sub
                ips = np.concatenate([ipcoords[e * nip:(e + 1) * nip] for e in sub])
                nodes = list(dict.fromkeys(int(n) for e in sub for n in conn[e]))      # unique, first-seen order
                m = ips.shape[0]
                if ndim == 3:
                    nt = 7 if m >= 7 else 4 if m >= 4 else 1
                else:
                    nt = 6 if m >= 6 else 4 if m >= 4 else 3 if m >= 3 else 1
                M = np.array([reg_terms(p[0], p[1], p[2], nt, ndim) for p in ips])
                invM = np.linalg.pinv(M)
                full_rank = np.linalg.matrix_rank(M) == nt
                N = np.array([reg_terms(coords[n, 0], coords[n, 1], coords[n, 2], nt, ndim) for n in nodes])
            W = np.array([dict(v)[f] for e in sub for v in vals[e]])
            Vn = N @ (invM @ W)
            V[nodes, fidx[f]] += Vn
            R[nodes, fidx[f]] += 1
            if not full_rank:
                deficient[nodes, fidx[f]] = True
    with np.errstate(invalid="ignore", divide="ignore"):
        V = V / R
    V[np.isnan(V)] = 0.0
    if return_deficient:
        return V, fields, deficient
    return V, fields
