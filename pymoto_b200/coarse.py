"""Level-1 Galerkin operator straight from the element densities (host-side set-up of ``pmb_galerkin_direct``).

The reference forms ``Ac = R^T A R`` with two sparse matrix products (pymoto/solvers/iterative.py:173).  On the finest
level ``A = P (sum_e s_e Ke) P + bcdiagval (I - P)`` (pymoto/modules/assembly.py:208-272) and every fine element lies
inside ONE coarse element, whose eight nodes interpolate all of its nodes.  Hence

    Ac = sum_E sum_{p < 8} s_{child(E, p)} * G_p  +  bcdiagval * R^T (I - P) R,      G_p = R_p^T Ke R_p,

with eight constant 24 x 24 matrices ``G_p`` (``R_p``: trilinear weights 1, 1/2, 1/4, 1/8 of child ``p``'s nodes with
respect to the parent's nodes).  Children that touch a Dirichlet dof use ``R_p^T (P_e Ke P_e) R_p`` instead: one extra
table entry per distinct (child position, mask pattern) -- a handful for face / edge supports -- looked up through a
per-coarse-element index.  The constant ``bcdiagval R^T (I - P) R`` term is a short list of (entry, value) pairs.
The fine matrix is never read: 33 MB of densities in, the coarse values out, instead of 8.2 GB + 9.7 GB of intermediates
at 256x128x128.  3-D grids; everything here runs once per (Ke, bc) pair.
"""
import numpy as np


def _w1(t, c):
    """1-D prolongation weight of fine position t in {0, 1, 2} inside a coarse element w.r.t. its node c in {0, 1}."""
    return 1.0 if t == 2 * c else (0.5 if abs(t - 2 * c) == 1 else 0.0)


def child_interpolation(ndof):
    """R[p]: (8 ndof) x (8 ndof) interpolation from the coarse element's dofs to the dofs of its child p = px + 2 py + 4 pz
    (node order of the reference: x fastest, pymoto/common/domain.py:132-142; weights of iterative.py:182-219)."""
    R = np.zeros((8, 8 * ndof, 8 * ndof))
    for p in range(8):
        px, py, pz = p & 1, (p >> 1) & 1, p >> 2
        for f in range(8):
            fx, fy, fz = f & 1, (f >> 1) & 1, f >> 2
            for c in range(8):
                cx, cy, cz = c & 1, (c >> 1) & 1, c >> 2
                w = _w1(px + fx, cx) * _w1(py + fy, cy) * _w1(pz + fz, cz)
                for d in range(ndof):
                    R[p, f * ndof + d, c * ndof + d] = w
    return R


def _table_layout(G, ndof):
    """(8 ndof)^2 row-major -> [a][b][d][c] (coarse row node, coarse column node, row dof, column dof)."""
    return np.ascontiguousarray(G.reshape(8, ndof, 8, ndof).transpose(0, 2, 1, 3))


def element_patterns(dims, ndof, bc):
    """Fine elements that touch a Dirichlet dof: (element numbers, bit patterns), bit a*ndof+d set for a masked local dof."""
    nx, ny, nz = dims
    NX, NY = nx + 1, ny + 1
    bc = np.asarray(bc, dtype=np.int64).ravel()
    if bc.size == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    node, d = bc // ndof, bc % ndof
    i, j, k = node % NX, (node // NX) % NY, node // (NX * NY)
    els, bits = [], []
    for a in range(8):  # the element in which the node is local node a sits at (i - ax, j - ay, k - az)
        ei, ej, ek = i - (a & 1), j - ((a >> 1) & 1), k - (a >> 2)
        ok = (ei >= 0) & (ei < nx) & (ej >= 0) & (ej < ny) & (ek >= 0) & (ek < nz)
        els.append(((ek * ny + ej) * nx + ei)[ok])
        bits.append((np.int64(1) << (a * ndof + d))[ok])
    els, bits = np.concatenate(els), np.concatenate(bits)
    order = np.argsort(els, kind="stable")
    els, bits = els[order], bits[order]
    uniq, start = np.unique(els, return_index=True)
    return uniq, np.bitwise_or.reduceat(bits, start)


def coarse_entry_offsets(cdims, ndof, I, J, K, CI, CJ, CK, d, c, k0c=0):
    """Offset (in doubles, relative to the first entry of coarse plane k0c) of entry (row node (I,J,K) dof d, column node
    (CI,CJ,CK) dof c) in the closed-form stencil-CSR layout (pmb_common.cuh: block_offset)."""
    NX, NY, NZ = (n + 1 for n in cdims)
    cnt = lambda i, M: 3 - (i == 0).astype(np.int64) - (i == M - 1).astype(np.int64)  # noqa: E731
    pre = lambda i, M: 3 * i - (i > 0).astype(np.int64) - (i == M).astype(np.int64)  # noqa: E731
    Sx, Sy = 3 * NX - 2, 3 * NY - 2
    cx, cy, cz = cnt(I, NX), cnt(J, NY), cnt(K, NZ)
    bo = pre(K, NZ) * Sy * Sx + cz * (pre(J, NY) * Sx + cy * pre(I, NX))
    k0 = np.asarray(k0c, dtype=np.int64)
    bo0 = pre(k0, NZ) * Sy * Sx
    ilo, jlo, klo = np.maximum(I - 1, 0), np.maximum(J - 1, 0), np.maximum(K - 1, 0)
    nbr = ((CK - klo) * cy + (CJ - jlo)) * cx + (CI - ilo)
    L = cx * cy * cz * ndof
    return ndof * ndof * (bo - bo0) + d * L + nbr * ndof + c


def dirichlet_term(dims, ndof, bc, bcdiag, k0c=0, k1c=None):
    """bcdiagval * R^T (I - P) R restricted to coarse planes [k0c, k1c): unique entry offsets and the values to add."""
    nx, ny, nz = dims
    NX, NY = nx + 1, ny + 1
    cdims = (nx // 2, ny // 2, nz // 2)
    k1c = cdims[2] + 1 if k1c is None else k1c
    bc = np.asarray(bc, dtype=np.int64).ravel()
    if bc.size == 0 or bcdiag == 0.0:
        return np.zeros(0, np.int64), np.zeros(0)
    node, d = bc // ndof, bc % ndof
    fi, fj, fk = node % NX, (node // NX) % NY, node // (NX * NY)

    def parents(f):  # two slots per dimension: (coarse index, weight); an even fine index has one parent (second weight 0)
        odd = (f & 1) == 1
        return (f >> 1, np.where(odd, 0.5, 1.0)), ((f >> 1) + 1, np.where(odd, 0.5, 0.0))

    px, py, pz = parents(fi), parents(fj), parents(fk)
    idx, val = [], []
    for a in range(8):
        (I, wi), (J, wj), (K, wk) = px[a & 1], py[(a >> 1) & 1], pz[a >> 2]
        wa = wi * wj * wk
        for b in range(8):
            (CI, vi), (CJ, vj), (CK, vk) = px[b & 1], py[(b >> 1) & 1], pz[b >> 2]
            w = wa * vi * vj * vk
            ok = (w != 0.0) & (K >= k0c) & (K < k1c)
            if not ok.any():
                continue
            idx.append(coarse_entry_offsets(cdims, ndof, I[ok], J[ok], K[ok], CI[ok], CJ[ok], CK[ok], d[ok], d[ok], k0c))
            val.append(bcdiag * w[ok])
    if not idx:
        return np.zeros(0, np.int64), np.zeros(0)
    idx, val = np.concatenate(idx), np.concatenate(val)
    uniq, inv = np.unique(idx, return_inverse=True)
    out = np.zeros(uniq.size)
    np.add.at(out, inv, val)
    return uniq, out


def build_tables(ke, ndof, dims, bc=None, bcdiag=0.0, k0c=0, k1c=None):
    """Host arrays for ``pmb_galerkin_direct``:

    ``Gtab`` (ntab, 8, 8, ndof, ndof): entries 0..7 = unmasked children, then one per (child, mask pattern) in use;
    ``cidx`` (coarse elements,) int32: -1 = no child touches a Dirichlet dof, else row of ``child_ids``;
    ``child_ids`` (m, 8) uint16: table entry of every child of such a coarse element;
    ``bc_idx``, ``bc_val``: the Dirichlet diagonal term for the coarse planes [k0c, k1c).
    """
    nx, ny, nz = dims
    assert nz > 0 and nx % 2 == 0 and ny % 2 == 0 and nz % 2 == 0
    ke = np.asarray(ke, dtype=np.float64).reshape(8 * ndof, 8 * ndof)
    R = child_interpolation(ndof)
    tabs = [_table_layout(R[p].T @ ke @ R[p], ndof) for p in range(8)]
    cx, cy, cz = nx // 2, ny // 2, nz // 2
    cidx = child_ids = None
    if bc is not None and np.size(bc) > 0:
        els, pats = element_patterns(dims, ndof, bc)
        ei, ej, ek = els % nx, (els // nx) % ny, els // (nx * ny)
        p = (ei & 1) + 2 * (ej & 1) + 4 * (ek & 1)
        key = p.astype(np.int64) << 32 | pats
        ukey, inv = np.unique(key, return_inverse=True)
        if 8 + ukey.size > 65535:
            raise ValueError("too many distinct Dirichlet patterns for the direct coarse-operator build")
        bitsel = np.arange(8 * ndof)
        for kk in ukey:
            pp, pat = int(kk >> 32), int(kk & 0xFFFFFFFF)
            keep = ((pat >> bitsel) & 1) == 0
            kem = ke * keep[:, None] * keep[None, :]
            tabs.append(_table_layout(R[pp].T @ kem @ R[pp], ndof))
        E = ((ek >> 1) * cy + (ej >> 1)) * cx + (ei >> 1)
        uE, invE = np.unique(E, return_inverse=True)
        child_ids = np.tile(np.arange(8, dtype=np.uint16), (uE.size, 1))
        child_ids[invE, p] = (8 + inv).astype(np.uint16)
        cidx = np.full(cx * cy * cz, -1, dtype=np.int32)
        cidx[uE] = np.arange(uE.size, dtype=np.int32)
    bc_idx, bc_val = dirichlet_term(dims, ndof, [] if bc is None else bc, float(bcdiag), k0c, k1c)
    return dict(Gtab=np.ascontiguousarray(np.stack(tabs)), cidx=cidx, child_ids=child_ids, bc_idx=bc_idx, bc_val=bc_val)


def emulate(tables, ndof, dims, s, k0c=0, k1c=None):
    """Plain-numpy evaluation of the direct formula on coarse planes [k0c, k1c): the stencil-CSR ``data`` array that
    ``pmb_galerkin_direct`` + ``pmb_scatter_add`` produce.  Test infrastructure for the host logic (small grids only)."""
    nx, ny, nz = dims
    cx, cy, cz = nx // 2, ny // 2, nz // 2
    k1c = cz + 1 if k1c is None else k1c
    NXc, NYc, NZc = cx + 1, cy + 1, cz + 1
    G, cidx, cids = tables["Gtab"], tables["cidx"], tables["child_ids"]
    s3 = np.asarray(s).reshape(nz, ny, nx)
    one = lambda v: np.array([v], dtype=np.int64)  # noqa: E731
    last = coarse_entry_offsets((cx, cy, cz), ndof, one(cx), one(cy), one(k1c - 1), one(cx), one(cy), one(min(k1c, cz)), one(ndof - 1),
                                one(ndof - 1), k0c)
    data = np.zeros(int(last[0]) + 1)
    for EK in range(max(k0c - 1, 0), min(k1c, cz)):
        for EJ in range(cy):
            for EI in range(cx):
                E = (EK * cy + EJ) * cx + EI
                ids = np.arange(8) if cidx is None or cidx[E] < 0 else cids[cidx[E]].astype(np.int64)
                AE = np.zeros((8, 8, ndof, ndof))
                for p in range(8):
                    AE += s3[2 * EK + (p >> 2), 2 * EJ + ((p >> 1) & 1), 2 * EI + (p & 1)] * G[ids[p]]
                for a in range(8):
                    I, J, K = EI + (a & 1), EJ + ((a >> 1) & 1), EK + (a >> 2)
                    if not (k0c <= K < k1c):
                        continue
                    for b in range(8):
                        CI, CJ, CK = EI + (b & 1), EJ + ((b >> 1) & 1), EK + (b >> 2)
                        for d in range(ndof):
                            for c in range(ndof):
                                off = coarse_entry_offsets((cx, cy, cz), ndof, one(I), one(J), one(K), one(CI), one(CJ), one(CK), one(d),
                                                           one(c), k0c)[0]
                                data[off] += AE[a, b, d, c]
    data[tables["bc_idx"]] += tables["bc_val"]
    return data
