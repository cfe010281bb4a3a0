"""Host-side builder of the tiled point-location index over the SOURCE mesh M0.

The reference interpolates every snapshot from the ORIGINAL mesh at every step
(/root/reference/Env2DAirfoil.py:556-568, quirk B6), so M0's index is built once per
environment and reused by every step.  DOLFIN answers the point queries with a
bounding-box tree over M0's cells; here M0 is cut into k-d leaves of ~`leaf_cells`
cells and everything a leaf's queries touch -- vertex coordinates, the P2/P1
coefficients of all T snapshots, the cell -> local-dof table and a micro-grid of
candidate lists -- is stored contiguously, so one CTA pulls a leaf into shared memory
with a handful of bulk copies and serves all target points inside the leaf's rectangle
from there (csrc/interp_tiled.cu).

Layout (flat arrays over leaves, per-leaf extents padded so that every section starts on
a 16-byte boundary and has a 16-byte-multiple size):

  tree      f64 [n_leaves-1]   split planes in heap order; the split axis is the mantissa LSB
  leaf_info i32 [n_leaves][16] vbase, nv, dbase, np2, cbase, nc, bbase, nbin, ebase, nent, gx, gy, ncl(real)
  leaf_rect f64 [n_leaves][4]  micro-grid origin x0, y0 and 1/dx, 1/dy
  coordsL   f64 [sum nv][2]    leaf-local vertex coordinates
  UL        f64 [T][sum np2][2]  leaf-local P2 coefficients (vertex dofs, then edge dofs)
  PL        f64 [T][sum nv]
  gidL      i32 [sum nc]       global cell id of each local cell (ascending inside a leaf)
  cvL       u16 [sum nc][6]    local dof ids: 3 vertices, 3 edges (index into the leaf's UL rows)
  binptrL   u16 [sum nbin]     micro-grid CSR pointer (gx*gy+1 valid entries)
  binsL     u16 [sum nent]     local cell ids per bin, ascending
  leaf_base i32 [n_leaves+1]   offsets of the per-leaf target-record buckets (fixed capacities)

A cell belongs to every leaf (and every micro-bin) its bounding box, inflated by `eps`, overlaps, so
any cell containing a query point to the barycentric tolerance is among the point's candidates and the
"lowest-index containing cell" pin is decided exactly as by a brute-force search.
Pure numpy: no GPU needed, unit-tested on CPU against the oracle's brute-force locate.
"""
from __future__ import annotations

import math

import numpy as np

GRID_EPS = 1e-9   # the same inflation csrc/geom.cu uses for its uniform grid
INFO_STRIDE = 16


def _pad(n, m):
    return (n + m - 1) // m * m


class TileIndex:
    __slots__ = ("n_leaves", "depth", "T", "tree", "leaf_info", "leaf_rect", "coordsL", "UL", "PL", "gidL", "cvL",
                 "binptrL", "binsL", "max_nv", "max_np2", "max_nc", "max_nbin", "max_nent", "u_stride", "p_stride",
                 "leaf_lo", "leaf_hi", "leaf_base", "total_cap", "leaf_bytes", "smem_bytes", "ctas_per_sm")

    def nbytes(self):
        return sum(getattr(self, k).nbytes for k in ("tree", "leaf_info", "leaf_rect", "coordsL", "UL", "PL", "gidL",
                                                      "cvL", "binptrL", "binsL"))

    def leaf_of(self, pts):
        """Leaf id of each query point -- the descent the classify kernel performs."""
        node = np.zeros(len(pts), dtype=np.int64)
        for _ in range(self.depth):
            s = self.tree[node]
            d = (s.view(np.int64) & 1).astype(np.int64)
            right = pts[np.arange(len(pts)), d] >= s
            node = 2 * node + 1 + right
        return node - (self.n_leaves - 1)


SM_SMEM_BYTES = 228 * 1024      # shared memory per SM (sm_100a); each resident CTA also costs 1 KB of system-reserved space
CTA_SMEM_MAX = 227 * 1024


def pick_smem(leaf_bytes, mean_points, fit_fraction=0.98):
    """Dynamic shared memory per CTA of the per-leaf kernel: the highest occupancy class (4, 3, 2, 1 CTAs per SM) in
    which at least `fit_fraction` of the leaves fit; the kernel serves the remaining (oversize) leaves from HBM.
    512-thread CTAs (leaves with more than ~288 target points on average, csrc/geom.cu) are register-limited to two
    per SM.  Returns (bytes, ctas_per_sm) or None when even one CTA per SM does not hold enough leaves."""
    ks = (4, 3, 2, 1) if mean_points <= 288 else (2, 1)
    for k in ks:
        lim = min((SM_SMEM_BYTES // k - 1024) // 16 * 16, CTA_SMEM_MAX)
        fit = leaf_bytes <= lim
        if fit.mean() >= fit_fraction:
            return int(leaf_bytes[fit].max()), k
    return None


def tri_rect_overlap(tri, rlo, rhi, eps, chunk=1 << 20):
    """Conservative triangle / axis-aligned rectangle overlap (separating axes: x, y and the three edge normals).

    tri [n,3,2], rlo / rhi [n,2].  False only when some axis separates the two by more than `eps`, so every cell that
    can contain a query point of the rectangle to the barycentric tolerance (a few ulps, far below eps) is kept; the
    bounding-box test alone keeps about twice as many (cell, rectangle) pairs for typical triangles."""
    n = len(tri)
    out = np.empty(n, dtype=bool)
    for s0 in range(0, n, chunk):
        t = tri[s0:s0 + chunk]
        px = [np.ascontiguousarray(t[:, i, 0]) for i in range(3)]      # component arrays: elementwise min / max of three
        py = [np.ascontiguousarray(t[:, i, 1]) for i in range(3)]      # columns is much cheaper than ufunc.reduce over axis 1
        lox, loy = rlo[s0:s0 + chunk, 0], rlo[s0:s0 + chunk, 1]
        hix, hiy = rhi[s0:s0 + chunk, 0], rhi[s0:s0 + chunk, 1]
        keep = (np.minimum(np.minimum(px[0], px[1]), px[2]) <= hix + eps) & \
               (np.maximum(np.maximum(px[0], px[1]), px[2]) >= lox - eps) & \
               (np.minimum(np.minimum(py[0], py[1]), py[2]) <= hiy + eps) & \
               (np.maximum(np.maximum(py[0], py[1]), py[2]) >= loy - eps)
        cx, cy = 0.5 * (lox + hix), 0.5 * (loy + hiy)
        hx, hy = 0.5 * (hix - lox), 0.5 * (hiy - loy)
        for i in range(3):
            j, k = (i + 1) % 3, (i + 2) % 3
            nx, ny = -(py[j] - py[i]), px[j] - px[i]
            nn = np.sqrt(nx * nx + ny * ny)
            pa = nx * px[i] + ny * py[i]                          # = projection of vertex j as well
            po = nx * px[k] + ny * py[k]
            tmin, tmax = np.minimum(pa, po), np.maximum(pa, po)
            pc = nx * cx + ny * cy
            r = np.abs(nx) * hx + np.abs(ny) * hy
            tol = eps * nn + 1e-14 * (np.abs(pc) + r + np.abs(tmin) + np.abs(tmax))   # eps in length units + rounding slack
            keep &= (pc - r <= tmax + tol) & (pc + r >= tmin - tol)
        out[s0:s0 + chunk] = keep
    return out


def build_tile_index(coords, cells, cell_edges, ne, U0, P0, leaf_cells=256, eps=GRID_EPS, bins_per_cell=4.0,
                     bucket_factor=2) -> TileIndex:
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int64)
    cell_edges = np.ascontiguousarray(cell_edges, dtype=np.int64)
    U0 = np.ascontiguousarray(U0, dtype=np.float64)
    P0 = np.ascontiguousarray(P0, dtype=np.float64)
    nv, nc, T = len(coords), len(cells), U0.shape[0]
    xy = coords[cells]                                   # [nc,3,2]
    cmin, cmax = xy.min(1) - eps, xy.max(1) + eps
    cen = (xy[:, 0] + xy[:, 1] + xy[:, 2]) / 3.0
    depth = max(0, int(math.ceil(math.log2(max(1.0, nc / float(leaf_cells))))))
    n_leaves = 1 << depth
    n_int = n_leaves - 1

    # ---- k-d tree by median splits of the cell centroids (heap order) ----
    tree = np.zeros(max(n_int, 1), dtype=np.float64)
    lo = np.empty((2 * n_leaves - 1, 2))
    hi = np.empty((2 * n_leaves - 1, 2))
    g_lo, g_hi = coords.min(0), coords.max(0)
    span = np.maximum(g_hi - g_lo, 1e-300)
    lo[0], hi[0] = g_lo - 1e-6 * span, g_hi + 1e-6 * span
    order = np.arange(nc)
    seg = {0: (0, nc)}
    for node in range(n_int):
        s, e = seg.pop(node)
        d = int(np.argmax(hi[node] - lo[node]))
        sub = order[s:e]
        m = (e - s) // 2
        if e - s >= 2:
            vals = cen[sub, d]
            part = np.argpartition(vals, m)
            order[s:e] = sub[part]
            sp = float(vals[part[m]])
        else:
            sp = float(0.5 * (lo[node, d] + hi[node, d]))
        sp = min(max(sp, lo[node, d]), hi[node, d])
        # the split axis rides in the mantissa LSB; cells and query points both use the tagged value
        bits = np.array([sp], dtype=np.float64).view(np.int64)
        bits[0] = (bits[0] & ~np.int64(1)) | d
        sp = float(bits.view(np.float64)[0])
        tree[node] = sp
        l, r = 2 * node + 1, 2 * node + 2
        lo[l], hi[l] = lo[node], hi[node].copy()
        hi[l, d] = sp
        lo[r], hi[r] = lo[node].copy(), hi[node]
        lo[r, d] = sp
        seg[l], seg[r] = (s, s + m), (s + m, e)
    leaf_lo, leaf_hi = lo[n_int:], hi[n_int:]

    # ---- (leaf, cell) pairs: a cell descends into every child its inflated bbox overlaps ----
    pc = np.arange(nc, dtype=np.int64)
    pn = np.zeros(nc, dtype=np.int64)
    for _ in range(depth):
        sp = tree[pn]
        d = sp.view(np.int64) & 1
        gl = cmin[pc, d] < sp
        gr = cmax[pc, d] >= sp
        pc, pn = np.concatenate([pc[gl], pc[gr]]), np.concatenate([2 * pn[gl] + 1, 2 * pn[gr] + 2])
    pl = pn - n_int
    # the descent used bounding boxes; keep only cells whose (inflated) TRIANGLE meets the leaf's rectangle: fewer halo
    # cells and dofs per leaf, so less HBM traffic and shared memory
    keep = tri_rect_overlap(xy[pc], leaf_lo[pl], leaf_hi[pl], eps)
    pc, pl = pc[keep], pl[keep]
    o = np.lexsort((pc, pl))
    pl, pc = pl[o], pc[o]
    npairs = len(pl)
    ncl = np.bincount(pl, minlength=n_leaves)
    cptr = np.concatenate([[0], np.cumsum(ncl)])

    # ---- leaf-local vertex / edge numbering ----
    def local_ids(glob, n_glob):
        key = (pl[:, None] * n_glob + glob).ravel()
        ukey, inv = np.unique(key, return_inverse=True)
        uleaf, uid = ukey // n_glob, ukey % n_glob
        cnt = np.bincount(uleaf, minlength=n_leaves)
        ptr = np.concatenate([[0], np.cumsum(cnt)])
        local = inv.reshape(-1, 3) - ptr[pl][:, None]
        return uleaf, uid, cnt, ptr, local

    vleaf, vid, nvl, vptr, lv = local_ids(cells[pc], nv)
    eleaf, eid, nel, eptr, le = local_ids(cell_edges[pc], ne)

    # ---- padded extents ----
    nv_pad = np.array([_pad(int(x), 2) for x in nvl], dtype=np.int64)
    np2 = nv_pad + nel                                  # edge dofs follow the PADDED vertex block
    np2_pad = np2                                       # 16-byte elements: no padding needed
    nc_pad = np.array([_pad(int(x), 4) for x in ncl], dtype=np.int64)
    if int(np2.max(initial=0)) > 65535 or int(nc_pad.max(initial=0)) > 65535:
        raise ValueError("leaf too large for 16-bit local ids; lower leaf_cells")
    vbase = np.concatenate([[0], np.cumsum(nv_pad)])
    dbase = np.concatenate([[0], np.cumsum(np2_pad)])
    cbase = np.concatenate([[0], np.cumsum(nc_pad)])

    # ---- micro-grid per leaf ----
    w = np.maximum(leaf_hi - leaf_lo, 1e-300)
    n_eff = np.maximum(ncl, 1).astype(np.float64) * float(bins_per_cell)
    gx = np.clip(np.rint(np.sqrt(n_eff * w[:, 0] / w[:, 1])), 1, 64).astype(np.int64)
    gy = np.clip(np.rint(n_eff / gx), 1, 64).astype(np.int64)
    inv_dx, inv_dy = gx / w[:, 0], gy / w[:, 1]
    x0, y0 = leaf_lo[:, 0], leaf_lo[:, 1]

    def bins(v, origin, inv, g):
        return np.clip(np.floor((v - origin[pl]) * inv[pl]), 0, g[pl] - 1).astype(np.int64)

    ix0, ix1 = bins(cmin[pc, 0], x0, inv_dx, gx), bins(cmax[pc, 0], x0, inv_dx, gx)
    iy0, iy1 = bins(cmin[pc, 1], y0, inv_dy, gy), bins(cmax[pc, 1], y0, inv_dy, gy)
    wx = ix1 - ix0 + 1
    cnt = wx * (iy1 - iy0 + 1)
    start = np.concatenate([[0], np.cumsum(cnt)])
    rep = np.repeat(np.arange(npairs), cnt)
    k = np.arange(int(start[-1])) - start[rep]
    bxi, byi = ix0[rep] + k % wx[rep], iy0[rep] + k // wx[rep]
    ebin = byi * gx[pl[rep]] + bxi
    eleaf_ = pl[rep]
    # candidate lists by triangle / bin-rectangle overlap, not bounding boxes: about half the entries, so a query walks
    # about half as many candidates before it reaches its cell
    dxl, dyl = w[:, 0] / gx, w[:, 1] / gy
    keep = np.empty(len(rep), dtype=bool)
    CH = 1 << 20                                            # chunked: the gathered triangles are 48 bytes per pair
    for s0 in range(0, len(rep), CH):
        sl = slice(s0, s0 + CH)
        lf = eleaf_[sl]
        dxb, dyb = dxl[lf], dyl[lf]
        blo = np.stack([x0[lf] + bxi[sl] * dxb, y0[lf] + byi[sl] * dyb], 1)
        bhi = np.stack([x0[lf] + (bxi[sl] + 1) * dxb, y0[lf] + (byi[sl] + 1) * dyb], 1)
        keep[sl] = tri_rect_overlap(xy[pc[rep[sl]]], blo, bhi, eps)
    rep, ebin, eleaf_ = rep[keep], ebin[keep], eleaf_[keep]
    elocal = rep - cptr[eleaf_]
    nbin = gx * gy
    bbase_real = np.concatenate([[0], np.cumsum(nbin + 1)])
    o = np.argsort(eleaf_ * 4096 + ebin, kind="stable")       # keeps local cell ids ascending inside a bin
    eleaf_, ebin, elocal = eleaf_[o], ebin[o], elocal[o]
    nent = np.bincount(eleaf_, minlength=n_leaves)
    if int(nent.max(initial=0)) > 65535:
        raise ValueError("leaf micro-grid has too many entries for 16-bit pointers; lower leaf_cells")
    nbin_pad = np.array([_pad(int(x) + 1, 8) for x in nbin], dtype=np.int64)
    nent_pad = np.array([_pad(max(int(x), 1), 8) for x in nent], dtype=np.int64)
    bbase = np.concatenate([[0], np.cumsum(nbin_pad)])
    ebase = np.concatenate([[0], np.cumsum(nent_pad)])
    eptr_real = np.concatenate([[0], np.cumsum(nent)])
    binptrL = np.zeros(int(bbase[-1]), dtype=np.uint16)
    per_bin = np.bincount(bbase_real[eleaf_] + ebin, minlength=int(bbase_real[-1]))
    for L in range(n_leaves):
        nb = int(nbin[L])
        c = per_bin[bbase_real[L]: bbase_real[L] + nb]
        binptrL[bbase[L] + 1: bbase[L] + nb + 1] = np.cumsum(c)
        binptrL[bbase[L] + nb + 1: bbase[L + 1]] = nent[L]
    binsL = np.zeros(int(ebase[-1]), dtype=np.uint16)
    binsL[(ebase[eleaf_] + np.arange(len(eleaf_)) - eptr_real[eleaf_])] = elocal

    # ---- gather the leaf-local payloads ----
    vdst = vbase[vleaf] + (np.arange(len(vid)) - vptr[vleaf])
    edst = dbase[eleaf] + nv_pad[eleaf] + (np.arange(len(eid)) - eptr[eleaf])
    coordsL = np.zeros((int(vbase[-1]), 2), dtype=np.float64)
    coordsL[vdst] = coords[vid]
    UL = np.zeros((T, int(dbase[-1]), 2), dtype=np.float64)
    UL[:, dbase[vleaf] + (np.arange(len(vid)) - vptr[vleaf])] = U0[:, vid]
    UL[:, edst] = U0[:, nv + eid]
    PL = np.zeros((T, int(vbase[-1])), dtype=np.float64)
    PL[:, vdst] = P0[:, vid]
    cdst = cbase[pl] + (np.arange(npairs) - cptr[pl])
    gidL = np.zeros(int(cbase[-1]), dtype=np.int32)
    gidL[cdst] = pc
    cvL = np.zeros((int(cbase[-1]), 6), dtype=np.uint16)
    cvL[cdst, :3] = lv
    cvL[cdst, 3:] = le + nv_pad[pl][:, None]

    info = np.zeros((n_leaves, INFO_STRIDE), dtype=np.int32)
    info[:, 0], info[:, 1] = vbase[:-1], nv_pad
    info[:, 2], info[:, 3] = dbase[:-1], np2_pad
    info[:, 4], info[:, 5] = cbase[:-1], nc_pad
    info[:, 6], info[:, 7] = bbase[:-1], nbin_pad
    info[:, 8], info[:, 9] = ebase[:-1], nent_pad
    info[:, 10], info[:, 11] = gx, gy
    info[:, 12] = ncl

    ti = TileIndex()
    ti.n_leaves, ti.depth, ti.T = n_leaves, depth, T
    ti.tree = tree
    ti.leaf_info = info
    ti.leaf_rect = np.ascontiguousarray(np.stack([x0, y0, inv_dx, inv_dy], 1))
    ti.coordsL, ti.UL, ti.PL, ti.gidL, ti.cvL, ti.binptrL, ti.binsL = coordsL, UL, PL, gidL, cvL, binptrL, binsL
    ti.max_nv, ti.max_np2 = int(nv_pad.max()), int(np2_pad.max())
    ti.max_nc, ti.max_nbin, ti.max_nent = int(nc_pad.max()), int(nbin_pad.max()), int(nent_pad.max())
    ti.u_stride, ti.p_stride = int(dbase[-1]), int(vbase[-1])
    ti.leaf_lo, ti.leaf_hi = leaf_lo, leaf_hi
    # target-record buckets: twice the leaf's own share of M0's P2 points (+ slack); a coarsened or smoothed copy of
    # M0 never exceeds it, anything denser spills to the kernel's overflow list
    mid = 0.5 * coords[cells[:, [1, 0, 0]]] + 0.5 * coords[cells[:, [2, 2, 1]]]
    own = np.bincount(ti.leaf_of(coords), minlength=n_leaves) + \
        np.bincount(ti.leaf_of(mid.reshape(-1, 2)), minlength=n_leaves) // 2
    cap = (int(bucket_factor) * own + 32 + 3) // 4 * 4
    ti.leaf_base = np.concatenate([[0], np.cumsum(cap)]).astype(np.int32)
    ti.total_cap = int(ti.leaf_base[-1])
    # shared memory each leaf's own sections need (csrc/interp_tiled.cuh: tile_leaf_bytes) and the per-CTA allocation
    ti.leaf_bytes = 16 + 16 * nv_pad + T * (16 * np2_pad + 8 * nv_pad) + 16 * nc_pad + 2 * nbin_pad + 2 * nent_pad
    picked = pick_smem(ti.leaf_bytes, ti.total_cap / (2.0 * n_leaves))
    ti.smem_bytes, ti.ctas_per_sm = picked if picked is not None else (0, 0)
    return ti


def emulate_locate(ti: TileIndex, pts, tol=1e-12):
    """numpy emulation of csrc/interp_tiled.cu's locate (candidate order and tests), for CPU unit tests.

    Returns (cell_of [n] int32 with -1 for misses, lam [n,3], local cell id, leaf id)."""
    leaf = ti.leaf_of(pts)
    cell_of = np.full(len(pts), -1, dtype=np.int32)
    lam = np.zeros((len(pts), 3))
    lcell = np.full(len(pts), -1, dtype=np.int64)
    for i, (L, (px, py)) in enumerate(zip(leaf, pts)):
        inf = ti.leaf_info[L]
        x0, y0, idx, idy = ti.leaf_rect[L]
        gx, gy = int(inf[10]), int(inf[11])
        bx = int(min(max(math.floor((px - x0) * idx), 0), gx - 1))
        by = int(min(max(math.floor((py - y0) * idy), 0), gy - 1))
        b = by * gx + bx
        bp = ti.binptrL[inf[6]: inf[6] + inf[7]]
        for s in range(int(bp[b]), int(bp[b + 1])):
            lc = int(ti.binsL[inf[8] + s])
            cv = ti.cvL[inf[4] + lc]
            (ax, ay), (bx_, by_), (cx, cy) = (ti.coordsL[inf[0] + int(cv[k])] for k in range(3))
            d1x, d1y, d2x, d2y = bx_ - ax, by_ - ay, cx - ax, cy - ay
            det = d1x * d2y - d2x * d1y
            qx, qy = px - ax, py - ay
            l1 = (qx * d2y - d2x * qy) / det
            l2 = (d1x * qy - qx * d1y) / det
            l0 = 1.0 - l1 - l2
            if min(l0, l1, l2) >= -tol:
                cell_of[i] = ti.gidL[inf[4] + lc]
                lam[i] = (l0, l1, l2)
                lcell[i] = lc
                break
    return cell_of, lam, lcell, leaf
