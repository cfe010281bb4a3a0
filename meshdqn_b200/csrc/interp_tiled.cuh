// interp_tiled.cuh -- tiled re-interpolation (included by geom.cu, so it is compiled with -fmad=false and shares
// InterpArgs / target_point / the closest-cell fallback with the uniform-grid path).
//
// Replaces DOLFIN's Function.interpolate + Function.__call__ (/root/reference/Env2DAirfoil.py:556-568,515-522)
// for large source meshes.  M0 is pre-cut on the host (meshdqn_b200/tile_index.py) into k-d leaves whose whole
// working set -- local vertex coordinates, the P2/P1 coefficients of all T snapshots, the cell->dof table and a
// micro-grid of candidate lists -- is contiguous in HBM.  Per step:
//   k_tile_classify : every target dof point descends the k-d tree and appends its (x, y, id) record to its leaf's
//                     bucket (warp-aggregated atomics; bucket capacities are fixed when the index is built, points
//                     beyond a bucket's capacity go to an overflow list);
//   k_tile_interp   : one CTA per leaf; one thread issues bulk (TMA) copies of the leaf's sections into shared
//                     memory while all threads load their records; point location walks the micro-bin's
//                     candidates in ascending cell id (first hit = lowest index), a division-free conservative
//                     reject skips cells that cannot contain the point, and the exact test and the P2/P1
//                     evaluation use the same operation sequence as the uniform-grid kernel and the oracle;
//   k_tile_overflow : the (normally empty) overflow list, served straight from the leaf arrays in HBM; also
//                     re-zeroes the overflow counter for the next call (each per-leaf CTA re-zeroes its own);
//   k_interp_miss   : closest-cell extrapolation for points outside every source cell (geom.cu).
// HBM traffic is ~ the algorithmic bytes (every source byte is read once per leaf that overlaps it); all gathers
// hit shared memory.
#pragma once

struct TileArgs {
    InterpArgs a;               // targets, outputs, miss list, tol, T (source arrays only for the fallback)
    const double *tree;
    const int *leaf_info;       // [n_leaves][16]
    const double *leaf_rect;    // [n_leaves][4]
    const double2 *coordsL;
    const double2 *UL;
    const double *PL;
    const int *gidL;
    const unsigned short *cvL, *binptrL, *binsL;
    const int *leaf_base;       // [n_leaves+1] bucket offsets (capacity prefix sums)
    long long u_stride, p_stride;
    int n_leaves, depth;
    int smem_cap;               // dynamic shared memory of this launch; a leaf that needs more is served from HBM
    int pf_ahead;               // L2 prefetch distance in leaves (CTAs resident on the GPU), 0 = off
    int *leaf_cnt;              // [n_leaves] zero on entry, zero again on exit
    int *ovf_count;             // zero on entry / exit
    unsigned int *ticket;       // zero on entry / exit
    int *ovf_list;              // [np]
    double2 *rec_xy;            // [total_cap] bucketed target coordinates
    int *rec_id;                // [total_cap] bucketed target ids
};

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase)
{
    unsigned ok;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    } while (!ok);
}

__device__ __forceinline__ int tile_leaf_of(const double *__restrict__ tree, int depth, double px, double py)
{
    int node = 0;
    for (int l = 0; l < depth; ++l) {
        const double s = __ldg(tree + node);
        const int d = (int)(__double_as_longlong(s) & 1LL);
        const double v = d ? py : px;
        node = 2 * node + 1 + (v >= s ? 1 : 0);
    }
    return node - ((1 << depth) - 1);
}

// A leaf's arrays, in shared memory (k_tile_interp) or in HBM (k_tile_overflow): same code, same bits.
struct LeafView {
    const double2 *xy, *U;
    const double *P;
    const int *gid;
    const unsigned *cv;             // 3 words per cell: v0|v1, v2|e0, e1|e2
    const unsigned short *bp, *bins;
    int u_rows, p_rows;             // rows between snapshots
    double x0, y0, inv_dx, inv_dy;
    int gx, gy;
};

struct TileHit {
    int lc;                         // leaf-local cell, -1: no cell contains the point
    double l0, l1, l2;
    unsigned w0, w1;                // the cell's first two dof words (v0|v1, v2|e0)
};

// Point location inside a leaf: the micro-bin's candidates in ascending cell id, first hit = lowest index.
__device__ __forceinline__ TileHit tile_locate(const LeafView &v, double px, double py, double margin, double tol)
{
    const int bx = min(max((int)floor((px - v.x0) * v.inv_dx), 0), v.gx - 1);
    const int by = min(max((int)floor((py - v.y0) * v.inv_dy), 0), v.gy - 1);
    const int b = by * v.gx + bx;
    const int s1 = v.bp[b + 1];
    int s = v.bp[b];
    TileHit h;
    h.lc = -1;
    h.l0 = h.l1 = h.l2 = 0.0;
    h.w0 = h.w1 = 0;
    while (true) {
        // phase A (division-free): advance to the next candidate the conservative test cannot reject
        double n1 = 0.0, n2 = 0.0, det = 1.0;
        int c = -1;
        while (s < s1) {
            const int cc = v.bins[s++];
            h.w0 = v.cv[3 * cc];
            h.w1 = v.cv[3 * cc + 1];
            const double2 A = v.xy[h.w0 & 0xffffu], B = v.xy[h.w0 >> 16], C = v.xy[h.w1 & 0xffffu];
            const double d1x = B.x - A.x, d1y = B.y - A.y, d2x = C.x - A.x, d2y = C.y - A.y;
            det = d1x * d2y - d2x * d1y;
            const double qx = px - A.x, qy = py - A.y;
            n1 = qx * d2y - d2x * qy;
            n2 = d1x * qy - qx * d1y;
            const double ad = fabs(det);
            const double m1 = det < 0.0 ? -n1 : n1, m2 = det < 0.0 ? -n2 : n2;
            const double lim = -margin * ad;
            // lambda_k < -margin cannot pass the exact test; a degenerate cell (det == 0) is never rejected here
            if (ad > 0.0 && (m1 < lim || m2 < lim || (ad - m1) - m2 < lim)) continue;
            c = cc;
            break;
        }
        if (c < 0) break;
        // phase B (the warp reconverges here): the exact test, same operations as bary() in geom.cu.
        // A zero numerator (target point on a source vertex / edge) would take the slow path of the
        // double-precision division; 0 * det has the quotient's value and sign.
        h.l1 = n1 == 0.0 ? n1 * det : n1 / det;
        h.l2 = n2 == 0.0 ? n2 * det : n2 / det;
        h.l0 = 1.0 - h.l1 - h.l2;
        if (fmin(h.l0, fmin(h.l1, h.l2)) >= -tol) {
            h.lc = c;
            break;
        }
    }
    return h;
}

// P2 velocity / P1 pressure of all T snapshots at the located point (same operation sequence as eval_point in geom.cu)
__device__ __forceinline__ void tile_eval(const InterpArgs &a, const LeafView &v, int i, const TileHit &h)
{
    if (h.lc < 0) {
        a.cell_of[i] = -1;
        a.miss_list[atomicAdd(a.miss_count, 1)] = i;
        return;
    }
    a.cell_of[i] = v.gid[h.lc];
    const unsigned w2 = v.cv[3 * h.lc + 2];
    const double l0 = h.l0, l1 = h.l1, l2 = h.l2;
    double phi[6];
    phi[0] = l0 * (2.0 * l0 - 1.0);
    phi[1] = l1 * (2.0 * l1 - 1.0);
    phi[2] = l2 * (2.0 * l2 - 1.0);
    phi[3] = 4.0 * l1 * l2;
    phi[4] = 4.0 * l0 * l2;
    phi[5] = 4.0 * l0 * l1;
    const int dof[6] = {(int)(h.w0 & 0xffffu), (int)(h.w0 >> 16), (int)(h.w1 & 0xffffu),
                        (int)(h.w1 >> 16),     (int)(w2 & 0xffffu),   (int)(w2 >> 16)};
    const double lam[3] = {l0, l1, l2};
    const int np2t = a.nv + a.ne;
    double2 *Uo = reinterpret_cast<double2 *>(a.U) + i;
    double *Po = a.P + i;
    for (int k = 0; k < a.T; ++k) {
        const double2 *Uk = v.U + k * v.u_rows;
        const double2 u0 = Uk[dof[0]];
        double ux = phi[0] * u0.x, uy = phi[0] * u0.y;   // == 0.0 + phi*u of the reference loop
#pragma unroll
        for (int q = 1; q < 6; ++q) {
            const double2 u = Uk[dof[q]];
            ux += phi[q] * u.x;
            uy += phi[q] * u.y;
        }
        __stcs(Uo, make_double2(ux, uy));
        Uo += np2t;
        if (i < a.nv) {
            const double *Pk = v.P + k * v.p_rows;
            double pv = lam[0] * Pk[dof[0]];
#pragma unroll
            for (int q = 1; q < 3; ++q) pv += lam[q] * Pk[dof[q]];
            __stcs(Po, pv);
            Po += a.nv;
        }
    }
}

// Four points per thread, interleaved: the k-d descent is a chain of `depth` dependent (L1-resident) loads and the slot
// allocation a chain of match -> atomic -> two bucket-offset loads, so one point per thread leaves the SM waiting on
// latency; four independent chains per thread keep four times as many loads and atomics in flight.
constexpr int CLS_PPT = 4;
__global__ void __launch_bounds__(256) k_tile_classify(const TileArgs t)
{
    const int np = t.a.nv + t.a.ne;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.a.miss_count = 0;
    for (int i0 = blockIdx.x * (256 * CLS_PPT); i0 < np; i0 += gridDim.x * (256 * CLS_PPT)) {
        int idx[CLS_PPT], node[CLS_PPT];
        double px[CLS_PPT], py[CLS_PPT];
#pragma unroll
        for (int q = 0; q < CLS_PPT; ++q) {
            idx[q] = i0 + q * 256 + (int)threadIdx.x;
            px[q] = py[q] = 0.0;
            node[q] = 0;
            if (idx[q] < np) target_point(t.a, idx[q], px[q], py[q]);
        }
        for (int l = 0; l < t.depth; ++l) {
#pragma unroll
            for (int q = 0; q < CLS_PPT; ++q) {
                const double s = __ldg(t.tree + node[q]);
                const int d = (int)(__double_as_longlong(s) & 1LL);
                const double v = d ? py[q] : px[q];
                node[q] = 2 * node[q] + 1 + (v >= s ? 1 : 0);
            }
        }
        // warp-aggregated slot allocation: one atomic per distinct leaf in the warp (spatially ordered targets
        // put most of a warp in one leaf); tail lanes carry distinct dummy keys and allocate nothing
        int leaf[CLS_PPT], base[CLS_PPT], rank[CLS_PPT], leader[CLS_PPT], b0[CLS_PPT], b1[CLS_PPT];
#pragma unroll
        for (int q = 0; q < CLS_PPT; ++q) {
            leaf[q] = idx[q] < np ? node[q] - ((1 << t.depth) - 1) : -1 - lane;
            const unsigned peers = __match_any_sync(FULL, leaf[q]);
            leader[q] = __ffs(peers) - 1;
            rank[q] = __popc(peers & ((1u << lane) - 1u));
            base[q] = 0;
            b0[q] = b1[q] = 0;
            if (leaf[q] >= 0) {
                if (lane == leader[q]) base[q] = atomicAdd(t.leaf_cnt + leaf[q], __popc(peers));
                b0[q] = __ldg(t.leaf_base + leaf[q]);
                b1[q] = __ldg(t.leaf_base + leaf[q] + 1);
            }
        }
#pragma unroll
        for (int q = 0; q < CLS_PPT; ++q) {
            const int bs = __shfl_sync(FULL, base[q], leader[q]);
            if (idx[q] < np) {
                const int slot = bs + rank[q];
                if (slot < b1[q] - b0[q]) {
                    t.rec_xy[b0[q] + slot] = make_double2(px[q], py[q]);
                    t.rec_id[b0[q] + slot] = idx[q];
                } else {
                    t.ovf_list[atomicAdd(t.ovf_count, 1)] = idx[q];
                }
            }
        }
    }
}

constexpr int TILE_THREADS = 512;

// Bytes of shared memory leaf L needs: mbarrier + its own (padded) sections, carved up in the order below.
__host__ __device__ inline unsigned tile_leaf_bytes(int nv, int np2, int nc, int nbin, int nent, int T)
{
    return 16u + 16u * nv + (unsigned)T * (16u * np2 + 8u * nv) + 16u * nc + 2u * nbin + 2u * nent;
}

template <int NT, int MINB>   // CTA size / CTAs per SM the register budget is set for
__global__ void __launch_bounds__(NT, MINB) k_tile_interp(const TileArgs t)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int L = blockIdx.x;
    const int *info = t.leaf_info + 16 * L;
    // everything the prologue needs is loaded up front (independent loads, one round trip)
    const int cnt_raw = __ldcg(t.leaf_cnt + L);
    const int b0 = __ldg(t.leaf_base + L), b1 = __ldg(t.leaf_base + L + 1);
    const int4 i0 = __ldg(reinterpret_cast<const int4 *>(info));      // vbase, nv, dbase, np2
    const int4 i1 = __ldg(reinterpret_cast<const int4 *>(info) + 1);  // cbase, nc, bbase, nbin
    const int4 i2 = __ldg(reinterpret_cast<const int4 *>(info) + 2);  // ebase, nent, gx, gy
    const double2 r0 = __ldg(reinterpret_cast<const double2 *>(t.leaf_rect + 4 * L));
    const double2 r1 = __ldg(reinterpret_cast<const double2 *>(t.leaf_rect + 4 * L) + 1);
    const int cnt = min(cnt_raw, b1 - b0);
    if (cnt == 0) return;
    if (threadIdx.x == 0) t.leaf_cnt[L] = 0;   // only this CTA reads the counter: left zeroed for the next call
    const InterpArgs &a = t.a;
    const int T = a.T;
    const int vbase = i0.x, nvl = i0.y, dbase = i0.z, np2l = i0.w, cbase = i1.x, ncl = i1.y, bbase = i1.z, nbin = i1.w;
    const int ebase = i2.x, nent = i2.y;
    // The carve-up follows the leaf's OWN section sizes (every section is a 16-byte multiple), so the launch is sized
    // for the typical leaf (four CTAs per SM) and not for the largest one; the rare leaf that does not fit
    // (large cells overlapping many leaves) is served by this CTA straight from the leaf arrays in HBM.
    const bool staged = tile_leaf_bytes(nvl, np2l, ncl, nbin, nent, T) <= (unsigned)t.smem_cap;
    const double margin = a.tol * (1.0 + 1e-6) + 1e-9;
    // this thread's first record travels while the leaf is being staged
    int j = (int)threadIdx.x;
    double2 q = make_double2(0.0, 0.0);
    int i = 0;
    if (j < cnt) {
        q = __ldcs(t.rec_xy + b0 + j);
        i = __ldcs(t.rec_id + b0 + j);
    }
    LeafView v;
    v.x0 = r0.x; v.y0 = r0.y; v.inv_dx = r1.x; v.inv_dy = r1.y; v.gx = i2.z; v.gy = i2.w;
    // two call sites on purpose: in the staged one every pointer of the view derives from `smem`, so the gathers
    // compile to LDS with 32-bit addresses instead of generic loads
    if (staged) {
        unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
        double2 *s_xy = reinterpret_cast<double2 *>(smem + 16);
        double2 *s_U = s_xy + nvl;
        double *s_P = reinterpret_cast<double *>(s_U + np2l * T);
        int *s_gid = reinterpret_cast<int *>(s_P + nvl * T);
        unsigned *s_cv = reinterpret_cast<unsigned *>(s_gid + ncl);
        unsigned short *s_bp = reinterpret_cast<unsigned short *>(s_cv + 3 * ncl);
        unsigned short *s_bins = s_bp + nbin;
        if (threadIdx.x == 0) {
            // two transactions: what point location needs lands first (bar[0]), the coefficients follow (bar[1]) while
            // the threads are already walking their candidates
            mbar_init(bar, 1);
            mbar_init(bar + 1, 1);
            mbar_expect_tx(bar, 16u * nvl + 12u * ncl + 2u * nbin + 2u * nent);
            bulk_g2s(s_bp, t.binptrL + bbase, 2u * nbin, bar);
            bulk_g2s(s_bins, t.binsL + ebase, 2u * nent, bar);
            bulk_g2s(s_cv, t.cvL + 6 * (size_t)cbase, 12u * ncl, bar);
            bulk_g2s(s_xy, t.coordsL + vbase, 16u * nvl, bar);
            mbar_expect_tx(bar + 1, (unsigned)T * (16u * np2l + 8u * nvl) + 4u * ncl);
            bulk_g2s(s_gid, t.gidL + cbase, 4u * ncl, bar + 1);
            for (int k = 0; k < T; ++k) {
                bulk_g2s(s_U + k * np2l, t.UL + (size_t)k * t.u_stride + dbase, 16u * np2l, bar + 1);
                bulk_g2s(s_P + k * nvl, t.PL + (size_t)k * t.p_stride + vbase, 8u * nvl, bar + 1);
            }
        }
        v.xy = s_xy; v.U = s_U; v.P = s_P; v.gid = s_gid; v.cv = s_cv; v.bp = s_bp; v.bins = s_bins;
        v.u_rows = np2l; v.p_rows = nvl;
        if (threadIdx.x == 64 && t.pf_ahead > 0) {
            // software pipelining across CTAs through L2: ask for the sections of the leaf that the CTA taking this
            // SM slot next (blockIdx + resident CTAs) will stage, so its bulk copies hit L2 instead of paying the HBM
            // latency inside its own short life; that CTA's descriptor row is fetched here, off everybody's critical path
            const int Ln = L + t.pf_ahead;
            if (Ln < t.n_leaves) {
                const int4 *inf = reinterpret_cast<const int4 *>(t.leaf_info + 16 * Ln);
                const int4 n0 = __ldg(inf), n1 = __ldg(inf + 1), n2 = __ldg(inf + 2);
                bulk_prefetch_l2(t.binptrL + n1.z, 2u * n1.w);
                bulk_prefetch_l2(t.binsL + n2.x, 2u * n2.y);
                bulk_prefetch_l2(t.cvL + 6 * (size_t)n1.x, 12u * n1.y);
                bulk_prefetch_l2(t.coordsL + n0.x, 16u * n0.y);
                bulk_prefetch_l2(t.gidL + n1.x, 4u * n1.y);
                for (int k = 0; k < T; ++k) {
                    bulk_prefetch_l2(t.UL + (size_t)k * t.u_stride + n0.z, 16u * n0.w);
                    bulk_prefetch_l2(t.PL + (size_t)k * t.p_stride + n0.x, 8u * n0.y);
                }
            }
            if (Ln + t.pf_ahead < t.n_leaves) {   // and the descriptor / counters of the one after that
                asm volatile("prefetch.global.L2 [%0];" ::"l"(t.leaf_info + 16 * (Ln + t.pf_ahead)));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(t.leaf_rect + 4 * (Ln + t.pf_ahead)));
            }
        }
        __syncthreads();  // the barriers are initialised before anyone polls them
        if (j >= cnt) return;
        mbar_wait(bar, 0);
        bool first = true;
        while (true) {
            const TileHit h = tile_locate(v, q.x, q.y, margin, a.tol);
            if (first) { mbar_wait(bar + 1, 0); first = false; }
            tile_eval(a, v, i, h);
            j += blockDim.x;
            if (j >= cnt) break;
            q = __ldcs(t.rec_xy + b0 + j);
            i = __ldcs(t.rec_id + b0 + j);
        }
    } else {
        v.xy = t.coordsL + vbase; v.U = t.UL + dbase; v.P = t.PL + vbase; v.gid = t.gidL + cbase;
        v.cv = reinterpret_cast<const unsigned *>(t.cvL + 6 * (size_t)cbase);
        v.bp = t.binptrL + bbase; v.bins = t.binsL + ebase;
        v.u_rows = (int)t.u_stride; v.p_rows = (int)t.p_stride;
        if (j >= cnt) return;
        while (true) {
            tile_eval(a, v, i, tile_locate(v, q.x, q.y, margin, a.tol));
            j += blockDim.x;
            if (j >= cnt) break;
            q = __ldcs(t.rec_xy + b0 + j);
            i = __ldcs(t.rec_id + b0 + j);
        }
    }
}

// Points that did not fit their leaf's bucket (a target mesh locally much denser than M0): located and evaluated
// from the leaf arrays in HBM, one thread per point.  Also leaves the counters zeroed for the next call.
__global__ void __launch_bounds__(256) k_tile_overflow(const TileArgs t)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const int novf = __ldcg(t.ovf_count);
    const InterpArgs &a = t.a;
    const double margin = a.tol * (1.0 + 1e-6) + 1e-9;
    for (int m = gtid; m < novf; m += gsz) {
        const int i = t.ovf_list[m];
        double px, py;
        target_point(a, i, px, py);
        const int L = tile_leaf_of(t.tree, t.depth, px, py);
        const int *info = t.leaf_info + 16 * L;
        LeafView v;
        v.xy = t.coordsL + info[0]; v.U = t.UL + info[2]; v.P = t.PL + info[0]; v.gid = t.gidL + info[4];
        v.cv = reinterpret_cast<const unsigned *>(t.cvL + 6 * (size_t)info[4]);
        v.bp = t.binptrL + info[6]; v.bins = t.binsL + info[8];
        v.u_rows = (int)t.u_stride; v.p_rows = (int)t.p_stride;
        v.x0 = t.leaf_rect[4 * L]; v.y0 = t.leaf_rect[4 * L + 1]; v.inv_dx = t.leaf_rect[4 * L + 2];
        v.inv_dy = t.leaf_rect[4 * L + 3]; v.gx = info[10]; v.gy = info[11];
        tile_eval(a, v, i, tile_locate(v, px, py, margin, a.tol));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(t.ticket, 1u) == gridDim.x - 1) {   // every CTA has read ovf_count
            *t.ovf_count = 0;
            *t.ticket = 0u;
        }
    }
}
