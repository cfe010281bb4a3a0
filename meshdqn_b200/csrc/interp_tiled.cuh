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
//                     re-zeroes the counters for the next call;
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
    int max_nv, max_np2, max_nc, max_nbin, max_nent;
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
    size_t u_rows, p_rows;          // rows between snapshots
    double x0, y0, inv_dx, inv_dy;
    int gx, gy;
};

__device__ __forceinline__ void tile_locate_eval(const InterpArgs &a, const LeafView &v, int i, double px, double py,
                                                 double margin)
{
    const int bx = min(max((int)floor((px - v.x0) * v.inv_dx), 0), v.gx - 1);
    const int by = min(max((int)floor((py - v.y0) * v.inv_dy), 0), v.gy - 1);
    const int b = by * v.gx + bx;
    const int s1 = v.bp[b + 1];
    int s = v.bp[b];
    int lc = -1;
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
    unsigned w0 = 0, w1 = 0;
    while (true) {
        // phase A (division-free): advance to the next candidate the conservative test cannot reject
        double n1 = 0.0, n2 = 0.0, det = 1.0;
        int c = -1;
        while (s < s1) {
            const int cc = v.bins[s++];
            w0 = v.cv[3 * cc];
            w1 = v.cv[3 * cc + 1];
            const double2 A = v.xy[w0 & 0xffffu], B = v.xy[w0 >> 16], C = v.xy[w1 & 0xffffu];
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
        l1 = n1 == 0.0 ? n1 * det : n1 / det;
        l2 = n2 == 0.0 ? n2 * det : n2 / det;
        l0 = 1.0 - l1 - l2;
        if (fmin(l0, fmin(l1, l2)) >= -a.tol) {
            lc = c;
            break;
        }
    }
    if (lc < 0) {
        a.cell_of[i] = -1;
        a.miss_list[atomicAdd(a.miss_count, 1)] = i;
        return;
    }
    a.cell_of[i] = v.gid[lc];
    const unsigned w2 = v.cv[3 * lc + 2];
    double phi[6];
    phi[0] = l0 * (2.0 * l0 - 1.0);
    phi[1] = l1 * (2.0 * l1 - 1.0);
    phi[2] = l2 * (2.0 * l2 - 1.0);
    phi[3] = 4.0 * l1 * l2;
    phi[4] = 4.0 * l0 * l2;
    phi[5] = 4.0 * l0 * l1;
    const int dof[6] = {(int)(w0 & 0xffffu), (int)(w0 >> 16), (int)(w1 & 0xffffu),
                        (int)(w1 >> 16),     (int)(w2 & 0xffffu), (int)(w2 >> 16)};
    const double lam[3] = {l0, l1, l2};
    const int np2t = a.nv + a.ne;
    for (int k = 0; k < a.T; ++k) {
        const double2 *Uk = v.U + (size_t)k * v.u_rows;
        const double2 u0 = Uk[dof[0]];
        double ux = phi[0] * u0.x, uy = phi[0] * u0.y;   // == 0.0 + phi*u of the reference loop
#pragma unroll
        for (int q = 1; q < 6; ++q) {
            const double2 u = Uk[dof[q]];
            ux += phi[q] * u.x;
            uy += phi[q] * u.y;
        }
        __stcs(reinterpret_cast<double2 *>(a.U) + (size_t)k * np2t + i, make_double2(ux, uy));
        if (i < a.nv) {
            const double *Pk = v.P + (size_t)k * v.p_rows;
            double pv = lam[0] * Pk[dof[0]];
#pragma unroll
            for (int q = 1; q < 3; ++q) pv += lam[q] * Pk[dof[q]];
            __stcs(a.P + (size_t)k * a.nv + i, pv);
        }
    }
}

__global__ void __launch_bounds__(256) k_tile_classify(const TileArgs t)
{
    const int np = t.a.nv + t.a.ne;
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *t.a.miss_count = 0;
    for (int i0 = blockIdx.x * blockDim.x; i0 < np; i0 += gridDim.x * blockDim.x) {
        const int i = i0 + threadIdx.x;
        int leaf = -1 - lane;  // distinct dummy keys for the tail lanes
        double px = 0.0, py = 0.0;
        if (i < np) {
            target_point(t.a, i, px, py);
            leaf = tile_leaf_of(t.tree, t.depth, px, py);
        }
        // warp-aggregated slot allocation: one atomic per distinct leaf in the warp (spatially ordered targets
        // put most of a warp in one leaf)
        const unsigned peers = __match_any_sync(FULL, leaf);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader && leaf >= 0) base = atomicAdd(t.leaf_cnt + leaf, __popc(peers));
        base = __shfl_sync(FULL, base, leader);
        if (i < np) {
            const int slot = base + __popc(peers & ((1u << lane) - 1u));
            const int b0 = __ldg(t.leaf_base + leaf), b1 = __ldg(t.leaf_base + leaf + 1);
            if (slot < b1 - b0) {
                t.rec_xy[b0 + slot] = make_double2(px, py);
                t.rec_id[b0 + slot] = i;
            } else {
                t.ovf_list[atomicAdd(t.ovf_count, 1)] = i;
            }
        }
    }
}

constexpr int TILE_THREADS = 512;

// shared-memory carve-up (bytes), identical on host and device
struct TileSmem {
    int off_coords, off_U, off_P, off_gid, off_cv, off_binptr, off_bins, total;
};
__host__ __device__ inline TileSmem tile_smem(int max_nv, int max_np2, int max_nc, int max_nbin, int max_nent, int T)
{
    TileSmem s;
    int o = 16;  // mbarrier
    s.off_coords = o; o += 16 * max_nv;
    s.off_U = o;      o += 16 * max_np2 * T;
    s.off_P = o;      o += 8 * max_nv * T;      // max_nv is even -> 16-byte multiples
    s.off_gid = o;    o += 4 * max_nc;          // max_nc multiple of 4
    s.off_cv = o;     o += 12 * max_nc;
    s.off_binptr = o; o += 2 * max_nbin;        // multiples of 8 entries
    s.off_bins = o;   o += 2 * max_nent;
    s.total = o;
    return s;
}

__global__ void __launch_bounds__(TILE_THREADS, 2) k_tile_interp(const TileArgs t)
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
    const InterpArgs &a = t.a;
    const int T = a.T;
    const TileSmem S = tile_smem(t.max_nv, t.max_np2, t.max_nc, t.max_nbin, t.max_nent, T);
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem);
    double2 *s_xy = reinterpret_cast<double2 *>(smem + S.off_coords);
    double2 *s_U = reinterpret_cast<double2 *>(smem + S.off_U);
    double *s_P = reinterpret_cast<double *>(smem + S.off_P);
    int *s_gid = reinterpret_cast<int *>(smem + S.off_gid);
    unsigned *s_cv = reinterpret_cast<unsigned *>(smem + S.off_cv);
    unsigned short *s_bp = reinterpret_cast<unsigned short *>(smem + S.off_binptr);
    unsigned short *s_bins = reinterpret_cast<unsigned short *>(smem + S.off_bins);
    const int nvl = i0.y, np2l = i0.w;
    if (threadIdx.x == 0) {
        const int vbase = i0.x, dbase = i0.z, cbase = i1.x, ncl = i1.y, bbase = i1.z, nbin = i1.w;
        const int ebase = i2.x, nent = i2.y;
        mbar_init(bar, 1);
        const unsigned bytes = 16u * nvl + (unsigned)T * (16u * np2l + 8u * nvl) + 16u * ncl + 2u * nbin + 2u * nent;
        mbar_expect_tx(bar, bytes);
        bulk_g2s(s_bp, t.binptrL + bbase, 2u * nbin, bar);
        bulk_g2s(s_bins, t.binsL + ebase, 2u * nent, bar);
        bulk_g2s(s_cv, t.cvL + 6 * (size_t)cbase, 12u * ncl, bar);
        bulk_g2s(s_xy, t.coordsL + vbase, 16u * nvl, bar);
        bulk_g2s(s_gid, t.gidL + cbase, 4u * ncl, bar);
        for (int k = 0; k < T; ++k) {
            bulk_g2s(s_U + (size_t)k * np2l, t.UL + (size_t)k * t.u_stride + dbase, 16u * np2l, bar);
            bulk_g2s(s_P + (size_t)k * nvl, t.PL + (size_t)k * t.p_stride + vbase, 8u * nvl, bar);
        }
    }
    LeafView v;
    v.xy = s_xy; v.U = s_U; v.P = s_P; v.gid = s_gid; v.cv = s_cv; v.bp = s_bp; v.bins = s_bins;
    v.u_rows = np2l; v.p_rows = nvl;
    v.x0 = r0.x; v.y0 = r0.y; v.inv_dx = r1.x; v.inv_dy = r1.y; v.gx = i2.z; v.gy = i2.w;
    const double margin = a.tol * (1.0 + 1e-6) + 1e-9;
    // this thread's first record travels while the leaf is being staged
    int j = (int)threadIdx.x;
    double2 q = make_double2(0.0, 0.0);
    int i = 0;
    if (j < cnt) {
        q = __ldcs(t.rec_xy + b0 + j);
        i = __ldcs(t.rec_id + b0 + j);
    }
    __syncthreads();  // the barrier is initialised before anyone polls it
    if (j >= cnt) return;
    mbar_wait(bar, 0);
    while (true) {
        tile_locate_eval(a, v, i, q.x, q.y, margin);
        j += blockDim.x;
        if (j >= cnt) break;
        q = __ldcs(t.rec_xy + b0 + j);
        i = __ldcs(t.rec_id + b0 + j);
    }
}

// Points that did not fit their leaf's bucket (a target mesh locally much denser than M0): located and evaluated
// from the leaf arrays in HBM, one thread per point.  Also leaves the counters zeroed for the next call.
__global__ void __launch_bounds__(256) k_tile_overflow(const TileArgs t)
{
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
    const int novf = __ldcg(t.ovf_count);
    for (int l = gtid; l < t.n_leaves; l += gsz) t.leaf_cnt[l] = 0;
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
        v.u_rows = (size_t)t.u_stride; v.p_rows = (size_t)t.p_stride;
        v.x0 = t.leaf_rect[4 * L]; v.y0 = t.leaf_rect[4 * L + 1]; v.inv_dx = t.leaf_rect[4 * L + 2];
        v.inv_dy = t.leaf_rect[4 * L + 3]; v.gx = info[10]; v.gy = info[11];
        tile_locate_eval(a, v, i, px, py, margin);
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
