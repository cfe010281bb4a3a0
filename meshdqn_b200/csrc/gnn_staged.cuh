// gnn_staged.cuh -- the replay / batched Q-path as STAGES on the 5th-gen tensor cores (tcgen05, TMEM, bulk TMA).
// Included by gnn_fused.cu inside its anonymous namespace (it reuses WDesc, the weight-gradient kernels and the loss).
//
// The fused kernel (qnet_kernel) walks one graph per CTA through ~40 barrier-separated phases and re-streams the
// 470 KB of weights from L2 for every graph.  Here the network of /root/reference/airfoilgcnn.py:85-145
// (SAGE, SAGE, GCN, GCN blocks with TopK pooling + 3-layer MLP) is cut where the row count collapses:
//
//   k_stage0  one CTA per graph: CSR + mean aggregation in shared memory, conv1 as a tcgen05 GEMM
//             (rows = nodes, [agg|x] split hi/lo in the canonical K-major layout, W1 resident, accumulator in TMEM),
//             TopK score in the TMEM epilogue (hidden rows never leave TMEM), rank, kept rows -> level 1.
//   k_stage1  GS graphs per CTA: conv2 as a "swapped" GEMM  D^T[128 out][rows] = W2[128][256] . cat2[rows][256]^T:
//             the weights are the M = 128 operand streamed by bulk TMA through a ring, the few activation rows are the
//             N operand, so the tensor-pipe time follows the row count instead of a 128-row tile.
//   k_stage2  GS2 graphs per CTA: conv4, conv5, lin1..3 (+ softmax/argmax) as a chain of swapped GEMMs on ONE weight
//             stream; in backward mode the same kernel continues with the Huber term and the transposed GEMMs down to
//             dX2, so the selected net's forward is never launched separately.
//   k_bwd1    conv2^T GEMM, scatter through the level-1 edges, pooling backward of levels 1 and 0.
//
// fp32 operands are split a = a_hi + a_lo (a_hi = a & 0xffffe000, exact in TF32) and every product is accumulated as
// a_hi.w_lo + a_lo.w_hi + a_hi.w_hi in fp32 inside TMEM (3xTF32): ~2^-21 relative per product -- BASELINE.json asks for
// 1e-5 on the Q-values, which plain TF32 (1e-3) cannot give.  The weights' hi / lo halves are pre-tiled once per
// weight update by k_wsplit into [M-tile][K-chunk][16 groups][8 rows][4 k] (K-major core matrices), both orientations.
#pragma once

namespace stg {

using namespace tcp;

constexpr int NTH = 256;
constexpr int KB = 32;                    // K elements per streamed weight block
constexpr int NCH = KB / 4;               // 16-byte chunks per row per block
constexpr int W_LBO = 2048;               // chunk stride of a 128-row weight tile (16 groups x 128 B)
constexpr int SBO = 128;                  // 8-row group stride
constexpr int W_HALF = NCH * W_LBO;       // bytes of one (hi or lo) weight block
constexpr int MAXBLK = 8;
constexpr int LDW = 132, LD2W = 260;      // padded row strides (floats) of 128- / 256-wide activation rows (= 4 mod 32)
constexpr int LDY2 = 68, LDY3 = 196;

enum { M_C1 = 0, M_C2F, M_COUNT };   // the tail layers and the backward GEMMs stream the flat fp32 weights (gnn_tail.cuh)

struct WMat {
    int off;          // float offset in wsplit: [M tile][K block of 32][hi (nch chunks) | lo (nch chunks)][16 groups][8 rows][4 k]
    int mpad, kpad;   // padded to 128 / 8
    int src, ld;      // flat-buffer offset and row length (out features) of the source matrix
    int M, K;         // valid extents in this orientation
    int mode;         // 0: forward (M = out, K = in)  1: transposed (M = in, K = out)  2: conv1 [agg(Fp) | x(Fp)] columns
};
enum { T_C2 = 0, T_C4, T_C5, T_L1, T_L2, T_L3, T_COUNT };
struct TMat { int off, src, K, C; };   // wsplit[off + c * K + k] = params[src + k * C + c]   (K = in, C = out; K % 4 == 0)
struct Plan {
    WMat m[M_COUNT];
    int item0[M_COUNT + 1];   // float4 items of k_wsplit (tiles)
    TMat t[T_COUNT];          // transposed fp32 copies for the backward layers (gnn_tail.cuh)
    int titem0[T_COUNT + 1];
    int total;                // floats
    int F, Fp, k1pad;
};
struct WBlk { unsigned off, half; };   // float offset of the block (hi then lo, contiguous), bytes of one half

__host__ __device__ inline int rup(int v, int m) { return (v + m - 1) / m * m; }

// Is this network / batch shape served by the staged path?  (everything else stays on the fused kernel)
inline bool supported(const mdq_net_t &net, int max_n, int max_e)
{
    if (net.n_blocks != 4 || net.width != 128) return false;
    if (net.blk[0].type != MDQ_BLOCK_SAGE || net.blk[1].type != MDQ_BLOCK_SAGE || net.blk[2].type != MDQ_BLOCK_GCN ||
        net.blk[3].type != MDQ_BLOCK_GCN)
        return false;
    if (net.blk[0].kin != net.in_dim || net.in_dim < 1 || net.in_dim > 24) return false;
    for (int b = 1; b < 4; ++b)
        if (net.blk[b].kin != 128) return false;
    if (net.lin_in[0] != 256 || net.lin_out[0] != 128 || net.lin_in[1] != 128 || net.lin_out[1] != 64 || net.lin_in[2] != 64 ||
        net.lin_out[2] != net.out_dim || net.out_dim < 1 || net.out_dim > 184)
        return false;
    if (max_n < 1 || max_n > 256 || max_e < 0 || max_e > 4096) return false;
    const int R1 = topk_count(net.ratio, max_n), R2 = topk_count(net.ratio, R1), R3 = topk_count(net.ratio, R2);
    if (R1 < 1 || R1 > 32 || R2 < 1 || R2 > 4 || R3 != 1 || topk_count(net.ratio, 1) != 1) return false;
    return true;
}

inline void build_plan(const mdq_net_t &net, Plan &P)
{
    memset(&P, 0, sizeof(P));
    P.F = net.in_dim;
    P.Fp = rup(P.F, 4);
    P.k1pad = rup(2 * P.Fp, 8);
    int off = 0, items = 0;
    auto add = [&](int id, int src, int ld, int M, int K, int mode) {
        WMat &m = P.m[id];
        m.src = src; m.ld = ld; m.M = M; m.K = K; m.mode = mode;
        m.mpad = rup(M, 128);
        m.kpad = mode == 2 ? P.k1pad : rup(K, 8);
        m.off = off;
        off += 2 * m.mpad * m.kpad;
        P.item0[id] = items;
        items += m.mpad * m.kpad / 4;
    };
    add(M_C1, net.blk[0].w_off, 128, 128, 2 * P.F, 2);
    add(M_C2F, net.blk[1].w_off, 128, 128, 256, 0);
    P.item0[M_COUNT] = items;
    int titems = 0;
    auto addt = [&](int id, int src, int K, int C) {
        P.t[id].off = off; P.t[id].src = src; P.t[id].K = K; P.t[id].C = C;
        off += rup(K * C, 4);
        P.titem0[id] = titems;
        titems += K * C / 4;
    };
    addt(T_C2, net.blk[1].w_off, 256, 128);
    addt(T_C4, net.blk[2].w_off, 128, 128);
    addt(T_C5, net.blk[3].w_off, 128, 128);
    addt(T_L1, net.lin_off[0], 256, 128);
    addt(T_L2, net.lin_off[1], 128, 64);
    addt(T_L3, net.lin_off[2], 64, net.out_dim);
    P.titem0[T_COUNT] = titems;
    P.total = off;
}

// weight blocks of M-tile `tile` of matrix m, in K order
inline int add_blocks(const WMat &m, int tile, WBlk *out, int n)
{
    const int nkb = (m.kpad + KB - 1) / KB;
    for (int kb = 0; kb < nkb; ++kb) {
        const int left = m.kpad / 4 - NCH * kb;
        const int nch = left < NCH ? left : NCH;
        out[n].off = (unsigned)(m.off + tile * (2 * m.kpad * 128) + kb * (2 * NCH * 512));
        out[n].half = (unsigned)(nch * W_LBO);
        ++n;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// k_wsplit: flat fp32 parameters -> hi / lo TF32 halves in the tiled K-major layout, every matrix of the plan
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wsplit(const Plan P, const float *__restrict__ params, float *__restrict__ ws)
{
    const int total = P.item0[M_COUNT];
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
        int id = 0;
        while (id + 1 < M_COUNT && it >= P.item0[id + 1]) ++id;
        const WMat m = P.m[id];
        int rem = it - P.item0[id];
        const int per_tile = m.kpad * 32;
        const int t = rem / per_tile;
        rem -= t * per_tile;
        const int c = rem >> 7, gq = (rem >> 3) & 15, r = rem & 7;
        const int Mi = 128 * t + 8 * gq + r;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int Ki = 4 * c + q;
            float x = 0.f;
            if (Mi < m.M) {
                if (m.mode == 0) {
                    if (Ki < m.K) x = __ldg(params + m.src + (size_t)Ki * m.ld + Mi);
                } else if (m.mode == 1) {
                    if (Ki < m.K) x = __ldg(params + m.src + (size_t)Mi * m.ld + Ki);
                } else {
                    int row = -1;
                    if (Ki < P.F) row = Ki;
                    else if (Ki >= P.Fp && Ki < P.Fp + P.F) row = P.F + Ki - P.Fp;
                    if (row >= 0) x = __ldg(params + m.src + (size_t)row * m.ld + Mi);
                }
            }
            v[q] = x;
        }
        float4 h, l;
        split_tf32(make_float4(v[0], v[1], v[2], v[3]), h, l);
        const int kb = c >> 3, cb = c & 7;
        const int left = m.kpad / 4 - NCH * kb;
        const int nchb = left < NCH ? left : NCH;
        float4 *dst = reinterpret_cast<float4 *>(ws + m.off) + (size_t)t * (2 * per_tile) + kb * (2 * NCH * 128) + cb * 128 + (rem & 127);
        dst[0] = h;
        dst[nchb * 128] = l;
    }
    const int ttotal = P.titem0[T_COUNT];
    for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < ttotal; it += gridDim.x * blockDim.x) {
        int id = 0;
        while (id + 1 < T_COUNT && it >= P.titem0[id + 1]) ++id;
        const TMat m = P.t[id];
        const int rem = it - P.titem0[id];
        const int k4 = rem % (m.K / 4), c = rem / (m.K / 4);
        float4 v;
        v.x = __ldg(params + m.src + (size_t)(4 * k4) * m.C + c);
        v.y = __ldg(params + m.src + (size_t)(4 * k4 + 1) * m.C + c);
        v.z = __ldg(params + m.src + (size_t)(4 * k4 + 2) * m.C + c);
        v.w = __ldg(params + m.src + (size_t)(4 * k4 + 3) * m.C + c);
        *reinterpret_cast<float4 *>(ws + m.off + (size_t)c * m.K + 4 * k4) = v;
    }
}

// profiling aid (mdq_qnet_set_trace): CTA 0 / thread 0 of each staged kernel stamps clock64() at its phase boundaries
#define STG_TRACE(buf, base, i)                                                         \
    do {                                                                                \
        if ((buf) && blockIdx.x == 0 && threadIdx.x == 0) (buf)[(base) + (i)] = clock64(); \
    } while (0)

// wall-clock (%globaltimer, ns) stamps of CTA 0 at slot 400 + i: a cross-kernel, cross-stream timeline of one step
#define STG_GT(buf, i)                                                                                     \
    do {                                                                                                   \
        if ((buf) && blockIdx.x == 0 && threadIdx.x == 0) {                                                \
            long long t_;                                                                                  \
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_));                                          \
            (buf)[400 + (i)] = t_;                                                                         \
        }                                                                                                  \
    } while (0)

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x))); }

// ------------------------------------------------------------------------------------------------
// k_stage0: block 0 of one graph per CTA (two CTAs per SM)
// ------------------------------------------------------------------------------------------------
struct S0Args {
    long long *trace;
    const float *params, *wsplit;
    unsigned w_off;                // float offset of conv1's tiles (K blocks of [hi | lo])
    unsigned pf_floats, n_params;  // L2 prefetch: the whole wsplit buffer (for the later stages) and the flat parameters
    int F, Fp, nch;                // nch = k1pad / 4 chunks
    int x_stride, col0, b_off, pool_off;
    float ratio;
    const float *x;
    const long long *esrc, *edst;  // int64 edge_index rows, or int32 rows when edge_i32 (half the bytes over PCIe)
    int edge_i32;
    const int *nptr, *eptr;
    int B, R1, EC1;
    float *x1;                     // [B][R1][128] kept rows x score
    unsigned short *e1;            // [B][EC1] level-1 edges (src | dst << 8), edge order kept
    int *e1n;                      // [B]
    float *r0;                     // [B][256] readout (max | mean)
    float *h1k, *c1k, *s1k, *z1k;  // backward saves: hidden rows [B][R1][128], inputs [B*R1][2F], score / z [B][R1]
    unsigned char *amax1;          // [B][128]
    int o_w, o_a, o_xs, o_key, o_es, o_ed, o_rowptr, o_cursor, o_score, o_z, o_newid, o_perm, o_bp, o_k64, total;
};

inline int s0_layout(S0Args &a, int max_n, int max_e, int R1)
{
    int o = 128;
    auto take = [&](int bytes) { int at = o; o += rup(bytes, 16); return at; };
    a.o_w = take(2 * a.nch * W_LBO);       // W1 hi | lo  (later: the kept hidden rows hk [R1][LDW])
    a.o_a = take(2 * a.nch * W_LBO);       // A tile hi | lo
    a.o_xs = take(max_n * a.F * 4);
    a.o_key = take((max_e > 0 ? max_e : 1) * 4);
    a.o_es = take(max_e > 0 ? max_e : 1);
    a.o_ed = take(max_e > 0 ? max_e : 1);
    a.o_rowptr = take((max_n + 1) * 4);
    a.o_cursor = take((max_n + 1) * 4);
    a.o_score = take(max_n * 4);
    a.o_z = take(max_n * 4);
    a.o_newid = take(max_n * 2);
    a.o_perm = take(R1 * 4);
    a.o_bp = take(256 * 4);                // conv1 bias | pool1 weight
    a.o_k64 = take((max_n + 2) * 8);       // (score, index) sort keys
    a.total = o;
    if (R1 * LDW * 4 > 2 * a.nch * W_LBO) return -1;
    return o;
}

template <bool SAVE>
__global__ void __launch_bounds__(NTH, 2) k_stage0(const __grid_constant__ S0Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.x;
    unsigned long long *wbar = reinterpret_cast<unsigned long long *>(sm), *accb = wbar + 1;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(sm + 32);
    const int nch = a.nch, F = a.F;
    unsigned char *w_hi = sm + a.o_w;
    unsigned char *a_hi = sm + a.o_a, *a_lo = a_hi + nch * W_LBO;
    float *xs = reinterpret_cast<float *>(sm + a.o_xs);
    unsigned *key = reinterpret_cast<unsigned *>(sm + a.o_key);
    unsigned char *es = sm + a.o_es, *ed = sm + a.o_ed;
    int *rowptr = reinterpret_cast<int *>(sm + a.o_rowptr), *cursor = reinterpret_cast<int *>(sm + a.o_cursor);
    float *score = reinterpret_cast<float *>(sm + a.o_score), *zs = reinterpret_cast<float *>(sm + a.o_z);
    short *newid = reinterpret_cast<short *>(sm + a.o_newid);
    int *perm = reinterpret_cast<int *>(sm + a.o_perm);
    float *hk = reinterpret_cast<float *>(w_hi);   // valid once the MMAs have consumed W1
    float *bias = reinterpret_cast<float *>(sm + a.o_bp), *pw = bias + 128;
    unsigned long long *k64 = reinterpret_cast<unsigned long long *>(sm + a.o_k64);

    STG_TRACE(a.trace, 0, 0);
    STG_GT(a.trace, SAVE ? 2 : 0);
    const int nb0 = a.nptr[g], n = a.nptr[g + 1] - nb0;
    const int eb0 = a.eptr[g], E = a.eptr[g + 1] - eb0;
    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(accb, 1);
        mbar_fence_init();
        mbar_expect_tx(wbar, 2u * nch * W_LBO);
        bulk_g2s(w_hi, a.wsplit + a.w_off, 2u * nch * W_LBO, wbar);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    for (int i = tid; i <= n; i += NTH) cursor[i] = 0;
    if (gridDim.x >= 64) {   // L2 prefetch for the later stages: 4 KB pieces of wsplit / params, spread over the grid
        const unsigned pieces_w = (a.pf_floats + 1023u) >> 10, pieces_p = (a.n_params + 1023u) >> 10;
        for (unsigned pc = blockIdx.x + gridDim.x * tid; pc < pieces_w + pieces_p; pc += gridDim.x * NTH) {
            const bool isw = pc < pieces_w;
            const unsigned f0 = (isw ? pc : pc - pieces_w) << 10;
            const unsigned lim = isw ? a.pf_floats : (a.n_params & ~3u);
            const unsigned nf = lim - f0 < 1024u ? lim - f0 : 1024u;
            if (f0 < lim && nf >= 4u) bulk_prefetch_l2((isw ? a.wsplit : a.params) + f0, (nf & ~3u) * 4u);
        }
    }
    __syncthreads();   // cursor zeroed: the edge loop below counts into it while the x rows are still in flight
    if (tid < 128) { bias[tid] = __ldg(a.params + a.b_off + tid); pw[tid] = __ldg(a.params + a.pool_off + tid); }
    for (int e = tid; e < E; e += NTH) {
        int s, d;
        if (a.edge_i32) {
            s = reinterpret_cast<const int *>(a.esrc)[eb0 + e] - nb0;
            d = reinterpret_cast<const int *>(a.edst)[eb0 + e] - nb0;
        } else {
            s = (int)(a.esrc[eb0 + e] - nb0);
            d = (int)(a.edst[eb0 + e] - nb0);
        }
        es[e] = (unsigned char)s;
        ed[e] = (unsigned char)d;
        atomicAdd(&cursor[d], 1);
    }
    {
        const float *xg = a.x + (size_t)nb0 * a.x_stride + a.col0;
        if (a.x_stride == F) {
            for (int idx = tid; idx < n * F; idx += NTH) xs[idx] = __ldg(xg + idx);
        } else {
            for (int idx = tid; idx < n * F; idx += NTH) {
                const int i = idx / F, f = idx - i * F;
                xs[idx] = __ldg(xg + (size_t)i * a.x_stride + f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;
    STG_TRACE(a.trace, 0, 1);   // x and edges loaded, edges counted, TMEM allocated
    // ---- CSR by destination, rows in edge order (torch_scatter's CPU loop order) ----
    STG_TRACE(a.trace, 0, 2);
    if (warp == 0) {
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const int v = (i < n) ? cursor[i] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (i < n) { rowptr[i] = carry + incl - v; cursor[i] = carry + incl - v; }
            carry += __shfl_sync(FULL, incl, 31);
        }
        if (lane == 0) rowptr[n] = carry;
    }
    __syncthreads();
    STG_TRACE(a.trace, 0, 3);   // scan
    for (int e = tid; e < E; e += NTH) {
        const int slot = atomicAdd(&cursor[ed[e]], 1);
        key[slot] = ((unsigned)e << 8) | es[e];
    }
    __syncthreads();
    STG_TRACE(a.trace, 0, 4);   // fill
    for (int i = tid; i < n; i += NTH) {   // rows are short: insertion sort by edge id restores the edge order
        const int s0 = rowptr[i], s1 = rowptr[i + 1];
        for (int p = s0 + 1; p < s1; ++p) {
            const unsigned kv = key[p];
            int q = p;
            while (q > s0 && key[q - 1] > kv) { key[q] = key[q - 1]; --q; }
            key[q] = kv;
        }
    }
    __syncthreads();
    STG_TRACE(a.trace, 0, 5);   // rows sorted

    // [mean over in-edges (Fp cols) | x (Fp cols) | 0...]: thread = row.  The row's mean stays in registers, so the rows of
    // the second tile (>= 128) are written from them once the first tile's MMAs have released the A buffer.
    constexpr int FMAX = 24;
    float mv[FMAX];
#pragma unroll
    for (int f = 0; f < FMAX; ++f) mv[f] = 0.f;
    if (tid < n) {
        const int s0 = rowptr[tid], s1 = rowptr[tid + 1];
        for (int sx = s0; sx < s1; ++sx) {
            const float *xr = xs + (key[sx] & 255u) * F;
#pragma unroll
            for (int f = 0; f < FMAX; ++f)
                if (f < F) mv[f] += xr[f];
        }
        const float inv = __frcp_rn((float)(s1 - s0 > 0 ? s1 - s0 : 1));   // mean = sum * (1 / deg): within 1 ulp of sum / deg
#pragma unroll
        for (int f = 0; f < FMAX; ++f) mv[f] = (f < F) ? mv[f] * inv : 0.f;
    }
    STG_TRACE(a.trace, 0, 14);  // (profiling) means in registers
    auto write_row = [&](int r) {   // this thread's row -> row r of the A tile, hi / lo, canonical layout
        const float *xi = xs + tid * F;
        const int cm = a.Fp >> 2;
        unsigned char *dh = a_hi + (r >> 3) * SBO + (r & 7) * 16, *dl = a_lo + (r >> 3) * SBO + (r & 7) * 16;
#pragma unroll
        for (int c = 0; c < FMAX / 4; ++c)
            if (c < cm) {
                float4 h, l;
                split_tf32(make_float4(mv[4 * c], mv[4 * c + 1], mv[4 * c + 2], mv[4 * c + 3]), h, l);
                *reinterpret_cast<float4 *>(dh + c * W_LBO) = h;
                *reinterpret_cast<float4 *>(dl + c * W_LBO) = l;
            }
        for (int c = cm; c < nch; ++c) {
            const int f0 = 4 * (c - cm);
            float4 v;
            v.x = f0 < F ? xi[f0] : 0.f;
            v.y = f0 + 1 < F ? xi[f0 + 1] : 0.f;
            v.z = f0 + 2 < F ? xi[f0 + 2] : 0.f;
            v.w = f0 + 3 < F ? xi[f0 + 3] : 0.f;
            float4 h, l;
            split_tf32(v, h, l);
            *reinterpret_cast<float4 *>(dh + c * W_LBO) = h;
            *reinterpret_cast<float4 *>(dl + c * W_LBO) = l;
        }
    };
    const unsigned idesc = idesc_tf32_m128(128);
    auto issue = [&](int tile) {
        tc_fence_after();
        const unsigned ah = s32(a_hi), al = s32(a_lo), wb = s32(w_hi);
        for (int ks = 0; ks < nch / 2; ++ks) {
            const unsigned long long dah = smem_desc(ah + 2 * ks * W_LBO, W_LBO, SBO), dal = smem_desc(al + 2 * ks * W_LBO, W_LBO, SBO);
            // W1 arrives as K blocks of 8 chunks, each [hi | lo]
            const int kb = ks >> 2, left = nch - NCH * kb, nchb = left < NCH ? left : NCH;
            const unsigned bh = wb + (unsigned)(kb * 2 * NCH + (ks & 3) * 2) * W_LBO, bl = bh + (unsigned)nchb * W_LBO;
            const unsigned long long dbh = smem_desc(bh, W_LBO, SBO), dbl = smem_desc(bl, W_LBO, SBO);
            mma_tf32(tmem + tile * 128, dah, dbl, idesc, ks ? 1u : 0u);
            mma_tf32(tmem + tile * 128, dal, dbh, idesc, 1u);
            mma_tf32(tmem + tile * 128, dah, dbh, idesc, 1u);
        }
        mma_commit(accb);
    };
    if (tid < n && tid < 128) write_row(tid);
    fence_async_smem();
    __syncthreads();
    STG_TRACE(a.trace, 0, 6);   // A tile 0 built
    if (tid == 0) {
        mbar_wait(wbar, 0);
        STG_TRACE(a.trace, 0, 7);   // W1 landed
        issue(0);
    }
    mbar_wait(accb, 0);
    STG_TRACE(a.trace, 0, 8);   // tile 0 MMAs complete
    if (n > 128) {   // second row tile through the same A buffer (two CTAs share an SM: no room for both tiles)
        if (tid >= 128 && tid < n) write_row(tid - 128);
        fence_async_smem();
        __syncthreads();
        if (tid == 0) issue(1);
        mbar_wait(accb, 1);
    }
    tc_fence_after();
    STG_TRACE(a.trace, 0, 9);   // tile 1 complete

    // ---- epilogue 1: TopK score of every row straight from TMEM (thread = row) ----
    float pn;
    {
        const float4 w = *reinterpret_cast<const float4 *>(pw + 4 * lane);
        pn = sqrtf(warp_sum(dot4(w, w)));
    }
    const int row = tid;
    const unsigned taddr = tmem + (((unsigned)(warp & 3) * 32u) << 16) + (unsigned)(warp >> 2) * 128u;
    const bool warp_active = warp * 32 < n;
    if (warp_active) {
        float dot = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float v[32];
            tmem_ld32(taddr + c0, v);
#pragma unroll
            for (int c = 0; c < 32; c += 4) {   // broadcast 16-byte shared-memory reads of bias / pool weight
                const float4 b4 = *reinterpret_cast<const float4 *>(bias + c0 + c), w4 = *reinterpret_cast<const float4 *>(pw + c0 + c);
                dot = fmaf(fmaxf(v[c] + b4.x, 0.f), w4.x, dot);
                dot = fmaf(fmaxf(v[c + 1] + b4.y, 0.f), w4.y, dot);
                dot = fmaf(fmaxf(v[c + 2] + b4.z, 0.f), w4.z, dot);
                dot = fmaf(fmaxf(v[c + 3] + b4.w, 0.f), w4.w, dot);
            }
        }
        if (row < n) {
            const float z = dot / pn;
            const float sc = tanhf(z) + 0.f;   // -0 -> +0: the integer key below must order like the float
            zs[row] = z;
            score[row] = sc;
            // larger key <=> ahead in (score desc, index asc) order
            const unsigned b = __float_as_uint(sc);
            k64[row] = ((unsigned long long)(b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u)) << 32) | (0xffffffffu - (unsigned)row);
        }
    }
    if (tid >= n && tid < n + 2) k64[tid] = 0ull;   // pad for the 16-byte reads of the ranking loop
    __syncthreads();
    STG_TRACE(a.trace, 0, 10);  // scores
    const int k = topk_count_dev(a.ratio, n);
    if (row < n) {   // rank = number of larger keys; broadcast 16-byte reads, one 64-bit compare per row
        const unsigned long long ki = k64[row];
        int r = 0;
        for (int j = 0; j < n; j += 2) {
            const ulonglong2 kk = *reinterpret_cast<const ulonglong2 *>(k64 + j);
            r += (kk.x > ki) + (kk.y > ki);
        }
        newid[row] = (short)(r < k ? r : -1);
        if (r < k) perm[r] = row;
    }
    __syncthreads();
    STG_TRACE(a.trace, 0, 11);  // ranked
    // ---- epilogue 2: the kept rows leave TMEM; the level-1 edge list ----
    if (warp_active) {
        const int myr = row < n ? (int)newid[row] : -1;
        if (__any_sync(FULL, myr >= 0)) {
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
                float v[32];
                tmem_ld32(taddr + c0, v);
                if (myr >= 0) {
                    float4 *dst = reinterpret_cast<float4 *>(hk + myr * LDW + c0);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(bias + c0 + 4 * c);
                        dst[c] = make_float4(fmaxf(v[4 * c] + b4.x, 0.f), fmaxf(v[4 * c + 1] + b4.y, 0.f),
                                             fmaxf(v[4 * c + 2] + b4.z, 0.f), fmaxf(v[4 * c + 3] + b4.w, 0.f));
                    }
                }
            }
        }
    }
    if (warp == 7) {   // ordered compaction of the edges whose endpoints both survive
        int cnt = 0;
        unsigned short *eo = a.e1 + (size_t)g * a.EC1;
        for (int base = 0; base < E; base += 32) {
            const int e = base + lane;
            const bool valid = e < E;
            const int s = valid ? (int)newid[es[e]] : -1, d = valid ? (int)newid[ed[e]] : -1;
            const bool keep = s >= 0 && d >= 0;
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) eo[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(s | (d << 8));
            cnt += __popc(m);
        }
        if (lane == 0) a.e1n[g] = cnt;
    }
    tc_fence_before();
    __syncthreads();
    STG_TRACE(a.trace, 0, 12);  // kept rows out of TMEM, edges filtered
    // ---- outputs ----
    const int R1 = a.R1;
    for (int idx = tid; idx < R1 * 128; idx += NTH) {
        const int r = idx >> 7, c = idx & 127;
        float h = 0.f, v = 0.f;
        if (r < k) { h = hk[r * LDW + c]; v = h * score[perm[r]]; }
        a.x1[((size_t)g * R1 + r) * 128 + c] = v;
        if (SAVE) a.h1k[((size_t)g * R1 + r) * 128 + c] = h;
    }
    if (tid < 128) {
        const int c = tid;
        float mx = -INFINITY, sum = 0.f;
        int am = 0;
        for (int r = 0; r < k; ++r) {
            const float v = hk[r * LDW + c] * score[perm[r]];
            if (v > mx) { mx = v; am = r; }
            sum += v;
        }
        a.r0[(size_t)g * 256 + c] = mx;
        a.r0[(size_t)g * 256 + 128 + c] = sum / (float)(k > 0 ? k : 1);
        if (SAVE) a.amax1[(size_t)g * 128 + c] = (unsigned char)am;
    }
    if (SAVE) {
        const int K2 = 2 * F;
        for (int idx = tid; idx < R1 * K2; idx += NTH) {   // inputs of the kept rows for the weight gradient: [agg | x]
            const int r = idx / K2, kk = idx - r * K2;
            float v = 0.f;
            if (r < k) {
                const int i = perm[r];
                if (kk < F) {
                    const int s0 = rowptr[i], s1 = rowptr[i + 1];
                    float sum = 0.f;
                    for (int s = s0; s < s1; ++s) sum += xs[(key[s] & 255u) * F + kk];
                    v = sum / (float)(s1 - s0 > 0 ? s1 - s0 : 1);
                } else {
                    v = xs[i * F + kk - F];
                }
            }
            a.c1k[((size_t)g * R1 + r) * K2 + kk] = v;
        }
        if (tid < R1) {
            a.s1k[(size_t)g * R1 + tid] = tid < k ? score[perm[tid]] : 0.f;
            a.z1k[(size_t)g * R1 + tid] = tid < k ? zs[perm[tid]] : 0.f;
        }
    }
    __syncthreads();
    STG_TRACE(a.trace, 0, 13);  // outputs written
    STG_GT(a.trace, SAVE ? 3 : 1);
    if (warp == 1) tmem_dealloc(tmem, 256);
}

// ------------------------------------------------------------------------------------------------
// swapped GEMM engine: D^T[128 x N] (TMEM) = Wtile[128 x K] . act[N x K]^T, weights streamed through a ring
// ------------------------------------------------------------------------------------------------
struct Pipe {
    unsigned long long *fullA, *emptyA, *emptyB, *accb;
    unsigned char *a_base, *b_base;
    unsigned b_stage;          // bytes of one activation stage (hi | lo)
    int sta, stb;
    const float *wsplit;
    const WBlk *blk;
    int nblk;
    unsigned j, nissued, acc_cnt, tmem;
    long long *trace;          // profiling: per block (B built, MMAs issued), per GEMM (accumulator ready, epilogue done)
    int tpos, tend;

    __device__ __forceinline__ void stamp()
    {
        if (trace && blockIdx.x == 0 && threadIdx.x == 0 && tpos < tend) trace[tpos++] = clock64();
    }

    __device__ __forceinline__ void init_barriers()   // one thread
    {
        for (int s = 0; s < sta; ++s) { mbar_init(fullA + s, 1); mbar_init(emptyA + s, 1); }
        for (int s = 0; s < stb; ++s) mbar_init(emptyB + s, 1);
        mbar_init(accb, 1);
        mbar_fence_init();
    }
    // warp 0: ask L2 for every weight block of this kernel (fire and forget) so the ring's copies hit L2, not DRAM
    __device__ __forceinline__ void prefetch_all()
    {
        if (threadIdx.x < 32)
            for (int i = (int)threadIdx.x; i < nblk; i += 32) {
                const WBlk b = blk[i];
                bulk_prefetch_l2(wsplit + b.off, 2u * b.half);
            }
    }
    // thread 0: start the bulk copies of every weight block below `upto`
    __device__ __forceinline__ void issue_upto(unsigned upto)
    {
        if (upto > (unsigned)nblk) upto = (unsigned)nblk;
        while (nissued < upto) {
            const int s = (int)(nissued % (unsigned)sta);
            const unsigned use = nissued / (unsigned)sta;
            if (use >= 1) mbar_wait(emptyA + s, (use - 1) & 1);
            const WBlk b = blk[nissued];
            mbar_expect_tx(fullA + s, 2u * b.half);
            bulk_g2s(a_base + (size_t)s * 2 * W_HALF, wsplit + b.off, 2u * b.half, fullA + s);
            ++nissued;
        }
    }
};

// act: fp32 rows in shared memory (row stride ld floats, readable and zero up to K4 = 4*ceil(K/4) columns);
// rows >= nvalid and columns >= K4 enter as zeros.  kpad: the weight matrix' padded K (its blocks are consumed here).
// epi(m, n, v): D^T[m][n] for this thread's feature m = TMEM lane, every row n < N.
// main loop: ONE copy of this code in each kernel (the kernels run a chain of up to 11 GEMMs; inlining it 11 times made
// k_stage2<bwd> 190 KB of straight-line SASS that has to come through a cold instruction cache)
__device__ __noinline__ void gemm_main(Pipe &p, int kpad, const float *act, int ld, int K4, int N, int nvalid)
{
    const int tid = threadIdx.x;
    const unsigned lbo_b = (unsigned)N * 16u;
    const unsigned idesc = idesc_tf32_m128(N);
    const int nkb = (kpad + KB - 1) / KB;
    for (int kb = 0; kb < nkb; ++kb, ++p.j) {
        const unsigned j = p.j;
        const int sa = (int)(j % (unsigned)p.sta), sb = (int)(j % (unsigned)p.stb);
        // keep sta-1 blocks in flight: the copy started here refills the stage block j-1 used, so thread 0 first waits for
        // that block's MMAs (issued one iteration ago, ~a few hundred cycles) -- deeper in flight beats waiting less
        if (tid == 0) p.issue_upto(j + (unsigned)p.sta);
        p.stamp();   // U: copies issued (waited for the stage of block j-1)
        if (j >= (unsigned)p.stb) {   // one poller per warp: 256 threads hammering try_wait slow the copy engine's signalling
            if ((tid & 31) == 0) mbar_wait(p.emptyB + sb, ((j / (unsigned)p.stb) - 1) & 1);
            __syncwarp();
        }
        p.stamp();   // W: activation stage free
        unsigned char *b_hi = p.b_base + (size_t)sb * p.b_stage, *b_lo = b_hi + (p.b_stage >> 1);
        const int k0 = kb * KB;
        const int nch = min(NCH, (kpad - k0) >> 2);
        for (int idx = tid; idx < N * nch; idx += NTH) {
            const int c = idx / N, r = idx - c * N;
            const int k = k0 + 4 * c;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nvalid && k < K4) v = *reinterpret_cast<const float4 *>(act + (size_t)r * ld + k);
            float4 h, l;
            split_tf32(v, h, l);
            const unsigned off = (unsigned)c * lbo_b + (unsigned)(r >> 3) * SBO + (unsigned)(r & 7) * 16u;
            *reinterpret_cast<float4 *>(b_hi + off) = h;
            *reinterpret_cast<float4 *>(b_lo + off) = l;
        }
        p.stamp();   // B: thread 0's share of the activation block written
        fence_async_smem();
        __syncthreads();
        p.stamp();   // S: all threads done
        if (tid == 0) {
            mbar_wait(p.fullA + sa, (j / (unsigned)p.sta) & 1);
            p.stamp();
            tc_fence_after();
            const unsigned ah = s32(p.a_base + (size_t)sa * 2 * W_HALF), al = ah + (unsigned)nch * W_LBO;
            const unsigned bh = s32(b_hi), bl = s32(b_lo);
            for (int ks = 0; ks < nch / 2; ++ks) {
                const unsigned long long dah = smem_desc(ah + 2 * ks * W_LBO, W_LBO, SBO), dal = smem_desc(al + 2 * ks * W_LBO, W_LBO, SBO);
                const unsigned long long dbh = smem_desc(bh + 2 * ks * lbo_b, lbo_b, SBO), dbl = smem_desc(bl + 2 * ks * lbo_b, lbo_b, SBO);
                mma_tf32(p.tmem, dah, dbl, idesc, (kb | ks) ? 1u : 0u);
                mma_tf32(p.tmem, dal, dbh, idesc, 1u);
                mma_tf32(p.tmem, dah, dbh, idesc, 1u);
            }
            mma_commit(p.emptyA + sa);
            mma_commit(p.emptyB + sb);
            if (kb == nkb - 1) mma_commit(p.accb);
            p.stamp();   // M: MMAs issued
        }
    }
    if ((tid & 31) == 0) mbar_wait(p.accb, p.acc_cnt & 1);
    __syncwarp();
    ++p.acc_cnt;
    tc_fence_after();
    p.stamp();
}

// Same product with the WHOLE activation operand already staged (hi at bfull, lo at bfull + half_bytes, chunk-major
// [K/4][N][16 B]): no per-block barrier -- thread 0 alone walks the weight ring (wait, issue, commit), everybody else
// goes straight to the accumulator barrier.  Measured: the per-block build / fence / __syncthreads of gemm_main cost
// ~2.6 k cycles per block against ~0.5 k for wait + issue.
__device__ __forceinline__ void gemm_stream(Pipe &p, int kpad, const unsigned char *bfull, unsigned half_bytes, unsigned lbo_b, int N)
{
    const int tid = threadIdx.x;
    const unsigned idesc = idesc_tf32_m128(N);
    const int nkb = (kpad + KB - 1) / KB;
    if (tid == 64) {
        // producer: keeps the weight ring full; a stage is refilled as soon as the MMAs that read it have committed
        p.issue_upto(p.j + (unsigned)nkb);
    } else if (tid == 0) {
        // MMA issuer: wait for a block, issue its 3 x (nch / 2) MMAs, commit the stage back to the producer
        const unsigned bh0 = s32(bfull), bl0 = bh0 + half_bytes;
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
            const unsigned j = p.j + (unsigned)kb;
            const int sa = (int)(j % (unsigned)p.sta);
            const int nch = min(NCH, (kpad - kb * KB) >> 2);
            mbar_wait(p.fullA + sa, (j / (unsigned)p.sta) & 1);
            tc_fence_after();
            const unsigned ah = s32(p.a_base + (size_t)sa * 2 * W_HALF), al = ah + (unsigned)nch * W_LBO;
            for (int ks = 0; ks < nch / 2; ++ks) {
                const unsigned boff = (unsigned)(kb * NCH + 2 * ks) * lbo_b;
                const unsigned long long dah = smem_desc(ah + 2 * ks * W_LBO, W_LBO, SBO), dal = smem_desc(al + 2 * ks * W_LBO, W_LBO, SBO);
                const unsigned long long dbh = smem_desc(bh0 + boff, lbo_b, SBO), dbl = smem_desc(bl0 + boff, lbo_b, SBO);
                mma_tf32(p.tmem, dah, dbl, idesc, (kb | ks) ? 1u : 0u);
                mma_tf32(p.tmem, dal, dbh, idesc, 1u);
                mma_tf32(p.tmem, dah, dbh, idesc, 1u);
            }
            mma_commit(p.emptyA + sa);
            if (kb == nkb - 1) mma_commit(p.accb);
            p.stamp();
        }
    }
    p.j += (unsigned)nkb;
    if ((tid & 31) == 0) mbar_wait(p.accb, p.acc_cnt & 1);
    __syncwarp();
    ++p.acc_cnt;
    tc_fence_after();
    p.stamp();
}

// epilogue of either main loop: this thread's feature m = TMEM lane, every row n < N
template <typename Epi>
__device__ __forceinline__ void gemm_epilogue(Pipe &p, int N, Epi &&epi)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = (warp & 3) * 32 + lane;
    const unsigned taddr = p.tmem + (((unsigned)(warp & 3) * 32u) << 16);
    for (int c0 = 8 * (warp >> 2); c0 < N; c0 += 16) {
        float v[8];
        tmem_ld8(taddr + (unsigned)c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) epi(m, c0 + i, v[i]);
    }
    tc_fence_before();
    __syncthreads();
    p.stamp();
}

template <typename Epi>
__device__ __forceinline__ void gemm_sw(Pipe &p, int kpad, const float *act, int ld, int K4, int N, int nvalid, Epi &&epi)
{
    gemm_main(p, kpad, act, ld, K4, N, nvalid);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = (warp & 3) * 32 + lane;
    const unsigned taddr = p.tmem + (((unsigned)(warp & 3) * 32u) << 16);
    for (int c0 = 8 * (warp >> 2); c0 < N; c0 += 16) {
        float v[8];
        tmem_ld8(taddr + (unsigned)c0, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) epi(m, c0 + i, v[i]);
    }
    tc_fence_before();
    __syncthreads();
    p.stamp();
}

// TopK score of rows H[row][0..128) (stride ld): z = h.w / ||w||, s = tanh(z); one warp per row
__device__ __forceinline__ void row_scores(const float *H, int ld, int nrows, const float *pw, float *score, float *zs)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float4 w = __ldg(reinterpret_cast<const float4 *>(pw) + lane);
    const float pn = sqrtf(warp_sum(dot4(w, w)));
    for (int r = warp; r < nrows; r += NTH / 32) {
        const float4 h = *reinterpret_cast<const float4 *>(H + (size_t)r * ld + 4 * lane);
        const float d = warp_sum(dot4(h, w));
        if (lane == 0) {
            const float z = d / pn;
            zs[r] = z;
            score[r] = tanhf(z);
        }
    }
}

// pooling backward of one kept row held by a warp (lane = 4 columns): gradient of x_out = h * s, s = tanh(h.w/||w||),
// through the ReLU that produced h.  Returns dP (gradient of the pre-activation); adds this row's pool-weight term.
__device__ __forceinline__ float4 pool_bwd_row(float4 dxo, float4 h, float s, float z, float4 w, float wn, float4 &dpool)
{
    const float ds = warp_sum(dot4(dxo, h));
    const float tds = ds * (1.f - s * s);
    const float wn2 = wn * wn;
    dpool.x += tds * (h.x / wn - z * w.x / wn2);
    dpool.y += tds * (h.y / wn - z * w.y / wn2);
    dpool.z += tds * (h.z / wn - z * w.z / wn2);
    dpool.w += tds * (h.w / wn - z * w.w / wn2);
    float4 dp;
    dp.x = h.x > 0.f ? dxo.x * s + tds * w.x / wn : 0.f;
    dp.y = h.y > 0.f ? dxo.y * s + tds * w.y / wn : 0.f;
    dp.z = h.z > 0.f ? dxo.z * s + tds * w.z / wn : 0.f;
    dp.w = h.w > 0.f ? dxo.w * s + tds * w.w / wn : 0.f;
    return dp;
}

// ------------------------------------------------------------------------------------------------
// k_stage1: block 1 (SAGEConv 256 -> 128, TopK, readout) of GS graphs per CTA
// ------------------------------------------------------------------------------------------------
struct S1Args {
    long long *trace;
    const float *params, *wsplit;
    WBlk blk[8];
    int nblk, b_off, pool_off;
    float ratio;
    const int *nptr;
    int B, GS, R1, EC1, R2, EC2, NRMAX, sta;
    const float *x1;
    const unsigned short *e1;
    const int *e1n;
    float *x2;                     // [B][R2][128]
    unsigned short *e2;            // [B][EC2]
    int *e2n;
    float *r1;                     // [B][256]
    float *h2k, *c2k, *s2k, *z2k;  // saves: [B][R2][128], [B*R2][256], [B][R2] x2
    unsigned char *perm2, *amax2;  // [B][R2], [B][128]
    int o_a, o_b, o_cat, o_es, o_int, o_score, o_z, o_newid, o_perm, total, tmem_cols;
};

inline int s1_layout(S1Args &a)
{
    int o = 256;
    auto take = [&](int bytes) { int at = o; o += rup(bytes, 128); return at; };
    const int npad = rup(a.NRMAX, 16);
    a.o_a = take(a.sta * 2 * W_HALF);                 // weight ring; afterwards the hidden rows H2 [NR][LDW]
    a.o_b = take(2 * 64 * (npad * 16 + 16));          // the whole [mean | x] operand: hi | lo, each [64 chunks][rows + 1][16 B]
    a.o_cat = take(a.NRMAX * LDW * 4);                // fp32 x rows (sources of the mean)
    a.o_es = take(a.GS * a.EC1 * 2);
    a.o_int = take((5 * a.GS + 4) * 4);               // rb[GS+1] n1[GS] ecnt[GS] k2[GS] ob[GS+1]
    a.o_score = take(a.NRMAX * 4);
    a.o_z = take(a.NRMAX * 4);
    a.o_newid = take(a.NRMAX * 2);
    a.o_perm = take(a.GS * a.R2 * 4);
    a.total = o;
    int cols = 32;
    while (cols < npad) cols <<= 1;
    a.tmem_cols = cols;
    if (a.NRMAX * LDW * 4 > a.sta * 2 * W_HALF) return -1;
    return o;
}

template <bool SAVE>
__global__ void __launch_bounds__(NTH, 1) k_stage1(const __grid_constant__ S1Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int GS = a.GS, R1 = a.R1, R2 = a.R2;
    const int g0 = blockIdx.x * GS, ng = min(GS, a.B - g0);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(sm + 160);
    float *xs = reinterpret_cast<float *>(sm + a.o_cat);
    float *H2 = reinterpret_cast<float *>(sm + a.o_a);
    unsigned short *es1 = reinterpret_cast<unsigned short *>(sm + a.o_es);
    int *rb = reinterpret_cast<int *>(sm + a.o_int), *n1s = rb + GS + 1, *ecnt = n1s + GS, *k2s = ecnt + GS, *ob = k2s + GS;
    float *score = reinterpret_cast<float *>(sm + a.o_score), *zs = reinterpret_cast<float *>(sm + a.o_z);
    short *newid = reinterpret_cast<short *>(sm + a.o_newid);
    int *perm = reinterpret_cast<int *>(sm + a.o_perm);
    const int npad = rup(a.NRMAX, 16);
    unsigned char *b_hi = sm + a.o_b;
    const unsigned bhalf = 64u * ((unsigned)npad * 16u + 16u);
    unsigned char *b_lo = b_hi + bhalf;

    Pipe p;
    p.fullA = bars; p.emptyA = bars + 8; p.emptyB = bars + 16; p.accb = bars + 18;
    p.a_base = sm + a.o_a; p.b_base = nullptr; p.b_stage = 0;
    p.sta = a.sta; p.stb = 2; p.wsplit = a.wsplit; p.blk = a.blk; p.nblk = a.nblk;
    p.j = 0; p.nissued = 0; p.acc_cnt = 0; p.trace = a.trace; p.tpos = 32 + 8; p.tend = 96;
    STG_TRACE(a.trace, 32, 0);
    STG_GT(a.trace, SAVE ? 6 : 4);
    if (tid == 0) {
        p.init_barriers();
        int r = 0, o2 = 0;
        for (int gi = 0; gi < ng; ++gi) {
            const int g = g0 + gi;
            const int n1 = topk_count_dev(a.ratio, a.nptr[g + 1] - a.nptr[g]);
            rb[gi] = r; n1s[gi] = n1; ecnt[gi] = a.e1n[g];
            ob[gi] = o2; k2s[gi] = topk_count_dev(a.ratio, n1);
            r += n1; o2 += k2s[gi];
        }
        rb[ng] = r; ob[ng] = o2;
    }
    if (warp == 1) tmem_alloc(tmem_slot, (unsigned)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    p.tmem = *tmem_slot;
    if (tid == 64) p.issue_upto((unsigned)p.sta);   // the producer thread (see gemm_stream) fills the weight ring
    STG_TRACE(a.trace, 32, 1);   // set-up done
    const int NR = rb[ng];
    const int N = (NR + 15) & ~15;
    const unsigned lbo = (unsigned)N * 16u + 16u;  // chunk stride: N rows + one pad slot, so the 32 lanes of a warp (one chunk
                                                   // each, same row) store to distinct bank groups
    auto graph_of = [&](int row) { int gi = 0; while (gi + 1 < ng && row >= rb[gi + 1]) ++gi; return gi; };
    // x half: global -> fp32 copy (mean sources) and the hi / lo operand chunks 32..63; loads batched 4 deep
    for (int base = 0; base < N * 32; base += NTH * 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * NTH + tid, row = idx >> 5, c4 = idx & 31;
            v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < NR) {
                const int gi = graph_of(row);
                v[u] = __ldg(reinterpret_cast<const float4 *>(a.x1 + ((size_t)(g0 + gi) * R1 + (row - rb[gi])) * 128) + c4);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * NTH + tid, row = idx >> 5, c4 = idx & 31;
            if (row < N) {
                if (row < NR) *reinterpret_cast<float4 *>(xs + (size_t)row * LDW + 4 * c4) = v[u];
                float4 h, l;
                split_tf32(v[u], h, l);
                const unsigned off = (unsigned)(32 + c4) * lbo + (unsigned)row * 16u;
                *reinterpret_cast<float4 *>(b_hi + off) = h;
                *reinterpret_cast<float4 *>(b_lo + off) = l;
            }
        }
    }
    p.stamp();   // x rows stored (thread 0's share)
    for (int gi = 0; gi < ng; ++gi)
        for (int j = tid; j < ecnt[gi]; j += NTH) es1[gi * a.EC1 + j] = a.e1[(size_t)(g0 + gi) * a.EC1 + j];
    p.stamp();   // edges copied
    __syncthreads();
    p.stamp();   // everyone there
    // mean over in-edges, in edge order; half-warp per row (16 lanes x 8 columns: two rows' serial edge scans run side
    // by side) -> operand chunks 0..31 (rows >= NR: zeros)
    {
        const int hl = lane & 15, hw = (tid >> 4);          // lane within the half-warp, half-warp id (0..15)
        for (int row = hw; row < N; row += NTH / 16) {
            float4 m0 = make_float4(0.f, 0.f, 0.f, 0.f), m1 = m0;
            if (row < NR) {
                const int gi = graph_of(row), r = row - rb[gi];
                const unsigned short *el = es1 + gi * a.EC1;
                const int ne = ecnt[gi];
                int cnt = 0;
                for (int j = 0; j < ne; ++j) {
                    const unsigned ev = el[j];
                    if ((int)(ev >> 8) == r) {
                        const float *xr = xs + (size_t)(rb[gi] + (ev & 255u)) * LDW;
                        m0 = f4_add(m0, *reinterpret_cast<const float4 *>(xr + 4 * hl));
                        m1 = f4_add(m1, *reinterpret_cast<const float4 *>(xr + 64 + 4 * hl));
                        ++cnt;
                    }
                }
                const float inv = __frcp_rn((float)(cnt > 0 ? cnt : 1));   // mean = sum * (1 / deg)
                m0 = make_float4(m0.x * inv, m0.y * inv, m0.z * inv, m0.w * inv);
                m1 = make_float4(m1.x * inv, m1.y * inv, m1.z * inv, m1.w * inv);
            }
            float4 h, l;
            split_tf32(m0, h, l);
            unsigned off = (unsigned)hl * lbo + (unsigned)row * 16u;
            *reinterpret_cast<float4 *>(b_hi + off) = h;
            *reinterpret_cast<float4 *>(b_lo + off) = l;
            split_tf32(m1, h, l);
            off += 16u * lbo;
            *reinterpret_cast<float4 *>(b_hi + off) = h;
            *reinterpret_cast<float4 *>(b_lo + off) = l;
        }
    }
    p.stamp();   // mean rows (thread 0's share)
    fence_async_smem();
    __syncthreads();
    STG_TRACE(a.trace, 32, 2);   // [mean | x] operand staged
    gemm_stream(p, 256, b_hi, bhalf, lbo, N);
    {
        const float bm = __ldg(a.params + a.b_off + (warp & 3) * 32 + lane);
        gemm_epilogue(p, N, [&](int m, int n, float v) {
            if (n < NR) H2[(size_t)n * LDW + m] = fmaxf(v + bm, 0.f);
        });
    }
    STG_TRACE(a.trace, 32, 3);   // GEMM + epilogue
    row_scores(H2, LDW, NR, a.params + a.pool_off, score, zs);
    __syncthreads();
    STG_TRACE(a.trace, 32, 4);   // scores
    if (tid < NR) {
        const int gi = graph_of(tid), r0 = rb[gi], r1 = rb[gi + 1], k = k2s[gi];
        const float si = score[tid];
        int r = 0;
        for (int j = r0; j < r1; ++j) {
            const float sj = score[j];
            r += (sj > si) || (sj == si && j < tid);
        }
        newid[tid] = (short)(r < k ? r : -1);
        if (r < k) perm[gi * R2 + r] = tid - r0;
    }
    __syncthreads();
    for (int idx = tid; idx < ng * R2 * 128; idx += NTH) {
        const int gi = idx / (R2 * 128), rem = idx - gi * R2 * 128, r = rem >> 7, c = rem & 127;
        float h = 0.f, v = 0.f;
        if (r < k2s[gi]) {
            const int i = rb[gi] + perm[gi * R2 + r];
            h = H2[(size_t)i * LDW + c];
            v = h * score[i];
        }
        a.x2[((size_t)(g0 + gi) * R2 + r) * 128 + c] = v;
        if (SAVE) a.h2k[((size_t)(g0 + gi) * R2 + r) * 128 + c] = h;
    }
    for (int idx = tid; idx < ng * 128; idx += NTH) {
        const int gi = idx >> 7, c = idx & 127, k = k2s[gi];
        float mx = -INFINITY, sum = 0.f;
        int am = 0;
        for (int r = 0; r < k; ++r) {
            const int i = rb[gi] + perm[gi * R2 + r];
            const float v = H2[(size_t)i * LDW + c] * score[i];
            if (v > mx) { mx = v; am = r; }
            sum += v;
        }
        a.r1[(size_t)(g0 + gi) * 256 + c] = mx;
        a.r1[(size_t)(g0 + gi) * 256 + 128 + c] = sum / (float)(k > 0 ? k : 1);
        if (SAVE) a.amax2[(size_t)(g0 + gi) * 128 + c] = (unsigned char)am;
    }
    for (int gi = warp; gi < ng; gi += NTH / 32) {   // level-2 edges: both endpoints kept, order kept
        const unsigned short *el = es1 + gi * a.EC1;
        unsigned short *eo = a.e2 + (size_t)(g0 + gi) * a.EC2;
        const int ne = ecnt[gi];
        int cnt = 0;
        for (int base = 0; base < ne; base += 32) {
            const int e = base + lane;
            const bool valid = e < ne;
            const unsigned ev = valid ? el[e] : 0u;
            const int s = valid ? (int)newid[rb[gi] + (ev & 255u)] : -1, d = valid ? (int)newid[rb[gi] + (ev >> 8)] : -1;
            const bool keep = s >= 0 && d >= 0;
            const unsigned m = __ballot_sync(FULL, keep);
            if (keep) eo[cnt + __popc(m & ((1u << lane) - 1u))] = (unsigned short)(s | (d << 8));
            cnt += __popc(m);
        }
        if (lane == 0) a.e2n[g0 + gi] = cnt;
    }
    if (SAVE) {
        // inputs [mean | x] of the kept rows for the weight gradient: hi + lo restores the fp32 value exactly
        for (int idx = tid; idx < ng * R2 * 64; idx += NTH) {
            const int gi = idx / (R2 * 64), rem = idx - gi * R2 * 64, r = rem >> 6, ch = rem & 63;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < k2s[gi]) {
                const unsigned off = (unsigned)ch * lbo + (unsigned)(rb[gi] + perm[gi * R2 + r]) * 16u;
                const float4 h = *reinterpret_cast<const float4 *>(b_hi + off), l = *reinterpret_cast<const float4 *>(b_lo + off);
                v = f4_add(h, l);
            }
            *reinterpret_cast<float4 *>(a.c2k + ((size_t)(g0 + gi) * R2 + r) * 256 + 4 * ch) = v;
        }
        for (int idx = tid; idx < ng * R2; idx += NTH) {
            const int gi = idx / R2, r = idx - gi * R2;
            const bool ok = r < k2s[gi];
            const int li = ok ? perm[gi * R2 + r] : 0;
            a.s2k[(size_t)(g0 + gi) * R2 + r] = ok ? score[rb[gi] + li] : 0.f;
            a.z2k[(size_t)(g0 + gi) * R2 + r] = ok ? zs[rb[gi] + li] : 0.f;
            a.perm2[(size_t)(g0 + gi) * R2 + r] = (unsigned char)li;
        }
    }
    __syncthreads();
    STG_TRACE(a.trace, 32, 5);   // outputs
    STG_GT(a.trace, SAVE ? 7 : 5);
    if (warp == 1) tmem_dealloc(p.tmem, (unsigned)a.tmem_cols);
}

}  // namespace stg
