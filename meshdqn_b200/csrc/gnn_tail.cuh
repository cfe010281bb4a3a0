// gnn_tail.cuh -- the row-starved tail of the staged Q-path on the CUDA cores (included after gnn_staged.cuh).
//
// After two TopK(0.1) poolings a graph holds <= 4 rows (conv4), then 1 (conv5, MLP): 8 graphs per CTA are 8..32 rows.
// Measured on B200 (tools/staged_trace.py, round 2): a tcgen05 chain over these layers pays ~3.6 k cycles of
// fence / barrier / issue latency per 32-deep K block (24 blocks forward, 48 with the backward chain) and streams the
// hi+lo weight tiles (2x the bytes) -- 129 k cycles per CTA for 52 MFLOP.  With so few rows the layers are bound by
// getting the 334 KB of fp32 weights into the SM once, not by math, so here:
//   * the flat fp32 weights stream through a bulk-TMA ring of [32 input rows][out] blocks (mbarrier full / empty, the
//     whole ring in flight, no CTA-wide barrier per block);
//   * forward layers: thread = output column, 4..8 rows in registers, k ascending fp32 FMA (the fused kernel's order);
//   * transposed layers (backward): warp per weight row, lanes across the row, butterfly reduction of 8 rows at once;
//   * the same blocks serve both directions, so k_tail<bwd> streams each weight matrix twice and nothing else.
// k_bwd1 (block 1 / block 0 backward) uses the same ring for conv2^T: 4..6 rows per CTA.
#pragma once

namespace stg {

constexpr int RST = 6;                    // ring stages
constexpr int RSTAGE = 24 * 1024;         // bytes per stage: 32 rows x (<= 192 floats)
constexpr int MAXRB = 64;

struct RBlk { unsigned off, bytes; };     // float offset into the flat parameters (bit 31 set: into wsplit), bytes (% 16 == 0)
constexpr unsigned RB_WSPLIT = 0x80000000u;

__host__ __device__ inline int rows_per_block(int C) { return C > 192 ? 16 : 32; }

// blocks of `rows` weight rows of length C starting at float offset `off` (of params, or of wsplit when from_wsplit)
inline int add_rblocks(int off, int rows, int C, RBlk *out, int n, bool from_wsplit = false)
{
    const int rpb = rows_per_block(C);
    for (int k0 = 0; k0 < rows; k0 += rpb) {
        const int nr = rows - k0 < rpb ? rows - k0 : rpb;
        out[n].off = (unsigned)(off + k0 * C) | (from_wsplit ? RB_WSPLIT : 0u);
        out[n].bytes = (unsigned)(((nr * C * 4) + 15) & ~15);
        ++n;
    }
    return n;
}

struct Ring {
    unsigned long long *full, *empty;
    unsigned char *base;
    const float *params, *wsplit;
    const RBlk *blk;
    int nblk;
    unsigned j, nissued;

    __device__ __forceinline__ void init()   // one thread
    {
        for (int s = 0; s < RST; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NTH / 32); }
        mbar_fence_init();
    }
    __device__ __forceinline__ void issue_upto(unsigned upto)   // thread 0
    {
        if (upto > (unsigned)nblk) upto = (unsigned)nblk;
        while (nissued < upto) {
            const int s = (int)(nissued % RST);
            const unsigned use = nissued / RST;
            if (use >= 1) mbar_wait(empty + s, (use - 1) & 1);
            const RBlk b = blk[nissued];
            mbar_expect_tx(full + s, b.bytes);
            bulk_g2s(base + (size_t)s * RSTAGE, (b.off & RB_WSPLIT) ? wsplit + (b.off & ~RB_WSPLIT) : params + b.off, b.bytes, full + s);
            ++nissued;
        }
    }
    // every thread: the next block's rows in shared memory
    __device__ __forceinline__ const float *acquire()
    {
        const int s = (int)(j % RST);
        if ((threadIdx.x & 31) == 0) mbar_wait(full + s, (j / RST) & 1);
        __syncwarp();
        return reinterpret_cast<const float *>(base + (size_t)s * RSTAGE);
    }
    // every compute thread, after its last read of the block: the warp hands the stage back to the producer warp
    __device__ __forceinline__ void release()
    {
        const int s = (int)(j % RST);
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
        ++j;
    }
};

// The kernels below run NTH compute threads plus ONE producer warp (threads NTH..NTH+31) whose lane 0 does nothing but
// keep the ring full (wait for a stage to come back, start its bulk copy); the compute warps never issue a copy and
// synchronise among themselves on a named barrier.
constexpr int NTH_TAIL = NTH + 32;
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void k_post(unsigned *sync)
{
    __threadfence();
    atomicAdd(sync, 1u);
}
__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NTH) : "memory"); }

// out[r][c] = sum_k act[r][k] * W[k][c] for a weight matrix of K rows (multiple of 32 rows per block, last block may be
// short) and C columns, rows r < 8*P.  CP = C padded to 64 / 128 / 256: 256 / CP row slots, RT = 8 * CP / 256 rows per
// thread and pass.  act rows beyond the valid ones must be readable (finite or not: their results are discarded).
template <int CP, int P>
__device__ __forceinline__ void dense_fwd(Ring &rg, int K, int C, const float *act, int ld, int np, float (&acc)[P][8 * CP / 256])
{
    constexpr int RT = 8 * CP / 256;
    const int c = threadIdx.x & (CP - 1), slot = threadIdx.x / CP;
    const bool cok = c < C;
#pragma unroll
    for (int p = 0; p < P; ++p)
#pragma unroll
        for (int q = 0; q < RT; ++q) acc[p][q] = 0.f;
    const int rpb = rows_per_block(C);
    for (int k0 = 0; k0 < K; k0 += rpb) {
        const float *w = rg.acquire();
        const int nk = K - k0 < rpb ? K - k0 : rpb;
        // 16 weights of this thread's column at a time (independent loads, all in flight), then the rows: per row 4
        // broadcast 16-byte loads feed 16 FMAs (k ascending per accumulator)
        for (int kk = 0; kk < nk; kk += 16) {
            float wv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) wv[i] = (cok && kk + i < nk) ? w[(kk + i) * C + c] : 0.f;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (p >= np) continue;
#pragma unroll
                for (int q = 0; q < RT; ++q) {
                    const float4 *ar = reinterpret_cast<const float4 *>(act + (size_t)(p * 8 + slot * RT + q) * ld + k0 + kk);
                    float s = acc[p][q];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float4 a4 = ar[u];
                        s = fmaf(a4.x, wv[4 * u], s);
                        s = fmaf(a4.y, wv[4 * u + 1], s);
                        s = fmaf(a4.z, wv[4 * u + 2], s);
                        s = fmaf(a4.w, wv[4 * u + 3], s);
                    }
                    acc[p][q] = s;
                }
            }
        }
        rg.release();
    }
}

// ------------------------------------------------------------------------------------------------
// k_tail: blocks 2 and 3, readout sum, MLP, softmax / argmax of GS graphs per CTA; BWD continues down to dX2
// ------------------------------------------------------------------------------------------------
struct TArgs {
    long long *trace;
    const float *params, *wsplit;
    RBlk blk[MAXRB];
    int nblk;
    int b4_off, p4_off, b5_off, p5_off, lb_off[3];
    int A, softmax;
    float ratio;
    const int *nptr;
    int B, GS, R2, EC2;
    const float *x2;
    const unsigned short *e2;
    const int *e2n;
    const float *r0, *r1;
    float *x3;                    // [B][128] level-3 rows (conv5 inputs)
    float *out, *emb;
    int *amax_out;
    // ---- backward ----
    int mode;                     // 0: grad_out given  1: Huber through Q(s)[a]  2: through max_a Q(s')
    const float *gout;
    const int *rp_action;
    const float *rp_reward;
    const int *rp_index;
    const float *rp_qother;
    float rp_gamma, rp_inv_batch;
    float *rp_scalar;
    float *lterm;                 // [n_graphs] Huber term of each graph's transition (modes 1 / 2); summed by k_bwd1
    float *lin_in[3], *lin_d[3];  // weight-gradient rows: inputs Rs / y1 / y2, deltas d1 / d2 / d3
    float *c5_d, *c4_d, *pool4_d, *pool5_d, *bias4_d, *bias5_d;
    float *dX2, *dR;              // [B][R2][128], [B][256]
    // early launch (BWD, modes 1 / 2): the kernel is started BEFORE Q_other exists and waits for it right where the loss needs
    // it -- sync[0] = posts (k_post after the other net's forward), sync[1] = tail launches completed, sync[2] = CTA ticket
    unsigned *sync;
    int o_ring, o_x2s, o_xw, o_x3s, o_h4, o_rs, o_y1, o_y2, o_y3, o_es, o_int, o_f, o_rowg, o_lg, total;
};

inline int tail_layout(TArgs &a)
{
    int o = 256;
    auto take = [&](int bytes) { int at = o; o += rup(bytes, 128); return at; };
    const int n2m = rup(a.GS * a.R2, 8), gsp = rup(a.GS, 8);
    a.o_ring = take(RST * RSTAGE);
    a.o_x2s = take(n2m * LDW * 4);      // X2 rows, then H3
    a.o_xw = take(n2m * LDW * 4);       // X2.W4, then dXW3
    a.o_x3s = take(gsp * LDW * 4);      // X3 rows, then dX3
    a.o_h4 = take(gsp * LDW * 4);       // H4, then dP4
    a.o_rs = take(gsp * LD2W * 4);      // readout sum, then dR
    a.o_y1 = take(gsp * LDW * 4);       // y1, then d1
    a.o_y2 = take(gsp * LDY2 * 4);      // y2, then d2
    a.o_y3 = take(gsp * LDY3 * 4);      // logits / softmax, then d3
    a.o_es = take(a.GS * a.EC2 * 2);
    a.o_int = take((4 * a.GS + 2) * 4);
    a.o_f = take((3 * n2m + 2 * a.GS) * 4);
    a.o_rowg = take(n2m);
    a.o_lg = take(a.GS * 4 * 4);        // loss-gradient inputs fetched at kernel start: sel, pred, nsv, rew per graph
    a.total = o;
    return o;
}

template <bool BWD>
__global__ void __launch_bounds__(NTH_TAIL, 1) k_tail(const __grid_constant__ TArgs a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int GS = a.GS, R2 = a.R2, A = a.A;
    const int g0 = blockIdx.x * GS, ng = min(GS, a.B - g0);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    float *X2s = reinterpret_cast<float *>(sm + a.o_x2s), *H3 = X2s;
    float *XW = reinterpret_cast<float *>(sm + a.o_xw);
    float *X3s = reinterpret_cast<float *>(sm + a.o_x3s), *H4 = reinterpret_cast<float *>(sm + a.o_h4);
    float *Rs = reinterpret_cast<float *>(sm + a.o_rs), *y1 = reinterpret_cast<float *>(sm + a.o_y1);
    float *y2 = reinterpret_cast<float *>(sm + a.o_y2), *y3 = reinterpret_cast<float *>(sm + a.o_y3);
    unsigned short *es2 = reinterpret_cast<unsigned short *>(sm + a.o_es);
    int *rb2 = reinterpret_cast<int *>(sm + a.o_int), *n2s = rb2 + GS + 1, *ecnt = n2s + GS, *sel3 = ecnt + GS;
    const int n2m = rup(GS * R2, 8);
    float *dis = reinterpret_cast<float *>(sm + a.o_f), *score3 = dis + n2m, *z3 = score3 + n2m, *s4 = z3 + n2m, *z4 = s4 + GS;
    unsigned char *rowg = sm + a.o_rowg;
    const float *P = a.params;

    Ring rg;
    rg.full = bars; rg.empty = bars + 8; rg.base = sm + a.o_ring; rg.params = P; rg.wsplit = a.wsplit; rg.blk = a.blk; rg.nblk = a.nblk;
    rg.j = 0; rg.nissued = 0;
    STG_TRACE(a.trace, 96, 0);
    STG_GT(a.trace, BWD ? 10 : 8);
    if (tid == 0) rg.init();
    __syncthreads();               // all NTH_TAIL threads: the barriers exist
    if (tid >= NTH) {              // producer warp
        if (tid == NTH) rg.issue_upto((unsigned)rg.nblk);
        return;
    }
    if (tid == 0) {
        int r = 0;
        for (int gi = 0; gi < ng; ++gi) {
            const int g = g0 + gi;
            const int n2 = topk_count_dev(a.ratio, topk_count_dev(a.ratio, a.nptr[g + 1] - a.nptr[g]));
            rb2[gi] = r; n2s[gi] = n2; ecnt[gi] = a.e2n[g];
            r += n2;
        }
        rb2[ng] = r;
    }
    // rows past the valid ones feed discarded accumulators only, but they must not be uninitialised NaN patterns that
    // trap nothing -- they are simply never read back; zero them once so every buffer read below is defined
    for (int i = tid; i < (a.o_es - a.o_x2s) / 4; i += NTH) reinterpret_cast<float *>(sm + a.o_x2s)[i] = 0.f;
    cta_sync();
    STG_TRACE(a.trace, 96, 1);
    const int N2 = rb2[ng];
    for (int idx = tid; idx < N2 * 32; idx += NTH) {
        const int row = idx >> 5, c4 = idx & 31;
        int gi = 0;
        while (gi + 1 < ng && row >= rb2[gi + 1]) ++gi;
        const int r = row - rb2[gi];
        *reinterpret_cast<float4 *>(X2s + (size_t)row * LDW + 4 * c4) =
            __ldg(reinterpret_cast<const float4 *>(a.x2 + ((size_t)(g0 + gi) * R2 + r) * 128) + c4);
        if (c4 == 0) rowg[row] = (unsigned char)gi;
    }
    for (int gi = 0; gi < ng; ++gi)
        for (int j = tid; j < ecnt[gi]; j += NTH) es2[gi * a.EC2 + j] = a.e2[(size_t)(g0 + gi) * a.EC2 + j];
    float *lgf = reinterpret_cast<float *>(sm + a.o_lg);   // [GS][4]: sel (as int bits), pred, nsv, rew
    // the transition data and Q_other row of each graph.  Normally fetched now, so that their (cold) latency hides under the
    // forward chain; an early-launched kernel (a.sync) fetches them after its wait instead, with L2-coherent loads
    auto fetch_loss_inputs = [&](bool coherent) {
        for (int gi = warp; gi < ng; gi += NTH / 32) {
            const int g = g0 + gi;
            int sel = 0;
            float pred = 0.f, nsv = 0.f, rew;
            if (a.mode == 1) {
                sel = a.rp_action[g];
                rew = a.rp_reward[g];
                const int slot = a.rp_index[g];
                float m = -INFINITY;
                if (slot >= 0) {
                    const float *q = a.rp_qother + (size_t)slot * A;
                    for (int c = lane; c < A; c += 32) m = fmaxf(m, coherent ? __ldcg(q + c) : __ldg(q + c));
#pragma unroll
                    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
                }
                nsv = slot >= 0 ? m : 0.f;
            } else {
                const int b = a.rp_index[g];
                const float *q = a.rp_qother + (size_t)b * A + a.rp_action[b];
                pred = coherent ? __ldcg(q) : __ldg(q);
                rew = a.rp_reward[b];
            }
            if (lane == 0) {
                lgf[gi * 4] = __int_as_float(sel);
                lgf[gi * 4 + 1] = pred;
                lgf[gi * 4 + 2] = nsv;
                lgf[gi * 4 + 3] = rew;
            }
        }
    };
    unsigned expected = 0;
    if (BWD && a.mode != 0) {
        if (a.sync) expected = *reinterpret_cast<volatile unsigned *>(a.sync + 1) + 1u;   // same value in every CTA (see the end)
        else fetch_loss_inputs(false);
    }
    cta_sync();
    STG_TRACE(a.trace, 96, 2);   // inputs loaded
    const int c128 = tid & 127, half = tid >> 7;
    // ---- block 2: GCNConv (conv4): XW = X2 . W4, then A_hat ----
    {
        float acc[4][4];
        dense_fwd<128, 4>(rg, 128, 128, X2s, LDW, (N2 + 7) >> 3, acc);
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = p * 8 + half * 4 + q;
                if (r < N2) XW[(size_t)r * LDW + c128] = acc[p][q];
            }
    }
    if (tid < N2) {
        const int gi = rowg[tid], r = tid - rb2[gi];
        const unsigned short *el = es2 + gi * a.EC2;
        int deg = 1;
        for (int j = 0; j < ecnt[gi]; ++j) {
            const unsigned ev = el[j];
            deg += ((int)(ev >> 8) == r && (int)(ev & 255u) != r);
        }
        dis[tid] = __fdiv_rn(1.f, __fsqrt_rn((float)deg));
    }
    cta_sync();
    for (int row = warp; row < N2; row += NTH / 32) {
        const int gi = rowg[row], r = row - rb2[gi];
        const unsigned short *el = es2 + gi * a.EC2;
        const float di = dis[row];
        float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j = 0; j < ecnt[gi]; ++j) {
            const unsigned ev = el[j];
            const int s = (int)(ev & 255u);
            if ((int)(ev >> 8) == r && s != r) {
                const float cf = dis[rb2[gi] + s] * di;
                const float4 xv = *reinterpret_cast<const float4 *>(XW + (size_t)(rb2[gi] + s) * LDW + 4 * lane);
                sum.x += cf * xv.x; sum.y += cf * xv.y; sum.z += cf * xv.z; sum.w += cf * xv.w;
            }
        }
        const float4 xs = *reinterpret_cast<const float4 *>(XW + (size_t)row * LDW + 4 * lane);
        const float4 b = __ldg(reinterpret_cast<const float4 *>(P + a.b4_off) + lane);
        const float dd = di * di;
        sum.x += dd * xs.x; sum.y += dd * xs.y; sum.z += dd * xs.z; sum.w += dd * xs.w;
        sum.x += b.x; sum.y += b.y; sum.z += b.z; sum.w += b.w;
        *reinterpret_cast<float4 *>(H3 + (size_t)row * LDW + 4 * lane) =
            make_float4(fmaxf(sum.x, 0.f), fmaxf(sum.y, 0.f), fmaxf(sum.z, 0.f), fmaxf(sum.w, 0.f));
    }
    cta_sync();
    row_scores(H3, LDW, N2, P + a.p4_off, score3, z3);
    cta_sync();
    if (tid < ng) {   // TopK with one survivor: highest score, ties -> lower index
        int best = rb2[tid];
        for (int j = rb2[tid] + 1; j < rb2[tid + 1]; ++j)
            if (score3[j] > score3[best]) best = j;
        sel3[tid] = best - rb2[tid];
    }
    cta_sync();
    for (int idx = tid; idx < ng * 128; idx += NTH) {
        const int gi = idx >> 7, c = idx & 127;
        const int i = rb2[gi] + sel3[gi];
        const float v = H3[(size_t)i * LDW + c] * score3[i];
        X3s[gi * LDW + c] = v;
        a.x3[(size_t)(g0 + gi) * 128 + c] = v;
    }
    cta_sync();
    STG_TRACE(a.trace, 96, 3);   // block 2 done
    // ---- block 3: GCNConv (conv5) on the single remaining node: A_hat = [1] ----
    {
        float acc[1][4];
        dense_fwd<128, 1>(rg, 128, 128, X3s, LDW, 1, acc);
        const float bm = __ldg(P + a.b5_off + c128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = half * 4 + q;
            if (r < ng) H4[r * LDW + c128] = fmaxf(acc[0][q] + bm, 0.f);
        }
    }
    cta_sync();
    row_scores(H4, LDW, ng, P + a.p5_off, s4, z4);
    cta_sync();
    // ---- readout: x1 + x2 + x4 + x5 (airfoilgcnn.py:134), each [max | mean] over its kept rows ----
    for (int idx = tid; idx < ng * 256; idx += NTH) {
        const int gi = idx >> 8, c = idx & 255, cc = c & 127;
        const size_t go = (size_t)(g0 + gi) * 256 + c;
        const float v = ((__ldg(a.r0 + go) + __ldg(a.r1 + go)) + X3s[gi * LDW + cc]) + H4[gi * LDW + cc] * s4[gi];
        Rs[gi * LD2W + c] = v;
        if (a.emb) a.emb[go] = v;
        if (BWD) a.lin_in[0][go] = v;
    }
    cta_sync();
    STG_TRACE(a.trace, 96, 4);   // block 3 + readout sum
    // ---- MLP ----
    {
        float acc[1][4];
        dense_fwd<128, 1>(rg, 256, 128, Rs, LD2W, 1, acc);
        const float bm = __ldg(P + a.lb_off[0] + c128);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = half * 4 + q;
            if (r < ng) {
                const float y = fmaxf(acc[0][q] + bm, 0.f);
                y1[r * LDW + c128] = y;
                if (BWD) a.lin_in[1][(size_t)(g0 + r) * 128 + c128] = y;
            }
        }
    }
    cta_sync();
    {
        float acc[1][2];
        dense_fwd<64, 1>(rg, 128, 64, y1, LDW, 1, acc);
        const int c = tid & 63, slot = tid >> 6;
        const float bm = __ldg(P + a.lb_off[1] + c);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int r = slot * 2 + q;
            if (r < ng) {
                const float y = fmaxf(acc[0][q] + bm, 0.f);
                y2[r * LDY2 + c] = y;
                if (BWD) a.lin_in[2][(size_t)(g0 + r) * 64 + c] = y;
            }
        }
    }
    cta_sync();
    {
        float acc[1][8];
        dense_fwd<256, 1>(rg, 64, A, y2, LDY2, 1, acc);
        if (tid < A) {
            const float bm = __ldg(P + a.lb_off[2] + tid);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < ng) y3[q * LDY3 + tid] = acc[0][q] + bm;
        }
    }
    cta_sync();
    STG_TRACE(a.trace, 96, 5);   // MLP
    if (BWD && a.mode != 0 && a.sync) {
        // Q_other is being computed by the other net's forward on another stream: wait for its post, then fetch
        STG_GT(a.trace, 12);
        if (tid == 0) {
            unsigned spins = 0;
            while ((int)(ld_acquire_gpu(a.sync) - expected) < 0) {
                __nanosleep(100);
                if (++spins > (1u << 24)) __trap();      // a post that never comes must fail loudly, not hang the device
            }
        }
        cta_sync();
        STG_GT(a.trace, 13);
        fetch_loss_inputs(true);
        cta_sync();
    }
    // ---- softmax, argmax (first maximum) ----
    for (int gi = warp; gi < ng; gi += NTH / 32) {
        float *y = y3 + gi * LDY3;
        const int g = g0 + gi;
        if (a.softmax) {
            float m = -INFINITY;
            for (int c = lane; c < A; c += 32) m = fmaxf(m, y[c]);
#pragma unroll
            for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
            float s = 0.f;
            for (int c = lane; c < A; c += 32) {
                const float e = expf(y[c] - m);
                y[c] = e;
                s += e;
            }
            s = warp_sum(s);
            for (int c = lane; c < A; c += 32) y[c] = y[c] / s;
        }
        __syncwarp();
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < A; c += 32) {
            const float v = y[c];
            if (a.out) a.out[(size_t)g * A + c] = v;
            if (v > bv) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (a.amax_out && lane == 0) a.amax_out[g] = bi;
        if (!BWD) continue;
        // ---- loss gradient w.r.t. the logits, in place (airfoil_dqn.py:264-304 for modes 1 / 2) ----
        if (a.mode == 0) {
            const float *go = a.gout + (size_t)g * A;
            float dot = 0.f;
            if (a.softmax) {
                for (int c = lane; c < A; c += 32) dot = fmaf(__ldg(go + c), y[c], dot);
                dot = warp_sum(dot);
            }
            for (int c = lane; c < A; c += 32) {
                const float gc = __ldg(go + c);
                y[c] = a.softmax ? y[c] * (gc - dot) : gc;
            }
        } else {
            int sel = __float_as_int(lgf[gi * 4]);
            float pred = lgf[gi * 4 + 1], nsv = lgf[gi * 4 + 2];
            const float rew = lgf[gi * 4 + 3];
            if (a.mode == 1) {
                pred = y[sel];
                if (lane == 0) a.rp_scalar[g] = pred;
            } else {
                sel = bi;
                nsv = bv;
                if (lane == 0) a.rp_scalar[g] = bv;
            }
            const float d = pred - (nsv * a.rp_gamma + rew);
            const float hd = (fabsf(d) < 1.f) ? d : (d > 0.f ? 1.f : -1.f);
            if (lane == 0) a.lterm[g] = (fabsf(d) < 1.f) ? 0.5f * d * d : (fabsf(d) - 0.5f);
            const float gsel = (a.mode == 1) ? hd * a.rp_inv_batch : -hd * a.rp_gamma * a.rp_inv_batch;
            const float ysel = y[sel];
            __syncwarp();
            const float dot = a.softmax ? gsel * ysel : 0.f;
            for (int c = lane; c < A; c += 32) {
                const float gc = (c == sel) ? gsel : 0.f;
                y[c] = a.softmax ? y[c] * (gc - dot) : gc;
            }
        }
        __syncwarp();
        for (int c = lane; c < A; c += 32) a.lin_d[2][(size_t)g * A + c] = y[c];
    }
    cta_sync();
    STG_TRACE(a.trace, 96, 6);   // softmax / loss gradient
    if (BWD) {
        // ---- MLP backward on the transposed weight copies: d2 = (d3.W3^T) relu'(y2), d1 = (d2.W2^T) relu'(y1), dR = d1.W1^T ----
        {
            float acc[1][2];
            dense_fwd<64, 1>(rg, A, 64, y3, LDY3, 1, acc);
            const int c = tid & 63, slot = tid >> 6;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int r = slot * 2 + q;
                if (r < ng) {
                    const float d = y2[r * LDY2 + c] > 0.f ? acc[0][q] : 0.f;
                    y2[r * LDY2 + c] = d;
                    a.lin_d[1][(size_t)(g0 + r) * 64 + c] = d;
                }
            }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 8);    // d2
        {
            float acc[1][4];
            dense_fwd<128, 1>(rg, 64, 128, y2, LDY2, 1, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = half * 4 + q;
                if (r < ng) {
                    const float d = y1[r * LDW + c128] > 0.f ? acc[0][q] : 0.f;
                    y1[r * LDW + c128] = d;
                    a.lin_d[0][(size_t)(g0 + r) * 128 + c128] = d;
                }
            }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 9);    // d1
        {
            float acc[1][8];
            dense_fwd<256, 1>(rg, 128, 256, y1, LDW, 1, acc);
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < ng) {
                    Rs[q * LD2W + tid] = acc[0][q];
                    a.dR[(size_t)(g0 + q) * 256 + tid] = acc[0][q];
                }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 10);   // dR
        // ---- block 3 backward: one row per graph, warp per graph ----
        {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(P + a.p5_off) + lane);
            const float wn = sqrtf(warp_sum(dot4(w, w)));
            for (int gi = warp; gi < ng; gi += NTH / 32) {
                const size_t go = (size_t)(g0 + gi) * 128 + 4 * lane;
                const float4 dmx = *reinterpret_cast<const float4 *>(Rs + gi * LD2W + 4 * lane);
                const float4 dmn = *reinterpret_cast<const float4 *>(Rs + gi * LD2W + 128 + 4 * lane);
                const float4 h = *reinterpret_cast<const float4 *>(H4 + gi * LDW + 4 * lane);
                float4 dpool = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 dp = pool_bwd_row(f4_add(dmn, dmx), h, s4[gi], z4[gi], w, wn, dpool);
                *reinterpret_cast<float4 *>(H4 + gi * LDW + 4 * lane) = dp;
                *reinterpret_cast<float4 *>(a.c5_d + go) = dp;
                *reinterpret_cast<float4 *>(a.bias5_d + go) = dp;
                *reinterpret_cast<float4 *>(a.pool5_d + go) = dpool;
            }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 11);   // pool5 backward
        {
            float acc[1][4];
            dense_fwd<128, 1>(rg, 128, 128, H4, LDW, 1, acc);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = half * 4 + q;
                if (r < ng) X3s[r * LDW + c128] = acc[0][q];
            }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 12);   // conv5^T
        // ---- block 2 backward: pooling through the one kept row, then A_hat^T ----
        {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(P + a.p4_off) + lane);
            const float wn = sqrtf(warp_sum(dot4(w, w)));
            for (int gi = warp; gi < ng; gi += NTH / 32) {
                const size_t go = (size_t)(g0 + gi) * 128 + 4 * lane;
                const float4 dmx = *reinterpret_cast<const float4 *>(Rs + gi * LD2W + 4 * lane);
                const float4 dmn = *reinterpret_cast<const float4 *>(Rs + gi * LD2W + 128 + 4 * lane);
                const float4 dxn = *reinterpret_cast<const float4 *>(X3s + gi * LDW + 4 * lane);
                const int r = sel3[gi], i = rb2[gi] + r;
                const float4 h = *reinterpret_cast<const float4 *>(H3 + (size_t)i * LDW + 4 * lane);
                float4 dpool = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 dp = pool_bwd_row(f4_add(f4_add(dxn, dmn), dmx), h, score3[i], z3[i], w, wn, dpool);
                *reinterpret_cast<float4 *>(a.bias4_d + go) = dp;
                *reinterpret_cast<float4 *>(a.pool4_d + go) = dpool;
                const unsigned short *el = es2 + gi * a.EC2;
                const float di = dis[i];
                for (int j = 0; j < R2; ++j) {
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j < n2s[gi]) {
                        if (j != r) {
                            const float cf = dis[rb2[gi] + j] * di;
                            for (int e = 0; e < ecnt[gi]; ++e) {
                                const unsigned ev = el[e];
                                if ((int)(ev >> 8) == r && (int)(ev & 255u) == j) {
                                    acc.x += cf * dp.x; acc.y += cf * dp.y; acc.z += cf * dp.z; acc.w += cf * dp.w;
                                }
                            }
                        } else {
                            const float dd = di * di;
                            acc.x += dd * dp.x; acc.y += dd * dp.y; acc.z += dd * dp.z; acc.w += dd * dp.w;
                        }
                        *reinterpret_cast<float4 *>(XW + (size_t)(rb2[gi] + j) * LDW + 4 * lane) = acc;
                    }
                    *reinterpret_cast<float4 *>(a.c4_d + ((size_t)(g0 + gi) * R2 + j) * 128 + 4 * lane) = acc;
                }
            }
        }
        cta_sync();
        STG_TRACE(a.trace, 96, 13);   // pool4 backward + A_hat^T
        {
            float acc[4][4];
            dense_fwd<128, 4>(rg, 128, 128, XW, LDW, (N2 + 7) >> 3, acc);
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = p * 8 + half * 4 + q;
                    if (r < N2) {
                        const int gi = rowg[r];
                        a.dX2[((size_t)(g0 + gi) * R2 + (r - rb2[gi])) * 128 + c128] = acc[p][q];
                    }
                }
        }
    }
    STG_TRACE(a.trace, 96, 7);   // end
    STG_GT(a.trace, BWD ? 11 : 9);
    if (BWD && a.mode != 0 && a.sync) {
        // the last CTA to finish counts this launch as completed; every CTA read sync[1] at its start, before any could finish last
        cta_sync();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(a.sync + 2, 1u) == gridDim.x - 1) {
                a.sync[2] = 0u;
                __threadfence();
                atomicAdd(a.sync + 1, 1u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// k_bwd1: block 1 backward (pooling, conv2^T, scatter through the level-1 edges) and block 0's pooling backward
// ------------------------------------------------------------------------------------------------
struct B1Args {
    long long *trace;
    const float *params, *wsplit;
    RBlk blk[8];
    int nblk, p1_off, p2_off;
    float ratio;
    const int *nptr;
    int B, GS, R1, EC1, R2;
    const unsigned short *e1;
    const int *e1n;
    const float *dX2, *dR;
    const float *h2k, *s2k, *z2k;
    const unsigned char *perm2, *amax2;
    const float *h1k, *s1k, *z1k;
    const unsigned char *amax1;
    float *c2_d, *pool2_d;         // weight-gradient deltas of conv2 [B*R2][128], pool2 [B][128]
    float *c1_d, *pool1_d;         // ... of conv1 [B*R1][128], pool1 [B][128]
    // loss = mean Huber over the `batch` transitions: the graphs' terms come from k_tail (lterm), mode 2 adds the terminal
    // transitions (no next-state graph); summed by CTA 0 in a fixed order.  mode 0: no loss.
    int mode, batch, A;
    const float *lterm, *rp_qother, *rp_reward;
    const int *rp_action, *next_slot;
    float *loss;
    int o_ring, o_dp2, o_dcat, o_dx1, o_h1, o_dr, o_am, o_es, o_int, o_f, total;
};

inline int b1_layout(B1Args &a)
{
    int o = 256;
    auto take = [&](int bytes) { int at = o; o += rup(bytes, 128); return at; };
    const int nk2 = rup(a.GS * a.R2, 8), nr1 = a.GS * a.R1;
    a.o_ring = take(RST * RSTAGE);
    a.o_dp2 = take(nk2 * LDW * 4);               // dX2 rows, then dP2
    a.o_dcat = take(nk2 * 2 * LDW * 4);                    // h2k rows + per-row pool terms first, then dcat
    a.o_dx1 = take(nr1 * LDW * 4);
    a.o_h1 = take(nr1 * LDW * 4);                // block 0's kept hidden rows
    a.o_dr = take(a.GS * 256 * 4);
    a.o_am = take(a.GS * 256);                   // amax2 | amax1
    a.o_es = take(a.GS * a.EC1 * 2);
    a.o_int = take((6 * a.GS + 4 + 2 * nk2) * 4);   // rb1[GS+1] ob2[GS+1] n1[GS] k2[GS] ecnt[GS] | deg[nk2] perm2[nk2]
    a.o_f = take((3 * nr1 + 2 * nk2) * 4);          // tds0, s1, z1 [nr1]; s2, z2 [nk2]
    a.total = o;
    return o;
}

__global__ void __launch_bounds__(NTH_TAIL, 1) k_bwd1(const __grid_constant__ B1Args a)
{
    extern __shared__ __align__(128) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int GS = a.GS, R1 = a.R1, R2 = a.R2;
    if (blockIdx.x * GS >= a.B) {
        // one extra CTA (launched when the loss is wanted): the batch loss from the per-transition Huber terms, beside the
        // working CTAs instead of at the end of CTA 0 (which made CTA 0 the last one to finish)
        if (tid >= NTH || a.mode == 0) return;
        float *red = reinterpret_cast<float *>(sm);
        float sum = 0.f;
        for (int g = tid; g < a.B; g += NTH) sum += a.lterm[g];
        if (a.mode == 2)
            for (int b = tid; b < a.batch; b += NTH)
                if (a.next_slot[b] < 0) {
                    const float d = a.rp_qother[(size_t)b * a.A + a.rp_action[b]] - a.rp_reward[b];
                    sum += (fabsf(d) < 1.f) ? 0.5f * d * d : (fabsf(d) - 0.5f);
                }
        red[tid] = sum;
        cta_sync();
        for (int o = NTH / 2; o; o >>= 1) {
            if (tid < o) red[tid] += red[tid + o];
            cta_sync();
        }
        if (tid == 0) *a.loss = red[0] / (float)a.batch;
        return;
    }
    const int g0 = blockIdx.x * GS, ng = min(GS, a.B - g0);
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    float *dP2 = reinterpret_cast<float *>(sm + a.o_dp2), *dcat = reinterpret_cast<float *>(sm + a.o_dcat);
    float *dX1 = reinterpret_cast<float *>(sm + a.o_dx1), *H1 = reinterpret_cast<float *>(sm + a.o_h1);
    float *dRs = reinterpret_cast<float *>(sm + a.o_dr);
    unsigned char *am2 = sm + a.o_am, *am1 = am2 + GS * 128;
    unsigned short *es1 = reinterpret_cast<unsigned short *>(sm + a.o_es);
    const int nk2m = rup(GS * R2, 8), nr1m = GS * R1;
    int *rb1 = reinterpret_cast<int *>(sm + a.o_int), *ob2 = rb1 + GS + 1, *n1s = ob2 + GS + 1, *k2s = n1s + GS, *ecnt = k2s + GS,
        *deg = ecnt + GS, *prm2 = deg + nk2m;
    float *tds0 = reinterpret_cast<float *>(sm + a.o_f), *s1 = tds0 + nr1m, *z1 = s1 + nr1m, *s2 = z1 + nr1m, *z2 = s2 + nk2m;
    float *H2 = dcat;                              // [nk2m][LDW] kept hidden rows of block 1 (dead before dcat is written)
    float *pt2 = dcat + (size_t)nk2m * LDW;        // [nk2m][LDW >= 128] per-row pool-weight terms (nk2m * 2 * LDW <= nk2m * LD2W)
    const float *P = a.params;

    Ring rg;
    rg.full = bars; rg.empty = bars + 8; rg.base = sm + a.o_ring; rg.params = P; rg.wsplit = a.wsplit; rg.blk = a.blk; rg.nblk = a.nblk;
    rg.j = 0; rg.nissued = 0;
    STG_TRACE(a.trace, 256, 0);
    STG_GT(a.trace, 14);
    if (tid == 0) rg.init();
    __syncthreads();               // all NTH_TAIL threads: the barriers exist
    if (tid >= NTH) {              // producer warp
        if (tid == NTH) rg.issue_upto((unsigned)rg.nblk);
        return;
    }
    if (tid == 0) {
        int r = 0, o2 = 0;
        for (int gi = 0; gi < ng; ++gi) {
            const int g = g0 + gi;
            const int n1 = topk_count_dev(a.ratio, a.nptr[g + 1] - a.nptr[g]);
            rb1[gi] = r; n1s[gi] = n1; ecnt[gi] = a.e1n[g];
            ob2[gi] = o2; k2s[gi] = topk_count_dev(a.ratio, n1);
            r += n1; o2 += k2s[gi];
        }
        rb1[ng] = r; ob2[ng] = o2;
    }
    for (int i = tid; i < nk2m * LDW; i += NTH) dP2[i] = 0.f;   // rows past NK2 are read (and discarded) by the GEMM
    cta_sync();
    const int NK2 = ob2[ng], NR1 = rb1[ng];
    // ---- everything this CTA needs from global memory, issued together (one latency, not one per use) ----
    for (int idx = tid; idx < NR1 * 32; idx += NTH) {     // block 0's kept hidden rows
        const int row = idx >> 5, c4 = idx & 31;
        int gi = 0;
        while (gi + 1 < ng && row >= rb1[gi + 1]) ++gi;
        *reinterpret_cast<float4 *>(H1 + (size_t)row * LDW + 4 * c4) =
            __ldg(reinterpret_cast<const float4 *>(a.h1k + ((size_t)(g0 + gi) * R1 + (row - rb1[gi])) * 128) + c4);
    }
    for (int idx = tid; idx < NK2 * 32; idx += NTH) {     // block 1's kept hidden rows and the incoming dX2 rows
        const int row = idx >> 5, c4 = idx & 31;
        int gi = 0;
        while (gi + 1 < ng && row >= ob2[gi + 1]) ++gi;
        const size_t ro = ((size_t)(g0 + gi) * R2 + (row - ob2[gi])) * 128;
        *reinterpret_cast<float4 *>(H2 + (size_t)row * LDW + 4 * c4) = __ldg(reinterpret_cast<const float4 *>(a.h2k + ro) + c4);
        *reinterpret_cast<float4 *>(dP2 + (size_t)row * LDW + 4 * c4) = __ldg(reinterpret_cast<const float4 *>(a.dX2 + ro) + c4);
    }
    for (int idx = tid; idx < ng * 64; idx += NTH)
        reinterpret_cast<float4 *>(dRs)[idx] = __ldg(reinterpret_cast<const float4 *>(a.dR + (size_t)g0 * 256) + idx);
    for (int idx = tid; idx < ng * 32; idx += NTH) {
        reinterpret_cast<unsigned *>(am2)[idx] = __ldg(reinterpret_cast<const unsigned *>(a.amax2 + (size_t)g0 * 128) + idx);
        reinterpret_cast<unsigned *>(am1)[idx] = __ldg(reinterpret_cast<const unsigned *>(a.amax1 + (size_t)g0 * 128) + idx);
    }
    for (int idx = tid; idx < NR1; idx += NTH) {
        int gi = 0;
        while (gi + 1 < ng && idx >= rb1[gi + 1]) ++gi;
        const size_t so = (size_t)(g0 + gi) * R1 + (idx - rb1[gi]);
        s1[idx] = a.s1k[so];
        z1[idx] = a.z1k[so];
    }
    for (int idx = tid; idx < NK2; idx += NTH) {
        int gi = 0;
        while (gi + 1 < ng && idx >= ob2[gi + 1]) ++gi;
        const size_t so = (size_t)(g0 + gi) * R2 + (idx - ob2[gi]);
        s2[idx] = a.s2k[so];
        z2[idx] = a.z2k[so];
        prm2[idx] = a.perm2[so];
    }
    for (int gi = 0; gi < ng; ++gi)
        for (int j = tid; j < ecnt[gi]; j += NTH) es1[gi * a.EC1 + j] = a.e1[(size_t)(g0 + gi) * a.EC1 + j];
    cta_sync();
    STG_TRACE(a.trace, 256, 1);  // inputs staged
    // ---- block 1 pooling backward: warp per kept row; the pool-weight terms are summed per graph in row order below ----
    {
        const float4 w = __ldg(reinterpret_cast<const float4 *>(P + a.p2_off) + lane);
        const float wn = sqrtf(warp_sum(dot4(w, w)));
        for (int row = warp; row < NK2; row += NTH / 32) {
            int gi = 0;
            while (gi + 1 < ng && row >= ob2[gi + 1]) ++gi;
            const int r = row - ob2[gi];
            const float fk = (float)k2s[gi];
            const float4 dmx = *reinterpret_cast<const float4 *>(dRs + gi * 256 + 4 * lane);
            const float4 dmn = *reinterpret_cast<const float4 *>(dRs + gi * 256 + 128 + 4 * lane);
            const uchar4 am = *reinterpret_cast<const uchar4 *>(am2 + gi * 128 + 4 * lane);
            float4 v = *reinterpret_cast<const float4 *>(dP2 + (size_t)row * LDW + 4 * lane);
            v.x += dmn.x / fk; v.y += dmn.y / fk; v.z += dmn.z / fk; v.w += dmn.w / fk;
            if (am.x == r) v.x += dmx.x;
            if (am.y == r) v.y += dmx.y;
            if (am.z == r) v.z += dmx.z;
            if (am.w == r) v.w += dmx.w;
            const float4 h = *reinterpret_cast<const float4 *>(H2 + (size_t)row * LDW + 4 * lane);
            float4 term = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 dp = pool_bwd_row(v, h, s2[row], z2[row], w, wn, term);
            *reinterpret_cast<float4 *>(dP2 + (size_t)row * LDW + 4 * lane) = dp;
            *reinterpret_cast<float4 *>(pt2 + (size_t)row * LDW + 4 * lane) = term;
        }
    }
    cta_sync();
    for (int idx = tid; idx < ng * R2 * 32; idx += NTH) {   // conv2's delta rows (zeros past the kept rows)
        const int gi = idx / (R2 * 32), rem = idx - gi * R2 * 32, r = rem >> 5, c4 = rem & 31;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < k2s[gi]) v = *reinterpret_cast<const float4 *>(dP2 + (size_t)(ob2[gi] + r) * LDW + 4 * c4);
        *reinterpret_cast<float4 *>(a.c2_d + ((size_t)(g0 + gi) * R2 + r) * 128 + 4 * c4) = v;
    }
    for (int idx = tid; idx < ng * 128; idx += NTH) {
        const int gi = idx >> 7, c = idx & 127;
        float acc = 0.f;
        for (int r = 0; r < k2s[gi]; ++r) acc += pt2[(size_t)(ob2[gi] + r) * LDW + c];
        a.pool2_d[(size_t)(g0 + gi) * 128 + c] = acc;
    }
    // in-degree of the kept rows' level-1 nodes (the mean's divisor)
    if (tid < NK2) {
        int gi = 0;
        while (gi + 1 < ng && tid >= ob2[gi + 1]) ++gi;
        const int i = prm2[tid];
        const unsigned short *el = es1 + gi * a.EC1;
        int d = 0;
        for (int e = 0; e < ecnt[gi]; ++e) d += ((int)(el[e] >> 8) == i);
        deg[tid] = d;
    }
    cta_sync();             // H2 / pt2 (aliasing dcat) are dead from here
    STG_TRACE(a.trace, 256, 2);  // block 1 pooling backward
    // ---- dcat = dP2 . W2^T on the transposed copy (rows = the 128 outputs of conv2, 256 columns = [mean | x] inputs) ----
    {
        float acc[2][8];
        dense_fwd<256, 2>(rg, 128, 256, dP2, LDW, (NK2 + 7) >> 3, acc);
#pragma unroll
        for (int p = 0; p < 2; ++p)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (p * 8 + q < NK2) dcat[(size_t)(p * 8 + q) * LD2W + tid] = acc[p][q];
    }
    cta_sync();
    STG_TRACE(a.trace, 256, 3);  // conv2^T
    // ---- dX1: the x half goes to the row itself, the mean half to its in-neighbours / in-degree; warp per level-1 row ----
    for (int row = warp; row < NR1; row += NTH / 32) {
        int gi = 0;
        while (gi + 1 < ng && row >= rb1[gi + 1]) ++gi;
        const int j = row - rb1[gi];
        const unsigned short *el = es1 + gi * a.EC1;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < k2s[gi]; ++r) {
            const int i = prm2[ob2[gi] + r];
            const float *dc = dcat + (size_t)(ob2[gi] + r) * LD2W;
            if (i == j) acc = f4_add(acc, *reinterpret_cast<const float4 *>(dc + 128 + 4 * lane));
            const int dg = deg[ob2[gi] + r];
            const float cnt = (float)(dg > 0 ? dg : 1);
            const float4 dm = *reinterpret_cast<const float4 *>(dc + 4 * lane);
            for (int e = 0; e < ecnt[gi]; ++e) {
                const unsigned ev = el[e];
                if ((int)(ev >> 8) == i && (int)(ev & 255u) == j) {
                    acc.x += dm.x / cnt; acc.y += dm.y / cnt; acc.z += dm.z / cnt; acc.w += dm.w / cnt;
                }
            }
        }
        *reinterpret_cast<float4 *>(dX1 + (size_t)row * LDW + 4 * lane) = acc;
    }
    cta_sync();
    STG_TRACE(a.trace, 256, 4);  // dX1
    // ---- block 0 pooling backward: warp per kept row; the pool-weight partial is summed per graph afterwards ----
    const float4 w1 = __ldg(reinterpret_cast<const float4 *>(P + a.p1_off) + lane);
    const float wn1 = sqrtf(warp_sum(dot4(w1, w1)));
    for (int row = warp; row < ng * R1; row += NTH / 32) {
        const int gi = row / R1, r = row - gi * R1, g = g0 + gi, k = n1s[gi];
        float4 dp = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < k) {
            const int lr = rb1[gi] + r;
            const float4 dmx = *reinterpret_cast<const float4 *>(dRs + gi * 256 + 4 * lane);
            const float4 dmn = *reinterpret_cast<const float4 *>(dRs + gi * 256 + 128 + 4 * lane);
            const uchar4 am = *reinterpret_cast<const uchar4 *>(am1 + gi * 128 + 4 * lane);
            const float fk = (float)k;
            float4 v = *reinterpret_cast<const float4 *>(dX1 + (size_t)lr * LDW + 4 * lane);
            v.x += dmn.x / fk; v.y += dmn.y / fk; v.z += dmn.z / fk; v.w += dmn.w / fk;
            if (am.x == r) v.x += dmx.x;
            if (am.y == r) v.y += dmx.y;
            if (am.z == r) v.z += dmx.z;
            if (am.w == r) v.w += dmx.w;
            const float4 h = *reinterpret_cast<const float4 *>(H1 + (size_t)lr * LDW + 4 * lane);
            const float s = s1[lr];
            const float ds = warp_sum(dot4(v, h));
            const float t = ds * (1.f - s * s);
            if (lane == 0) tds0[lr] = t;
            dp.x = h.x > 0.f ? v.x * s + t * w1.x / wn1 : 0.f;
            dp.y = h.y > 0.f ? v.y * s + t * w1.y / wn1 : 0.f;
            dp.z = h.z > 0.f ? v.z * s + t * w1.z / wn1 : 0.f;
            dp.w = h.w > 0.f ? v.w * s + t * w1.w / wn1 : 0.f;
        }
        *reinterpret_cast<float4 *>(a.c1_d + ((size_t)g * R1 + r) * 128 + 4 * lane) = dp;
    }
    cta_sync();
    for (int idx = tid; idx < ng * 128; idx += NTH) {
        const int gi = idx >> 7, c = idx & 127;
        const float wc = __ldg(P + a.p1_off + c);
        const float wn2 = wn1 * wn1;
        float acc = 0.f;
        for (int r = 0; r < n1s[gi]; ++r) {
            const int lr = rb1[gi] + r;
            acc += tds0[lr] * (H1[(size_t)lr * LDW + c] / wn1 - z1[lr] * wc / wn2);
        }
        a.pool1_d[(size_t)(g0 + gi) * 128 + c] = acc;
    }
    STG_TRACE(a.trace, 256, 5);  // end
    STG_GT(a.trace, 15);
}

}  // namespace stg
