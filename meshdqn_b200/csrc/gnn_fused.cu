// gnn_fused.cu -- fused Q-network kernels (sm_100a).
//
// Replaces the torch_geometric call sites of /root/reference/airfoilgcnn.py:85-145 (NodeRemovalNet.forward)
// and :170-209 (AirfoilGCNN.forward): SAGEConv / GCNConv message passing, TopKPooling, global max/mean
// readout, the 3-layer MLP, softmax and the argmax of airfoil_dqn.py:209 -- in ONE kernel launch, one CTA
// per `gpc` graphs, every intermediate in shared memory.  The backward kernel recomputes the forward per
// graph (nothing is saved between launches), back-propagates through the kept rows only (TopK makes every
// other row's gradient zero) and emits (delta, input) rows; a split-K weight-gradient pass reduces them in
// a fixed order, so gradients are deterministic and atomics-free.
//
// Message passing is a stable CSR-by-destination gather (edges keep their edge_index order inside a row,
// as torch_scatter's CPU loop does), mean / GCN-normalised, fp32, sequential per row.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>

#include "mdq_common.cuh"
#include "tc_prims.cuh"

namespace {

constexpr int NT = 256;  // 2 CTAs per SM (<= 113 KB shared memory, <= 128 registers each)
constexpr int NWARP = NT / 32;
constexpr int MAXB = MDQ_MAX_BLOCKS;
constexpr int MLP_SPLIT = 4;
constexpr unsigned FULL = 0xffffffffu;
using eid_t = unsigned short;  // CTA-local node/row id in shared memory
constexpr int MAX_STAGE = 8;
constexpr int STAGE_WORDS = 4096;  // 16 KB per stage
constexpr int MAXCH = 128;

struct QLay {
    int G, n_max, e_max, KC1, W, nb, bwd;
    int ncap[MAXB + 1];    // per-graph row cap entering block b; ncap[nb] = rows after the last pool
    int rowoff[MAXB + 1];  // b>=1: first row (in xbuf / per-row arrays) of block b's input rows
    int hoff[MAXB + 1];    // b>=1: first row in hbuf of block b's hidden rows
    int eoff[MAXB + 1];    // b>=1: offset of block b's edge list in e2s/e2d
    int xrows, e2cap, nrow_all, ymax;
    int o_w1, o_cat1, o_big, o_cat2, o_xbuf, o_hbuf, o_e1s, o_e1d, o_csr, o_e2s, o_e2d;
    int o_rowptr, o_cursor, o_score, o_z, o_newid, o_parent, o_dis, o_seg, o_ecnt, o_racc, o_y, o_part;
    int o_dx, o_dp, o_dcat, o_dr, o_amax, o_tds, o_h1k, o_c1k;
    // backward of a graph whose saved activations overflow shared memory (AirfoilGCNN, TopK 0.5, 180 nodes): the hidden
    // rows (hbuf, h1k) live in a per-CTA slice of the global workspace instead (sp_words floats, L2-resident)
    int spill, sp_h1k, sp_hbuf, sp_words;
    // weight stream: blocks >= 1 and the MLP read their weights from 16 KB shared-memory stages that one
    // thread fills with cp.async.bulk (TMA) in layer order, several chunks ahead of the consumers
    int nstage, o_stage[MAX_STAGE], o_mbar, nchunks;  // nstage == 0: weights are read straight from global memory
    int fused;  // width 128: block 0 computes its TopK scores in the GEMM epilogue and never stores the hidden rows
    int ck_blk[MAXB], ck_lin[3], ck_blin[3], ck_bblk[MAXB];
    int total;  // 4-byte words
};

struct WChunks {
    int off[MAXCH];              // float offset into the flat parameter buffer
    unsigned short rows[MAXCH];  // weight rows (k values) in the chunk
    unsigned short cols[MAXCH];  // row length (outputs)
};

struct WLayer {
    int K, C, rpg, w_off, b_off, d_off, i_off;
};
constexpr int MAXWL = 3 * MAXB + 3;
struct WDesc {
    int nl;
    int total;  // floats of workspace
    WLayer l[MAXWL];
    int tstart[MAXWL + 1];  // weight-gradient tasks of layer i are [tstart[i], tstart[i+1]) (filled at launch)
};

struct QArgs {
    mdq_net_t net;
    QLay L;
    const float *params;
    const float *x;
    const long long *esrc, *edst;  // int64 edge_index rows (PyG layout)
    const int *nptr, *eptr;
    int B;
    float *out, *emb;
    int *amax_out;
    const float *gout;
    float *ws;
    float *spill;            // backward: [B][L.sp_words] hidden rows when L.spill
    WDesc wd;
    WChunks ck;
    // fused replay gradient (rp_mode 0: grad_out given; 1: Huber through this net's Q(s)[a]; 2: through max_a Q(s'))
    int rp_mode;
    const int *rp_action;    // [B]   action of transition b
    const float *rp_reward;  // [B]
    const int *rp_index;     // mode 1: next_slot [B] (row of rp_qother or -1); mode 2: owner [n_next] (transition of row g)
    const float *rp_qother;  // mode 1: Q2(s') [n_next, A]; mode 2: Q1(s) [B, A]
    float rp_gamma, rp_inv_batch;
    float *rp_scalar;        // [n_graphs] out: mode 1 pred_b, mode 2 max_a Q(s'_g)
    long long *trace;  // optional: clock64() of CTA 0 / thread 0 at phase boundaries (profiling aid)
};

// ------------------------------------------------------------------------------------------------
// host: shared-memory layout
// ------------------------------------------------------------------------------------------------
int topk_count(float ratio, int n) { return (int)ceilf(ratio * (float)n); }

int build_layout_impl(const mdq_net_t &net, int max_n, int max_e, int G, int bwd, QLay &L, WChunks *ck, int spill)
{
    memset(&L, 0, sizeof(L));
    L.spill = spill;
    if (net.n_blocks < 1 || net.n_blocks > MAXB) return MDQ_EINVAL;
    if (net.width < 4 || net.width > 256 || (net.width & 3)) return MDQ_EINVAL;
    if (net.blk[0].type != MDQ_BLOCK_SAGE || net.blk[0].kin != net.in_dim) return MDQ_EINVAL;
    for (int b = 1; b < net.n_blocks; ++b)
        if (net.blk[b].kin != net.width) return MDQ_EINVAL;
    if (net.lin_in[0] != 2 * net.width || net.lin_in[1] != net.lin_out[0] || net.lin_in[2] != net.lin_out[1] ||
        net.lin_out[2] != net.out_dim)
        return MDQ_EINVAL;
    for (int i = 0; i < 3; ++i)
        if (net.lin_in[i] % MLP_SPLIT) return MDQ_EINVAL;
    if (max_n < 1 || max_e < 0 || G < 1) return MDQ_EINVAL;
    if (bwd && G != 1) return MDQ_EINVAL;
    const int W = net.width, nb = net.n_blocks;
    L.G = G; L.n_max = max_n; L.e_max = max_e > 0 ? max_e : 1; L.W = W; L.nb = nb; L.bwd = bwd;
    L.KC1 = mdq::pad4(2 * net.in_dim);
    L.ncap[0] = max_n;
    for (int b = 0; b < nb; ++b) L.ncap[b + 1] = topk_count(net.ratio, L.ncap[b]);
    const int gcap1 = G * L.ncap[1];
    if (bwd) {
        int r = 0;
        for (int b = 1; b <= nb; ++b) { L.rowoff[b] = r; L.hoff[b] = r; r += G * L.ncap[b]; }
        L.xrows = r;
        for (int b = 1; b < nb; ++b) L.eoff[b] = (b - 1) * G * L.e_max;
        L.e2cap = (nb > 1 ? nb - 1 : 1) * G * L.e_max;
    } else {
        for (int b = 1; b <= nb; ++b) { L.rowoff[b] = ((b - 1) & 1) * gcap1; L.hoff[b] = 0; }
        L.xrows = 2 * gcap1;
        for (int b = 1; b < nb; ++b) L.eoff[b] = ((b - 1) & 1) * G * L.e_max;
        L.e2cap = 2 * G * L.e_max;
    }
    L.nrow_all = max_n + L.xrows;
    int ymax = net.lin_out[0];
    if (net.lin_out[1] > ymax) ymax = net.lin_out[1];
    if (net.out_dim > ymax) ymax = net.out_dim;
    L.ymax = mdq::pad4(ymax);
    int o = 0;
    auto take = [&](int words) { int at = o; o += mdq::pad4(words); return at; };
    // the dense stages of blocks >= 1 give every thread one 4x4 output tile per pass; more rows than one pass
    // covers are handled by re-reading the weights straight from global memory (no staged stream)
    const bool multi_pass = ((gcap1 + 3) / 4) * (W / 4) > NT;
    const int dcat_a = G * L.ncap[nb > 1 ? 2 : 1] * 2 * W, dcat_c = gcap1 * W;
    const int dcat_words = dcat_a > dcat_c ? dcat_a : dcat_c;
    // scratch of blocks >= 1 (and of the backward pass): cat2 [+ dx, dp, dcat]
    int alias_words = mdq::pad4(gcap1 * 2 * W);
    if (bwd) alias_words += mdq::pad4(L.xrows * W) + mdq::pad4(gcap1 * W) + mdq::pad4(dcat_words);
    const int w1_words = mdq::pad4((L.KC1 + 1) * W), cat1_words = mdq::pad4(max_n * L.KC1);
    L.fused = (W == 128) ? 1 : 0;
    L.nstage = 0;
    int n_dedicated = 0;
    // The cp.async stage ring is kept for experiments (MDQ_QNET_STAGED=1).  Measured on B200 at 256 threads and
    // 2 CTAs/SM, reading the weights straight from L2 with 16-byte loads (8+ in flight per thread) is faster than
    // staging them (per-chunk barrier + LDGSTS issue cost ~800 cycles per 16 KB), so direct mode is the default.
    static const bool staged_env = [] { const char *e = getenv("MDQ_QNET_STAGED"); return e && e[0] == '1'; }();
    const bool want_stages = staged_env && !bwd && !multi_pass;
    if (L.fused) {
        // block 0 keeps no hidden rows: its staged weights and [agg|x] rows share one region with the scratch
        // of the later blocks (they are never live at the same time)
        const int l1_words = w1_words + cat1_words;
        const int region = l1_words > alias_words ? l1_words : alias_words;
        const int base = take(region);
        L.o_w1 = base;
        L.o_cat1 = base + w1_words;
        L.o_big = base;  // unused
        L.o_cat2 = base;
        if (want_stages) {  // forward: weight-stream stages = the region's tail + dedicated ones up to the 2-CTA/SM budget
            for (int at = base + alias_words; at + STAGE_WORDS <= base + region && L.nstage < MAX_STAGE; at += STAGE_WORDS)
                L.o_stage[L.nstage++] = at;
            n_dedicated = L.nstage >= 1 ? 1 : 2;
        }
    } else {
        L.o_w1 = take(w1_words);
        L.o_cat1 = take(cat1_words);
        // `big` holds block 0's hidden rows H1 [n][W]; afterwards the same words hold the later blocks' scratch
        const int big_words = mdq::pad4(max_n * W > alias_words ? max_n * W : alias_words);
        L.o_big = take(big_words);
        L.o_cat2 = L.o_big;
        if (want_stages) {
            if (w1_words >= STAGE_WORDS) L.o_stage[L.nstage++] = L.o_w1;
            if (cat1_words >= STAGE_WORDS) L.o_stage[L.nstage++] = L.o_cat1;
            for (int at = L.o_big + alias_words; at + STAGE_WORDS <= L.o_big + big_words && L.nstage < MAX_STAGE; at += STAGE_WORDS)
                L.o_stage[L.nstage++] = at;
            n_dedicated = L.nstage < 2 ? 2 - L.nstage : 0;
        }
    }
    if (bwd) {
        L.o_dx = L.o_cat2 + mdq::pad4(gcap1 * 2 * W);
        L.o_dp = L.o_dx + mdq::pad4(L.xrows * W);
        L.o_dcat = L.o_dp + mdq::pad4(gcap1 * W);
        L.sp_h1k = 0;
        L.sp_hbuf = mdq::pad4(gcap1 * W);
        L.sp_words = L.sp_hbuf + mdq::pad4(L.xrows * W);
        L.o_h1k = spill ? 0 : take(gcap1 * W);      // block 0's kept hidden rows
        L.o_c1k = take(gcap1 * L.KC1);  // ... and their input rows [agg | x]
    }
    for (int i = 0; i < n_dedicated && L.nstage < MAX_STAGE; ++i) L.o_stage[L.nstage++] = take(STAGE_WORDS);
    L.o_mbar = 0;
    L.o_xbuf = take(L.xrows * W);
    L.o_hbuf = (bwd && spill) ? 0 : take((bwd ? L.xrows : gcap1) * W);
    // edge endpoints / CSR columns are CTA-local row ids < 65536: stored as 16-bit, two per word
    if (max_n > 65535 || gcap1 > 65535) return MDQ_ESMEM;
    L.o_e1s = take((L.e_max + 1) / 2);
    L.o_e1d = take((L.e_max + 1) / 2);
    L.o_csr = take((G * L.e_max + 1) / 2);
    L.o_e2s = take((L.e2cap + 1) / 2);
    L.o_e2d = take((L.e2cap + 1) / 2);
    const int rmax = max_n > gcap1 ? max_n : gcap1;
    L.o_rowptr = take(rmax + 2);
    L.o_cursor = take(rmax + 2);
    L.o_score = take(L.nrow_all);
    L.o_z = take(L.nrow_all);
    L.o_newid = take(L.nrow_all);
    L.o_parent = take(L.nrow_all);
    L.o_dis = take(rmax);
    L.o_seg = take((nb + 1) * (G + 1));
    L.o_ecnt = take(nb + 1);
    L.o_racc = take(G * 2 * W);
    L.o_y = take(G * 3 * L.ymax);
    L.o_part = take(8 * G * L.ymax);
    if (bwd) {
        L.o_dr = take(2 * W + 3 * L.ymax);
        L.o_amax = take(nb * G * W);
        L.o_tds = take(gcap1 * 2);
    }
    L.total = o;
    // ---- weight chunks in consumption order ----
    int nck = 0;
    bool overflow = false;
    auto add_layer = [&](int w_off, int K, int C) {
        const int start = nck;
        int per = (STAGE_WORDS / C) & ~3;
        if (per < 4) { overflow = true; return start; }
        for (int k0 = 0; k0 < K; k0 += per) {
            if (nck >= MAXCH) { overflow = true; return start; }
            if (ck) {
                ck->off[nck] = w_off + k0 * C;
                ck->rows[nck] = (unsigned short)(K - k0 < per ? K - k0 : per);
                ck->cols[nck] = (unsigned short)C;
            }
            ++nck;
        }
        return start;
    };
    auto blk_k = [&](int b) { return net.blk[b].type == MDQ_BLOCK_SAGE ? 2 * W : W; };
    for (int b = 1; b < nb; ++b) L.ck_blk[b] = add_layer(net.blk[b].w_off, blk_k(b), W);
    for (int i = 0; i < 3; ++i) L.ck_lin[i] = add_layer(net.lin_off[i], net.lin_in[i], net.lin_out[i]);
    if (bwd) {
        for (int i = 2; i >= 0; --i) L.ck_blin[i] = add_layer(net.lin_off[i], net.lin_in[i], net.lin_out[i]);
        for (int b = nb - 1; b >= 1; --b) L.ck_bblk[b] = add_layer(net.blk[b].w_off, blk_k(b), W);
    }
    if (overflow) return MDQ_ESMEM;
    L.nchunks = nck;
    return MDQ_OK;
}

// Shared memory first; a backward layout that does not fit the 227 KB of one CTA moves its hidden rows to the workspace.
int build_layout(const mdq_net_t &net, int max_n, int max_e, int G, int bwd, QLay &L, WChunks *ck = nullptr)
{
    int rc = build_layout_impl(net, max_n, max_e, G, bwd, L, ck, 0);
    if (rc == MDQ_OK && bwd && (size_t)L.total * 4 > 227 * 1024) rc = build_layout_impl(net, max_n, max_e, G, bwd, L, ck, 1);
    return rc;
}

int build_wdesc(const mdq_net_t &net, int B, int max_n, WDesc &wd)
{
    memset(&wd, 0, sizeof(wd));
    int ncap[MAXB + 1];
    ncap[0] = max_n;
    for (int b = 0; b < net.n_blocks; ++b) ncap[b + 1] = topk_count(net.ratio, ncap[b]);
    int nl = 0, off = 0;
    auto add = [&](int K, int C, int rpg, int w_off, int b_off) {
        WLayer &l = wd.l[nl++];
        l.K = K; l.C = C; l.rpg = rpg; l.w_off = w_off; l.b_off = b_off;
        l.d_off = off; off += mdq::pad4(B * rpg * C);
        l.i_off = off; off += mdq::pad4(B * rpg * K);
    };
    const int W = net.width;
    for (int b = 0; b < net.n_blocks; ++b) {  // layers [0, nb): conv weights
        const mdq_block_t &k = net.blk[b];
        if (k.type == MDQ_BLOCK_SAGE) add(2 * k.kin, W, ncap[b + 1], k.w_off, k.b_off);
        else add(k.kin, W, ncap[b], k.w_off, -1);
    }
    for (int b = 0; b < net.n_blocks; ++b) add(0, W, 1, -1, net.blk[b].pool_off);  // [nb, 2nb): pool weights
    for (int i = 0; i < 3; ++i) add(net.lin_in[i], net.lin_out[i], 1, net.lin_off[i], net.lin_boff[i]);  // [2nb, 2nb+3)
    for (int b = 0; b < net.n_blocks; ++b)  // [2nb+3, 3nb+3): GCN bias (K=0), unused (rpg=0) for SAGE
        add(0, W, net.blk[b].type == MDQ_BLOCK_GCN ? 1 : 0, -1, net.blk[b].b_off);
    wd.nl = nl;
    wd.total = off;
    return MDQ_OK;
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ int topk_count_dev(float ratio, int n) { return __float2int_ru(__fmul_rn(ratio, (float)n)); }

// Stable CSR by destination: rowptr[n+1], csr[E] = source of each in-edge, in edge_index order per row.
// E <= 4*NT: every edge counts the earlier edges with its destination (its slot inside the row) in parallel;
// larger graphs fall back to one warp walking the edge list with match_any.
__device__ void build_csr(int n, int E, const eid_t *es, const eid_t *ed, int *rowptr, int *cursor, eid_t *csr)
{
    const int tid = threadIdx.x;
    for (int i = tid; i <= n; i += NT) cursor[i] = 0;
    __syncthreads();
    const bool par = E <= 4 * NT;
    int myrank[4] = {0, 0, 0, 0};
    if (par) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int e = tid + it * NT;
            if (e < E) {
                const int d = ed[e];
                int r = 0;
                // eight 16-bit destinations per 16-byte load (ed is 16-byte aligned): xor with d replicated, count
                // the zero halves -- ~1.5 instructions per earlier edge instead of a dependent LDS per edge
                const unsigned dd = (unsigned)d * 0x10001u;
                const uint4 *ed4 = reinterpret_cast<const uint4 *>(ed);
                const int full = ((reinterpret_cast<uintptr_t>(ed) & 15) == 0) ? (e >> 3) : 0;   // later blocks' lists may start unaligned
                for (int q = 0; q < full; ++q) {
                    const uint4 v = ed4[q];
                    const unsigned x0 = v.x ^ dd, x1 = v.y ^ dd, x2 = v.z ^ dd, x3 = v.w ^ dd;
                    r += !(x0 & 0xffffu) + !(x0 >> 16) + !(x1 & 0xffffu) + !(x1 >> 16) +
                         !(x2 & 0xffffu) + !(x2 >> 16) + !(x3 & 0xffffu) + !(x3 >> 16);
                }
                for (int e2 = full << 3; e2 < e; ++e2) r += (ed[e2] == d);
                myrank[it] = r;
                atomicAdd(&cursor[d], 1);
            }
        }
    } else {
        for (int e = tid; e < E; e += NT) atomicAdd(&cursor[ed[e]], 1);
    }
    __syncthreads();
    if (tid < 32) {
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int i = base + tid;
            const int v = (i < n) ? cursor[i] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (tid >= o) incl += t;
            }
            if (i < n) rowptr[i] = carry + incl - v;
            carry += __shfl_sync(FULL, incl, 31);
        }
        if (tid == 0) rowptr[n] = carry;
        if (!par) {
            __syncwarp();
            for (int i = tid; i < n; i += 32) cursor[i] = rowptr[i];
            __syncwarp();
            for (int base = 0; base < E; base += 32) {
                const int e = base + tid;
                const bool valid = e < E;
                const int d = valid ? (int)ed[e] : (-1 - tid);
                const unsigned m = __match_any_sync(FULL, d);
                const int rank = __popc(m & ((1u << tid) - 1u));
                const int pos = valid ? cursor[d] + rank : 0;
                __syncwarp();
                if (valid && (31 - __clz(m)) == tid) cursor[d] += __popc(m);
                __syncwarp();
                if (valid) csr[pos] = es[e];
            }
        }
    }
    __syncthreads();
    if (par) {
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int e = tid + it * NT;
            if (e < E) csr[rowptr[ed[e]] + myrank[it]] = es[e];
        }
        __syncthreads();
    }
}

// out[r][c] = act(bias[c] + sum_k A[r][k] * WT[k][c]); 4x4 register tile, k ascending.
template <bool W_SMEM>
__device__ void dense_rows(int n, int K, const float *A, int lda, const float *WT, const float *bias, float *out,
                           int ldo, int W, bool relu, const int *rowidx = nullptr)
{
    const int q = W >> 2;
    const int ngroups = (n + 3) >> 2;
    for (int item = threadIdx.x; item < ngroups * q; item += NT) {
        const int rg = item / q;
        const int c4 = (item - rg * q) << 2;
        const int r0 = rg << 2;
        const float *a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = min(r0 + j, n - 1);
            a[j] = A + (size_t)(rowidx ? rowidx[rr] : rr) * lda;
        }
        float acc[4][4];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) b4 = W_SMEM ? *reinterpret_cast<const float4 *>(bias + c4)
                              : __ldg(reinterpret_cast<const float4 *>(bias + c4));
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w; }
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            float4 w[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                w[kk] = W_SMEM ? *reinterpret_cast<const float4 *>(WT + (size_t)(k + kk) * W + c4)
                               : __ldg(reinterpret_cast<const float4 *>(WT + (size_t)(k + kk) * W + c4));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 xv = *reinterpret_cast<const float4 *>(a[j] + k);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[j][0] = fmaf(xs[kk], w[kk].x, acc[j][0]);
                    acc[j][1] = fmaf(xs[kk], w[kk].y, acc[j][1]);
                    acc[j][2] = fmaf(xs[kk], w[kk].z, acc[j][2]);
                    acc[j][3] = fmaf(xs[kk], w[kk].w, acc[j][3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (r0 + j < n) {
                float4 o4 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                if (relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
                *reinterpret_cast<float4 *>(out + (size_t)(r0 + j) * ldo + c4) = o4;
            }
        }
    }
}

// Block 0 with W == 128: h = relu([agg|x] W + b) is never stored; each warp owns a 4-row group (32 lanes x 4
// channels), folds h with the pooling weights and reduces over the warp:  z = (h . w) / ||w||, score = tanh(z).
__device__ void dense1_score(int n, int K, const float *A, int lda, const float *WT, const float *bias,
                             const float *__restrict__ pw_g, float *score, float *zval)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c4 = lane << 2;
    const float4 pw = __ldg(reinterpret_cast<const float4 *>(pw_g + c4));
    const float wnorm = sqrtf(warp_sum(fmaf(pw.x, pw.x, fmaf(pw.y, pw.y, fmaf(pw.z, pw.z, pw.w * pw.w)))));
    const float4 b4 = *reinterpret_cast<const float4 *>(bias + c4);
    const int ngroups = (n + 3) >> 2;
    for (int rg = warp; rg < ngroups; rg += NWARP) {
        const int r0 = rg << 2;
        const float *a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = A + (size_t)min(r0 + j, n - 1) * lda;
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w; }
#pragma unroll 3
        for (int k = 0; k < K; k += 4) {
            float4 w[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) w[kk] = *reinterpret_cast<const float4 *>(WT + (size_t)(k + kk) * 128 + c4);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 xv = *reinterpret_cast<const float4 *>(a[j] + k);
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    acc[j][0] = fmaf(xs[kk], w[kk].x, acc[j][0]);
                    acc[j][1] = fmaf(xs[kk], w[kk].y, acc[j][1]);
                    acc[j][2] = fmaf(xs[kk], w[kk].z, acc[j][2]);
                    acc[j][3] = fmaf(xs[kk], w[kk].w, acc[j][3]);
                }
            }
        }
        float p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            p[j] = fmaxf(acc[j][0], 0.f) * pw.x;
            p[j] = fmaf(fmaxf(acc[j][1], 0.f), pw.y, p[j]);
            p[j] = fmaf(fmaxf(acc[j][2], 0.f), pw.z, p[j]);
            p[j] = fmaf(fmaxf(acc[j][3], 0.f), pw.w, p[j]);
            p[j] = warp_sum(p[j]);
        }
        if (lane < 4 && r0 + lane < n) {
            const float dot = lane == 0 ? p[0] : (lane == 1 ? p[1] : (lane == 2 ? p[2] : p[3]));
            const float z = dot / wnorm;
            zval[r0 + lane] = z;
            score[r0 + lane] = tanhf(z);
        }
    }
}

// ---- weight stream: cp.async (LDGSTS, 16 B per thread) global -> shared, ns-1 chunks in flight ----
// (A single-thread cp.async.bulk / TMA version measured ~12 B/clk per SM here -- latency-bound on its few
// outstanding requests -- so every thread issues its own 16-byte asynchronous copies instead.)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct WStream {
    const float *params;
    float *smem;
    const int *o_stage;
    const WChunks *ck;
    int ns, nchunks;
    long long *ftrace;  // optional fine trace (CTA 0, thread 0): 3 stamps per chunk from slot 128 on

    __device__ __forceinline__ void stamp(int j, int which) const
    {
        if (ftrace && blockIdx.x == 0 && threadIdx.x == 0 && j < 100) ftrace[128 + 3 * j + which] = clock64();
    }
    __device__ __forceinline__ int rows(int j) const { return ck->rows[j]; }
    // every thread copies its 16-byte units of chunk j and commits one group (empty past the last chunk)
    __device__ __forceinline__ void issue(int j) const
    {
        if (ns == 0) return;
        if (j < nchunks) {
            const int units = (ck->rows[j] * ck->cols[j]) >> 2;
            const float *src = params + ck->off[j];
            float *dst = smem + o_stage[j % ns];
            for (int i = threadIdx.x; i < units; i += NT) cp_async16(dst + 4 * i, src + 4 * i);
        }
        cp_async_commit();
    }
    __device__ __forceinline__ void start() const  // once the stage regions are free; all threads call
    {
        for (int j = 0; j < ns - 1; ++j) issue(j);
    }
    // chunk j is complete and visible to all threads on return; refills the stage chunk j-1 used
    __device__ __forceinline__ const float *wait(int j) const
    {
        stamp(j, 0);
        if (ns == 0) return params + ck->off[j];  // direct mode: the consumer reads global memory (L1/L2)
        switch (ns) {
            case 2: cp_async_wait<0>(); break;
            case 3: cp_async_wait<1>(); break;
            case 4: cp_async_wait<2>(); break;
            case 5: cp_async_wait<3>(); break;
            case 6: cp_async_wait<4>(); break;
            case 7: cp_async_wait<5>(); break;
            default: cp_async_wait<6>(); break;
        }
        __syncthreads();
        issue(j + ns - 1);
        stamp(j, 1);
        return smem + o_stage[j % ns];
    }
    __device__ __forceinline__ void release(int j) const { stamp(j, 2); }
};

// Direct-mode dense layer for the small pooled levels (18, 2, 1 rows per graph): out = act(bias + A . WT), WT [K][W]
// read straight from L2.  A TR x 4 register tile per thread over (row group, column quad) items, and the K range
// split over KS = NT / items thread slices so that all eight warps issue FMAs (one 4 x 4 tile per thread used 160
// threads for 18 rows and 32 for <= 4 rows: the phase was bound by the issue rate of the one or two busy schedulers).
// The slices' partial sums are added into `out` in slice order (fixed, so results are reproducible), bias first.
template <int TR>
__device__ void dense_direct(const float *__restrict__ WT, int n, int K, const float *A, int lda, const float *bias,
                             float *out, int ldo, int W, bool relu)
{
    const int q = W >> 2;
    const int items = ((n + TR - 1) / TR) * q;   // <= NT (caller)
    const int ks_raw = NT / items;
    const int KS = ks_raw >= 8 ? 8 : (ks_raw >= 4 ? 4 : (ks_raw >= 2 ? 2 : 1));
    const int slice = threadIdx.x / items, item = threadIdx.x - slice * items;
    const bool active = slice < KS;
    const int rg = active ? item / q : 0;
    const int c4 = active ? (item - rg * q) << 2 : 0;
    const int r0 = rg * TR;
    const int kper = ((K + 4 * KS - 1) / (4 * KS)) * 4;
    const int k0 = min(K, slice * kper), k1 = active ? min(K, k0 + kper) : k0;
    const float *a[TR];
#pragma unroll
    for (int j = 0; j < TR; ++j) a[j] = A + (size_t)min(r0 + j, n - 1) * lda;
    float acc[TR][4];
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias && active && slice == 0) b4 = __ldg(reinterpret_cast<const float4 *>(bias + c4));
#pragma unroll
    for (int j = 0; j < TR; ++j) { acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w; }
    const float *wp = WT + (size_t)k0 * W + c4;
#pragma unroll 2
    for (int k = k0; k < k1; k += 4, wp += 4 * (size_t)W) {
        float4 w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) w[t] = __ldg(reinterpret_cast<const float4 *>(wp + (size_t)t * W));
#pragma unroll
        for (int j = 0; j < TR; ++j) {
            const float4 xv = *reinterpret_cast<const float4 *>(a[j] + k);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                acc[j][0] = fmaf(xs[t], w[t].x, acc[j][0]);
                acc[j][1] = fmaf(xs[t], w[t].y, acc[j][1]);
                acc[j][2] = fmaf(xs[t], w[t].z, acc[j][2]);
                acc[j][3] = fmaf(xs[t], w[t].w, acc[j][3]);
            }
        }
    }
    for (int s = 0; s < KS; ++s) {
        if (active && slice == s) {
#pragma unroll
            for (int j = 0; j < TR; ++j) {
                if (r0 + j < n) {
                    float4 *o = reinterpret_cast<float4 *>(out + (size_t)(r0 + j) * ldo + c4);
                    float4 v = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    if (s > 0) {
                        const float4 p = *o;
                        v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
                    }
                    if (relu && s == KS - 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    *o = v;
                }
            }
        }
        __syncthreads();
    }
}

// out[r][c] = act(bias[c] + sum_k A[r][k] * WT[k][c]) with WT streamed through shared-memory stages.
// One 4x4 output tile per thread (host guarantees ceil(n/4) * W/4 <= NT); k ascending.
__device__ void dense_stream(const WStream &ws, int j0, int n, int K, const float *A, int lda, const float *bias, float *out,
                             int ldo, int W, bool relu)
{
    const int q = W >> 2;
    const int ngroups = (n + 3) >> 2;
    if (ws.ns == 0 && (K & 3) == 0) {   // direct mode: weights contiguous [K][W] in the parameter buffer
        const float *WT = ws.params + ws.ck->off[j0];
        const int g5 = (n + 4) / 5;
        if (g5 * q * 2 <= NT && ngroups * q * 2 > NT) { dense_direct<5>(WT, n, K, A, lda, bias, out, ldo, W, relu); return; }
        if (ngroups * q <= NT) { dense_direct<4>(WT, n, K, A, lda, bias, out, ldo, W, relu); return; }
    }
    const int gpp = NT / q;  // row groups per pass (a second pass only happens in direct mode, ws.ns == 0)
    for (int g0 = 0; g0 < ngroups; g0 += gpp) {
        const int item = threadIdx.x;
        const bool active = item < min(gpp, ngroups - g0) * q;
        const int rg = g0 + (active ? item / q : 0);
        const int c4 = active ? (item % q) << 2 : 0;
        const int r0 = rg << 2;
        const float *a[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = A + (size_t)min(r0 + j, n - 1) * lda;
        float acc[4][4];
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias && active) b4 = __ldg(reinterpret_cast<const float4 *>(bias + c4));
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[j][0] = b4.x; acc[j][1] = b4.y; acc[j][2] = b4.z; acc[j][3] = b4.w; }
        int k = 0;
        for (int jc = j0; k < K; ++jc) {
            const float *st = ws.wait(jc);
            const int rows = ws.rows(jc);
            if (active) {
#pragma unroll 2
                for (int kk = 0; kk < rows; kk += 4) {
                    float4 w[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) w[t] = *reinterpret_cast<const float4 *>(st + (size_t)(kk + t) * W + c4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 xv = *reinterpret_cast<const float4 *>(a[j] + k + kk);
                        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            acc[j][0] = fmaf(xs[t], w[t].x, acc[j][0]);
                            acc[j][1] = fmaf(xs[t], w[t].y, acc[j][1]);
                            acc[j][2] = fmaf(xs[t], w[t].z, acc[j][2]);
                            acc[j][3] = fmaf(xs[t], w[t].w, acc[j][3]);
                        }
                    }
                }
            }
            k += rows;
            ws.release(jc);
        }
        if (active) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (r0 + j < n) {
                    float4 o4 = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    if (relu) { o4.x = fmaxf(o4.x, 0.f); o4.y = fmaxf(o4.y, 0.f); o4.z = fmaxf(o4.z, 0.f); o4.w = fmaxf(o4.w, 0.f); }
                    *reinterpret_cast<float4 *>(out + (size_t)(r0 + j) * ldo + c4) = o4;
                }
            }
        }
    }
    __syncthreads();
}

// rows x K times WT[K][O] (+ bias): every chunk is split over SL k-slices, partial sums combined in slice order.
__device__ void mlp_stream(const WStream &ws, int j0, int rows, int K, int O, const float *in, int ldi,
                           const float *__restrict__ b, float *out, int ldo, bool relu, float *part)
{
    if (ws.ns == 0 && rows == 1 && (O & 3) == 0 && (O >> 2) <= NT) {
        // direct mode, one row: thread = (4 output columns, k-slice); 16-byte coalesced weight loads, 8 in flight
        const float *WT = ws.params + ws.ck->off[j0];
        const int Q = O >> 2;
        int SLd = NT / Q;
        SLd = SLd >= 8 ? 8 : (SLd >= 4 ? 4 : (SLd >= 2 ? 2 : 1));
        while (K % SLd) SLd >>= 1;
        const int per = K / SLd;
        if (threadIdx.x < Q * SLd) {
            const int sl = threadIdx.x / Q, qd = threadIdx.x - sl * Q;
            const float4 *wp = reinterpret_cast<const float4 *>(WT + (size_t)(sl * per) * O) + qd;
            const float *xi = in + sl * per;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
            for (int kk = 0; kk < per; ++kk) {
                const float4 w = __ldg(wp + (size_t)kk * Q);
                const float xv = xi[kk];
                acc.x = fmaf(xv, w.x, acc.x); acc.y = fmaf(xv, w.y, acc.y);
                acc.z = fmaf(xv, w.z, acc.z); acc.w = fmaf(xv, w.w, acc.w);
            }
            *reinterpret_cast<float4 *>(part + (size_t)sl * O + 4 * qd) = acc;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < O; c += NT) {
            float v = __ldg(b + c);
            for (int sl = 0; sl < SLd; ++sl) v += part[sl * O + c];
            out[c] = relu ? fmaxf(v, 0.f) : v;
        }
        __syncthreads();
        return;
    }
    const int RO = rows * O;
    const int fit = NT / RO;
    const int SL = fit >= 4 ? 4 : (fit >= 2 ? 2 : 1);
    const int total = SL * RO;  // host guarantees total <= 2 * NT
    float acc[2] = {0.f, 0.f};
    int s_[2], r_[2], c_[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
        const int item = threadIdx.x + it * NT;
        const int ii = item < total ? item : 0;
        s_[it] = ii / RO;
        const int rem = ii - s_[it] * RO;
        r_[it] = rem / O;
        c_[it] = rem - r_[it] * O;
    }
    int k = 0;
    const bool whole = ws.ns == 0 && (K % SL) == 0;   // direct mode: the weights are one contiguous [K][O] block, so the
    for (int jc = j0; k < K; ++jc) {                  // k loop runs once over all of K (clean 16-deep load batches)
        const float *st = ws.wait(jc);
        const int cr = whole ? K : ws.rows(jc);
        const int per = cr / SL;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            if (threadIdx.x + it * NT < total) {
                const float *xi = in + r_[it] * ldi + k + s_[it] * per;
                const float *wp = st + (size_t)(s_[it] * per) * O + c_[it];
                float av = acc[it];
#pragma unroll 16
                for (int kk = 0; kk < per; ++kk) av = fmaf(xi[kk], wp[(size_t)kk * O], av);
                acc[it] = av;
            }
        }
        k += cr;
        ws.release(jc);
    }
#pragma unroll
    for (int it = 0; it < 2; ++it)
        if (threadIdx.x + it * NT < total) part[threadIdx.x + it * NT] = acc[it];
    __syncthreads();
    for (int idx = threadIdx.x; idx < RO; idx += NT) {
        const int r = idx / O, c = idx - r * O;
        float v = __ldg(b + c);
        for (int sidx = 0; sidx < SL; ++sidx) v += part[sidx * RO + idx];
        out[r * ldo + c] = relu ? fmaxf(v, 0.f) : v;
    }
    __syncthreads();
}

// out[r][k] = gate(k) * sum_c D[r][c] * WT[k][c] for one or two delta rows: a warp per weight row, lanes along c (NCI
// 32-float pieces per row), RB rows in flight per warp, butterfly reduction per row.
template <int NCI, int RB, bool TWO>
__device__ __forceinline__ void matmul_t_rows(const float *__restrict__ WT, int K, int C, const float *D, int ldd, float *out,
                                              int ldo, const float *gate)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float d0[NCI], d1[NCI];
#pragma unroll
    for (int i = 0; i < NCI; ++i) {
        const int c = lane + 32 * i;
        d0[i] = c < C ? D[c] : 0.f;
        d1[i] = (TWO && c < C) ? D[ldd + c] : 0.f;
    }
    for (int kb = warp * RB; kb < K; kb += NWARP * RB) {
        float w[RB][NCI];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            const float *wrow = WT + (size_t)min(kb + j, K - 1) * C;
#pragma unroll
            for (int i = 0; i < NCI; ++i) {
                const int c = lane + 32 * i;
                w[j][i] = c < C ? __ldg(wrow + c) : 0.f;
            }
        }
        float a0[RB], a1[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            a0[j] = 0.f;
            a1[j] = 0.f;
#pragma unroll
            for (int i = 0; i < NCI; ++i) {
                a0[j] = fmaf(d0[i], w[j][i], a0[j]);
                if (TWO) a1[j] = fmaf(d1[i], w[j][i], a1[j]);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                a0[j] += __shfl_xor_sync(FULL, a0[j], o);
                if (TWO) a1[j] += __shfl_xor_sync(FULL, a1[j], o);
            }
        }
        float v0 = a0[0], v1 = a1[0];
#pragma unroll
        for (int j = 1; j < RB; ++j) {
            if (lane == j) { v0 = a0[j]; v1 = a1[j]; }
        }
        if (lane < RB && kb + lane < K) {
            const int k = kb + lane;
            const bool on = (gate == nullptr || gate[k] > 0.f);
            out[k] = on ? v0 : 0.f;
            if (TWO) out[ldo + k] = on ? v1 : 0.f;
        }
    }
}

// out[r][k] = gate(k) * sum_c D[r][c] * WT[k][c]: input gradients; one warp per (r, k), lanes over c.
__device__ void matmul_t_stream(const WStream &ws, int j0, int K, int C, int nr, const float *D, int ldd, float *out, int ldo,
                                const float *gate)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (ws.ns == 0 && nr <= 2 && C <= 256) {
        // direct mode, MLP-sized (one or two delta rows): a warp per weight row k, lanes along c -- every load is a
        // coalesced piece of the row (a thread per row touched 32 different lines per request and was bound by the L1
        // tag rate; the generic path below issued its loads one dependent round trip at a time)
        const float *WT = ws.params + ws.ck->off[j0];
        const int nci = (C + 31) >> 5;
        const bool two = nr == 2;
        if (nci <= 2) { if (two) matmul_t_rows<2, 8, true>(WT, K, C, D, ldd, out, ldo, gate); else matmul_t_rows<2, 8, false>(WT, K, C, D, ldd, out, ldo, gate); }
        else if (nci <= 4) { if (two) matmul_t_rows<4, 8, true>(WT, K, C, D, ldd, out, ldo, gate); else matmul_t_rows<4, 8, false>(WT, K, C, D, ldd, out, ldo, gate); }
        else if (nci <= 6) { if (two) matmul_t_rows<6, 4, true>(WT, K, C, D, ldd, out, ldo, gate); else matmul_t_rows<6, 4, false>(WT, K, C, D, ldd, out, ldo, gate); }
        else { if (two) matmul_t_rows<8, 4, true>(WT, K, C, D, ldd, out, ldo, gate); else matmul_t_rows<8, 4, false>(WT, K, C, D, ldd, out, ldo, gate); }
        __syncthreads();
        return;
    }
    if (ws.ns == 0 && (C & 3) == 0) {
        // direct mode: one thread per weight row k (rows are contiguous in memory, 16-byte loads, many in flight),
        // delta rows broadcast from shared memory; two delta rows per pass
        const float *WT = ws.params + ws.ck->off[j0];
        for (int r0 = 0; r0 < nr; r0 += 2) {
            const bool two = r0 + 1 < nr;
            const float *d0 = D + (size_t)r0 * ldd, *d1 = D + (size_t)(two ? r0 + 1 : r0) * ldd;
            for (int k = threadIdx.x; k < K; k += NT) {
                const float4 *wrow = reinterpret_cast<const float4 *>(WT + (size_t)k * C);
                float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
                for (int c4 = 0; c4 < (C >> 2); ++c4) {
                    const float4 w = __ldg(wrow + c4);
                    const float4 x0 = *reinterpret_cast<const float4 *>(d0 + 4 * c4);
                    const float4 x1 = *reinterpret_cast<const float4 *>(d1 + 4 * c4);
                    a0 = fmaf(x0.x, w.x, a0); a0 = fmaf(x0.y, w.y, a0); a0 = fmaf(x0.z, w.z, a0); a0 = fmaf(x0.w, w.w, a0);
                    a1 = fmaf(x1.x, w.x, a1); a1 = fmaf(x1.y, w.y, a1); a1 = fmaf(x1.z, w.z, a1); a1 = fmaf(x1.w, w.w, a1);
                }
                const bool on = (gate == nullptr || gate[k] > 0.f);
                out[(size_t)r0 * ldo + k] = on ? a0 : 0.f;
                if (two) out[(size_t)(r0 + 1) * ldo + k] = on ? a1 : 0.f;
            }
        }
        __syncthreads();
        return;
    }
    int k = 0;
    for (int jc = j0; k < K; ++jc) {
        const float *st = ws.wait(jc);
        const int rows = ws.rows(jc);
        for (int t = warp; t < nr * rows; t += NWARP) {
            const int r = t / rows, kk = t - r * rows;
            float acc = 0.f;
            for (int c = lane; c < C; c += 32) acc = fmaf(D[r * ldd + c], st[(size_t)kk * C + c], acc);
            acc = warp_sum(acc);
            if (lane == 0) out[r * ldo + k + kk] = (gate == nullptr || gate[k + kk] > 0.f) ? acc : 0.f;
        }
        k += rows;
        ws.release(jc);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// the fused kernel
// ------------------------------------------------------------------------------------------------
template <bool BWD>
__global__ void __launch_bounds__(NT, 2) qnet_kernel(const __grid_constant__ QArgs a)
{
    extern __shared__ __align__(16) float smem[];
    const QLay &L = a.L;
    const mdq_net_t &net = a.net;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int W = L.W, nb = L.nb, G = L.G, KC1 = L.KC1, F = net.in_dim;
    const int g0 = blockIdx.x * G;
    const int ng = min(G, a.B - g0);
    if (ng <= 0) return;

    float *w1s = smem + L.o_w1;
    float *cat1 = smem + L.o_cat1;
    float *big = smem + L.o_big;
    float *cat2 = smem + L.o_cat2;
    float *xbuf = smem + L.o_xbuf;
    float *spill = (BWD && L.spill) ? a.spill + (size_t)blockIdx.x * L.sp_words : nullptr;
    float *hbuf = spill ? spill + L.sp_hbuf : smem + L.o_hbuf;
    float *h1k = spill ? spill + L.sp_h1k : smem + L.o_h1k;
    eid_t *e1s = reinterpret_cast<eid_t *>(smem + L.o_e1s);
    eid_t *e1d = reinterpret_cast<eid_t *>(smem + L.o_e1d);
    eid_t *csr = reinterpret_cast<eid_t *>(smem + L.o_csr);
    eid_t *e2s = reinterpret_cast<eid_t *>(smem + L.o_e2s);
    eid_t *e2d = reinterpret_cast<eid_t *>(smem + L.o_e2d);
    int *rowptr = reinterpret_cast<int *>(smem + L.o_rowptr);
    int *cursor = reinterpret_cast<int *>(smem + L.o_cursor);
    float *score = smem + L.o_score;
    float *zval = smem + L.o_z;
    int *newid = reinterpret_cast<int *>(smem + L.o_newid);
    int *parent = reinterpret_cast<int *>(smem + L.o_parent);
    float *dis = smem + L.o_dis;
    int *seg = reinterpret_cast<int *>(smem + L.o_seg);  // [nb+1][G+1]
    int *ecnt = reinterpret_cast<int *>(smem + L.o_ecnt);
    float *racc = smem + L.o_racc;
    float *ybuf = smem + L.o_y;
    float *part = smem + L.o_part;
    int *amaxs = BWD ? reinterpret_cast<int *>(smem + L.o_amax) : nullptr;

    const float *P = a.params;
    int trace_n = 0;
#define MDQ_TRACE()                                                                              \
    do {                                                                                         \
        if (a.trace && blockIdx.x == 0 && tid == 0 && trace_n < 120) a.trace[trace_n] = clock64(); \
        ++trace_n;                                                                               \
    } while (0)
    MDQ_TRACE();  // 0: start

    // ---- stage conv1 weights (+bias row) in shared memory, zero padded to KC1 rows ----
    {
        const float *wg = P + net.blk[0].w_off;
        const int rows = 2 * F;
        for (int idx = tid; idx < KC1 * W; idx += NT) w1s[idx] = (idx < rows * W) ? __ldg(wg + idx) : 0.f;
        for (int c = tid; c < W; c += NT) w1s[KC1 * W + c] = __ldg(P + net.blk[0].b_off + c);
        if (tid <= G) seg[1 * (G + 1) + tid] = 0;
        if (tid <= nb) ecnt[tid] = 0;
    }
    // (no barrier here: the first graph's loads below are followed by one)
    WStream wst;
    wst.params = P; wst.smem = smem; wst.o_stage = L.o_stage;
    wst.ck = &a.ck; wst.ns = L.nstage; wst.nchunks = L.nchunks; wst.ftrace = a.trace;

    // pool weight norm helper: each warp recomputes it (W <= 256 -> <= 8 values per lane)
    auto pool_weights = [&](int b, float (&pw)[8], float &wnorm) {
        const float *wp = P + net.blk[b].pool_off;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = lane + 32 * j;
            pw[j] = (c < W) ? __ldg(wp + c) : 0.f;
            ss = fmaf(pw[j], pw[j], ss);
        }
        wnorm = sqrtf(warp_sum(ss));
    };

    // scores of `n` rows H[i][:] -> score/z at per-row index base+i
    auto compute_scores = [&](int b, int n, const float *H, int base) {
        float pw[8], wnorm;
        pool_weights(b, pw, wnorm);
        for (int i = warp; i < n; i += NWARP) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = lane + 32 * j;
                if (c < W) acc = fmaf(H[(size_t)i * W + c], pw[j], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                const float z = acc / wnorm;
                zval[base + i] = z;
                score[base + i] = tanhf(z);
            }
        }
    };

    // ordered compaction of kept edges (both endpoints kept), ids mapped through newid (+base of the row arrays)
    auto filter_edges = [&](int E, const eid_t *es, const eid_t *ed, int nbase, eid_t *os, eid_t *od, int *count) {
        if (warp == 0) {
            int cnt = *count;
            for (int base = 0; base < E; base += 32) {
                const int e = base + lane;
                const bool valid = e < E;
                const int s = valid ? newid[nbase + es[e]] : -1;
                const int d = valid ? newid[nbase + ed[e]] : -1;
                const bool keep = valid && s >= 0 && d >= 0;
                const unsigned m = __ballot_sync(FULL, keep);
                if (keep) {
                    const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                    os[pos] = (eid_t)s;
                    od[pos] = (eid_t)d;
                }
                cnt += __popc(m);
            }
            __syncwarp();
            if (lane == 0) *count = cnt;
        }
    };

    // ================================ block 0: one graph at a time ================================
    for (int gi = 0; gi < ng; ++gi) {
        const int g = g0 + gi;
        const int nb0 = a.nptr[g], n = a.nptr[g + 1] - nb0;
        const int eb0 = a.eptr[g], E = a.eptr[g + 1] - eb0;
        {
            const float *xg = a.x + (size_t)nb0 * net.x_stride + net.in_col0;
            for (int idx = tid; idx < n * F; idx += NT) {
                const int i = idx / F, f = idx - i * F;
                cat1[i * KC1 + F + f] = __ldg(xg + (size_t)i * net.x_stride + f);
            }
            const int padc = KC1 - 2 * F;
            for (int idx = tid; idx < n * padc; idx += NT) {
                const int i = idx / padc, f = idx - i * padc;
                cat1[i * KC1 + 2 * F + f] = 0.f;
            }
            for (int e = tid; e < E; e += NT) {
                e1s[e] = (eid_t)(a.esrc[eb0 + e] - nb0);
                e1d[e] = (eid_t)(a.edst[eb0 + e] - nb0);
            }
        }
        __syncthreads();
        MDQ_TRACE();  // L1: inputs loaded
        build_csr(n, E, e1s, e1d, rowptr, cursor, csr);
        MDQ_TRACE();  // L1: csr
        // mean aggregation, edge order per row; thread per (row, feature) so that all 32 lanes work (a warp per row
        // used 17 of them) and a thread's items are independent chains
        for (int idx = tid; idx < n * F; idx += NT) {
            const int i = idx / F, f = idx - i * F;
            const int s0 = rowptr[i], s1 = rowptr[i + 1];
            float sum = 0.f;
            for (int s = s0; s < s1; ++s) sum += cat1[csr[s] * KC1 + F + f];
            cat1[i * KC1 + f] = sum / (float)(s1 - s0 > 0 ? s1 - s0 : 1);
        }
        __syncthreads();
        MDQ_TRACE();  // L1: aggregated
        if (L.fused) {
            dense1_score(n, KC1, cat1, KC1, w1s, w1s + KC1 * W, P + net.blk[0].pool_off, score, zval);
            __syncthreads();
            MDQ_TRACE();  // L1: dense
        } else {
            dense_rows<true>(n, KC1, cat1, KC1, w1s, w1s + KC1 * W, big, W, W, true);
            __syncthreads();
            MDQ_TRACE();  // L1: dense
            compute_scores(0, n, big, 0);
            __syncthreads();
        }
        MDQ_TRACE();  // L1: scores
        const int k = topk_count_dev(net.ratio, n);
        const int obase = seg[1 * (G + 1) + gi];
        // rank = number of rows ahead of i in (score desc, index asc) order.  Thread per row, every thread walks all
        // scores: the reads are shared-memory broadcasts (same address across the warp) and independent, so the loop
        // is ~4 instructions per j at full ILP instead of a shuffle-reduction chain per row.
        for (int i = tid; i < n; i += NT) {
            const float si = score[i];
            int r = 0;
            int j = 0;
            for (; j + 4 <= n; j += 4) {                       // score[] is 16-byte aligned (shared-memory carve-up)
                const float4 s4 = *reinterpret_cast<const float4 *>(score + j);
                r += (s4.x > si) || (s4.x == si && j < i);
                r += (s4.y > si) || (s4.y == si && j + 1 < i);
                r += (s4.z > si) || (s4.z == si && j + 2 < i);
                r += (s4.w > si) || (s4.w == si && j + 3 < i);
            }
            for (; j < n; ++j) {
                const float sj = score[j];
                r += (sj > si) || (sj == si && j < i);
            }
            newid[i] = (r < k) ? (obase + r) : -1;
            if (r < k) parent[L.n_max + L.rowoff[1] + obase + r] = i;
        }
        if (tid == 0) seg[1 * (G + 1) + gi + 1] = obase + k;
        __syncthreads();
        MDQ_TRACE();  // L1: ranked
        {
            float *xo = xbuf + (size_t)(L.rowoff[1] + obase) * W;
            const int *par = parent + L.n_max + L.rowoff[1] + obase;
            if (L.fused) {
                // recompute the hidden rows of the k kept nodes only (same code path, bit-identical)
                float *hk = (BWD ? h1k : hbuf) + (size_t)obase * W;
                dense_rows<true>(k, KC1, cat1, KC1, w1s, w1s + KC1 * W, hk, W, W, true, par);
                __syncthreads();
                for (int idx = tid; idx < k * W; idx += NT) xo[idx] = hk[idx] * score[par[idx / W]];
            } else {
                for (int idx = tid; idx < k * W; idx += NT) {
                    const int r = idx / W, c = idx - r * W;
                    const int i = par[r];
                    const float h = big[(size_t)i * W + c];
                    xo[idx] = h * score[i];
                    if (BWD) h1k[(size_t)(obase + r) * W + c] = h;
                }
            }
            if (BWD)
                for (int idx = tid; idx < k * KC1; idx += NT) {
                    const int r = idx / KC1, j = idx - r * KC1;
                    (smem + L.o_c1k)[(size_t)(obase + r) * KC1 + j] = cat1[par[r] * KC1 + j];
                }
        }
        __syncthreads();
        for (int c = tid; c < W; c += NT) {  // readout: max / mean over the kept rows
            const float *xo = xbuf + (size_t)(L.rowoff[1] + obase) * W + c;
            float mx = -INFINITY, sum = 0.f;
            int am = 0;
            for (int r = 0; r < k; ++r) {
                const float v = xo[(size_t)r * W];
                if (v > mx) { mx = v; am = r; }
                sum += v;
            }
            racc[gi * 2 * W + c] = mx;
            racc[gi * 2 * W + W + c] = sum / (float)(k > 0 ? k : 1);
            if (BWD) amaxs[(0 * G + gi) * W + c] = am;
        }
        if (nb > 1) filter_edges(E, e1s, e1d, 0, e2s + L.eoff[1], e2d + L.eoff[1], &ecnt[1]);
        __syncthreads();
        MDQ_TRACE();  // L1: gathered, read out, edges filtered
    }

    // block 0 is done: conv1's staged weights, cat1 and the tail of `big` become weight-stream stages
    wst.start();

    // ================================ blocks >= 1: rows of all `ng` graphs together ================================
    for (int b = 1; b < nb; ++b) {
        const int *sg = seg + b * (G + 1);
        int *sgn = seg + (b + 1) * (G + 1);
        const int nrows = sg[ng];
        const int E = ecnt[b];
        const float *X = xbuf + (size_t)L.rowoff[b] * W;
        float *H = hbuf + (size_t)L.hoff[b] * W;
        const eid_t *es = e2s + L.eoff[b], *ed = e2d + L.eoff[b];
        const int rbase = L.n_max + L.rowoff[b];
        build_csr(nrows, E, es, ed, rowptr, cursor, csr);
        MDQ_TRACE();  // Bk: csr
        if (net.blk[b].type == MDQ_BLOCK_SAGE) {
            for (int idx = tid; idx < nrows * W; idx += NT) {
                const int i = idx / W, f = idx - i * W;
                const int s0 = rowptr[i], s1 = rowptr[i + 1];
                float sum = 0.f;
                for (int s = s0; s < s1; ++s) sum += X[(size_t)csr[s] * W + f];
                const int cnt = s1 - s0;
                cat2[(size_t)i * 2 * W + f] = sum / (float)(cnt > 0 ? cnt : 1);
                cat2[(size_t)i * 2 * W + W + f] = X[idx];
            }
            __syncthreads();
            dense_stream(wst, L.ck_blk[b], nrows, 2 * W, cat2, 2 * W, P + net.blk[b].b_off, H, W, W, true);
        } else {
            for (int i = tid; i < nrows; i += NT) {
                int deg = 1;
                for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) deg += (csr[s] != i);
                dis[i] = __fdiv_rn(1.f, __fsqrt_rn((float)deg));
            }
            __syncthreads();
            dense_stream(wst, L.ck_blk[b], nrows, W, X, W, nullptr, cat2, W, W, false);
            __syncthreads();
            const float *bias = P + net.blk[b].b_off;
            for (int idx = tid; idx < nrows * W; idx += NT) {
                const int i = idx / W, c = idx - i * W;
                const float di = dis[i];
                float sum = 0.f;
                for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) {
                    const int j = csr[s];
                    if (j != i) sum += (dis[j] * di) * cat2[(size_t)j * W + c];
                }
                sum += (di * di) * cat2[idx];
                sum += __ldg(bias + c);
                H[idx] = fmaxf(sum, 0.f);
            }
        }
        __syncthreads();
        MDQ_TRACE();  // Bk: conv done
        compute_scores(b, nrows, H, rbase);
        if (tid == 0) {
            int acc = 0;
            sgn[0] = 0;
            for (int gi = 0; gi < ng; ++gi) {
                acc += topk_count_dev(net.ratio, sg[gi + 1] - sg[gi]);
                sgn[gi + 1] = acc;
            }
        }
        __syncthreads();
        for (int i = tid; i < nrows; i += NT) {
            int gi = 0;
            while (gi + 1 < ng && i >= sg[gi + 1]) ++gi;
            const int r0 = sg[gi], r1 = sg[gi + 1];
            const int k = sgn[gi + 1] - sgn[gi];
            const float si = score[rbase + i];
            int r = 0;
            for (int j = r0; j < r1; ++j) {
                const float sj = score[rbase + j];
                r += (sj > si) || (sj == si && j < i);
            }
            newid[rbase + i] = (r < k) ? (sgn[gi] + r) : -1;
            if (r < k) parent[L.n_max + L.rowoff[b + 1] + sgn[gi] + r] = i;
        }
        __syncthreads();
        {
            const int nk = sgn[ng];
            float *xo = xbuf + (size_t)L.rowoff[b + 1] * W;
            const int *par = parent + L.n_max + L.rowoff[b + 1];
            for (int idx = tid; idx < nk * W; idx += NT) {
                const int r = idx / W, c = idx - r * W;
                const int i = par[r];
                xo[idx] = H[(size_t)i * W + c] * score[rbase + i];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < ng * W; idx += NT) {
            const int gi = idx / W, c = idx - gi * W;
            const int r0 = sgn[gi], r1 = sgn[gi + 1];
            const float *xo = xbuf + (size_t)L.rowoff[b + 1] * W + c;
            float mx = -INFINITY, sum = 0.f;
            int am = 0;
            for (int r = r0; r < r1; ++r) {
                const float v = xo[(size_t)r * W];
                if (v > mx) { mx = v; am = r - r0; }
                sum += v;
            }
            const int k = r1 - r0;
            racc[gi * 2 * W + c] += mx;
            racc[gi * 2 * W + W + c] += sum / (float)(k > 0 ? k : 1);
            if (BWD) amaxs[(b * G + gi) * W + c] = am;
        }
        if (b + 1 < nb) filter_edges(E, es, ed, rbase, e2s + L.eoff[b + 1], e2d + L.eoff[b + 1], &ecnt[b + 1]);
        __syncthreads();
        MDQ_TRACE();  // Bk: pooled, read out, edges filtered
    }

    // ================================ MLP + softmax + argmax ================================
    const int YM = L.ymax;
    float *y1 = ybuf, *y2 = ybuf + G * YM, *y3 = ybuf + 2 * G * YM;
    if (a.emb)
        for (int idx = tid; idx < ng * 2 * W; idx += NT) a.emb[(size_t)g0 * 2 * W + idx] = racc[idx];
    mlp_stream(wst, L.ck_lin[0], ng, net.lin_in[0], net.lin_out[0], racc, 2 * W, P + net.lin_boff[0], y1, YM, true, part);
    MDQ_TRACE();  // lin1
    mlp_stream(wst, L.ck_lin[1], ng, net.lin_in[1], net.lin_out[1], y1, YM, P + net.lin_boff[1], y2, YM, true, part);
    MDQ_TRACE();  // lin2
    mlp_stream(wst, L.ck_lin[2], ng, net.lin_in[2], net.lin_out[2], y2, YM, P + net.lin_boff[2], y3, YM, false, part);
    MDQ_TRACE();  // MLP done
    const int A = net.out_dim;
    for (int gi = warp; gi < ng; gi += NWARP) {
        float *y = y3 + gi * YM;
        if (net.softmax) {
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
            if (!BWD) a.out[(size_t)(g0 + gi) * A + c] = v;
            if (v > bv) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bv, o);
            const int oi = __shfl_xor_sync(FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (!BWD && a.amax_out && lane == 0) a.amax_out[g0 + gi] = bi;
    }
    MDQ_TRACE();  // softmax / argmax done
    if (!BWD) return;

    // =====================================================================================
    // backward (G == 1): rows of graph g0 only
    // =====================================================================================
    __syncthreads();
    const int g = g0;
    const WDesc &wd = a.wd;
    float *ws = a.ws;
    float *dx = smem + L.o_dx;
    float *dp = smem + L.o_dp;
    float *dcat = smem + L.o_dcat;
    float *dr = smem + L.o_dr;          // [2W]
    float *d1 = dr + 2 * W;             // [ymax] x3
    float *d2 = d1 + YM;
    float *d3 = d2 + YM;
    float *tds = smem + L.o_tds;        // [gcap1][2] : ds*(1-s^2), unused

    auto emit_row = [&](const WLayer &l, int r, const float *delta, const float *in, int in_n) {
        float *dd = ws + l.d_off + ((size_t)g * l.rpg + r) * l.C;
        for (int c = tid; c < l.C; c += NT) dd[c] = delta ? delta[c] : 0.f;
        if (l.K > 0) {
            float *ii = ws + l.i_off + ((size_t)g * l.rpg + r) * l.K;
            for (int k = tid; k < l.K; k += NT) ii[k] = (in && k < in_n) ? in[k] : 0.f;
        }
    };
    // all rpg rows of a layer in one flattened pass: rows >= nvalid are zero-filled
    auto emit_rows = [&](const WLayer &l, int nvalid, const float *D, int ldd, const float *IN, int ldin, int in_n) {
        float *dd = ws + l.d_off + (size_t)g * l.rpg * l.C;
        for (int idx = tid; idx < l.rpg * l.C; idx += NT) {
            const int r = idx / l.C, c = idx - r * l.C;
            dd[idx] = (r < nvalid) ? D[(size_t)r * ldd + c] : 0.f;
        }
        if (l.K > 0) {
            float *ii = ws + l.i_off + (size_t)g * l.rpg * l.K;
            for (int idx = tid; idx < l.rpg * l.K; idx += NT) {
                const int r = idx / l.K, k = idx - r * l.K;
                ii[idx] = (r < nvalid && k < in_n) ? IN[(size_t)r * ldin + k] : 0.f;
            }
        }
    };

    // ---- MLP backward ----
    {
        if (warp == 0 && a.rp_mode == 0) {
            const float *go = a.gout + (size_t)g * A;
            float dot = 0.f;
            if (net.softmax) {
                for (int c = lane; c < A; c += 32) dot = fmaf(__ldg(go + c), y3[c], dot);
                dot = warp_sum(dot);
            }
            for (int c = lane; c < A; c += 32) {
                const float gc = __ldg(go + c);
                d3[c] = net.softmax ? y3[c] * (gc - dot) : gc;
            }
        } else if (warp == 0) {
            // Huber(delta=1, mean) on pred = Q1(s)[a] vs target = r + gamma * max_a Q2(s'), airfoil_dqn.py:264-304,
            // evaluated here so the selected net's forward is not launched twice
            int sel;
            float pred, nsv, rew;
            if (a.rp_mode == 1) {
                sel = a.rp_action[g];
                pred = y3[sel];
                rew = a.rp_reward[g];
                const int slot = a.rp_index[g];
                float m = -INFINITY;
                if (slot >= 0) {
                    const float *q = a.rp_qother + (size_t)slot * A;
                    for (int c = lane; c < A; c += 32) m = fmaxf(m, __ldg(q + c));
#pragma unroll
                    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
                }
                nsv = slot >= 0 ? m : 0.f;
                if (lane == 0) a.rp_scalar[g] = pred;
            } else {
                const int b = a.rp_index[g];
                pred = __ldg(a.rp_qother + (size_t)b * A + a.rp_action[b]);
                rew = a.rp_reward[b];
                float bv = -INFINITY;
                int bi = 0x7fffffff;
                for (int c = lane; c < A; c += 32) {
                    const float v = y3[c];
                    if (v > bv) { bv = v; bi = c; }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) {
                    const float ov = __shfl_xor_sync(FULL, bv, o);
                    const int oi = __shfl_xor_sync(FULL, bi, o);
                    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                }
                sel = bi;
                nsv = bv;
                if (lane == 0) a.rp_scalar[g] = bv;
            }
            const float d = pred - (nsv * a.rp_gamma + rew);
            const float hd = (fabsf(d) < 1.f) ? d : (d > 0.f ? 1.f : -1.f);
            const float gsel = (a.rp_mode == 1) ? hd * a.rp_inv_batch : -hd * a.rp_gamma * a.rp_inv_batch;
            const float dot = net.softmax ? gsel * y3[sel] : 0.f;
            for (int c = lane; c < A; c += 32) {
                const float gc = (c == sel) ? gsel : 0.f;
                d3[c] = net.softmax ? y3[c] * (gc - dot) : gc;
            }
        }
        __syncthreads();
        MDQ_TRACE();  // backward: Huber / softmax gradient
        emit_row(wd.l[2 * nb + 2], 0, d3, y2, net.lin_in[2]);
        matmul_t_stream(wst, L.ck_blin[2], net.lin_in[2], net.lin_out[2], 1, d3, 0, d2, 0, y2);
        MDQ_TRACE();  // backward: lin3^T
        emit_row(wd.l[2 * nb + 1], 0, d2, y1, net.lin_in[1]);
        matmul_t_stream(wst, L.ck_blin[1], net.lin_in[1], net.lin_out[1], 1, d2, 0, d1, 0, y1);
        MDQ_TRACE();  // backward: lin2^T
        emit_row(wd.l[2 * nb + 0], 0, d1, racc, net.lin_in[0]);
        matmul_t_stream(wst, L.ck_blin[0], net.lin_in[0], net.lin_out[0], 1, d1, 0, dr, 0, nullptr);
    }

    MDQ_TRACE();  // backward: MLP done
    // ---- conv blocks, last to first ----
    for (int b = nb - 1; b >= 0; --b) {
        MDQ_TRACE();  // backward: block b starts
        const int n_b = (b == 0) ? (a.nptr[g + 1] - a.nptr[g]) : seg[b * (G + 1) + 1];
        const int k = seg[(b + 1) * (G + 1) + 1];  // kept rows (G == 1)
        // hidden rows of block b: block 0 keeps only its kept rows (compact, row r); blocks >= 1 keep all rows
        const float *H = (b == 0) ? h1k : hbuf + (size_t)L.hoff[b] * W;
        const bool compactH = (b == 0);
        const int rbase = (b == 0) ? 0 : L.n_max + L.rowoff[b];
        const int *par = parent + L.n_max + L.rowoff[b + 1];
        const float *dxn = dx + (size_t)L.rowoff[b + 1] * W;  // valid when b+1 < nb
        const int *am = amaxs + (b * G) * W;
        const WLayer &lw = wd.l[b];
        const bool sage = net.blk[b].type == MDQ_BLOCK_SAGE;
        // dXn -> dp (temporarily holds dXn), ds per kept row
        for (int idx = tid; idx < k * W; idx += NT) {
            const int r = idx / W, c = idx - r * W;
            float v = (b + 1 < nb) ? dxn[idx] : 0.f;
            v += dr[W + c] / (float)k;
            if (am[c] == r) v += dr[c];
            dp[idx] = v;
        }
        __syncthreads();
        {
            float pw[8], wnorm;
            pool_weights(b, pw, wnorm);
            for (int r = warp; r < k; r += NWARP) {
                const int i = par[r];
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = lane + 32 * j;
                    if (c < W) acc = fmaf(dp[r * W + c], H[(size_t)(compactH ? r : i) * W + c], acc);
                }
                acc = warp_sum(acc);
                const float s = score[rbase + i];
                if (lane == 0) tds[r] = acc * (1.f - s * s);
            }
            __syncthreads();
            // pool weight gradient (per-graph partial, sequential over kept rows) and dP
            const float *wp = P + net.blk[b].pool_off;
            float *dpool = ws + wd.l[nb + b].d_off + (size_t)g * W;
            for (int c = tid; c < W; c += NT) {
                const float wc = __ldg(wp + c);
                float accw = 0.f;
                for (int r = 0; r < k; ++r) {
                    const int i = par[r];
                    accw += tds[r] * (H[(size_t)(compactH ? r : i) * W + c] / wnorm - zval[rbase + i] * wc / (wnorm * wnorm));
                }
                dpool[c] = accw;
            }
            for (int idx = tid; idx < k * W; idx += NT) {
                const int r = idx / W, c = idx - r * W;
                const int i = par[r];
                const float h = H[(size_t)(compactH ? r : i) * W + c];
                const float dh = dp[idx] * score[rbase + i] + tds[r] * __ldg(wp + c) / wnorm;
                dp[idx] = (h > 0.f) ? dh : 0.f;
            }
        }
        __syncthreads();
        if (b == 0) {
            emit_rows(lw, k, dp, W, smem + L.o_c1k, KC1, 2 * F);
            break;
        }
        const float *X = xbuf + (size_t)L.rowoff[b] * W;
        const int E = ecnt[b];
        build_csr(n_b, E, e2s + L.eoff[b], e2d + L.eoff[b], rowptr, cursor, csr);
        float *dxb = dx + (size_t)L.rowoff[b] * W;
        if (sage) {
            // inputs of the kept rows: [mean agg | x] recomputed into cat2 rows 0..k-1
            for (int idx = tid; idx < k * W; idx += NT) {
                const int r = idx / W, f = idx - r * W;
                const int i = par[r];
                const int s0 = rowptr[i], s1 = rowptr[i + 1];
                float sum = 0.f;
                for (int s = s0; s < s1; ++s) sum += X[(size_t)csr[s] * W + f];
                const int cnt = s1 - s0;
                cat2[(size_t)r * 2 * W + f] = sum / (float)(cnt > 0 ? cnt : 1);
                cat2[(size_t)r * 2 * W + W + f] = X[(size_t)i * W + f];
            }
            __syncthreads();
            emit_rows(lw, k, dp, W, cat2, 2 * W, 2 * W);
            // dcat[r][kk] = sum_c dP[r][c] * WT[kk][c]
            matmul_t_stream(wst, L.ck_bblk[b], 2 * W, W, k, dp, W, dcat, 2 * W, nullptr);
            for (int idx = tid; idx < n_b * W; idx += NT) {
                const int j = idx / W, f = idx - j * W;
                float acc = 0.f;
                for (int r = 0; r < k; ++r) {
                    const int i = par[r];
                    if (i == j) acc += dcat[(size_t)r * 2 * W + W + f];
                    const int s0 = rowptr[i], s1 = rowptr[i + 1];
                    const float cnt = (float)(s1 - s0 > 0 ? s1 - s0 : 1);
                    for (int s = s0; s < s1; ++s)
                        if (csr[s] == j) acc += dcat[(size_t)r * 2 * W + f] / cnt;
                }
                dxb[idx] = acc;
            }
        } else {
            for (int i = tid; i < n_b; i += NT) {
                int deg = 1;
                for (int s = rowptr[i]; s < rowptr[i + 1]; ++s) deg += (csr[s] != i);
                dis[i] = __fdiv_rn(1.f, __fsqrt_rn((float)deg));
            }
            // GCN bias gradient: sum of dP over the kept rows
            {
                float *db = ws + wd.l[2 * nb + 3 + b].d_off + (size_t)g * W;
                for (int c = tid; c < W; c += NT) {
                    float acc = 0.f;
                    for (int r = 0; r < k; ++r) acc += dp[r * W + c];
                    db[c] = acc;
                }
            }
            __syncthreads();
            // dXW[j][c] into dcat [n_b][W]
            for (int idx = tid; idx < n_b * W; idx += NT) {
                const int j = idx / W, c = idx - j * W;
                float acc = 0.f;
                for (int r = 0; r < k; ++r) {
                    const int i = par[r];
                    const float di = dis[i];
                    for (int s = rowptr[i]; s < rowptr[i + 1]; ++s)
                        if (csr[s] == j && j != i) acc += (dis[j] * di) * dp[r * W + c];
                    if (i == j) acc += (di * di) * dp[r * W + c];
                }
                dcat[idx] = acc;
            }
            __syncthreads();
            emit_rows(lw, n_b, dcat, W, X, W, W);
            __syncthreads();
            matmul_t_stream(wst, L.ck_bblk[b], W, W, n_b, dcat, W, dxb, W, nullptr);
        }
        __syncthreads();
    }
    MDQ_TRACE();  // backward done
#undef MDQ_TRACE
}

// ------------------------------------------------------------------------------------------------
// weight gradient: partial[s][k][c] = sum_{rows in chunk s} in[row][k] * delta[row][c]   (k == K -> bias)
// ------------------------------------------------------------------------------------------------
constexpr int WG_ROWS = 32;   // rows per chunk: short dependent-load chains (the kernel is latency-, not FLOP-bound: ~0.1 GFLOP),
                              // the chunk partials are summed in fixed order by wgrad_reduce_kernel
constexpr int WG_KG = 8;      // k values per thread
constexpr int WG_SUB = 8;     // 32-row sub-chunks a task walks: 256 rows per partial (the 32-row tasks of round 1 wrote 6 MB of
                              // partials and re-read 8 MB; the accumulators simply stay in registers across sub-chunks)
constexpr int WG_TASK = WG_ROWS * WG_SUB;

// layer li: S chunks of WG_ROWS rows, nkg groups of WG_KG k-values; partial offset = sum over earlier layers
__device__ __forceinline__ void wg_layer_geom(const WDesc &wd, int B, int li, int &S, int &nkg, int &poff)
{
    int off = 0;
    for (int i = 0; i < li; ++i) {
        const int rows = B * wd.l[i].rpg;
        if (rows == 0) continue;
        off += ((rows + WG_TASK - 1) / WG_TASK) * ((wd.l[i].K + 1 + WG_KG - 1) / WG_KG) * WG_KG * wd.l[i].C;
    }
    const int rows = B * wd.l[li].rpg;
    S = (rows + WG_TASK - 1) / WG_TASK;
    nkg = (wd.l[li].K + 1 + WG_KG - 1) / WG_KG;
    poff = off;
}

__global__ void __launch_bounds__(256, 3) wgrad_partial_kernel(const WDesc wd, int B, const float *__restrict__ ws,
                                                            float *__restrict__ partial)
{
    __shared__ __align__(16) float xin[WG_ROWS][WG_KG];  // the chunk's input rows, these WG_KG k-values
    // one CTA per task, tasks of all layers flattened (a 2-D grid sized by the largest layer launched ~5x more CTAs
    // than there are tasks, and their scheduling alone cost more than the arithmetic)
    int li = 0;
    while (li + 1 < wd.nl && (int)blockIdx.x >= wd.tstart[li + 1]) ++li;
    const WLayer l = wd.l[li];
    if (l.rpg == 0) return;
    int S, nkg, poff;
    wg_layer_geom(wd, B, li, S, nkg, poff);
    const int KB = l.K + 1;  // last "k" is the bias column (input == 1)
    {
        const int t = (int)blockIdx.x - wd.tstart[li];
        if (t >= S * nkg) return;
        const int kg = t / S, chunk = t - kg * S;
        const int rows = B * l.rpg;
        const int k0 = kg * WG_KG;
        float *po = partial + poff + (size_t)t * WG_KG * l.C;
        // C <= 256 in every layer: one column per thread, its WG_KG accumulators live across the task's sub-chunks
        const int c = threadIdx.x;
        float acc[WG_KG];
#pragma unroll
        for (int j = 0; j < WG_KG; ++j) acc[j] = 0.f;
        for (int sub = 0; sub < WG_SUB; ++sub) {
            const int r0 = chunk * WG_TASK + sub * WG_ROWS, r1 = min(rows, r0 + WG_ROWS);
            if (r0 >= rows) break;
            __syncthreads();
            for (int idx = threadIdx.x; idx < WG_ROWS * WG_KG; idx += blockDim.x) {
                const int rr = idx / WG_KG, j = idx - rr * WG_KG;
                const int r = r0 + rr, k = k0 + j;
                float v = 0.f;
                if (r < r1) v = (k < l.K) ? __ldg(ws + l.i_off + (size_t)r * l.K + k) : (k == l.K ? 1.f : 0.f);
                xin[rr][j] = v;
            }
            __syncthreads();
            if (c < l.C) {
                const float *dl = ws + l.d_off + c;
                const int nr = r1 - r0;
                // all WG_ROWS delta loads of the sub-chunk are issued together; rows past its end contribute d = 0
                float dv[WG_ROWS];
#pragma unroll
                for (int rr = 0; rr < WG_ROWS; ++rr) dv[rr] = rr < nr ? __ldg(dl + (size_t)(r0 + rr) * l.C) : 0.f;
#pragma unroll
                for (int rr = 0; rr < WG_ROWS; ++rr) {
                    const float d = dv[rr];
                    const float4 xa = *reinterpret_cast<const float4 *>(&xin[rr][0]);
                    const float4 xb = *reinterpret_cast<const float4 *>(&xin[rr][4]);
                    acc[0] = fmaf(xa.x, d, acc[0]); acc[1] = fmaf(xa.y, d, acc[1]);
                    acc[2] = fmaf(xa.z, d, acc[2]); acc[3] = fmaf(xa.w, d, acc[3]);
                    acc[4] = fmaf(xb.x, d, acc[4]); acc[5] = fmaf(xb.y, d, acc[5]);
                    acc[6] = fmaf(xb.z, d, acc[6]); acc[7] = fmaf(xb.w, d, acc[7]);
                }
            }
        }
        if (c < l.C) {
#pragma unroll
            for (int j = 0; j < WG_KG; ++j)
                if (k0 + j < KB) po[(size_t)j * l.C + c] = acc[j];
        }
    }
}

// grad[w_off + k*C + c] = sum_s partial[kg][s][kj][c] in chunk order; k == K is the bias / pool-weight row
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WDesc wd, int B, const float *__restrict__ partial,
                                                           float *__restrict__ grad)
{
    const int li = blockIdx.y;
    const WLayer l = wd.l[li];
    if (l.rpg == 0) return;
    int S, nkg, poff;
    wg_layer_geom(wd, B, li, S, nkg, poff);
    const int KB = l.K + 1;
    const float *pl = partial + poff;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < KB * l.C; idx += gridDim.x * blockDim.x) {
        const int k = idx / l.C, c = idx - k * l.C;
        const int kg = k / WG_KG, kj = k - kg * WG_KG;
        float acc = 0.f;
        const float *pp = pl + ((size_t)kg * S * WG_KG + kj) * l.C + c;
        const size_t sstep = (size_t)WG_KG * l.C;
#pragma unroll 8
        for (int s = 0; s < S; ++s) acc += __ldg(pp + s * sstep);   // independent loads, fixed summation order
        if (k < l.K) {
            if (l.w_off >= 0) grad[l.w_off + (size_t)k * l.C + c] = acc;
        } else if (l.b_off >= 0) {
            grad[l.b_off + c] = acc;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Huber replay loss + Adam
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) huber_kernel(const float *__restrict__ q1, const float *__restrict__ q2,
                                                    const int *__restrict__ action, const float *__restrict__ reward,
                                                    const int *__restrict__ next_slot, int B, int n_next, int A,
                                                    float gamma, int select, float *loss, float *gq1, float *gq2)
{
    __shared__ float red[256];
    const int tid = threadIdx.x;
    // zero the gradient buffer this call owns
    if (select) { for (int i = tid; i < B * A; i += 256) gq1[i] = 0.f; }
    else { for (int i = tid; i < n_next * A; i += 256) gq2[i] = 0.f; }
    __syncthreads();
    float lsum = 0.f;
    for (int b = tid; b < B; b += 256) {
        const int act = action[b];
        const float pred = q1[(size_t)b * A + act];
        const int slot = next_slot[b];
        float nsv = 0.f;
        int am = 0;
        if (slot >= 0) {
            const float *q = q2 + (size_t)slot * A;
            float m = q[0];
            for (int c = 1; c < A; ++c)
                if (q[c] > m) { m = q[c]; am = c; }
            nsv = m;
        }
        const float target = nsv * gamma + reward[b];
        const float d = pred - target;
        const float ad = fabsf(d);
        lsum += (ad < 1.f) ? 0.5f * d * d : (ad - 0.5f);
        const float gd = ((ad < 1.f) ? d : (d > 0.f ? 1.f : -1.f)) / (float)B;
        if (select) gq1[(size_t)b * A + act] = gd;
        else if (slot >= 0) gq2[(size_t)slot * A + am] = -gd * gamma;
    }
    red[tid] = lsum;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (tid < o) red[tid] += red[tid + o];
        __syncthreads();
    }
    if (tid == 0) *loss = red[0] / (float)B;
}

// loss = mean_b huber(pred_b - (r_b + gamma * nsv_b)) from the scalars the fused backward left behind;
// one warp per transition (max over A by lanes), fixed-shape reduction -> deterministic
__global__ void __launch_bounds__(1024) replay_loss_kernel(int mode, const float *__restrict__ scalar,
                                                           const float *__restrict__ qother, const int *__restrict__ action,
                                                           const float *__restrict__ reward, const int *__restrict__ next_slot,
                                                           int B, int A, float gamma, float *loss)
{
    __shared__ float red[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = tid & 3;  // 4 lanes per transition: 256 transitions per pass, their q rows read concurrently
    float lsum = 0.f;         // meaningful where sub == 0
    for (int b0 = 0; b0 < B; b0 += 256) {
        const int b = b0 + (tid >> 2);
        const bool ok = b < B;
        const int slot = ok ? next_slot[b] : -1;
        float pred = 0.f, nsv = 0.f;
        if (mode == 1) {
            if (ok) pred = scalar[b];
            float m = -INFINITY;
            if (slot >= 0) {
                const float *q = qother + (size_t)slot * A;
#pragma unroll 8
                for (int c = sub; c < A; c += 4) m = fmaxf(m, __ldg(q + c));
            }
            m = fmaxf(m, __shfl_xor_sync(FULL, m, 1));
            m = fmaxf(m, __shfl_xor_sync(FULL, m, 2));
            nsv = slot >= 0 ? m : 0.f;
        } else if (ok) {
            pred = qother[(size_t)b * A + action[b]];
            if (slot >= 0) nsv = scalar[slot];
        }
        if (ok && sub == 0) {
            const float d = pred - (nsv * gamma + reward[b]);
            const float ad = fabsf(d);
            lsum += (ad < 1.f) ? 0.5f * d * d : (ad - 0.5f);
        }
    }
    lsum = warp_sum(lsum);
    if (lane == 0) red[warp] = lsum;
    __syncthreads();
    if (warp == 0) {
        float v = red[lane];
#pragma unroll
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) *loss = v / (float)B;
    }
}

__global__ void adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                            float *__restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            float gscale, float step_size, float bc2_sqrt)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        const float pi = p[i];
        gi = fmaf(wd, pi, gi);
        const float m0 = m[i];
        const float mi = fmaf(gi - m0, 1.f - b1, m0);  // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

// Adam with the step count on the device (CUDA-graph friendly: no host-computed bias corrections in the arguments).
// step_dev[0] = steps taken so far; this launch is step t = step_dev[0] + 1 and advances the counter when its last block ends.
// The last block to finish advances the device step counter (step_dev[0]; step_dev[1] is the block ticket): every block
// has read the counter by then, and no separate launch sits between the optimizer and what waits for it.
__device__ __forceinline__ void step_done(int *step_dev, int t0)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(reinterpret_cast<unsigned *>(step_dev + 1), 1u) == gridDim.x - 1) {
            step_dev[1] = 0;
            step_dev[0] = t0 + 1;
        }
    }
}

__global__ void adam_dev_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                                float *__restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                                float gscale, int *__restrict__ step_dev)
{
    __shared__ float sh[2];
    const int t0 = step_dev[0];
    if (threadIdx.x == 0) {
        const double t = (double)(t0 + 1);
        const double bc1 = 1.0 - pow((double)b1, t), bc2 = 1.0 - pow((double)b2, t);
        sh[0] = (float)((double)lr / bc1);
        sh[1] = (float)sqrt(bc2);
    }
    __syncthreads();
    const float step_size = sh[0], bc2_sqrt = sh[1];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        const float pi = p[i];
        gi = fmaf(wd, pi, gi);
        const float m0 = m[i];
        const float mi = fmaf(gi - m0, 1.f - b1, m0);
        const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - step_size * (mi / denom);
    }
    step_done(step_dev, t0);
}

// ------------------------------------------------------------------------------------------------
// Fused gradient all-reduce + Adam over NVLink peer memory (data-parallel replay training).
// Every rank owns a "stage" buffer that all peers have mapped (torch symmetric memory does the mapping; the
// arithmetic and the synchronisation are here); its layout and the two-shot exchange are described at the kernel.
// Flags carry the step number, so they never need resetting.  One launch replaces ncclAllReduce (latency-bound at
// 0.5 MB) + the Adam launch.  A buffer parity is rewritten two steps later, which a rank can only reach after every
// peer has finished with it (it must have seen their flags of the step in between).  All blocks are co-resident
// (grid <= 132 blocks of 256 threads): they wait for each other's phases through a block counter.
// ------------------------------------------------------------------------------------------------
constexpr int AR_MAX_WORLD = 16;
struct ArArgs {
    float *stage[AR_MAX_WORLD];       // peer-mapped stage buffers, index = rank
    int rank, world;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Two-shot, push-only exchange (all remote traffic is posted stores, every wait polls LOCAL memory):
//   1. every rank writes slice j of its gradient into rank j's inbox [parity][sender][chunk], then raises flag1 on all ranks;
//   2. rank j adds its inbox in RANK ORDER (the one summation order every element sees, on whichever rank it is reduced)
//      and writes the reduced slice into every rank's result vector [parity][world * chunk], then raises flag2;
//   3. every rank runs Adam over the whole vector from its local copy of the reduced gradient.
// Per rank 2 * (world - 1) / world * 4n bytes cross NVLink (one-shot pulling: (world - 1) * 4n).  Stage layout in floats:
// inbox at 0 (2 * npad), result at 2 * npad (2 * npad), flag words at 4 * npad (2 x AR_MAX_WORLD), npad = world * chunk.
__device__ __forceinline__ float4 ld_cv4(const float *p)
{
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned *p, unsigned v)
{
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// End of a phase: the LAST block to get here (all of the grid's stores of the phase are then ordered before its fence)
// raises this rank's flag on every peer -- one fence, then `world` independent relaxed stores from `world` threads -- and
// every block waits until all ranks' flags for this step have arrived in local memory.
__device__ __forceinline__ void ar_signal_and_wait(const ArArgs &ar, unsigned *counter, int which, int64_t npad, unsigned seq)
{
    __shared__ unsigned last;
    __syncthreads();                 // the block's stores happen-before thread 0's fence (cumulative), one fence per block
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned done = atomicAdd(counter, 1u) == gridDim.x - 1;
        if (done) *counter = 0u;
        last = done;
        __threadfence_system();
    }
    __syncthreads();
    if (last && (int)threadIdx.x < ar.world)
        st_relaxed_sys(reinterpret_cast<unsigned *>(ar.stage[threadIdx.x] + 4 * npad) + which * AR_MAX_WORLD + ar.rank, seq);
    if ((int)threadIdx.x < ar.world) {
        const unsigned *flag = reinterpret_cast<const unsigned *>(ar.stage[ar.rank] + 4 * npad) + which * AR_MAX_WORLD + threadIdx.x;
        unsigned spins = 0;
        while ((int)(ld_acquire_sys(flag) - seq) < 0)
            if (++spins > (1u << 23)) __trap();     // (~10 s) a lost peer must fail loudly, never hang the device
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) allreduce_adam_kernel(const ArArgs ar, float *__restrict__ p, const float *__restrict__ g,
                                                           float *__restrict__ m, float *__restrict__ v, int64_t n, float lr,
                                                           float b1, float b2, float eps, float wd,
                                                           int *__restrict__ step_dev, unsigned *block_counter)
{
    __shared__ float sh[2];
    const int t = step_dev[0];
    const unsigned seq = (unsigned)t + 1u;
    const int world = ar.world, rank = ar.rank;
    const int64_t chunk = ((n + 4 * world - 1) / (4 * world)) * 4, npad = chunk * world;
    const int64_t par = (int64_t)(t & 1) * npad;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsz = (int64_t)gridDim.x * blockDim.x;
    long long *stamp = reinterpret_cast<long long *>(ar.stage[rank] + 4 * npad + 2 * AR_MAX_WORLD);
#define AR_STAMP(k)                                                                             \
    if (gtid == 0) { long long tn; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tn)); stamp[k] = tn; }
    AR_STAMP(0)
    if (threadIdx.x == 0) {
        const double tt = (double)(t + 1);
        const double bc1 = 1.0 - pow((double)b1, tt), bc2 = 1.0 - pow((double)b2, tt);
        sh[0] = (float)((double)lr / bc1);
        sh[1] = (float)sqrt(bc2);
    }
    // 1. scatter: float4 i of my gradient belongs to rank i / (chunk / 4)
    const int64_t c4 = chunk / 4;
    for (int64_t i = gtid; i < npad / 4; i += gsz) {
        const int64_t e = 4 * i;
        float4 val;
        if (e + 3 < n) val = *reinterpret_cast<const float4 *>(g + e);     // g is 16-byte aligned (a whole torch tensor)
        else { val.x = e < n ? g[e] : 0.f; val.y = e + 1 < n ? g[e + 1] : 0.f; val.z = e + 2 < n ? g[e + 2] : 0.f; val.w = 0.f; }
        const int owner = (int)(i / c4);
        const int64_t off = (i - owner * c4) * 4;
        *reinterpret_cast<float4 *>(ar.stage[owner] + par + (int64_t)rank * chunk + off) = val;
    }
    AR_STAMP(1)
    ar_signal_and_wait(ar, block_counter, 0, npad, seq);
    AR_STAMP(2)
    // 2. reduce my slice in rank order, publish it to everyone
    const float *inbox = ar.stage[rank] + par;
    for (int64_t i = gtid; i < c4; i += gsz) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 a = ld_cv4(inbox + (int64_t)r * chunk + 4 * i);
            s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
        }
        for (int r = 0; r < world; ++r)
            *reinterpret_cast<float4 *>(ar.stage[r] + 2 * npad + par + (int64_t)rank * chunk + 4 * i) = s;
    }
    AR_STAMP(3)
    ar_signal_and_wait(ar, block_counter + 1, 1, npad, seq);
    AR_STAMP(4)
    // 3. Adam over the whole vector
    const float step_size = sh[0], bc2_sqrt = sh[1];
    const float gscale = 1.f / (float)world;
    const float *red = ar.stage[rank] + 2 * npad + par;
    for (int64_t i = gtid; i < npad / 4; i += gsz) {
        const float4 gs4 = ld_cv4(red + 4 * i);
        const float gsv[4] = {gs4.x, gs4.y, gs4.z, gs4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t e = 4 * i + u;
            if (e >= n) break;
            float gi = gsv[u] * gscale;
            const float pi = p[e];
            gi = fmaf(wd, pi, gi);
            const float m0 = m[e];
            const float mi = fmaf(gi - m0, 1.f - b1, m0);
            const float vi = fmaf(b2, v[e], (1.f - b2) * gi * gi);
            m[e] = mi;
            v[e] = vi;
            const float denom = sqrtf(vi) / bc2_sqrt + eps;
            p[e] = pi - step_size * (mi / denom);
        }
    }
    AR_STAMP(5)
#undef AR_STAMP
    step_done(step_dev, t);
}

int setup_launch(const mdq_net_t *net, int max_n, int max_e, int G, int bwd, QLay &L, WChunks *ck, void (*kern)(const QArgs))
{
    int rc = build_layout(*net, max_n, max_e, G, bwd, L, ck);
    if (rc == MDQ_ESMEM) {
        mdq::set_error("qnet: graphs of %d nodes / %d edges do not fit the fused kernel's shared memory tiles "
                       "(%d graph(s)/CTA); a single graph of this size runs through mdq_qnet_forward_layered", max_n, max_e, G);
        return rc;
    }
    if (rc != MDQ_OK) { mdq::set_error("qnet: unsupported network/size (rc=%d)", rc); return rc; }
    const size_t bytes = (size_t)L.total * 4;
    if (bytes > 227 * 1024) {
        mdq::set_error("qnet: graphs of %d nodes / %d edges need %zu B of shared memory (> 227 KB) with %d graph(s)/CTA",
                       max_n, max_e, bytes, G);
        return MDQ_ESMEM;
    }
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess)  // ask for the largest shared-memory carve-out so two ~110 KB CTAs fit one SM
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { mdq::set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return MDQ_ECUDA; }
    return MDQ_OK;
}

long long *g_trace = nullptr;

#include "gnn_staged.cuh"
#include "gnn_tail.cuh"

}  // namespace

#include "gnn_staged_api.cuh"

// ================================================================================================
// C ABI
// ================================================================================================
extern "C" {

void mdq_qnet_set_trace(int64_t *device_buf) { g_trace = reinterpret_cast<long long *>(device_buf); }

int64_t mdq_qnet_smem_bytes(const mdq_net_t *net, int max_n, int max_e, int gpc, int backward)
{
    QLay L;
    if (build_layout(*net, max_n, max_e, gpc, backward, L) != MDQ_OK) return -1;
    return (int64_t)L.total * 4;
}

int mdq_qnet_occupancy(const mdq_net_t *net, int max_n, int max_e, int backward)
{
    QLay L;
    WChunks ck;
    int rc = backward ? setup_launch(net, max_n, max_e, 1, 1, L, &ck, qnet_kernel<true>)
                      : setup_launch(net, max_n, max_e, 1, 0, L, &ck, qnet_kernel<false>);
    if (rc != MDQ_OK) return rc;
    int nblk = 0;
    cudaError_t e = backward ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, qnet_kernel<true>, NT, (size_t)L.total * 4)
                             : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nblk, qnet_kernel<false>, NT, (size_t)L.total * 4);
    if (e != cudaSuccess) { mdq::set_error("occupancy query: %s", cudaGetErrorString(e)); return MDQ_ECUDA; }
    return nblk;
}

int mdq_qnet_pick_gpc(const mdq_net_t *net, int n_graphs, int max_n, int max_e)
{
    // fewest waves over 148 SMs; CTAs/SM follows from shared memory (228 KB/SM, 1 KB reserved per CTA) and the
    // 128-register launch bound (<= 2); ties -> fewer graphs per CTA (shorter dependent chain per CTA)
    int best = 1;
    long best_waves = -1;
    for (int G = 1; G <= 4; ++G) {
        QLay L;
        if (build_layout(*net, max_n, max_e, G, 0, L) != MDQ_OK) break;
        const size_t bytes = (size_t)L.total * 4;
        if (bytes > 227 * 1024) break;
        int per_sm = (int)((228 * 1024) / (bytes + 1024));
        if (per_sm > 2) per_sm = 2;
        if (per_sm < 1) per_sm = 1;
        const long ctas = (n_graphs + G - 1) / G;
        const long waves = (ctas + 148L * per_sm - 1) / (148L * per_sm);
        // cost ~ waves * graphs per CTA (a CTA walks its graphs' first block one after the other)
        const long cost = waves * G;
        if (best_waves < 0 || cost < best_waves) { best_waves = cost; best = G; }
    }
    return best;
}

int mdq_qnet_forward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                     const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                     int max_n, int max_e, float *out, float *embedding, int32_t *argmax, void *stream)
{
    if (!net || !params || !x || !node_ptr || !edge_ptr || !out || n_graphs < 1) {
        mdq::set_error("mdq_qnet_forward: null argument or empty batch");
        return MDQ_EINVAL;
    }
    const int G = mdq_qnet_pick_gpc(net, n_graphs, max_n, max_e);
    QArgs a;
    memset(&a, 0, sizeof(a));
    int rc = setup_launch(net, max_n, max_e, G, 0, a.L, &a.ck, qnet_kernel<false>);
    if (rc != MDQ_OK) return rc;
    a.net = *net;
    a.params = params; a.x = x; a.esrc = (const long long *)edge_src; a.edst = (const long long *)edge_dst; a.nptr = node_ptr; a.eptr = edge_ptr;
    a.B = n_graphs; a.out = out; a.emb = embedding; a.amax_out = argmax; a.trace = g_trace;
    const int grid = (n_graphs + G - 1) / G;
    qnet_kernel<false><<<grid, NT, (size_t)a.L.total * 4, (cudaStream_t)stream>>>(a);
    return mdq::check_launch("qnet_kernel<fwd>");
}

int64_t mdq_qnet_bwd_workspace_floats(const mdq_net_t *net, int n_graphs, int max_n)
{
    WDesc wd;
    if (build_wdesc(*net, n_graphs, max_n, wd) != MDQ_OK) return -1;
    // rows workspace + split-K partials + task tables (ints stored in the same buffer)
    int64_t partial = 0, tasks = 0;
    for (int i = 0; i < wd.nl; ++i) {
        const WLayer &l = wd.l[i];
        if (l.rpg == 0) continue;
        const int rows = n_graphs * l.rpg;
        const int S = (rows + WG_TASK - 1) / WG_TASK;
        const int nkg = (l.K + 1 + WG_KG - 1) / WG_KG;
        partial += (int64_t)S * nkg * WG_KG * l.C;
        tasks += (int64_t)S * nkg;
    }
    (void)tasks;
    // + every graph's spill slice (hidden rows of a backward layout that overflows shared memory; max_e does not enter)
    QLay L;
    if (build_layout(*net, max_n, 1, 1, 1, L) != MDQ_OK) return -1;
    return (int64_t)wd.total + partial + 64 + (int64_t)n_graphs * L.sp_words;
}

static int64_t bwd_spill_offset(const WDesc &wd, int n_graphs)
{
    int64_t partial = 0;
    for (int i = 0; i < wd.nl; ++i) {
        const WLayer &l = wd.l[i];
        if (l.rpg == 0) continue;
        const int S = (n_graphs * l.rpg + WG_TASK - 1) / WG_TASK;
        const int nkg = (l.K + 1 + WG_KG - 1) / WG_KG;
        partial += (int64_t)S * nkg * WG_KG * l.C;
    }
    return (int64_t)wd.total + partial + 64;
}

static int qnet_backward_launch(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                                const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                                int max_n, int max_e, QArgs &a, float *grad, float *workspace, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int rc = setup_launch(net, max_n, max_e, 1, 1, a.L, &a.ck, qnet_kernel<true>);
    if (rc != MDQ_OK) return rc;
    build_wdesc(*net, n_graphs, max_n, a.wd);
    a.net = *net;
    a.params = params; a.x = x; a.esrc = (const long long *)edge_src; a.edst = (const long long *)edge_dst; a.nptr = node_ptr; a.eptr = edge_ptr;
    a.B = n_graphs; a.ws = workspace; a.trace = g_trace;
    a.spill = workspace + bwd_spill_offset(a.wd, n_graphs);
    const WDesc &wd = a.wd;
    int n_tasks = 0;
    for (int i = 0; i < wd.nl; ++i) {
        const WLayer &l = wd.l[i];
        a.wd.tstart[i] = n_tasks;
        if (l.rpg == 0) continue;
        const int S = (n_graphs * l.rpg + WG_TASK - 1) / WG_TASK;
        const int nkg = (l.K + 1 + WG_KG - 1) / WG_KG;
        n_tasks += S * nkg;
    }
    a.wd.tstart[wd.nl] = n_tasks;
    for (int i = 0; i < wd.nl; ++i)
        if (wd.l[i].C > 256) { mdq::set_error("qnet backward: layers wider than 256 outputs are not supported (%d)", wd.l[i].C); return MDQ_EINVAL; }
    float *d_partial = workspace + wd.total;
    cudaError_t e = cudaMemsetAsync(grad, 0, (size_t)net->n_params * sizeof(float), st);
    if (e != cudaSuccess) { mdq::set_error("memset grad: %s", cudaGetErrorString(e)); return MDQ_ECUDA; }
    qnet_kernel<true><<<n_graphs, NT, (size_t)a.L.total * 4, st>>>(a);
    rc = mdq::check_launch("qnet_kernel<bwd>");
    if (rc != MDQ_OK) return rc;
    wgrad_partial_kernel<<<n_tasks > 0 ? n_tasks : 1, 256, 0, st>>>(a.wd, n_graphs, workspace, d_partial);
    rc = mdq::check_launch("wgrad_partial_kernel");
    if (rc != MDQ_OK) return rc;
    dim3 rg(32, wd.nl);
    wgrad_reduce_kernel<<<rg, 256, 0, st>>>(a.wd, n_graphs, d_partial, grad);
    return mdq::check_launch("wgrad_reduce_kernel");
}

int mdq_qnet_backward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                      const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                      int max_n, int max_e, const float *grad_out, float *grad, float *workspace, void *stream)
{
    if (!net || !params || !x || !grad_out || !grad || !workspace || n_graphs < 1) {
        mdq::set_error("mdq_qnet_backward: null argument or empty batch");
        return MDQ_EINVAL;
    }
    QArgs a;
    memset(&a, 0, sizeof(a));
    a.gout = grad_out;
    return qnet_backward_launch(net, params, x, edge_src, edge_dst, node_ptr, edge_ptr, n_graphs, max_n, max_e, a, grad,
                                workspace, stream);
}

int mdq_qnet_replay_backward(const mdq_net_t *net, const float *params, const float *x, const int64_t *edge_src,
                             const int64_t *edge_dst, const int32_t *node_ptr, const int32_t *edge_ptr, int n_graphs,
                             int max_n, int max_e, int mode, const int32_t *action, const float *reward,
                             const int32_t *index, const int32_t *next_slot, const float *q_other, int batch, float gamma,
                             float *scalar, float *loss, float *grad, float *workspace, void *stream)
{
    if (!net || !params || !x || !grad || !workspace || !action || !reward || !index || !next_slot || !scalar || !loss ||
        n_graphs < 1 || batch < 1 || (mode != 1 && mode != 2) || (mode == 2 && !q_other)) {
        mdq::set_error("mdq_qnet_replay_backward: bad argument");
        return MDQ_EINVAL;
    }
    QArgs a;
    memset(&a, 0, sizeof(a));
    a.rp_mode = mode; a.rp_action = action; a.rp_reward = reward; a.rp_index = index; a.rp_qother = q_other;
    a.rp_gamma = gamma; a.rp_inv_batch = 1.f / (float)batch; a.rp_scalar = scalar;
    int rc = qnet_backward_launch(net, params, x, edge_src, edge_dst, node_ptr, edge_ptr, n_graphs, max_n, max_e, a, grad,
                                  workspace, stream);
    if (rc != MDQ_OK) return rc;
    replay_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mode, scalar, q_other, action, reward, next_slot, batch,
                                                           net->out_dim, gamma, loss);
    return mdq::check_launch("replay_loss_kernel");
}

int mdq_huber_replay(const float *q1, const float *q2, const int32_t *action, const float *reward,
                     const int32_t *next_slot, int batch, int n_next, int out_dim, float gamma, int select,
                     float *loss, float *grad_q1, float *grad_q2, void *stream)
{
    if (!q1 || !action || !reward || !next_slot || !loss || batch < 1 || (select ? !grad_q1 : !grad_q2)) {
        mdq::set_error("mdq_huber_replay: null argument");
        return MDQ_EINVAL;
    }
    huber_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(q1, q2, action, reward, next_slot, batch, n_next, out_dim, gamma,
                                                      select, loss, grad_q1, grad_q2);
    return mdq::check_launch("huber_kernel");
}

int mdq_adam_step(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, int step, void *stream)
{
    if (!params || !grad || !exp_avg || !exp_avg_sq || n < 1 || step < 1) {
        mdq::set_error("mdq_adam_step: bad argument");
        return MDQ_EINVAL;
    }
    const double bc1d = 1.0 - pow((double)beta1, (double)step);
    const double bc2d = 1.0 - pow((double)beta2, (double)step);
    const int threads = 256;
    int blocks = (int)((n + threads - 1) / threads);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                              weight_decay, grad_scale, (float)((double)lr / bc1d), (float)sqrt(bc2d));
    return mdq::check_launch("adam_kernel");
}

int64_t mdq_allreduce_stage_floats(int64_t n, int world)
{
    if (n < 1 || world < 1 || world > AR_MAX_WORLD) return -1;
    const int64_t chunk = ((n + 4 * world - 1) / (4 * world)) * 4;
    return 4 * chunk * world + 2 * AR_MAX_WORLD + 16;      // + 8 x 64-bit phase stamps of the last call (diagnostics)
}

int mdq_allreduce_adam(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                       float beta1, float beta2, float eps, float weight_decay, int32_t *step_dev,
                       const uint64_t *h_peer_stage, int rank, int world, uint32_t *block_counter, void *stream)
{
    if (!params || !grad || !exp_avg || !exp_avg_sq || !step_dev || !h_peer_stage || !block_counter || n < 1 || world < 1 ||
        world > AR_MAX_WORLD || rank < 0 || rank >= world) {
        mdq::set_error("mdq_allreduce_adam: bad argument");
        return MDQ_EINVAL;
    }
    ArArgs ar;
    memset(&ar, 0, sizeof(ar));
    for (int r = 0; r < world; ++r) ar.stage[r] = reinterpret_cast<float *>(h_peer_stage[r]);
    ar.rank = rank; ar.world = world;
    int blocks = (int)((n / 4 + 255) / 256);
    if (blocks > 132) blocks = 132;          // every block must be resident while it waits for the peers' flags
    allreduce_adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(ar, params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                    eps, weight_decay, step_dev, block_counter);
    return mdq::check_launch("allreduce_adam_kernel");
}

int mdq_adam_step_dev(float *params, const float *grad, float *exp_avg, float *exp_avg_sq, int64_t n, float lr,
                      float beta1, float beta2, float eps, float weight_decay, float grad_scale, int32_t *step_dev,
                      void *stream)
{
    if (!params || !grad || !exp_avg || !exp_avg_sq || !step_dev || n < 1) {
        mdq::set_error("mdq_adam_step_dev: bad argument");
        return MDQ_EINVAL;
    }
    const int threads = 256;
    int blocks = (int)((n + threads - 1) / threads);
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_dev_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(params, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                                  weight_decay, grad_scale, step_dev);
    return mdq::check_launch("adam_dev_kernel");
}

}  // extern "C"
