// gnn_layered.cu -- layer-by-layer Q-network forward for ONE large graph (sm_100a).
//
// The fused kernel (gnn_fused.cu) keeps a whole state graph in one CTA's shared memory; a graph with
// ~0.5M nodes (BASELINE.json config 4: the state graph of a ~1M-triangle mesh) does not fit, so the same
// network (/root/reference/airfoilgcnn.py:85-145, :170-209) runs here as a sequence of grid-wide kernels:
//
//   k_csr_*        CSR-by-destination build, rows kept in edge order (torch_scatter's CPU summation order)
//   k_sage_rows    SAGEConv message passing: A[i] = [ mean_{j->i} x_j | x_i ]  -- warp per row, lanes over the
//                  (vectorised) feature row, deterministic, atomics-free, HBM-bound
//   node GEMM      C = A . W (+ bias, ReLU, TopK score) -- fp32 FFMA kernel here; tcgen05 3xTF32 kernel in
//                  node_gemm_tc.cuh (tensor cores, TMEM accumulator)
//   k_gcn_rows     GCNConv aggregation with self loops, D^-1/2 (A+I) D^-1/2, + bias, ReLU, TopK score
//   radix sort     TopKPooling's descending stable sort (ties -> lower node index), 8-bit LSD passes
//   k_pool_*       gather * score, inverse map, ordered edge filter, global max / mean readout
//   k_mlp_head     lin1/lin2/lin3 + softmax + first-max argmax (airfoil_dqn.py:208-209)
//
// One graph per call (B = 1); every size the launches need is known on the host (n_{l+1} = ceil(ratio * n_l) in
// float32, as PyG computes it), edge counts stay on the device.  No host synchronisation inside.
#include <math.h>

#include <algorithm>

#include "mdq_common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int grid_for(long long n, int per, int cap = 148 * 32)
{
    long long b = (n + per - 1) / per;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_scan_excl_1024(int v, int *wtmp, int &total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wtmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = wtmp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, wi, o);
            if (lane >= o) wi += t;
        }
        wtmp[lane] = wi - w;
        if (lane == 31) wtmp[32] = wi;
    }
    __syncthreads();
    const int res = wtmp[warp] + incl - v;
    total = wtmp[32];
    __syncthreads();
    return res;
}

// exclusive scan of in[0..n) -> out[0..n], out[n] = total; one CTA of 1024 threads (n is at most a few 100k here)
__global__ void __launch_bounds__(1024) k_scan(const int *__restrict__ in, int *__restrict__ out, int n)
{
    __shared__ int wtmp[33];
    int carry = 0;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = (i < n) ? in[i] : 0;
        int total;
        const int ex = block_scan_excl_1024(v, wtmp, total);
        if (i < n) out[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// Large scans (row pointers of 0.5M-node graphs): per-CTA partial sums, scan of the partials, local scan + offset.
constexpr int SCAN_TILE = 4096;  // elements per CTA (1024 threads x 4)
__global__ void __launch_bounds__(1024) k_scan_partials(const int *__restrict__ in, int n, int *__restrict__ part)
{
    __shared__ int wtmp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) s += (base + q < n) ? in[base + q] : 0;
    int total;
    block_scan_excl_1024(s, wtmp, total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_scan_apply(const int *__restrict__ in, int n, const int *__restrict__ part_ex,
                                                     int *__restrict__ out)
{
    __shared__ int wtmp[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * 4;
    int v[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        v[q] = (base + q < n) ? in[base + q] : 0;
        s += v[q];
    }
    int total;
    int ex = block_scan_excl_1024(s, wtmp, total) + part_ex[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (base + q < n) out[base + q] = ex;
        ex += v[q];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = ex;
}

// out[0..n] = exclusive scan of in[0..n); part: scratch of cdiv(n, SCAN_TILE) * 2 + 2 ints
int scan_i32(const int *in, int *out, int n, int *part, cudaStream_t st)
{
    if (n <= 8192) {
        k_scan<<<1, 1024, 0, st>>>(in, out, n);
        return mdq::check_launch("k_scan");
    }
    const int nb = cdiv(n, SCAN_TILE);
    int rc;
    k_scan_partials<<<nb, 1024, 0, st>>>(in, n, part);
    if ((rc = mdq::check_launch("k_scan_partials"))) return rc;
    k_scan<<<1, 1024, 0, st>>>(part, part + nb + 1, nb);
    if ((rc = mdq::check_launch("k_scan"))) return rc;
    k_scan_apply<<<nb, 1024, 0, st>>>(in, n, part + nb + 1, out);
    return mdq::check_launch("k_scan_apply");
}

// ------------------------------------------------------------------------------------------------
// CSR by destination, rows in edge order
// ------------------------------------------------------------------------------------------------
__global__ void k_edges_to_i32(const long long *__restrict__ src, const long long *__restrict__ dst, int ne,
                               int *__restrict__ s32, int *__restrict__ d32, int *__restrict__ ecount)
{
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
        s32[e] = (int)src[e];
        d32[e] = (int)dst[e];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *ecount = ne;
}

__global__ void k_csr_count(const int *__restrict__ dst, const int *__restrict__ ecount, int *__restrict__ deg)
{
    const int ne = *ecount;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) atomicAdd(deg + dst[e], 1);
}

// Each row receives its edge ids in arbitrary order (atomics), then one thread per row sorts them ascending, so
// the summation order -- and therefore every bit of the aggregate -- is reproducible and equals edge order.
__global__ void k_csr_fill(const int *__restrict__ dst, const int *__restrict__ ecount, const int *__restrict__ row_ptr,
                           int *__restrict__ cursor, int *__restrict__ eid)
{
    const int ne = *ecount;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < ne; e += gridDim.x * blockDim.x) {
        const int d = dst[e];
        eid[row_ptr[d] + atomicAdd(cursor + d, 1)] = e;
    }
}
__global__ void k_csr_sort_rows(const int *__restrict__ src, const int *__restrict__ row_ptr, int n, int *__restrict__ eid,
                                int *__restrict__ col)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const int a = row_ptr[r], b = row_ptr[r + 1];
        const int d = b - a;
        if (d <= 16) {   // mesh vertex degrees: sort in registers (odd-even transposition network)
            int v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = (i < d) ? eid[a + i] : 0x7fffffff;
#pragma unroll
            for (int pass = 0; pass < 16; ++pass)
#pragma unroll
                for (int i = pass & 1; i + 1 < 16; i += 2) {
                    const int lo = min(v[i], v[i + 1]), hi = max(v[i], v[i + 1]);
                    v[i] = lo;
                    v[i + 1] = hi;
                }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i < d) col[a + i] = src[v[i]];
        } else {
            for (int i = a + 1; i < b; ++i) {  // insertion sort in place for the rare long row
                const int v = eid[i];
                int j = i - 1;
                while (j >= a && eid[j] > v) {
                    eid[j + 1] = eid[j];
                    --j;
                }
                eid[j + 1] = v;
            }
            for (int i = a; i < b; ++i) col[i] = src[eid[i]];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SAGEConv message passing: A[r] = [ mean_{j -> i} X[j] | X[i] | 0-pad ], i = rows ? rows[r] : r
// ------------------------------------------------------------------------------------------------
// SAGEConv message passing.  Output row layout (the node GEMM's A operand):
//     A[i] = [ x_i (F, zero-padded to Fp = 4*ceil(F/4)) | mean_{j->i} x_j (F, padded to Fp) ],   lda >= 2 Fp.
// Narrow rows (F <= 32, e.g. the 17 state features -> Fp = 20) take two passes: k_pad_rows copies x into the
// 16-byte-aligned left half, then k_sage_rows_vec4 gathers neighbours' left halves with LDG.128 -- one thread per
// (row, 4-feature chunk), up to eight neighbour chunks in flight -- and writes the mean into the right half.
// (Reading the 68-byte unpadded rows directly costs one 4-byte load per feature and is instruction-bound.)
// Each output element is summed sequentially in CSR (= edge) order, so the bits equal torch_scatter's CPU loop;
// no atomics, deterministic.
// (32-bit index arithmetic throughout: a 64-bit division per thread costs more instructions than the kernel's work)
// Two alternatives were measured on the 0.5M-node / 3.0M-edge graph and dropped (ncu, same bits): a single-pass
// shared-memory tile kernel with one thread per (row, feature) reading the unpadded rows -- 87 us, issue-bound at
// 50 M warp instructions (one LDG + FADD per lane and edge) against 72 us for this pair (19 + 53, 35 M); and a
// thread-per-row kernel with the row in registers -- 11.6 M instructions but 61 us after the same pad pass, bound
// by the L1 sector rate of fully divergent 16-byte loads (32 rows per request, l1tex 60 %).
template <int CH>   // CH > 0: compile-time chunks per row (constant division), 0: runtime
__global__ void __launch_bounds__(256) k_pad_rows(const float *__restrict__ X, int ldx, int col0, int F, int n_rows, int ch_rt,
                                                  float *__restrict__ A, int lda)
{
    const int ch = CH > 0 ? CH : ch_rt;
    const unsigned total = (unsigned)n_rows * (unsigned)ch;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned r = idx / (unsigned)ch, c = idx - r * (unsigned)ch;
        const float *xr = X + (size_t)r * ldx + col0 + 4 * c;
        float4 v;
        v.x = __ldg(xr);
        v.y = (4 * c + 1 < (unsigned)F) ? __ldg(xr + 1) : 0.f;
        v.z = (4 * c + 2 < (unsigned)F) ? __ldg(xr + 2) : 0.f;
        v.w = (4 * c + 3 < (unsigned)F) ? __ldg(xr + 3) : 0.f;
        reinterpret_cast<float4 *>(A + (size_t)r * lda)[c] = v;
    }
}

template <int CH>
__global__ void __launch_bounds__(256) k_sage_rows_vec4(const int *__restrict__ row_ptr, const int *__restrict__ col,
                                                        int n_rows, int ch_rt, float *__restrict__ A, int lda)
{
    const int ch = CH > 0 ? CH : ch_rt;
    const unsigned total = (unsigned)n_rows * (unsigned)ch;
    const unsigned ld4 = (unsigned)lda >> 2;
    const float4 *A4 = reinterpret_cast<const float4 *>(A);
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned r = idx / (unsigned)ch, c = idx - r * (unsigned)ch;
        const int a = __ldg(row_ptr + r), b = __ldg(row_ptr + r + 1);
        const float4 *Ac = A4 + c;      // plain loads below: the left half of A was written by k_pad_rows
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = a;
        for (; e + 4 <= b; e += 4) {
            const unsigned j0 = __ldg(col + e), j1 = __ldg(col + e + 1), j2 = __ldg(col + e + 2), j3 = __ldg(col + e + 3);
            const float4 v0 = Ac[(size_t)j0 * ld4], v1 = Ac[(size_t)j1 * ld4], v2 = Ac[(size_t)j2 * ld4], v3 = Ac[(size_t)j3 * ld4];
            s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
            s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
            s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
            s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
        }
        if (e < b) {   // 1..3 left: clamped (always valid) loads, weight 0 for the clamped repeats -- s + 0*v == s exactly
            const unsigned j0 = __ldg(col + e), j1 = __ldg(col + min(e + 1, b - 1)), j2 = __ldg(col + min(e + 2, b - 1));
            const float4 v0 = Ac[(size_t)j0 * ld4], v1 = Ac[(size_t)j1 * ld4], v2 = Ac[(size_t)j2 * ld4];
            const float w1 = (e + 1 < b) ? 1.f : 0.f, w2 = (e + 2 < b) ? 1.f : 0.f;
            s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
            s.x = fmaf(w1, v1.x, s.x); s.y = fmaf(w1, v1.y, s.y); s.z = fmaf(w1, v1.z, s.z); s.w = fmaf(w1, v1.w, s.w);
            s.x = fmaf(w2, v2.x, s.x); s.y = fmaf(w2, v2.y, s.y); s.z = fmaf(w2, v2.z, s.z); s.w = fmaf(w2, v2.w, s.w);
        }
        const float deg = (float)max(b - a, 1);
        reinterpret_cast<float4 *>(A + (size_t)r * lda)[ch + c] = make_float4(s.x / deg, s.y / deg, s.z / deg, s.w / deg);
        // columns beyond 2 Fp (lda rounded up for the tensor-core K step) are zeroed by the chunk-0 thread
        if (c == 0)
            for (int k = 8 * ch; k < lda; ++k) A[(size_t)r * lda + k] = 0.f;
    }
}

// F == 128: one warp per row, lane owns a float4 (512-byte rows, fully coalesced 16-byte vector loads), four
// neighbour rows in flight; A[i] = [ x_i | mean ] with lda = 256.
__global__ void __launch_bounds__(256) k_sage_rows_128(const float *__restrict__ X, const int *__restrict__ row_ptr,
                                                       const int *__restrict__ col, int n_rows, float *__restrict__ A)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float4 *X4 = reinterpret_cast<const float4 *>(X) + lane;
    for (int r = warp; r < n_rows; r += nwarps) {
        const int a = __ldg(row_ptr + r), b = __ldg(row_ptr + r + 1);
        const float4 self = __ldg(X4 + (size_t)r * 32);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        int e = a;
        for (; e + 4 <= b; e += 4) {
            const int j0 = __ldg(col + e), j1 = __ldg(col + e + 1), j2 = __ldg(col + e + 2), j3 = __ldg(col + e + 3);
            const float4 v0 = __ldg(X4 + (size_t)j0 * 32), v1 = __ldg(X4 + (size_t)j1 * 32);
            const float4 v2 = __ldg(X4 + (size_t)j2 * 32), v3 = __ldg(X4 + (size_t)j3 * 32);
            s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
            s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
            s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
            s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
        }
        for (; e < b; ++e) {
            const float4 v = __ldg(X4 + (size_t)__ldg(col + e) * 32);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const float deg = (float)max(b - a, 1);
        float4 *Ar = reinterpret_cast<float4 *>(A + (size_t)r * 256);
        Ar[lane] = self;
        Ar[32 + lane] = make_float4(s.x / deg, s.y / deg, s.z / deg, s.w / deg);
    }
}

// ------------------------------------------------------------------------------------------------
// node GEMM, fp32 FFMA: C[M, 128-wide tiles] = epilogue(A[rows][K] . W[K][N])
//   epilogue: + bias (nullable), ReLU (flag), TopK score tanh(h.p / ||p||) (pool nullable), row scale (nullable,
//   multiplies the stored row: x[perm] * score[perm]), store C (nullable)
// ------------------------------------------------------------------------------------------------
struct GemmArgs {
    const float *A;        // [*, lda]
    const int *rows;       // nullable: A row of output row r is rows[r]
    int lda, K, M, N;      // N = width (multiple of 4, <= 256)
    const float *W;        // [K][N]
    const float *bias;     // [N] or null
    const float *pool;     // [N] or null
    const float *row_scale;  // indexed like A rows (score of the source row) or null
    int relu;
    float *C;              // [M, N] or null
    float *score;          // [M] or null (requires pool)
    int sage_F, sage_Fp;   // > 0: A rows are [x (Fp) | mean (Fp)] while W rows are [lin_l (F) ; lin_r (F)]: W row k reads
                           // A column (k < F ? Fp + k : k - F).  The tensor-core path gets W pre-permuted instead.
};

constexpr int GT_M = 64, GT_K = 16;
template <int N>
__global__ void __launch_bounds__(256) k_node_gemm_f32(const GemmArgs g)
{
    // 256 threads: thread (ty, tx) -> rows ty*4..+3 (16 row groups), cols tx + 16*c (N/16 columns)
    constexpr int CPT = N / 16;
    __shared__ float As[GT_K][GT_M + 4];
    __shared__ float Ws[GT_K][N];
    __shared__ float red[GT_M][17];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float pnorm = 0.f;
    if (g.pool) {
        for (int c = 0; c < N; ++c) pnorm += g.pool[c] * g.pool[c];
        pnorm = sqrtf(pnorm);
    }
    for (int m0 = blockIdx.x * GT_M; m0 < g.M; m0 += gridDim.x * GT_M) {
        float acc[4][CPT];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
        for (int k0 = 0; k0 < g.K; k0 += GT_K) {
            // A tile: 64 rows x 16 k
            for (int idx = threadIdx.x; idx < GT_M * GT_K; idx += 256) {
                const int r = idx / GT_K, k = idx % GT_K;
                const int row = m0 + r;
                float v = 0.f;
                if (row < g.M && k0 + k < g.K) {
                    const int ar = g.rows ? g.rows[row] : row;
                    const int kk = k0 + k;
                    const int ca = g.sage_F > 0 ? (kk < g.sage_F ? g.sage_Fp + kk : kk - g.sage_F) : kk;
                    v = g.A[(size_t)ar * g.lda + ca];
                }
                As[k][r] = v;
            }
            for (int idx = threadIdx.x; idx < GT_K * N; idx += 256) {
                const int k = idx / N, c = idx % N;
                Ws[k][c] = (k0 + k < g.K) ? g.W[(size_t)(k0 + k) * N + c] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < GT_K; ++k) {
                float a[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    const float w = Ws[k][tx + 16 * c];
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i][c] = fmaf(a[i], w, acc[i][c]);
                }
            }
            __syncthreads();
        }
        // epilogue
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = m0 + ty * 4 + i;
            float dot = 0.f;
            float sc = 1.f;
            if (g.row_scale && row < g.M) sc = g.row_scale[g.rows ? g.rows[row] : row];
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                const int col = tx + 16 * c;
                float v = acc[i][c] + (g.bias ? g.bias[col] : 0.f);
                if (g.relu) v = fmaxf(v, 0.f);
                if (g.pool) dot += v * g.pool[col];
                if (g.C && row < g.M) g.C[(size_t)row * N + col] = v * sc;
            }
            if (g.pool) red[ty * 4 + i][tx] = dot;
        }
        if (g.pool) {
            __syncthreads();
            if (threadIdx.x < GT_M) {
                const int row = m0 + threadIdx.x;
                float d = 0.f;
#pragma unroll
                for (int t = 0; t < 16; ++t) d += red[threadIdx.x][t];
                if (row < g.M) g.score[row] = tanhf(d / pnorm);
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// GCNConv aggregation on XW: H[i] = relu( sum_{j->i, j != i} dis_j dis_i XW[j] + dis_i^2 XW[i] + b ), score
// deg_i = 1 + #{j -> i, j != i};  warp per row, lane owns a float4 (width 128) -- generic width via loop
// ------------------------------------------------------------------------------------------------
__global__ void k_gcn_deg(const int *__restrict__ row_ptr, const int *__restrict__ col, int n, float *__restrict__ dis)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        int d = 1;
        for (int i = row_ptr[r]; i < row_ptr[r + 1]; ++i) d += (col[i] != r);
        dis[r] = powf((float)d, -0.5f);
    }
}

__global__ void __launch_bounds__(256) k_gcn_rows(const float *__restrict__ XW, int W, const int *__restrict__ row_ptr,
                                                  const int *__restrict__ col, const float *__restrict__ dis, int n,
                                                  const float *__restrict__ bias, const float *__restrict__ pool,
                                                  float *__restrict__ H, float *__restrict__ score)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float pn = 0.f;
    for (int c = lane; c < W; c += 32) pn += pool[c] * pool[c];
#pragma unroll
    for (int o = 16; o; o >>= 1) pn += __shfl_xor_sync(FULL, pn, o);
    pn = sqrtf(pn);
    for (int r = warp; r < n; r += nwarps) {
        const int a = row_ptr[r], b = row_ptr[r + 1];
        const float di = dis[r];
        float dot = 0.f;
        for (int c = lane; c < W; c += 32) {
            float s = 0.f;
            for (int i = a; i < b; ++i) {
                const int j = col[i];
                if (j != r) s += (dis[j] * di) * XW[(size_t)j * W + c];
            }
            s += (di * di) * XW[(size_t)r * W + c];
            float v = fmaxf(s + bias[c], 0.f);
            H[(size_t)r * W + c] = v;
            dot += v * pool[c];
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(FULL, dot, o);
        if (lane == 0) score[r] = tanhf(dot / pn);
    }
}

// ------------------------------------------------------------------------------------------------
// TopK: stable LSD radix sort of (descending score, ascending index), 8-bit digits
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned desc_key(float s)
{
    s += 0.0f;  // -0 -> +0: torch.sort treats them as equal
    unsigned u = __float_as_uint(s);
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;  // ascending order-preserving
    return ~u;                                   // descending
}

constexpr int RS_THREADS = 256, RS_ITEMS = 2, RS_TILE = RS_THREADS * RS_ITEMS;  // 512 keys per CTA: short per-warp chains (2 match_any steps), more CTAs in flight
// table[digit][cta]
__global__ void __launch_bounds__(RS_THREADS) k_radix_hist(const unsigned *__restrict__ key, int n, int shift,
                                                           int *__restrict__ table, int ncta)
{
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
    for (int q = 0; q < RS_ITEMS; ++q) {
        const int i = base + q * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(key[i] >> shift) & 255], 1);
    }
    __syncthreads();
    table[threadIdx.x * ncta + blockIdx.x] = h[threadIdx.x];
}

// Stable scatter: warp w of the CTA owns the contiguous chunk [w*512, w*512+512) of the tile and walks it 32 keys at
// a time; ranks inside a 32-key group come from match_any, offsets across groups / warps from counters.
template <bool INLINE_SCAN>
__global__ void __launch_bounds__(RS_THREADS) k_radix_scatter(const unsigned *__restrict__ key, const int *__restrict__ val,
                                                              int n, int shift, const int *__restrict__ table_ex, int ncta,
                                                              unsigned *__restrict__ key_out, int *__restrict__ val_out)
{
    constexpr int NW = RS_THREADS / 32, PER_WARP = RS_TILE / NW;
    __shared__ int cnt[NW][256];     // per-warp digit counts, then per-warp exclusive bases
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = lane; d < 256; d += 32) cnt[warp][d] = 0;
    __syncwarp();
    const int wbase = blockIdx.x * RS_TILE + warp * PER_WARP;
    for (int q = 0; q < PER_WARP; q += 32) {
        if (wbase + q >= n) break;                       // warp-uniform: nothing left in this chunk
        const int i = wbase + q + lane;
        const int d = (i < n) ? (int)((key[i] >> shift) & 255) : 256 + lane;
        const unsigned peers = __match_any_sync(FULL, d);
        if (i < n && (peers & ((1u << lane) - 1u)) == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // bases: global (table_ex[digit][cta]) + digits counted by lower warps of this CTA
    {
        const int d = threadIdx.x;  // 256 threads <-> 256 digits
        int run;
        if (INLINE_SCAN) {
            // table_ex holds the RAW counts table[digit][cta]: every CTA forms its own exclusive bases (digit-major
            // prefix) from it -- ncta loads per thread and one block scan -- instead of a scan launch per pass
            __shared__ int wsum[NW];
            int tot = 0, before = 0;
            for (int c = 0; c < ncta; ++c) {
                const int v = table_ex[d * ncta + c];
                before += (c < (int)blockIdx.x) ? v : 0;
                tot += v;
            }
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            int wbase2 = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) wbase2 += (w < warp) ? wsum[w] : 0;
            run = wbase2 + incl - tot + before;
        } else {
            run = table_ex[d * ncta + blockIdx.x];
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int q = 0; q < PER_WARP; q += 32) {
        if (wbase + q >= n) break;
        const int i = wbase + q + lane;
        unsigned k = 0;
        int d = 256 + lane;
        if (i < n) {
            k = key[i];
            d = (int)((k >> shift) & 255);
        }
        const unsigned peers = __match_any_sync(FULL, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (i < n) {
            const int pos = cnt[warp][d] + rank;
            key_out[pos] = k;
            val_out[pos] = val[i];
        }
        __syncwarp();
        if (i < n && rank == 0) cnt[warp][d] += __popc(peers);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// TopK selection: only the k = ceil(ratio n) best rows need ordering.  A 4-pass MSD radix SELECT finds the k-th key T,
// an ordered compaction keeps {key < T} plus the first (k - #{key < T}) rows with key == T in index order (the
// tie rule), and the LSD sort above orders just those k pairs.
// ------------------------------------------------------------------------------------------------
struct SelState {
    unsigned prefix, mask;
    int remaining, count_lt;
};
__global__ void k_select_init(const float *__restrict__ score, int n, unsigned *__restrict__ key, SelState *st, int k,
                              int *__restrict__ hist)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) key[i] = desc_key(score[i]);
    if (blockIdx.x == 0) {
        hist[threadIdx.x & 255] = 0;
        if (threadIdx.x == 0) {
            st->prefix = 0u; st->mask = 0u; st->remaining = k; st->count_lt = 0;
            reinterpret_cast<int *>(st)[4] = 0;   // ticket of k_select_hist's last-CTA pick
        }
    }
}
// One launch per 8-bit digit: every CTA histograms the keys still matching the prefix; the LAST CTA to finish (ticket)
// picks the digit whose cumulative count first reaches `remaining`, extends the prefix and re-zeroes the histogram
// (a separate single-CTA pick launch per pass cost as much as the histogram itself on the small levels).
__global__ void __launch_bounds__(256) k_select_hist(const unsigned *__restrict__ key, int n, int shift, SelState *st,
                                                     int *__restrict__ hist)
{
    __shared__ int h[256];
    __shared__ int wsum[8];
    __shared__ int is_last;
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned prefix = st->prefix, mask = st->mask;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned k = key[i];
        if ((k & mask) == prefix) atomicAdd(&h[(k >> shift) & 255], 1);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(hist + threadIdx.x, h[threadIdx.x]);
    __threadfence();
    __syncthreads();
    int *ticket = reinterpret_cast<int *>(st) + 4;
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = __ldcg(hist + threadIdx.x);
    hist[threadIdx.x] = 0;   // ready for the next pass
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += wsum[w];
    incl += off;
    const int need = st->remaining;
    __syncthreads();
    // the digit whose cumulative count first reaches `need` (the last digit if rounding left it short)
    const bool hit = (incl >= need && incl - c < need) || (threadIdx.x == 255 && incl < need);
    if (hit) {
        const int cum = incl - c;
        st->prefix = prefix | ((unsigned)threadIdx.x << shift);
        st->mask = mask | (255u << shift);
        st->remaining = need - cum;
        st->count_lt += cum;
    }
    if (threadIdx.x == 0) *ticket = 0;
}
constexpr int SC_TILE = 4096;
__global__ void __launch_bounds__(1024) k_select_count(const unsigned *__restrict__ key, int n, const SelState *__restrict__ st,
                                                       int *__restrict__ part_lt, int *__restrict__ part_eq)
{
    __shared__ int wtmp[33];
    const unsigned T = st->prefix;
    const int base = blockIdx.x * SC_TILE + threadIdx.x * 4;
    int lt = 0, eq = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (base + q < n) {
            const unsigned k = key[base + q];
            lt += (k < T);
            eq += (k == T);
        }
    int tl, te;
    block_scan_excl_1024(lt, wtmp, tl);
    block_scan_excl_1024(eq, wtmp, te);
    if (threadIdx.x == 0) {
        part_lt[blockIdx.x] = tl;
        part_eq[blockIdx.x] = te;
    }
}
__global__ void __launch_bounds__(1024) k_select_write(const unsigned *__restrict__ key, int n, const SelState *__restrict__ st,
                                                       const int *__restrict__ lt_ex, const int *__restrict__ eq_ex,
                                                       unsigned *__restrict__ key_out, int *__restrict__ val_out)
{
    __shared__ int wtmp[33];
    const unsigned T = st->prefix;
    const int m = st->remaining;             // rows with key == T to keep (lowest indices first)
    const int base = blockIdx.x * SC_TILE + threadIdx.x * 4;
    unsigned k[4];
    int lt = 0, eq = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        k[q] = (base + q < n) ? key[base + q] : 0xffffffffu;
        if (base + q < n) {
            lt += (k[q] < T);
            eq += (k[q] == T);
        }
    }
    int tl, te;
    int lb = block_scan_excl_1024(lt, wtmp, tl) + lt_ex[blockIdx.x];
    int eb = block_scan_excl_1024(eq, wtmp, te) + eq_ex[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (base + q < n) {
            const bool isl = k[q] < T, ise = k[q] == T;
            if (isl || (ise && eb < m)) {
                const int pos = lb + min(eb, m);
                key_out[pos] = k[q];
                val_out[pos] = base + q;
            }
            lb += isl;
            eb += ise;
        }
}

// ------------------------------------------------------------------------------------------------
// pooling: inverse map, gather * score, ordered edge filter, readout
// ------------------------------------------------------------------------------------------------
__global__ void k_fill_i32(int *__restrict__ p, int n, int v)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_inverse_perm(const int *__restrict__ perm, int k, int *__restrict__ inv)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < k; p += gridDim.x * blockDim.x) inv[perm[p]] = p;
}
// xp[p] = H[perm[p]] * score[perm[p]]   (float4 lanes; W multiple of 4)
__global__ void k_pool_gather(const float *__restrict__ H, int W, const int *__restrict__ perm, const float *__restrict__ score,
                              int k, float *__restrict__ xp)
{
    const int w4 = W / 4;
    const unsigned total = (unsigned)k * (unsigned)w4;
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int p = (int)(idx / (unsigned)w4), c = (int)(idx % (unsigned)w4);
        const int src = perm[p];
        const float s = score[src];
        float4 v = reinterpret_cast<const float4 *>(H + (size_t)src * W)[c];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        reinterpret_cast<float4 *>(xp + (size_t)p * W)[c] = v;
    }
}

// ordered compaction of the edges whose endpoints both survive: per-CTA counts, scan, write
constexpr int EF_TILE = 1024 * 4;
__global__ void __launch_bounds__(1024) k_edge_count(const int *__restrict__ src, const int *__restrict__ dst,
                                                     const int *__restrict__ ecount, const int *__restrict__ inv,
                                                     int *__restrict__ part)
{
    __shared__ int wtmp[33];
    const int ne = *ecount;
    const int base = blockIdx.x * EF_TILE + threadIdx.x * 4;
    int c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = base + q;
        if (e < ne) c += (inv[src[e]] >= 0 && inv[dst[e]] >= 0);
    }
    int total;
    block_scan_excl_1024(c, wtmp, total);
    if (threadIdx.x == 0) part[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) k_edge_write(const int *__restrict__ src, const int *__restrict__ dst,
                                                     const int *__restrict__ ecount, const int *__restrict__ inv,
                                                     const int *__restrict__ part_ex, int nparts, int *__restrict__ src_out,
                                                     int *__restrict__ dst_out, int *__restrict__ ecount_out)
{
    __shared__ int wtmp[33];
    const int ne = *ecount;
    const int base = blockIdx.x * EF_TILE + threadIdx.x * 4;
    int s2[4], d2[4], c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int e = base + q;
        s2[q] = d2[q] = -1;
        if (e < ne) {
            s2[q] = inv[src[e]];
            d2[q] = inv[dst[e]];
        }
        c += (s2[q] >= 0 && d2[q] >= 0);
    }
    int total;
    int pos = block_scan_excl_1024(c, wtmp, total) + part_ex[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (s2[q] >= 0 && d2[q] >= 0) {
            src_out[pos] = s2[q];
            dst_out[pos] = d2[q];
            ++pos;
        }
    if (blockIdx.x == 0 && threadIdx.x == 0) *ecount_out = part_ex[nparts];
}

// readout: acc[0:W] += max over rows, acc[W:2W] += mean over rows.  Two stages, fixed shape -> deterministic.
constexpr int RO_ROWS = 128;   // rows per CTA in stage 1
// 256 threads: lane group c4 = tid % (W/4) owns a float4 column group, the 256/(W/4) row slices interleave rows
__global__ void __launch_bounds__(256) k_readout_partial(const float *__restrict__ X, int W, int n, float *__restrict__ pmax,
                                                         float *__restrict__ psum)
{
    __shared__ float4 smx[256], ssm[256];
    const int w4 = W / 4, slices = 256 / w4;
    const int c4 = threadIdx.x % w4, sl = threadIdx.x / w4;
    const int r0 = blockIdx.x * RO_ROWS, r1 = min(n, r0 + RO_ROWS);
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), sm = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sl < slices)
        for (int r = r0 + sl; r < r1; r += slices) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(X + (size_t)r * W) + c4);
            mx.x = fmaxf(mx.x, v.x); mx.y = fmaxf(mx.y, v.y); mx.z = fmaxf(mx.z, v.z); mx.w = fmaxf(mx.w, v.w);
            sm.x += v.x; sm.y += v.y; sm.z += v.z; sm.w += v.w;
        }
    smx[threadIdx.x] = mx;
    ssm[threadIdx.x] = sm;
    __syncthreads();
    if (threadIdx.x < w4) {
        for (int q = 1; q < slices; ++q) {      // fixed order -> deterministic
            const float4 a = smx[q * w4 + c4], b = ssm[q * w4 + c4];
            mx.x = fmaxf(mx.x, a.x); mx.y = fmaxf(mx.y, a.y); mx.z = fmaxf(mx.z, a.z); mx.w = fmaxf(mx.w, a.w);
            sm.x += b.x; sm.y += b.y; sm.z += b.z; sm.w += b.w;
        }
        reinterpret_cast<float4 *>(pmax + (size_t)blockIdx.x * W)[c4] = mx;
        reinterpret_cast<float4 *>(psum + (size_t)blockIdx.x * W)[c4] = sm;
    }
}
__global__ void __launch_bounds__(1024) k_readout_final(const float *__restrict__ pmax, const float *__restrict__ psum, int W,
                                                        int nparts, int n, int first, float *__restrict__ acc)
{
    __shared__ float smx[1024], ssm[1024];
    const int slices = 1024 / W;
    const int c = threadIdx.x % W, sl = threadIdx.x / W;
    float mx = -INFINITY, sm = 0.f;
    if (sl < slices)
        for (int p = sl; p < nparts; p += slices) {
            mx = fmaxf(mx, pmax[(size_t)p * W + c]);
            sm += psum[(size_t)p * W + c];
        }
    smx[threadIdx.x] = mx;
    ssm[threadIdx.x] = sm;
    __syncthreads();
    if (threadIdx.x < W) {
        for (int q = 1; q < slices; ++q) {
            mx = fmaxf(mx, smx[q * W + c]);
            sm += ssm[q * W + c];
        }
        const float mean = sm / (float)max(n, 1);
        acc[c] = first ? mx : acc[c] + mx;
        acc[W + c] = first ? mean : acc[W + c] + mean;
    }
}

// ------------------------------------------------------------------------------------------------
// MLP head: lin1 -> ReLU -> lin2 -> ReLU -> lin3 -> softmax -> argmax (one CTA, B = 1)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mlp_head(const mdq_net_t net, const float *__restrict__ params,
                                                  const float *__restrict__ acc, float *__restrict__ out,
                                                  float *__restrict__ embedding, int *__restrict__ argmax)
{
    __shared__ float a[512], b[512];
    __shared__ float red[256];
    __shared__ int redi[256];
    const int in0 = net.lin_in[0];
    for (int i = threadIdx.x; i < in0; i += blockDim.x) {
        a[i] = acc[i];
        if (embedding) embedding[i] = acc[i];
    }
    __syncthreads();
    float *cur = a, *nxt = b;
    __shared__ float part[8][512];
    const int wq = threadIdx.x >> 5, lq = threadIdx.x & 31;
    for (int l = 0; l < 3; ++l) {
        const int K = net.lin_in[l], O = net.lin_out[l];
        const float *Wt = params + net.lin_off[l];   // [K][O]
        const float *bs = params + net.lin_boff[l];
        // warp wq takes every 8th k; lanes run over the outputs (coalesced weight rows), 4 outputs per lane in flight
        for (int o0 = 0; o0 < O; o0 += 128) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
            const int o = o0 + lq;
            for (int k = wq; k < K; k += 8) {
                const float x = cur[k];
                const float *wr = Wt + (size_t)k * O;
                if (o < O) s0 = fmaf(x, __ldg(wr + o), s0);
                if (o + 32 < O) s1 = fmaf(x, __ldg(wr + o + 32), s1);
                if (o + 64 < O) s2 = fmaf(x, __ldg(wr + o + 64), s2);
                if (o + 96 < O) s3 = fmaf(x, __ldg(wr + o + 96), s3);
            }
            part[wq][o] = s0;
            if (o + 32 < 512) part[wq][o + 32] = s1;
            if (o + 64 < 512) part[wq][o + 64] = s2;
            if (o + 96 < 512) part[wq][o + 96] = s3;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < O; o += blockDim.x) {
            float s = part[0][o];
#pragma unroll
            for (int q = 1; q < 8; ++q) s += part[q][o];
            s += bs[o];
            nxt[o] = (l < 2) ? fmaxf(s, 0.f) : s;
        }
        __syncthreads();
        float *t = cur; cur = nxt; nxt = t;
    }
    const int O = net.out_dim;
    // softmax + first-max argmax
    float mx = -INFINITY;
    int mi = 0x7fffffff;
    for (int o = threadIdx.x; o < O; o += blockDim.x)
        if (cur[o] > mx) { mx = cur[o]; mi = o; }
    red[threadIdx.x] = mx;
    redi[threadIdx.x] = mi;
    __syncthreads();
    for (int s = 128; s; s >>= 1) {
        if (threadIdx.x < s) {
            const float om = red[threadIdx.x + s];
            const int oi = redi[threadIdx.x + s];
            if (om > red[threadIdx.x] || (om == red[threadIdx.x] && oi < redi[threadIdx.x])) {
                red[threadIdx.x] = om;
                redi[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    mx = red[0];
    const int best = redi[0];
    __syncthreads();
    if (net.softmax) {
        float s = 0.f;
        for (int o = threadIdx.x; o < O; o += blockDim.x) s += expf(cur[o] - mx);
        red[threadIdx.x] = s;
        __syncthreads();
        for (int st = 128; st; st >>= 1) {
            if (threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
            __syncthreads();
        }
        const float tot = red[0];
        for (int o = threadIdx.x; o < O; o += blockDim.x) out[o] = expf(cur[o] - mx) / tot;
    } else {
        for (int o = threadIdx.x; o < O; o += blockDim.x) out[o] = cur[o];
    }
    if (argmax && threadIdx.x == 0) *argmax = best;
}

#include "node_gemm_tc.cuh"

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int launch_gemm(const GemmArgs &g, int mode, const float *wsplit, cudaStream_t st)
{
    if (g.M <= 0) return MDQ_OK;
    if (mode == 1 && g.N == 128) return launch_gemm_tc(g, wsplit, st);
    const int grid = grid_for(g.M, GT_M, 148 * 8);
    switch (g.N) {
    case 64: k_node_gemm_f32<64><<<grid, 256, 0, st>>>(g); break;
    case 128: k_node_gemm_f32<128><<<grid, 256, 0, st>>>(g); break;
    default:
        mdq::set_error("node GEMM: conv width %d not supported by the layered path (64 or 128)", g.N);
        return MDQ_EINVAL;
    }
    return mdq::check_launch("k_node_gemm_f32");
}

int build_csr(const int *src, const int *dst, const int *ecount, int ecap, int n, int *deg, int *row_ptr, int *cursor,
              int *eid, int *col, int *scan_part, cudaStream_t st)
{
    int rc;
    cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)(n + 1), st);
    cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)(n + 1), st);
    k_csr_count<<<grid_for(ecap, 256), 256, 0, st>>>(dst, ecount, deg);
    if ((rc = mdq::check_launch("k_csr_count"))) return rc;
    if ((rc = scan_i32(deg, row_ptr, n, scan_part, st))) return rc;
    k_csr_fill<<<grid_for(ecap, 256), 256, 0, st>>>(dst, ecount, row_ptr, cursor, eid);
    if ((rc = mdq::check_launch("k_csr_fill"))) return rc;
    k_csr_sort_rows<<<grid_for(n, 128), 128, 0, st>>>(src, row_ptr, n, eid, col);
    return mdq::check_launch("k_csr_sort_rows");
}

int sage_rows(const float *X, int ldx, int col0, int F, const int *row_ptr, const int *col, int n, float *A, int lda,
              cudaStream_t st)
{
    if (n <= 0) return MDQ_OK;
    const int Fp = (F + 3) & ~3;
    if (lda < 2 * Fp || (lda & 3) || (reinterpret_cast<uintptr_t>(A) & 15)) {
        mdq::set_error("SAGE aggregation: lda %d must be a multiple of 4 and >= 2*%d, A 16-byte aligned", lda, Fp);
        return MDQ_EINVAL;
    }
    if (F <= 32) {
        const int ch = Fp >> 2;
        if ((long long)n * ch >= (1LL << 31)) {
            mdq::set_error("SAGE aggregation: %d rows x %d chunks exceeds the 32-bit element index", n, ch);
            return MDQ_EINVAL;
        }
        int rc;
        const int grid = grid_for((long long)n * ch, 256, 148 * 64);
        if (ch == 5) k_pad_rows<5><<<grid, 256, 0, st>>>(X, ldx, col0, F, n, ch, A, lda);
        else k_pad_rows<0><<<grid, 256, 0, st>>>(X, ldx, col0, F, n, ch, A, lda);
        if ((rc = mdq::check_launch("k_pad_rows"))) return rc;
        if (ch == 5) k_sage_rows_vec4<5><<<grid, 256, 0, st>>>(row_ptr, col, n, ch, A, lda);
        else k_sage_rows_vec4<0><<<grid, 256, 0, st>>>(row_ptr, col, n, ch, A, lda);
        return mdq::check_launch("k_sage_rows_vec4");
    }
    if (F == 128 && ldx == 128 && col0 == 0 && lda == 256 && !(reinterpret_cast<uintptr_t>(X) & 15)) {
        k_sage_rows_128<<<grid_for((long long)n * 32, 256, 148 * 16), 256, 0, st>>>(X, row_ptr, col, n, A);
        return mdq::check_launch("k_sage_rows_128");
    }
    mdq::set_error("SAGE aggregation: %d input features not supported by the layered path (<= 32 or 128)", F);
    return MDQ_EINVAL;
}

// LSD radix sort of n (key, val) pairs already in key_a / val_a; the ordered values end up in *val_sorted
int radix_sort_pairs(int n, unsigned *key_a, unsigned *key_b, int *val_a, int *val_b, int *table, int *scan_part,
                     int **val_sorted, cudaStream_t st)
{
    int rc;
    const int ncta = cdiv(n, RS_TILE);
    unsigned *kin = key_a, *kout = key_b;
    int *vin = val_a, *vout = val_b;
    for (int shift = 0; shift < 32; shift += 8) {
        k_radix_hist<<<ncta, RS_THREADS, 0, st>>>(kin, n, shift, table, ncta);
        if ((rc = mdq::check_launch("k_radix_hist"))) return rc;
        if (ncta <= 256) {   // every CTA scans the small count table itself: no scan launch
            k_radix_scatter<true><<<ncta, RS_THREADS, 0, st>>>(kin, vin, n, shift, table, ncta, kout, vout);
        } else {
            if ((rc = scan_i32(table, table + 256 * ncta + 8, 256 * ncta, scan_part, st))) return rc;
            k_radix_scatter<false><<<ncta, RS_THREADS, 0, st>>>(kin, vin, n, shift, table + 256 * ncta + 8, ncta, kout, vout);
        }
        if ((rc = mdq::check_launch("k_radix_scatter"))) return rc;
        unsigned *tk = kin; kin = kout; kout = tk;
        int *tv = vin; vin = vout; vout = tv;
    }
    *val_sorted = vin;
    return MDQ_OK;
}

// TopKPooling's perm: the k best rows by (score desc, index asc), in that order.
//   key_all [n]; key_a/key_b/val_a/val_b [>= k]; sel: SelState + 256-int histogram + 4 * (cdiv(n, SC_TILE) + 2) ints
// Small levels (n <= TS_MAX rows): ONE single-CTA launch instead of the ~17 of select + sort -- bitonic sort of the
// 64-bit composites (descending-score key << 32 | row index) in shared memory, the first k of the ascending order are
// the kept rows in exactly the order select + stable LSD sort produce (score descending, ties by lower index).
constexpr int TS_MAX = 8192, TS_THREADS = 1024;
__global__ void __launch_bounds__(TS_THREADS) k_topk_small(const float *__restrict__ score, int n, int k, int n2,
                                                           int *__restrict__ perm)
{
    extern __shared__ unsigned long long ts_arr[];
    for (int i = threadIdx.x; i < n2; i += TS_THREADS)
        ts_arr[i] = i < n ? (((unsigned long long)desc_key(score[i]) << 32) | (unsigned)i) : ~0ull;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (n2 >> 1); t += TS_THREADS) {
                const int lo = ((t & ~(stride - 1)) << 1) | (t & (stride - 1));   // element whose `stride` bit is 0
                const int hi = lo | stride;
                const bool up = (lo & size) == 0;
                const unsigned long long a = ts_arr[lo], b = ts_arr[hi];
                if ((a > b) == up) { ts_arr[lo] = b; ts_arr[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int r = threadIdx.x; r < k; r += TS_THREADS) perm[r] = (int)(unsigned)(ts_arr[r] & 0xffffffffull);
}

int topk_select_sort(const float *score, int n, int k, unsigned *key_all, unsigned *key_a, unsigned *key_b, int *val_a,
                     int *val_b, int *table, int *scan_part, int *sel, int **perm, cudaStream_t st)
{
    int rc;
    if (n <= TS_MAX) {
        int n2 = 2;
        while (n2 < n) n2 <<= 1;
        static bool configured = false;
        if (!configured) {
            cudaFuncSetAttribute(k_topk_small, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_MAX * 8);
            configured = true;
        }
        k_topk_small<<<1, TS_THREADS, (size_t)n2 * 8, st>>>(score, n, k, n2, val_a);
        *perm = val_a;
        return mdq::check_launch("k_topk_small");
    }
    SelState *state = reinterpret_cast<SelState *>(sel);
    int *hist = sel + 8;
    const int nparts = cdiv(n, SC_TILE);
    int *part_lt = hist + 256, *part_eq = part_lt + nparts + 1, *lt_ex = part_eq + nparts + 1, *eq_ex = lt_ex + nparts + 1;
    k_select_init<<<grid_for(n, 256), 256, 0, st>>>(score, n, key_all, state, k, hist);
    if ((rc = mdq::check_launch("k_select_init"))) return rc;
    for (int shift = 24; shift >= 0; shift -= 8) {
        k_select_hist<<<grid_for(n, 256 * 8, 148 * 4), 256, 0, st>>>(key_all, n, shift, state, hist);
        if ((rc = mdq::check_launch("k_select_hist"))) return rc;
    }
    k_select_count<<<nparts, 1024, 0, st>>>(key_all, n, state, part_lt, part_eq);
    if ((rc = mdq::check_launch("k_select_count"))) return rc;
    k_scan<<<1, 1024, 0, st>>>(part_lt, lt_ex, nparts);
    if ((rc = mdq::check_launch("k_scan"))) return rc;
    k_scan<<<1, 1024, 0, st>>>(part_eq, eq_ex, nparts);
    if ((rc = mdq::check_launch("k_scan"))) return rc;
    k_select_write<<<nparts, 1024, 0, st>>>(key_all, n, state, lt_ex, eq_ex, key_a, val_a);
    if ((rc = mdq::check_launch("k_select_write"))) return rc;
    return radix_sort_pairs(k, key_a, key_b, val_a, val_b, table, scan_part, perm, st);
}

struct Bump {
    char *base;
    size_t off, cap;
    template <class T> T *take(size_t n)
    {
        off = (off + 255) & ~(size_t)255;
        T *p = reinterpret_cast<T *>(base ? base + off : nullptr);
        off += n * sizeof(T);
        return p;
    }
};

int pool_count(float ratio, int n)
{
    // PyG: (ratio * num_nodes.to(torch.float)).ceil().long() -- float32 arithmetic
    const float k = ceilf(ratio * (float)n);
    int ki = (int)k;
    return ki < 0 ? 0 : (ki > n ? n : ki);
}

// Runs the forward when ws.base != nullptr; with a null base it only sizes the workspace.
int forward_layered(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                    const int64_t *edge_src, const int64_t *edge_dst, int n0, int e0, int mode, float *out,
                    float *embedding, int32_t *argmax, Bump &ws, cudaStream_t st)
{
    const int W = net->width;
    const bool run = ws.base != nullptr;
    int rc;
    // persistent across levels
    float *acc = ws.take<float>(2 * W);
    int *scan_part = ws.take<int>(2 * cdiv(std::max(n0 + 1, 256 * cdiv(n0, RS_TILE) + 8), SCAN_TILE) + 8);
    int *src[2], *dst[2], *ecount[2];
    for (int i = 0; i < 2; ++i) {
        src[i] = ws.take<int>(e0 + 1);
        dst[i] = ws.take<int>(e0 + 1);
        ecount[i] = ws.take<int>(1);
    }
    int *deg = ws.take<int>(n0 + 2), *row_ptr = ws.take<int>(n0 + 2), *cursor = ws.take<int>(n0 + 2);
    int *eid = ws.take<int>(e0 + 1), *col = ws.take<int>(e0 + 1);
    float *score = ws.take<float>(n0 + 1);
    unsigned *key_all = ws.take<unsigned>(n0 + 1);
    int *sel = ws.take<int>(8 + 256 + 4 * (cdiv(n0, SC_TILE) + 2));
    unsigned *key_a = ws.take<unsigned>(n0 + 1), *key_b = ws.take<unsigned>(n0 + 1);
    int *val_a = ws.take<int>(n0 + 1), *val_b = ws.take<int>(n0 + 1);
    int *table = ws.take<int>(2 * (256 * cdiv(n0, RS_TILE) + 8) + 8);
    int *inv = ws.take<int>(n0 + 1);
    int *epart = ws.take<int>(2 * cdiv(e0 + 1, EF_TILE) + 4);
    float *dis = ws.take<float>(n0 + 1);
    const int n1 = pool_count(net->ratio, n0);
    float *pmax = ws.take<float>((size_t)cdiv(std::max(n1, 1), RO_ROWS) * W);
    float *psum = ws.take<float>((size_t)cdiv(std::max(n1, 1), RO_ROWS) * W);
    // level buffers: A (message-passing output / GEMM input), H (conv output), xp ping-pong (pooled features)
    const int kin0 = net->blk[0].kin;
    const int lda0 = (net->blk[0].type == MDQ_BLOCK_SAGE) ? ((2 * ((kin0 + 3) & ~3) + 7) / 8 * 8) : ((kin0 + 7) / 8 * 8);
    float *A0 = ws.take<float>((size_t)n0 * lda0);
    float *A1 = ws.take<float>((size_t)n1 * 2 * W);       // SAGE input of levels >= 1
    float *H = ws.take<float>((size_t)n1 * W);            // conv output of levels >= 1 / XW of GCN levels
    float *H2 = ws.take<float>((size_t)n1 * W);
    float *xp[2] = {ws.take<float>((size_t)n1 * W), ws.take<float>((size_t)n1 * W)};
    if (!run) return MDQ_OK;

    k_edges_to_i32<<<grid_for(e0, 256), 256, 0, st>>>(reinterpret_cast<const long long *>(edge_src),
                                                      reinterpret_cast<const long long *>(edge_dst), e0, src[0], dst[0],
                                                      ecount[0]);
    if ((rc = mdq::check_launch("k_edges_to_i32"))) return rc;

    int n = n0, ecap = e0, cur = 0;
    const float *X = x;       // current node features
    int ldx = net->x_stride, col0 = net->in_col0;
    for (int l = 0; l < net->n_blocks; ++l) {
        const mdq_block_t &b = net->blk[l];
        const float *Wl = params + b.w_off, *bias = params + b.b_off, *pool = params + b.pool_off;
        const float *wsp = wsplit ? wsplit + 3 * (size_t)b.w_off : nullptr;   // see pack_wsplit()
        const int k = pool_count(net->ratio, n);
        if ((rc = build_csr(src[cur], dst[cur], ecount[cur], ecap, n, deg, row_ptr, cursor, eid, col, scan_part, st)))
            return rc;
        int *perm = nullptr;
        float *xnext = xp[l & 1];
        if (b.type == MDQ_BLOCK_SAGE) {
            const int F = b.kin;
            float *A = (l == 0) ? A0 : A1;
            const int lda = (l == 0) ? lda0 : 2 * W;
            if (l > 0 && F != W) { mdq::set_error("layered path: SAGE block %d has kin %d != width", l, F); return MDQ_EINVAL; }
            if ((rc = sage_rows(X, ldx, col0, F, row_ptr, col, n, A, lda, st))) return rc;
            // pass 1: scores of all rows (the conv output itself is only needed for the kept rows)
            GemmArgs g{};
            g.A = A; g.rows = nullptr; g.lda = lda; g.K = 2 * F; g.M = n; g.N = W; g.W = Wl; g.bias = bias; g.pool = pool;
            g.sage_F = F; g.sage_Fp = (F + 3) & ~3;
            g.row_scale = nullptr; g.relu = 1; g.C = nullptr; g.score = score;
            if ((rc = launch_gemm(g, mode, wsp, st))) return rc;
            if ((rc = topk_select_sort(score, n, k, key_all, key_a, key_b, val_a, val_b, table, scan_part, sel, &perm, st))) return rc;
            // pass 2: conv output of the kept rows, scaled by their score: x[perm] * score[perm]
            g.rows = perm; g.M = k; g.pool = nullptr; g.score = nullptr; g.row_scale = score; g.C = xnext;
            if ((rc = launch_gemm(g, mode, wsp, st))) return rc;
        } else {
            if (b.kin != W && l > 0) { mdq::set_error("layered path: GCN block %d has kin %d != width", l, b.kin); return MDQ_EINVAL; }
            // XW = X . W (no bias / activation), then normalised aggregation with self loops
            GemmArgs g{};
            const float *Ain = X;
            int lda = ldx;
            if (col0 != 0) { mdq::set_error("layered path: GCN first block with a column offset is not supported"); return MDQ_EINVAL; }
            g.A = Ain; g.rows = nullptr; g.lda = lda; g.K = b.kin; g.M = n; g.N = W; g.W = Wl; g.bias = nullptr; g.pool = nullptr;
            g.row_scale = nullptr; g.relu = 0; g.C = H; g.score = nullptr;
            if ((rc = launch_gemm(g, (lda % 4 == 0) ? mode : 0, wsp, st))) return rc;
            k_gcn_deg<<<grid_for(n, 256), 256, 0, st>>>(row_ptr, col, n, dis);
            if ((rc = mdq::check_launch("k_gcn_deg"))) return rc;
            k_gcn_rows<<<grid_for((long long)n * 32, 256, 148 * 16), 256, 0, st>>>(H, W, row_ptr, col, dis, n, bias, pool, H2, score);
            if ((rc = mdq::check_launch("k_gcn_rows"))) return rc;
            if ((rc = topk_select_sort(score, n, k, key_all, key_a, key_b, val_a, val_b, table, scan_part, sel, &perm, st))) return rc;
            k_pool_gather<<<grid_for((long long)k * (W / 4), 256), 256, 0, st>>>(H2, W, perm, score, k, xnext);
            if ((rc = mdq::check_launch("k_pool_gather"))) return rc;
        }
        // readout of the pooled level
        const int nparts = cdiv(std::max(k, 1), RO_ROWS);
        k_readout_partial<<<nparts, 256, 0, st>>>(xnext, W, k, pmax, psum);
        if ((rc = mdq::check_launch("k_readout_partial"))) return rc;
        k_readout_final<<<1, 1024, 0, st>>>(pmax, psum, W, nparts, k, l == 0, acc);
        if ((rc = mdq::check_launch("k_readout_final"))) return rc;
        if (l + 1 < net->n_blocks) {
            // re-index the surviving edges (ordered)
            k_fill_i32<<<grid_for(n, 256), 256, 0, st>>>(inv, n, -1);
            if ((rc = mdq::check_launch("k_fill_i32"))) return rc;
            k_inverse_perm<<<grid_for(k, 256), 256, 0, st>>>(perm, k, inv);
            if ((rc = mdq::check_launch("k_inverse_perm"))) return rc;
            const int np = cdiv(ecap + 1, EF_TILE);
            k_edge_count<<<np, 1024, 0, st>>>(src[cur], dst[cur], ecount[cur], inv, epart);
            if ((rc = mdq::check_launch("k_edge_count"))) return rc;
            k_scan<<<1, 1024, 0, st>>>(epart, epart + np + 1, np);
            if ((rc = mdq::check_launch("k_scan"))) return rc;
            k_edge_write<<<np, 1024, 0, st>>>(src[cur], dst[cur], ecount[cur], inv, epart + np + 1, np, src[cur ^ 1],
                                              dst[cur ^ 1], ecount[cur ^ 1]);
            if ((rc = mdq::check_launch("k_edge_write"))) return rc;
            cur ^= 1;
        }
        X = xnext; ldx = W; col0 = 0; n = k;
    }
    k_mlp_head<<<1, 256, 0, st>>>(*net, params, acc, out, embedding, argmax);
    return mdq::check_launch("k_mlp_head");
}

}  // namespace

extern "C" {

int64_t mdq_qnet_layered_workspace_bytes(const mdq_net_t *net, int n_nodes, int n_edges)
{
    if (!net || n_nodes < 1 || n_edges < 0) return -1;
    Bump ws{nullptr, 0, 0};
    forward_layered(net, nullptr, nullptr, nullptr, nullptr, nullptr, n_nodes, n_edges, 0, nullptr, nullptr, nullptr, ws,
                    nullptr);
    return (int64_t)ws.off + 256;
}

int mdq_qnet_forward_layered(const mdq_net_t *net, const float *params, const float *wsplit, const float *x,
                             const int64_t *edge_src, const int64_t *edge_dst, int n_nodes, int n_edges, int gemm_mode,
                             float *out, float *embedding, int32_t *argmax, void *workspace, int64_t workspace_bytes,
                             void *stream)
{
    if (!net || !params || !x || !out || !workspace || n_nodes < 1 || n_edges < 0 || (n_edges > 0 && (!edge_src || !edge_dst)) ||
        net->n_blocks < 1 || net->n_blocks > MDQ_MAX_BLOCKS || (net->width != 64 && net->width != 128) ||
        net->lin_in[0] > 512 || net->lin_out[0] > 512 || net->lin_out[1] > 512 || net->out_dim > 512 ||
        (gemm_mode == 1 && !wsplit)) {
        mdq::set_error("mdq_qnet_forward_layered: bad argument");
        return MDQ_EINVAL;
    }
    if (workspace_bytes < mdq_qnet_layered_workspace_bytes(net, n_nodes, n_edges)) {
        mdq::set_error("mdq_qnet_forward_layered: workspace too small");
        return MDQ_EINVAL;
    }
    Bump ws{reinterpret_cast<char *>(workspace), 0, (size_t)workspace_bytes};
    return forward_layered(net, params, wsplit, x, edge_src, edge_dst, n_nodes, n_edges, gemm_mode, out, embedding, argmax,
                           ws, (cudaStream_t)stream);
}

/* standalone entry points (tests, roofline measurement) */
int mdq_csr_build(const int32_t *src, const int32_t *dst, const int32_t *ecount, int ecap, int n, int32_t *row_ptr,
                  int32_t *col, int32_t *scratch, void *stream)
{
    if ((ecap > 0 && (!src || !dst || !col)) || !ecount || !row_ptr || !scratch || n < 1 || ecap < 0) {
        mdq::set_error("mdq_csr_build: bad argument");
        return MDQ_EINVAL;
    }
    // scratch: deg [n+2] | cursor [n+2] | eid [ecap+1] | scan partials
    int *deg = scratch, *cursor = scratch + (n + 2), *eid = cursor + (n + 2), *part = eid + (ecap + 1);
    return build_csr(src, dst, ecount, ecap, n, deg, row_ptr, cursor, eid, col, part, (cudaStream_t)stream);
}

int64_t mdq_csr_build_scratch_words(int ecap, int n) { return 2LL * (n + 2) + ecap + 1 + 2LL * cdiv(n + 1, SCAN_TILE) + 16; }

int mdq_sage_aggregate(const float *x, int ldx, int col0, int F, const int32_t *row_ptr, const int32_t *col, int n,
                       float *A, int lda, void *stream)
{
    if (!x || !row_ptr || !A || n < 1 || F < 1) {
        mdq::set_error("mdq_sage_aggregate: bad argument");
        return MDQ_EINVAL;
    }
    return sage_rows(x, ldx, col0, F, row_ptr, col, n, A, lda, (cudaStream_t)stream);
}

int mdq_node_gemm(const float *A, const int32_t *rows, int lda, int K, int M, int N, const float *W, const float *wsplit,
                  const float *bias, const float *pool, const float *row_scale, int relu, int gemm_mode, float *C,
                  float *score, void *stream)
{
    if (!A || !W || M < 0 || K < 1 || (pool && !score) || (gemm_mode == 1 && !wsplit)) {
        mdq::set_error("mdq_node_gemm: bad argument");
        return MDQ_EINVAL;
    }
    GemmArgs g{};
    g.A = A; g.rows = rows; g.lda = lda; g.K = K; g.M = M; g.N = N; g.W = W; g.bias = bias; g.pool = pool;
    g.row_scale = row_scale; g.relu = relu; g.C = C; g.score = score;
    return launch_gemm(g, gemm_mode, wsplit, (cudaStream_t)stream);
}

}  // extern "C"
