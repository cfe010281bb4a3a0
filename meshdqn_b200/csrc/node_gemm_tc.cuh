// node_gemm_tc.cuh -- dense node-feature GEMM on the 5th-gen tensor cores (tcgen05, TMEM accumulator), 3xTF32.
//
//   C[M,128] = epilogue( A[rows][K] . W[K][128] )      (the per-vertex Linear layers of SAGEConv / GCNConv,
//                                                        /root/reference/airfoilgcnn.py:94,100,112,118)
//
// fp32 operands are split a = a_hi + a_lo with a_hi = a & 0xffffe000 (exactly representable in TF32), and the
// product is accumulated as a_hi.w_lo + a_lo.w_hi + a_hi.w_hi in fp32 inside TMEM: three kind::tf32 MMAs per
// k-step recover fp32-level accuracy (the dropped a_lo.w_lo term is ~2^-22 relative), which plain TF32 (10-bit
// mantissa, ~1e-3) cannot -- BASELINE.json asks for 1e-5 on the Q-values.
//
// One CTA = 256 threads, 128-row tiles, persistent over tiles.  Per K-block of 16 (two k-steps):
//   * W_hi / W_lo blocks (pre-split and pre-tiled into the canonical K-major core-matrix layout on the host,
//     meshdqn_b200/airfoilgcnn.py) arrive by two bulk TMA copies signalled on an mbarrier;
//   * all threads load their 16-byte pieces of the A block (coalesced LDG.128, optional row gather), split them
//     and store hi / lo into the canonical layout (8-row x 16-byte core matrices, SWIZZLE_NONE);
//   * one thread issues the tcgen05.mma instructions and commits them to the stage's "empty" mbarrier, so the
//     next block's loads overlap the tensor pipe.
// Epilogue: tcgen05.ld (32 lanes x 32 columns per warp), + bias, ReLU, TopK score tanh(h.p/||p||), row scale, store.
#pragma once

namespace tc {

constexpr int TM = 128, TN = 128, KB = 16, STAGES = 2;
constexpr int NCH = KB / 4;                    // 16-byte chunks per row per K-block
constexpr int A_LBO = 2048 + 32;               // chunk stride of the A tiles (padded: conflict-free STS.128)
constexpr int B_LBO = 2048;                    // chunk stride of the W tiles (as laid out in HBM)
constexpr int SBO = 128;                       // 8-row group stride
constexpr int A_TILE = NCH * A_LBO, B_TILE = NCH * B_LBO;
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
constexpr int SMEM_BYTES = 128 + STAGES * STAGE_BYTES + 128 * 2 * 4;

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned phase)
{
    unsigned ok, spins = 0;
    do {
        asm volatile(
            "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
            : "=r"(ok)
            : "r"(s32(bar)), "r"(phase)
            : "memory");
        if (!ok && ++spins > (1u << 22)) __trap();   // a protocol error must fail loudly, never hang the device
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
                 "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(s32(bar))
                 : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell)
__device__ __forceinline__ unsigned long long smem_desc(unsigned addr, unsigned lbo, unsigned sbo)
{
    return (unsigned long long)((addr & 0x3ffffu) >> 4) | ((unsigned long long)(lbo >> 4) << 16) |
           ((unsigned long long)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128 (cute::UMMA::InstrDescriptor)
constexpr unsigned IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(TN >> 3) << 17) | ((unsigned)(TM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float *v)
{
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __launch_bounds__(256) k_node_gemm_tc(const GemmArgs g, const float *__restrict__ wsplit, int kpad)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem);   // full[3], empty[3], acc
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + 96);
    unsigned char *stages = smem + 128;
    float *red = reinterpret_cast<float *>(smem + 128 + STAGES * STAGE_BYTES);  // [128][2]
    unsigned long long *full = bars, *empty = bars + STAGES, *accb = bars + 2 * STAGES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(accb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;

    float pnorm = 1.f;
    if (g.pool) {
        float s = 0.f;
        for (int c = 0; c < TN; ++c) s += g.pool[c] * g.pool[c];
        pnorm = sqrtf(s);
    }
    const int nblocks_k = (kpad + KB - 1) / KB;
    const float *whi = wsplit, *wlo = wsplit + (size_t)kpad * TN;
    const int n_tiles = (g.M + TM - 1) / TM;
    const int my_tiles = (n_tiles > (int)blockIdx.x) ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_it = my_tiles * nblocks_k;
    // A blocks are fetched two K-blocks ahead into registers (the LDG latency overlaps the STS / MMA of the blocks
    // in between); iteration j of this CTA is (tile = blockIdx.x + (j / nblocks_k) * gridDim.x, kb = j % nblocks_k)
    auto fetch = [&](int j, float4 (&dst)[2]) {
        dst[0] = dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= total_it) return;
        const int t = j / nblocks_k, kb = j - t * nblocks_k;
        const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TM, k0 = kb * KB;
        const int nch = min(NCH, (kpad - k0) >> 2);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int idx = (int)threadIdx.x + 256 * q;
            if (idx < TM * nch) {
                const int r = idx / nch, c = idx - r * nch;
                const int row = m0 + r;
                if (row < g.M) {
                    const int ar = g.rows ? __ldg(g.rows + row) : row;
                    dst[q] = __ldg(reinterpret_cast<const float4 *>(g.A + (size_t)ar * g.lda + k0) + c);
                }
            }
        }
    };
    float4 pre0[2], pre1[2];
    fetch(0, pre0);
    fetch(1, pre1);
    unsigned tile_cnt = 0;
    auto issue_b = [&](int j) {
        const int s = j % STAGES;
        const unsigned use = (unsigned)j / STAGES;
        if (use >= 1) mbar_wait(empty + s, (use - 1) & 1);      // every thread: the stage is free for A stores too
        if (threadIdx.x == 0) {
            const int kb = j % nblocks_k, k0 = kb * KB;
            const int nch = min(NCH, (kpad - k0) >> 2);
            unsigned char *b_hi = stages + s * STAGE_BYTES + 2 * A_TILE, *b_lo = b_hi + B_TILE;
            mbar_expect_tx(full + s, 2u * nch * B_LBO);
            bulk_g2s(b_hi, whi + (size_t)(k0 >> 2) * (B_LBO / 4), nch * B_LBO, full + s);
            bulk_g2s(b_lo, wlo + (size_t)(k0 >> 2) * (B_LBO / 4), nch * B_LBO, full + s);
        }
    };
    if (total_it > 0) issue_b(0);
    auto step = [&](int it, float4 (&cur)[2]) {
        const int t = it / nblocks_k, kb = it - t * nblocks_k;
        const int m0 = ((int)blockIdx.x + t * (int)gridDim.x) * TM;
        {
            const int s = it % STAGES;
            const unsigned use = (unsigned)it / STAGES;
            const int k0 = kb * KB;
            const int nch = min(NCH, (kpad - k0) >> 2);       // chunks in this block (kpad is a multiple of 8)
            unsigned char *st = stages + s * STAGE_BYTES;
            unsigned char *a_hi = st, *a_lo = st + A_TILE, *b_hi = st + 2 * A_TILE, *b_lo = b_hi + B_TILE;
            // one block ahead: free the next stage (its MMAs were committed two iterations ago) and start the bulk
            // copies of its W_hi / W_lo block, so their latency hides behind this block's stores and MMAs
            if (it + 1 < total_it) issue_b(it + 1);
            // A block: 128 rows x nch chunks of 16 bytes, split hi / lo into the canonical layout
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int idx = (int)threadIdx.x + 256 * q;
                if (idx < TM * nch) {
                    const int r = idx / nch, c = idx - r * nch;
                    const float4 v = cur[q];
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                    const int off = c * A_LBO + (r >> 3) * SBO + (r & 7) * 16;
                    *reinterpret_cast<float4 *>(a_hi + off) = h;
                    *reinterpret_cast<float4 *>(a_lo + off) = l;
                }
            }
            fetch(it + 2, cur);                                            // refill the registers two blocks ahead
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
            __syncthreads();
            if (threadIdx.x == 0) {
                mbar_wait(full + s, use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned ah = s32(a_hi), al = s32(a_lo), bh = s32(b_hi), bl = s32(b_lo);
                for (int ks = 0; ks < nch / 2; ++ks) {        // one k-step = 8 tf32 = two chunks
                    const unsigned long long dah = smem_desc(ah + 2 * ks * A_LBO, A_LBO, SBO);
                    const unsigned long long dal = smem_desc(al + 2 * ks * A_LBO, A_LBO, SBO);
                    const unsigned long long dbh = smem_desc(bh + 2 * ks * B_LBO, B_LBO, SBO);
                    const unsigned long long dbl = smem_desc(bl + 2 * ks * B_LBO, B_LBO, SBO);
                    mma_tf32(tmem, dah, dbl, (kb | ks) ? 1u : 0u);
                    mma_tf32(tmem, dal, dbh, 1u);
                    mma_tf32(tmem, dah, dbh, 1u);
                }
                mma_commit(empty + s);
                if (kb == nblocks_k - 1) mma_commit(accb);
            }
        }
        if (kb != nblocks_k - 1) return;
        // ---- epilogue ----
        mbar_wait(accb, tile_cnt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int r = (warp & 3) * 32 + lane;            // TMEM lane == tile row
        const int row = m0 + r;
        const int cbase = (warp >> 2) * 64;
        float sc = 1.f;
        if (g.row_scale && row < g.M) sc = g.row_scale[g.rows ? g.rows[row] : row];
        float dot = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c0 = cbase + half * 32;
            float v[32];
            tmem_ld32(tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float x = v[c] + (g.bias ? __ldg(g.bias + c0 + c) : 0.f);
                if (g.relu) x = fmaxf(x, 0.f);
                if (g.pool) dot += x * __ldg(g.pool + c0 + c);
                v[c] = x * sc;
            }
            if (g.C && row < g.M) {
                float4 *dst = reinterpret_cast<float4 *>(g.C + (size_t)row * TN + c0);
#pragma unroll
                for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
        }
        if (g.pool) red[r * 2 + (warp >> 2)] = dot;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();     // every warp has read its accumulator slice before the next tile overwrites it
        if (g.pool && threadIdx.x < TM && m0 + (int)threadIdx.x < g.M)
            g.score[m0 + threadIdx.x] = tanhf((red[threadIdx.x * 2] + red[threadIdx.x * 2 + 1]) / pnorm);
        ++tile_cnt;
    };
    for (int it = 0; it < total_it; it += 2) {
        step(it, pre0);
        if (it + 1 < total_it) step(it + 1, pre1);
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}


// ---- K <= 48 (block 0: K = 2 x 17 features -> 40): the whole W stays in shared memory, one K block per tile ----
// Per tile: registers (prefetched one tile ahead) -> hi/lo split -> canonical smem -> 3 x (kpad/8) MMAs -> epilogue.
// No stage ring: the accumulator barrier also says the A tile is free again; two CTAs per SM overlap each other.
constexpr int SK_MAX_CH = 12;                                  // kpad <= 48
__host__ __device__ constexpr int sk_smem_bytes(int nch) { return 128 + 2 * nch * A_LBO + 2 * nch * B_LBO + 128 * 2 * 4; }

__global__ void __launch_bounds__(256, 2) k_node_gemm_tc_smallk(const GemmArgs g, const float *__restrict__ wsplit, int kpad)
{
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *wbar = reinterpret_cast<unsigned long long *>(smem), *accb = wbar + 1;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(smem + 64);
    const int nch = kpad >> 2;
    unsigned char *a_hi = smem + 128, *a_lo = a_hi + nch * A_LBO, *b_hi = a_lo + nch * A_LBO, *b_lo = b_hi + nch * B_LBO;
    float *red = reinterpret_cast<float *>(b_lo + nch * B_LBO);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(wbar, 1);
        mbar_init(accb, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(wbar, 2u * nch * B_LBO);
        bulk_g2s(b_hi, wsplit, nch * B_LBO, wbar);
        bulk_g2s(b_lo, wsplit + (size_t)kpad * TN, nch * B_LBO, wbar);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(128u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = *tmem_slot;
    float pnorm = 1.f;
    if (g.pool) {
        float s = 0.f;
        for (int c = 0; c < TN; ++c) s += __ldg(g.pool + c) * __ldg(g.pool + c);
        pnorm = sqrtf(s);
    }
    // this thread's chunks of a tile: idx = tid + 256 q -> (row r, chunk c), fixed for the whole kernel
    constexpr int QMAX = (TM * SK_MAX_CH + 255) / 256;   // 6
    int rr[QMAX], cc[QMAX];
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
        const int idx = (int)threadIdx.x + 256 * q;
        rr[q] = idx / nch;
        cc[q] = idx - rr[q] * nch;
        if (idx >= TM * nch) rr[q] = -1;
    }
    const int n_tiles = (g.M + TM - 1) / TM;
    float4 pre[QMAX];
    auto fetch = [&](int tile) {
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            pre[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            const int row = tile * TM + rr[q];
            if (rr[q] >= 0 && tile < n_tiles && row < g.M) {
                const int ar = g.rows ? __ldg(g.rows + row) : row;
                pre[q] = __ldg(reinterpret_cast<const float4 *>(g.A + (size_t)ar * g.lda) + cc[q]);
            }
        }
    };
    fetch(blockIdx.x);
    unsigned tile_cnt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tile_cnt) {
        const int m0 = tile * TM;
        // the previous tile's epilogue waited on accb, so its MMAs have finished reading the A tile
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
            if (rr[q] >= 0) {
                const float4 v = pre[q];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
                const int off = cc[q] * A_LBO + (rr[q] >> 3) * SBO + (rr[q] & 7) * 16;
                *reinterpret_cast<float4 *>(a_hi + off) = h;
                *reinterpret_cast<float4 *>(a_lo + off) = l;
            }
        fetch(tile + gridDim.x);                                           // next tile's rows travel during MMA + epilogue
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            if (tile_cnt == 0) mbar_wait(wbar, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned ah = s32(a_hi), al = s32(a_lo), bh = s32(b_hi), bl = s32(b_lo);
            for (int ks = 0; ks < nch / 2; ++ks) {
                const unsigned long long dah = smem_desc(ah + 2 * ks * A_LBO, A_LBO, SBO);
                const unsigned long long dal = smem_desc(al + 2 * ks * A_LBO, A_LBO, SBO);
                const unsigned long long dbh = smem_desc(bh + 2 * ks * B_LBO, B_LBO, SBO);
                const unsigned long long dbl = smem_desc(bl + 2 * ks * B_LBO, B_LBO, SBO);
                mma_tf32(tmem, dah, dbl, ks ? 1u : 0u);
                mma_tf32(tmem, dal, dbh, 1u);
                mma_tf32(tmem, dah, dbh, 1u);
            }
            mma_commit(accb);
        }
        mbar_wait(accb, tile_cnt & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int r = (warp & 3) * 32 + lane;
        const int row = m0 + r;
        const int cbase = (warp >> 2) * 64;
        float sc = 1.f;
        if (g.row_scale && row < g.M) sc = g.row_scale[g.rows ? g.rows[row] : row];
        float dot = 0.f;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int c0 = cbase + half * 32;
            float v[32];
            tmem_ld32(tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)c0, v);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float x = v[c] + (g.bias ? __ldg(g.bias + c0 + c) : 0.f);
                if (g.relu) x = fmaxf(x, 0.f);
                if (g.pool) dot += x * __ldg(g.pool + c0 + c);
                v[c] = x * sc;
            }
            if (g.C && row < g.M) {
                float4 *dst = reinterpret_cast<float4 *>(g.C + (size_t)row * TN + c0);
#pragma unroll
                for (int c = 0; c < 8; ++c) dst[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
            }
        }
        if (g.pool) red[r * 2 + (warp >> 2)] = dot;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (g.pool && threadIdx.x < TM && m0 + (int)threadIdx.x < g.M)
            g.score[m0 + threadIdx.x] = tanhf((red[threadIdx.x * 2] + red[threadIdx.x * 2 + 1]) / pnorm);
    }
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u) : "memory");
    }
}

}  // namespace tc

int launch_gemm_tc(const GemmArgs &g, const float *wsplit, cudaStream_t st)
{
    if (g.N != tc::TN || (g.lda & 3) || (reinterpret_cast<uintptr_t>(g.A) & 15) || (reinterpret_cast<uintptr_t>(wsplit) & 15)) {
        mdq::set_error("tcgen05 node GEMM needs width 128 and 16-byte aligned rows (lda %d)", g.lda);
        return MDQ_EINVAL;
    }
    // SAGE layout: A rows are [x (Fp) | mean (Fp)] and wsplit holds W permuted / zero-padded to those columns
    const int kpad = g.sage_F > 0 ? ((2 * g.sage_Fp + 7) & ~7) : ((g.K + 7) & ~7);
    if (kpad > g.lda) {
        mdq::set_error("tcgen05 node GEMM: A rows must be zero-padded to a multiple of 8 columns (K %d, lda %d)", g.K, g.lda);
        return MDQ_EINVAL;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc::k_node_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
        if (e != cudaSuccess) {
            mdq::set_error("k_node_gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MDQ_ECUDA;
        }
        configured = true;
    }
    const int n_tiles = (g.M + tc::TM - 1) / tc::TM;
    if (kpad <= 4 * tc::SK_MAX_CH) {
        const int smem = tc::sk_smem_bytes(kpad >> 2);
        static int configured_sk = 0;
        if (configured_sk < smem) {
            cudaError_t e = cudaFuncSetAttribute(tc::k_node_gemm_tc_smallk, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 tc::sk_smem_bytes(tc::SK_MAX_CH));
            if (e != cudaSuccess) {
                mdq::set_error("k_node_gemm_tc_smallk: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return MDQ_ECUDA;
            }
            configured_sk = tc::sk_smem_bytes(tc::SK_MAX_CH);
        }
        const int grid_sk = n_tiles < 148 * 2 ? n_tiles : 148 * 2;
        tc::k_node_gemm_tc_smallk<<<grid_sk, 256, smem, st>>>(g, wsplit, kpad);
        return mdq::check_launch("k_node_gemm_tc_smallk");
    }
    const int grid = n_tiles < 148 * 3 ? n_tiles : 148 * 3;
    tc::k_node_gemm_tc<<<grid, 256, tc::SMEM_BYTES, st>>>(g, wsplit, kpad);
    return mdq::check_launch("k_node_gemm_tc");
}
