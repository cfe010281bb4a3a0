// Microbenchmark: bulk (TMA) global->shared copy rate of ONE CTA and of many CTAs reading the same buffer,
// cold / warm L2, different depths in flight.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../meshdqn_b200/csrc/tc_prims.cuh"
using namespace tcp;

__global__ void __launch_bounds__(256) k_stream(const float *src, int nblk, int blk_bytes, int depth, int ncopy, long long *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    unsigned char *buf = sm + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int issued = 0;
        for (int j = 0; j < nblk; ++j) {
            while (issued < nblk && issued < j + depth) {
                const int s = issued % depth;
                mbar_expect_tx(bars + s, blk_bytes);
                const int piece = blk_bytes / ncopy;
                for (int c = 0; c < ncopy; ++c)
                    bulk_g2s(buf + (size_t)s * blk_bytes + c * piece, (const char *)src + (size_t)issued * blk_bytes + c * piece, piece, bars + s);
                ++issued;
            }
            mbar_wait(bars + (j % depth), (j / depth) & 1);
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
}

// same data through plain loads (LDG.128 by all threads -> STS)
__global__ void __launch_bounds__(256) k_ldg(const float4 *src, int n4, long long *out, float *sink)
{
    extern __shared__ __align__(128) unsigned char sm[];
    float4 *buf = reinterpret_cast<float4 *>(sm);
    __syncthreads();
    long long t0 = clock64();
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < n4; i += 256 * 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (i + u * 256 < n4) ? __ldg(src + i + u * 256) : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) buf[(threadIdx.x + u * 256) & 2047] = v[u];
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc.x == 123.f) sink[0] = buf[threadIdx.x].x;
}

int main()
{
    const size_t total = 1 << 20;   // 1 MB of "weights"
    float *d; long long *out; float *sink; char *flush;
    cudaMalloc(&d, total); cudaMalloc(&out, 64); cudaMalloc(&sink, 64); cudaMalloc(&flush, 256 << 20);
    cudaMemset(d, 0, total);
    cudaFuncSetAttribute(k_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int grid : {1, 32, 128}) {
        for (int cold = 0; cold < 2; ++cold) {
            for (int blk_kb : {16, 32}) {
                for (int depth : {1, 2, 4}) {
                    for (int ncopy : {1, 2, 8}) {
                        const int blk_bytes = blk_kb * 1024, nblk = (int)(total / blk_bytes);
                        long long best = 1LL << 60;
                        for (int rep = 0; rep < 3; ++rep) {
                            if (cold) cudaMemset(flush, rep, 256 << 20);
                            k_stream<<<grid, 256, 128 + depth * blk_bytes>>>(d, nblk, blk_bytes, depth, ncopy, out);
                            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                            if (h < best) best = h;
                        }
                        printf("grid %3d %s blk %2d KB depth %d ncopy %d: %8lld cycles  %.1f B/clk\n", grid, cold ? "cold" : "warm", blk_kb,
                               depth, ncopy, best, (double)total / best);
                    }
                }
            }
            long long best = 1LL << 60;
            for (int rep = 0; rep < 3; ++rep) {
                if (cold) cudaMemset(flush, rep, 256 << 20);
                k_ldg<<<grid, 256, 64 * 1024>>>((const float4 *)d, (int)(total / 16), out, sink);
                long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
                if (h < best) best = h;
            }
            printf("grid %3d %s LDG.128 x8 -> STS: %8lld cycles  %.1f B/clk\n", grid, cold ? "cold" : "warm", best, (double)total / best);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
