// Follow-up: is the ~800-cycle cost per bulk copy a per-instruction service time?  Sizes up to 192 KB per copy,
// issue from one thread vs several warps (own barrier each), and time of the issue instructions alone.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../meshdqn_b200/csrc/tc_prims.cuh"
using namespace tcp;

__global__ void __launch_bounds__(256) k_one(const char *src, int copy_bytes, int ncopies, int issuers, long long *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sm);
    unsigned char *buf = sm + 128;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int w = threadIdx.x >> 5;
    long long t0 = clock64(), t1 = 0, t2 = 0;
    if ((threadIdx.x & 31) == 0 && w < issuers) {
        const int mine = ncopies / issuers;
        mbar_expect_tx(bars + w, (unsigned)(mine * copy_bytes));
        for (int c = 0; c < mine; ++c) {
            const int id = w * mine + c;
            bulk_g2s(buf + (size_t)id * copy_bytes, src + (size_t)id * copy_bytes, copy_bytes, bars + w);
        }
        t1 = clock64();
        mbar_wait(bars + w, 0);
        t2 = clock64();
        if (w == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
}

int main()
{
    const size_t total = 192 * 1024;
    char *d; long long *out;
    cudaMalloc(&d, total); cudaMalloc(&out, 64);
    cudaMemset(d, 1, total);
    cudaFuncSetAttribute(k_one, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 + (int)total);
    for (int copy_kb : {1, 4, 16, 32, 64, 96, 192}) {
        for (int issuers : {1, 2, 4}) {
            const int ncopies = (int)(total / (copy_kb * 1024));
            if (ncopies < issuers) continue;
            long long best[2] = {1LL << 60, 1LL << 60};
            for (int rep = 0; rep < 4; ++rep) {
                k_one<<<1, 256, 128 + total>>>(d, copy_kb * 1024, ncopies, issuers, out);
                long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                if (h[1] < best[1]) { best[0] = h[0]; best[1] = h[1]; }
            }
            printf("192 KB as %3d copies of %3d KB, %d issuing warps: issue %6lld cycles, all landed %6lld cycles (%.1f B/clk)\n", ncopies, copy_kb,
                   issuers, best[0], best[1], (double)total / best[1]);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
