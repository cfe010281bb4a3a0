// fp64_lat.cu -- dependent-chain latencies (cycles) of the operations on k_smooth's critical path, one warp / 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double seed, int n)
{
    __shared__ double sh[64];
    const int tid = threadIdx.x;
    if (tid < 64) sh[tid] = seed + tid;
    __syncthreads();
    double a = seed + 1.5, b = seed + 2.25;
    long long t0, t1;
    int slot = 0;
#define RUN(body)                                                   \
    __syncthreads();                                                \
    t0 = clock64();                                                 \
    for (int i = 0; i < n; ++i) { body; }                           \
    t1 = clock64();                                                 \
    if (tid == 0) cyc[slot] = (t1 - t0);                            \
    ++slot;
    RUN(a = a + b)
    RUN(a = fma(a, 0.999, b))
    RUN(a = b / a + 1.0)          // div + add
    RUN(a = sqrt(a) + 3.0)        // sqrt + add
    RUN(a = __shfl_xor_sync(0xffffffffu, a, 1) + 1.0)
    RUN({ int idx = ((int)a) & 63; a = sh[idx] + 1.0; })
    RUN({ unsigned h = __reduce_min_sync(0xffffffffu, (unsigned)__double2hiint(a)); a = a + (double)(h & 1); })
    RUN(__syncthreads())
    RUN({ sh[tid & 63] = a; __syncwarp(); a = sh[(tid + 1) & 63] + 1.0; __syncwarp(); })
    RUN({ double d = sqrt(a * a + b * b); a = fabs(a * b) / d + 1.0; })                 // len + distance
    RUN({ double r = sqrt(a * a + b * b); a = a + 0.5 * a / r; })                        // r + move
    out[tid] = a;
}
int main()
{
    double *out; long long *cyc, h[16];
    cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8 * 16);
    const char *names[] = {"dadd", "dfma", "ddiv+dadd", "dsqrt+dadd", "shfl64+dadd", "lds+cvt+dadd", "redux+dadd", "syncthreads",
                           "sts/syncwarp/lds/syncwarp", "len+dist", "r+move"};
    for (int threads : {32, 256}) {
        const int n = 2000;
        k<<<1, threads>>>(out, cyc, 1.0, n);
        k<<<1, threads>>>(out, cyc, 1.0, n);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads=%d\n", threads);
        for (int i = 0; i < 11; ++i) printf("  %-28s %7.1f cycles\n", names[i], (double)h[i] / n);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
