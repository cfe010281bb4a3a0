// Shared helpers for libmeshdqn_b200.so (error reporting, launch accounting).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "meshdqn_b200.h"

namespace mdq {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);

inline int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return MDQ_ECUDA;
    }
    count_launch();
    return MDQ_OK;
}

inline int pad4(int v) { return (v + 3) & ~3; }

}  // namespace mdq
