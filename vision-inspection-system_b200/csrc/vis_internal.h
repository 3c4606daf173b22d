// vis_internal.h — shared helpers of libvis_b200.so (not part of the public ABI)
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "vis_b200.h"

#define VIS_PRECISION_BITS 22          // Pillow: 32 - 8 - 2 (libImaging/Resample.c)
#define VIS_PATCH 14
#define VIS_ROW_FLOATS 1176            // 3 channels * 2 temporal * 14 * 14

namespace vis {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);   // records the error text, returns VIS_E_CUDA

inline int check_launch(const char* where) {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? VIS_OK : cuda_fail(e, where);
}

}  // namespace vis
