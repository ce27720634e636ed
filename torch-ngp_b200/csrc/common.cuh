// common.cuh — shared host/device helpers for the sm_100a hot-path library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/ngp_b200.h"

namespace ngp {

// ---- error plumbing (thread-local message, launch counter) -----------------------------------
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

int fail(int code, const char* fmt, ...);

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(NGP_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    }
    return NGP_OK;
}

template <typename T>
__host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

inline cudaStream_t as_stream(ngp_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// number of SMs of the current device (cached)
int sm_count();

}  // namespace ngp
