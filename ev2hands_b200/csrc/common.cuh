// Shared helpers for libev2h.so (host-side status handling, small device utilities).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/ev2h.h"

namespace ev2h {

// Thread-local message for ev2h_last_error(); defined in capi.cu.
void set_error(const char *fmt, ...);

inline int fail(ev2h_status st, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    set_error("%s", buf);
    return (int)st;
}

inline int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(EV2H_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return EV2H_OK;
}

#define EV2H_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return ev2h::fail(EV2H_ERR_BAD_ARGUMENT, __VA_ARGS__); \
    } while (0)

inline cudaStream_t as_stream(ev2h_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

}  // namespace ev2h
