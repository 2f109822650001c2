// Shared helpers for the telogator2_b200 CUDA translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "telogator2_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "telogator2_b200 is written for sm_100a (B200) only"
#endif

#define TG_EXPORT extern "C" __attribute__((visibility("default")))

char *tg_err_buf();                      // thread-local message buffer (tg_api.cu)
int tg_set_err(int code, const char *fmt, ...);

#define TG_CUDA_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess)                                                                  \
            return tg_set_err(TG_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                              __FILE__, __LINE__);                                               \
    } while (0)

#define TG_FULL_MASK 0xffffffffu

static inline int tg_sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}
