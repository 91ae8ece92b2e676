// common.cuh -- context, device-memory bookkeeping and launch helpers shared by all modules of
// libsindyn_cuda.  Device code in this library targets sm_100a only (no multi-arch dispatch).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/sindyn.h"

#define SINDYN_NUM_SMS_B200 148

struct sindyn_ctx;

#define CU_CHECK(ctx, expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            (ctx)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_e), __FILE__, __LINE__); \
            return SINDYN_ERR_CUDA;                                                            \
        }                                                                                      \
    } while (0)

#define SD_CHECK(expr)                   \
    do {                                 \
        int _s = (expr);                 \
        if (_s != SINDYN_OK) return _s;  \
    } while (0)

// Base for both handle types: stream, error string, allocation list, launch counter.
struct sindyn_base {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    unsigned long long launches = 0;
    std::vector<void *> dev_allocs;
    std::vector<void *> host_allocs;

    void set_error(const std::string &m, const char *file, int line)
    {
        char buf[64];
        snprintf(buf, sizeof buf, " (%s:%d)", strrchr(file, '/') ? strrchr(file, '/') + 1 : file, line);
        err = m + buf;
    }
    template <class T> int dalloc(T **p, size_t count)
    {
        void *q = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = sizeof(T);
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) {
            set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e), __FILE__, __LINE__);
            return SINDYN_ERR_CUDA;
        }
        cudaMemsetAsync(q, 0, bytes, stream);
        dev_allocs.push_back(q);
        *p = (T *)q;
        return SINDYN_OK;
    }
    template <class T> int halloc(T **p, size_t count)
    {
        void *q = nullptr;
        cudaError_t e = cudaMallocHost(&q, count * sizeof(T));
        if (e != cudaSuccess) {
            set_error(std::string("cudaMallocHost: ") + cudaGetErrorString(e), __FILE__, __LINE__);
            return SINDYN_ERR_CUDA;
        }
        host_allocs.push_back(q);
        *p = (T *)q;
        return SINDYN_OK;
    }
    void free_all()
    {
        for (void *p : dev_allocs) cudaFree(p);
        for (void *p : host_allocs) cudaFreeHost(p);
        dev_allocs.clear();
        host_allocs.clear();
    }
};

// Every kernel launch of the library goes through this macro so that gpu_launches is a count,
// not an estimate.
#define LAUNCH(ctx, kern, grid, block, smem, ...)                          \
    do {                                                                   \
        kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);     \
        (ctx)->launches++;                                                 \
    } while (0)

#define LAUNCH_CHECK(ctx) CU_CHECK(ctx, cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// pitched host <-> dense device copies
static inline cudaError_t copy_in_2d(void *dst, const void *src, size_t src_step, size_t row_bytes, int rows, cudaStream_t s)
{
    return cudaMemcpy2DAsync(dst, row_bytes, src, src_step ? src_step : row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s);
}
static inline cudaError_t copy_out_2d(void *dst, size_t dst_step, const void *src, size_t row_bytes, int rows, cudaStream_t s)
{
    return cudaMemcpy2DAsync(dst, dst_step ? dst_step : row_bytes, src, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, s);
}
