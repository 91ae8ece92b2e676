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

// Programmatic dependent launch (sm_90+): a kernel launched with LAUNCH_PDL may become resident and run its input-independent
// prologue while the previous kernel of the stream is still executing; it must call pdl_wait() before it touches anything
// the previous kernel wrote (the wait returns when that grid has completed and its writes are visible), and calls
// pdl_trigger() right after so that its own successor can be scheduled.  Without the launch attribute both are no-ops.
// Works in stream capture (the graph gets programmatic dependency edges).
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
#define LAUNCH_PDL(ctx, kern, grid, block, smem, ...)                                        \
    do {                                                                                     \
        cudaLaunchConfig_t cfg_ = {};                                                        \
        cfg_.gridDim = dim3(grid);                                                           \
        cfg_.blockDim = dim3(block);                                                         \
        cfg_.dynamicSmemBytes = (smem);                                                      \
        cfg_.stream = (ctx)->stream;                                                         \
        cudaLaunchAttribute at_[1];                                                          \
        at_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                      \
        at_[0].val.programmaticStreamSerializationAllowed = 1;                               \
        cfg_.attrs = at_;                                                                    \
        cfg_.numAttrs = 1;                                                                   \
        cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__);                                        \
        (ctx)->launches++;                                                                   \
    } while (0)

#define LAUNCH_CHECK(ctx) CU_CHECK(ctx, cudaGetLastError())

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Pageable host memory makes cudaMemcpyAsync a slow, synchronous, driver-staged copy (~0.5 ms for one 640x480 BGR frame).
// Callers of the reference interface hand over cv::Mat buffers (pageable), so the per-frame entry points stage them
// through pinned bounce buffers owned by the handle: one host memcpy at memory bandwidth, then a true async DMA.
static inline bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
// host (pitched) -> device (dense)
static inline cudaError_t stage_in_2d(void *dst_dev, const void *src, size_t src_step, size_t row_bytes, int rows, void *pinned, cudaStream_t s)
{
    if (!src_step) src_step = row_bytes;
    if (!pinned || host_ptr_is_pinned(src))
        return cudaMemcpy2DAsync(dst_dev, row_bytes, src, src_step, row_bytes, rows, cudaMemcpyHostToDevice, s);
    if (src_step == row_bytes) memcpy(pinned, src, row_bytes * rows);
    else for (int r = 0; r < rows; ++r) memcpy((char *)pinned + r * row_bytes, (const char *)src + r * src_step, row_bytes);
    return cudaMemcpyAsync(dst_dev, pinned, row_bytes * rows, cudaMemcpyHostToDevice, s);
}
// device (dense) -> pinned bounce buffer; finish with stage_out_finish after the stream has been synchronised
static inline cudaError_t stage_out_begin(void *pinned, const void *src_dev, size_t bytes, cudaStream_t s)
{
    return cudaMemcpyAsync(pinned, src_dev, bytes, cudaMemcpyDeviceToHost, s);
}
static inline void stage_out_finish(void *dst, size_t dst_step, const void *pinned, size_t row_bytes, int rows)
{
    if (!dst_step) dst_step = row_bytes;
    if (dst_step == row_bytes) memcpy(dst, pinned, row_bytes * rows);
    else for (int r = 0; r < rows; ++r) memcpy((char *)dst + r * dst_step, (const char *)pinned + r * row_bytes, row_bytes);
}

// pitched host <-> dense device copies
static inline cudaError_t copy_in_2d(void *dst, const void *src, size_t src_step, size_t row_bytes, int rows, cudaStream_t s)
{
    return cudaMemcpy2DAsync(dst, row_bytes, src, src_step ? src_step : row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s);
}
static inline cudaError_t copy_out_2d(void *dst, size_t dst_step, const void *src, size_t row_bytes, int rows, cudaStream_t s)
{
    return cudaMemcpy2DAsync(dst, dst_step ? dst_step : row_bytes, src, row_bytes, row_bytes, rows, cudaMemcpyDeviceToHost, s);
}
