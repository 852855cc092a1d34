// common.h -- host-side helpers shared by the translation units of libsabc_b200.so
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/sabc_b200.h"

namespace sabc {

char* last_error_buf();                       // thread-local, 1024 bytes
int set_error(int code, const char* fmt, ...);

#define SABC_CUDA(call)                                                                                   \
    do {                                                                                                  \
        cudaError_t err__ = (call);                                                                       \
        if (err__ != cudaSuccess)                                                                         \
            return ::sabc::set_error(SABC_ERR_CUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(err__), \
                                     __FILE__, __LINE__, cudaGetErrorString(err__));                      \
    } while (0)

#define SABC_TRY(expr)                  \
    do {                                \
        int rc__ = (expr);              \
        if (rc__ != 0) return rc__;     \
    } while (0)

struct ModelVTable;
const ModelVTable* find_model(const char* name);

// RAII device buffer
template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    cudaError_t alloc(size_t count) {
        release();
        if (count == 0) count = 1;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    cudaError_t ensure(size_t count) { return count <= n ? cudaSuccess : alloc(count); }
};

}  // namespace sabc
