// Shared host-side helpers: error handling without exceptions across the C ABI, device buffers.
#pragma once
#include <cuda.h>

#include "../../include/pix2pose_b200.h"
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <stdexcept>
#include <string>

namespace p2p {

// Status codes (P2P_OK, P2P_ERR_*) are the macros of include/pix2pose_b200.h.

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

inline std::string fmt(const char* f, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    return std::string(buf);
}

#define P2P_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess)                                                                      \
            throw ::p2p::Error(P2P_ERR_CUDA, ::p2p::fmt("%s failed: %s (%s:%d)", #expr,      \
                                                                cudaGetErrorString(_e), __FILE__, __LINE__)); \
    } while (0)

#define P2P_CHECK(cond, ...)                                                                   \
    do {                                                                                       \
        if (!(cond)) throw ::p2p::Error(P2P_ERR_INVALID, ::p2p::fmt(__VA_ARGS__));      \
    } while (0)

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (count) P2P_CUDA(cudaMalloc(&p, count * sizeof(T)));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void upload(const T* host, size_t count, cudaStream_t s = 0) {
        if (count > n) alloc(count);
        if (count) P2P_CUDA(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, s));
    }
};

void set_last_error(const std::string& m);
const char* get_last_error();
void require_device();  // throws P2P_ERR_NO_DEVICE unless an sm_100 GPU is current

// Runs f(), mapping exceptions to status codes (no exception crosses the C ABI).
template <typename F>
int guarded(F&& f) {
    try {
        f();
        return P2P_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return P2P_ERR_INTERNAL;
    } catch (...) {
        set_last_error("unknown error");
        return P2P_ERR_INTERNAL;
    }
}

}  // namespace p2p
