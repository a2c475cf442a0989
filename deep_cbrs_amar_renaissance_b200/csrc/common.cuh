// Internal helpers shared by the sm_100a kernels.  Not part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cbrs_b200.h"

namespace cbrs {

void set_error(const char *fmt, ...);

#define CBRS_REQUIRE(cond, code, ...)            \
    do {                                         \
        if (!(cond)) {                           \
            ::cbrs::set_error(__VA_ARGS__);      \
            return (code);                       \
        }                                        \
    } while (0)

// launch check: catches configuration errors without synchronising
#define CBRS_CHECK_LAUNCH(what)                                                         \
    do {                                                                                \
        cudaError_t e__ = cudaGetLastError();                                           \
        if (e__ != cudaSuccess) {                                                       \
            ::cbrs::set_error("%s: %s", (what), cudaGetErrorString(e__));               \
            return CBRS_E_CUDA;                                                         \
        }                                                                               \
    } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over the caller's workspace
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t n) : base((char *)p), cap(n), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = align_up(count * sizeof(T));
        if (off + bytes > cap) return nullptr;
        T *r = (T *)(base + off);
        off += bytes;
        return r;
    }
};

// ---- scan / sort primitives (scan_sort.cu) --------------------------------
size_t scan_u32_workspace_bytes(int64_t n);
// in-place exclusive scan; *total_out (device uint32, may be null) receives the sum
int scan_u32_exclusive(uint32_t *data, int64_t n, uint32_t *total_out, void *ws, size_t ws_bytes,
                       cudaStream_t s);
size_t scan_i64_workspace_bytes(int64_t n);
int scan_i64_exclusive(int64_t *data, int64_t n, int64_t *total_out, void *ws, size_t ws_bytes,
                       cudaStream_t s);
size_t sort_workspace_bytes(int64_t n);
// stable LSD radix sort; result is left in keys/payload (copied back if needed)
int sort_pairs_u64(uint64_t *keys, uint32_t *payload, int64_t n, int key_bits, void *ws,
                   size_t ws_bytes, cudaStream_t s);

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// streaming (read-once) loads: keep them out of L1 so gathered rows stay cached
__device__ __forceinline__ int ld_stream_i32(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

}  // namespace cbrs
