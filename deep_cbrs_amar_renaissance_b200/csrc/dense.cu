// Dense layer with fused row gather + two-source concat + bias + row-op + activation (fp32).
// Rows P1 (X.W transform), P2 ([x || agg].W + l2-normalise), P3 (X.W + attention logits),
// S1-S4 (embedding lookup + Concatenate + Dense towers / classifier).
//
// Replaces tf.nn.embedding_lookup + layers.Concatenate + keras Dense at
// /root/reference/src/models/basic.py:31-37,72-75, src/models/hybrid.py:72-89,136-140,
// src/models/dense.py:4-17 and the K.dot(x, kernel) inside the spektral layers.
//
// fp32 FFMA register-tiled GEMM (the reference computes in fp32 and parity is 1e-5, which
// rules bf16 tensor cores out for this path; the bf16 tcgen05 scorer is a separate kernel).
// CTA tile 128 x BN, k-step 32, 8 x TN micro-tile; A rows are fetched through the optional
// index vectors, so a gathered, concatenated batch is never materialised in HBM.
#include "common.cuh"

#include <cuda_bf16.h>

namespace cbrs {

struct DenseParams {
    const float *x1; int64_t ld1; const int64_t *idx1; int32_t f1;
    const float *x2; int64_t ld2; const int64_t *idx2; int32_t f2;
    const float *w; const float *b;
    int64_t m; int32_t n;
    int act; int rowop;
    const float *a_self; const float *a_neigh; float *p_out; float *q_out;
    float *out; int64_t ldo;
    // multi-GPU: finished rows (and the attention logit q) are also stored into these peer-mapped copies
    float *out_peer[CBRS_MAX_PEERS - 1];
    float *q_peer[CBRS_MAX_PEERS - 1];
    int n_peer;
    int out_bf16;  // out / out_peer hold bf16 (round to nearest even), ldo in elements
    // grouped (relational) output: column c of row m is stored at row (c / group_h) * group_rows + m, column c % group_h,
    // i.e. X . [W_0 | ... | W_{R-1}] lands directly in the stacked [R * group_rows, group_h] operand of the relational
    // sparse kernel.  group_h == 0: the ordinary [m, n] output.
    int32_t group_h;
    int64_t group_rows;
};

__device__ __forceinline__ int64_t out_index(const DenseParams &p, int64_t m, int ng) {
    if (p.group_h == 0) return m * p.ldo + ng;
    const int r = ng / p.group_h;
    return ((int64_t)r * p.group_rows + m) * p.ldo + (ng - r * p.group_h);
}

__device__ __forceinline__ void store_out1(float *base, int64_t idx, float v, int bf16) {
    if (bf16) reinterpret_cast<__nv_bfloat16 *>(base)[idx] = __float2bfloat16_rn(v);
    else base[idx] = v;
}
__device__ __forceinline__ void store_out4(float *base, int64_t idx, float a, float b, float c, float d, int bf16) {
    if (bf16) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned int *>(&lo);
        r.y = *reinterpret_cast<const unsigned int *>(&hi);
        *reinterpret_cast<uint2 *>(reinterpret_cast<__nv_bfloat16 *>(base) + idx) = r;
    } else {
        *reinterpret_cast<float4 *>(base + idx) = make_float4(a, b, c, d);
    }
}

constexpr int kBM = 128, kBK = 32, kTM = 8;
constexpr int kDenseThreads = 256;

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case CBRS_ACT_RELU: return fmaxf(v, 0.f);
        case CBRS_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case CBRS_ACT_TANH: return tanhf(v);
        default: return v;
    }
}

template <int BN>
__global__ void __launch_bounds__(kDenseThreads) dense_kernel(const DenseParams p) {
    constexpr int TN = BN / 16;  // 16 threads span the tile's columns
    __shared__ float As[kBK][kBM + 1];
    __shared__ float Bs[kBK][BN];
    __shared__ int64_t src1[kBM];
    __shared__ int64_t src2[kBM];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int K = p.f1 + p.f2;

    if (tid < kBM) {
        const int64_t m = m0 + tid;
        int64_t r1 = -1, r2 = -1;
        if (m < p.m) {
            r1 = (p.idx1 ? p.idx1[m] : m) * p.ld1;
            if (p.x2) r2 = (p.idx2 ? p.idx2[m] : m) * p.ld2;
        }
        src1[tid] = r1;
        src2[tid] = r2;
    }
    __syncthreads();

    float acc[kTM][TN];
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < K; k0 += kBK) {
        // A tile: consecutive threads walk k (128 B per row), rows strided by 8
#pragma unroll
        for (int it = 0; it < (kBM * kBK) / kDenseThreads; ++it) {
            const int e = tid + it * kDenseThreads;
            const int mm = e / kBK, kk = e % kBK;
            const int kg = k0 + kk;
            float v = 0.f;
            if (kg < K) {
                if (kg < p.f1) {
                    const int64_t r = src1[mm];
                    if (r >= 0) v = __ldg(p.x1 + r + kg);
                } else {
                    const int64_t r = src2[mm];
                    if (r >= 0) v = __ldg(p.x2 + r + (kg - p.f1));
                }
            }
            As[kk][mm] = v;
        }
#pragma unroll
        for (int it = 0; it < (kBK * BN + kDenseThreads - 1) / kDenseThreads; ++it) {
            const int e = tid + it * kDenseThreads;
            if (e < kBK * BN) {
                const int kk = e / BN, nn = e % BN;
                const int kg = k0 + kk, ng = n0 + nn;
                Bs[kk][nn] = (kg < K && ng < p.n) ? __ldg(p.w + (int64_t)kg * p.n + ng) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float a[kTM], bv[TN];
#pragma unroll
            for (int i = 0; i < kTM; ++i) a[i] = As[kk][ty * kTM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < kTM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: bias, row-op over the 16 lanes that share a row, activation
#pragma unroll
    for (int i = 0; i < kTM; ++i) {
        const int64_t m = m0 + ty * kTM + i;
        float ss = 0.f, ps = 0.f, qs = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ng = n0 + tx * TN + j;
            float v = acc[i][j];
            if (ng < p.n) {
                if (p.b) v += __ldg(p.b + ng);
                if (p.rowop == CBRS_ROWOP_L2NORM) ss = fmaf(v, v, ss);
                if (p.rowop == CBRS_ROWOP_ATTN) {
                    ps = fmaf(v, __ldg(p.a_self + ng), ps);
                    qs = fmaf(v, __ldg(p.a_neigh + ng), qs);
                }
            } else {
                v = 0.f;
            }
            acc[i][j] = v;
        }
        if (p.rowop != CBRS_ROWOP_NONE) {  // uniform branch; all 32 lanes shuffle
#pragma unroll
            for (int o = 8; o; o >>= 1) {
                ss += __shfl_xor_sync(0xffffffffu, ss, o, 16);
                ps += __shfl_xor_sync(0xffffffffu, ps, o, 16);
                qs += __shfl_xor_sync(0xffffffffu, qs, o, 16);
            }
        }
        if (m >= p.m) continue;
        float scale = 1.f;
        if (p.rowop == CBRS_ROWOP_L2NORM) scale = 1.f / sqrtf(fmaxf(ss, 1e-12f));
        if (p.rowop == CBRS_ROWOP_ATTN && tx == 0) {
            p.p_out[m] = ps;
            p.q_out[m] = qs;
            for (int q = 0; q < p.n_peer; ++q) p.q_peer[q][m] = qs;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ng = n0 + tx * TN + j;
            if (ng < p.n) {
                const float v = apply_act(acc[i][j] * scale, p.act);
                store_out1(p.out, out_index(p, m, ng), v, p.out_bf16);
                for (int q = 0; q < p.n_peer; ++q) store_out1(p.out_peer[q], out_index(p, m, ng), v, p.out_bf16);
            }
        }
    }
}

// stand-alone row l2-normalise + activation for widths the fused epilogue cannot span
__global__ void __launch_bounds__(256) row_l2norm_act_kernel(float *x, int64_t ld, int64_t m, int32_t n, int act) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    float ss = 0.f;
    for (int c = lane; c < n; c += 32) { const float v = x[row * ld + c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float scale = 1.f / sqrtf(fmaxf(ss, 1e-12f));
    for (int c = lane; c < n; c += 32) x[row * ld + c] = apply_act(x[row * ld + c] * scale, act);
}


// ---------------------------------------------------------------------------------------
// Fast path: K (both sources) a multiple of 16, 16-byte aligned rows, n % 4 == 0.
// 3-stage cp.async (LDGSTS, L2-only) ring for the A and W tiles, 128 x BN x 16 CTA tile,
// 8 x TN micro-tile read from shared memory with 128-bit loads along k (A) and n (W):
// one LDS.128 per 16 FFMA.  Rows of a thread are interleaved (ty + 16 i) and its columns
// split in two 4-wide groups so neither operand read has a bank conflict.
constexpr int kFBK = 16, kFStages = 3, kFAS = kFBK + 4;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int BN>
__global__ void __launch_bounds__(kDenseThreads, 2) dense_fast_kernel(const DenseParams p) {
    constexpr int TN = BN / 16;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *As = reinterpret_cast<float *>(smem_raw);                    // [stages][kBM][kFAS]
    float *Bs = As + kFStages * kBM * kFAS;                             // [stages][kFBK][BN]
    int64_t *src1 = reinterpret_cast<int64_t *>(Bs + kFStages * kFBK * BN);
    int64_t *src2 = src1 + kBM;

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = (int64_t)blockIdx.x * kBM;
    const int n0 = blockIdx.y * BN;
    const int K = p.f1 + p.f2;
    const int KT = K / kFBK;

    if (tid < kBM) {
        const int64_t m = m0 + tid;
        int64_t r1 = -1, r2 = -1;
        if (m < p.m) {
            r1 = (p.idx1 ? p.idx1[m] : m) * p.ld1;
            if (p.x2) r2 = (p.idx2 ? p.idx2[m] : m) * p.ld2;
        }
        src1[tid] = r1;
        src2[tid] = r2;
    }
    __syncthreads();

    auto load_tile = [&](int stage, int kt) {
        const int k0 = kt * kFBK;
        const bool first = k0 < p.f1;
        const float *base = first ? p.x1 + k0 : p.x2 + (k0 - p.f1);
        const int64_t *src = first ? src1 : src2;
        float *as = As + stage * kBM * kFAS;
#pragma unroll
        for (int it = 0; it < (kBM * 4) / kDenseThreads; ++it) {
            const int e = tid + it * kDenseThreads;
            const int mm = e >> 2, c = e & 3;
            const int64_t r = src[mm];
            cp_async16(as + mm * kFAS + c * 4, r >= 0 ? (const void *)(base + r + c * 4) : (const void *)p.x1, r >= 0 ? 16 : 0);
        }
        float *bs = Bs + stage * kFBK * BN;
        constexpr int chunks = kFBK * BN / 4;
#pragma unroll
        for (int it = 0; it < (chunks + kDenseThreads - 1) / kDenseThreads; ++it) {
            const int e = tid + it * kDenseThreads;
            if (e < chunks) {
                const int kk = e / (BN / 4), c = e % (BN / 4);
                const int ng = n0 + c * 4;
                const bool ok = ng < p.n;
                cp_async16(bs + kk * BN + c * 4, ok ? (const void *)(p.w + (int64_t)(k0 + kk) * p.n + ng) : (const void *)p.w,
                           ok ? 16 : 0);
            }
        }
    };

    float acc[kTM][TN];
#pragma unroll
    for (int i = 0; i < kTM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < kFStages - 1; ++s) {
        if (s < KT) load_tile(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<kFStages - 2>();
        __syncthreads();
        if (kt + kFStages - 1 < KT) load_tile((kt + kFStages - 1) % kFStages, kt + kFStages - 1);
        cp_async_commit();
        const float *as = As + (kt % kFStages) * kBM * kFAS;
        const float *bs = Bs + (kt % kFStages) * kFBK * BN;
#pragma unroll
        for (int kq = 0; kq < kFBK / 4; ++kq) {
            float4 a[kTM];
#pragma unroll
            for (int i = 0; i < kTM; ++i) a[i] = *reinterpret_cast<const float4 *>(as + (ty + 16 * i) * kFAS + kq * 4);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float bv[TN];
                const float *brow = bs + (kq * 4 + kk) * BN;
                if (TN == 8) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(brow + tx * 4);
                    const float4 b1 = *reinterpret_cast<const float4 *>(brow + 64 + tx * 4);
                    bv[0] = b0.x; bv[1 % TN] = b0.y; bv[2 % TN] = b0.z; bv[3 % TN] = b0.w;
                    bv[4 % TN] = b1.x; bv[5 % TN] = b1.y; bv[6 % TN] = b1.z; bv[7 % TN] = b1.w;
                } else if (TN == 4) {
                    const float4 b0 = *reinterpret_cast<const float4 *>(brow + tx * 4);
                    bv[0] = b0.x; bv[1 % TN] = b0.y; bv[2 % TN] = b0.z; bv[3 % TN] = b0.w;
                } else {
                    const float2 b0 = *reinterpret_cast<const float2 *>(brow + tx * 2);
                    bv[0] = b0.x; bv[1 % TN] = b0.y;
                }
#pragma unroll
                for (int i = 0; i < kTM; ++i) {
                    const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av, bv[j], acc[i][j]);
                }
            }
        }
    }
    cp_async_wait<0>();

    // column of micro-tile slot j
    auto col_of = [&](int j) -> int {
        if (TN == 8) return n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        return n0 + tx * TN + j;
    };
    const bool vec_out = (p.ldo % 4 == 0) && (((uintptr_t)p.out) % (p.out_bf16 ? 8 : 16) == 0) && TN >= 4 &&
                         (p.group_h % 4 == 0);  // a 4-wide group must not straddle two relation blocks
#pragma unroll
    for (int i = 0; i < kTM; ++i) {
        const int64_t m = m0 + ty + 16 * i;
        float ss = 0.f, ps = 0.f, qs = 0.f;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ng = col_of(j);
            float v = acc[i][j];
            if (ng < p.n) {
                if (p.b) v += __ldg(p.b + ng);
                if (p.rowop == CBRS_ROWOP_L2NORM) ss = fmaf(v, v, ss);
                if (p.rowop == CBRS_ROWOP_ATTN) {
                    ps = fmaf(v, __ldg(p.a_self + ng), ps);
                    qs = fmaf(v, __ldg(p.a_neigh + ng), qs);
                }
            } else {
                v = 0.f;
            }
            acc[i][j] = v;
        }
        if (p.rowop != CBRS_ROWOP_NONE) {
#pragma unroll
            for (int o = 8; o; o >>= 1) {
                ss += __shfl_xor_sync(0xffffffffu, ss, o, 16);
                ps += __shfl_xor_sync(0xffffffffu, ps, o, 16);
                qs += __shfl_xor_sync(0xffffffffu, qs, o, 16);
            }
        }
        if (m >= p.m) continue;
        float scale = 1.f;
        if (p.rowop == CBRS_ROWOP_L2NORM) scale = 1.f / sqrtf(fmaxf(ss, 1e-12f));
        if (p.rowop == CBRS_ROWOP_ATTN && tx == 0) {
            p.p_out[m] = ps;
            p.q_out[m] = qs;
            for (int q = 0; q < p.n_peer; ++q) p.q_peer[q][m] = qs;
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = apply_act(acc[i][j] * scale, p.act);
        // copy -1 is the local output; 0 .. n_peer-1 are the peers' (same layout, NVLink stores)
        for (int q = -1; q < p.n_peer; ++q) {
            float *obase = q < 0 ? p.out : p.out_peer[q];
            if (vec_out) {
#pragma unroll
                for (int j0 = 0; j0 < TN; j0 += 4) {
                    const int ng = col_of(j0);
                    if (ng < p.n)  // n % 4 == 0: a 4-wide group is entirely in or out
                        store_out4(obase, out_index(p, m, ng), acc[i][j0], acc[i][(j0 + 1) % TN], acc[i][(j0 + 2) % TN],
                                   acc[i][(j0 + 3) % TN], p.out_bf16);
                }
            } else {
#pragma unroll
                for (int j = 0; j < TN; ++j) {
                    const int ng = col_of(j);
                    if (ng < p.n) store_out1(obase, out_index(p, m, ng), acc[i][j], p.out_bf16);
                }
            }
        }
    }
}

template <int BN>
static int launch_dense_fast(const DenseParams &p, cudaStream_t s) {
    constexpr size_t smem = (size_t)kFStages * (kBM * kFAS + kFBK * BN) * sizeof(float) + 2 * kBM * sizeof(int64_t);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(dense_fast_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("dense_fast: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return CBRS_E_CUDA;
        }
        configured = true;
    }
    dim3 grid((unsigned)cdiv(p.m, kBM), (unsigned)cdiv(p.n, BN));
    dense_fast_kernel<BN><<<grid, kDenseThreads, smem, s>>>(p);
    CBRS_CHECK_LAUNCH("dense_fast");
    return CBRS_OK;
}

static bool fast_ok(const DenseParams &p) {
    auto al = [](const void *q) { return ((uintptr_t)q % 16) == 0; };
    if (p.n % 4 || p.n < 32 || p.f1 % kFBK || p.f2 % kFBK || p.ld1 % 4 || !al(p.x1) || !al(p.w)) return false;
    if (p.x2 && (p.ld2 % 4 || !al(p.x2))) return false;
    return true;
}

template <int BN>
static int launch_dense(const DenseParams &p, cudaStream_t s) {
    dim3 grid((unsigned)cdiv(p.m, kBM), (unsigned)cdiv(p.n, BN));
    dense_kernel<BN><<<grid, kDenseThreads, 0, s>>>(p);
    CBRS_CHECK_LAUNCH("dense");
    return CBRS_OK;
}

}  // namespace cbrs

using namespace cbrs;

static int dense_impl(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2, int64_t ld2,
                      const int64_t *idx2, int32_t f2, const float *w, const float *b, int64_t m, int32_t n, int act,
                      int rowop, const float *a_self, const float *a_neigh, float *p_out, float *q_out, float *out,
                      int64_t ldo, void *const *out_peers_host, void *const *q_peers_host, int n_peers, int out_dtype,
                      void *stream, int32_t group_h = 0, int64_t group_rows = 0) {
    CBRS_REQUIRE(x1 && w && out, CBRS_E_INVALID, "dense: null argument");
    CBRS_REQUIRE(group_h == 0 || (group_h > 0 && n % group_h == 0 && group_rows >= m && rowop == CBRS_ROWOP_NONE &&
                                  ldo >= group_h),
                 CBRS_E_INVALID, "dense: grouped output needs n %% group_h == 0, group_rows >= m, no row-op");
    CBRS_REQUIRE(out_dtype == CBRS_DTYPE_F32 || out_dtype == CBRS_DTYPE_BF16, CBRS_E_INVALID, "dense: out_dtype=%d", out_dtype);
    CBRS_REQUIRE(out_dtype == CBRS_DTYPE_F32 || !(rowop == CBRS_ROWOP_L2NORM && n > 128), CBRS_E_UNSUPPORTED,
                 "dense: the two-pass l2-normalise (n > 128) writes float32 only");
    CBRS_REQUIRE(n_peers >= 0 && n_peers < CBRS_MAX_PEERS && (n_peers == 0 || out_peers_host), CBRS_E_INVALID,
                 "dense: n_peers=%d (at most %d peer copies)", n_peers, CBRS_MAX_PEERS - 1);
    CBRS_REQUIRE(n_peers == 0 || rowop != CBRS_ROWOP_ATTN || q_peers_host, CBRS_E_INVALID,
                 "dense: the attention row-op with peers needs q_peers");
    CBRS_REQUIRE(n_peers == 0 || !(rowop == CBRS_ROWOP_L2NORM && n > 128), CBRS_E_UNSUPPORTED,
                 "dense: the two-pass l2-normalise (n > 128) has no peer form");
    CBRS_REQUIRE(m >= 0 && n > 0 && f1 > 0 && f2 >= 0 && ld1 >= f1 && (ldo >= n || group_h > 0), CBRS_E_INVALID,
                 "dense: m=%lld n=%d f1=%d f2=%d ld1=%lld ldo=%lld", (long long)m, n, f1, f2, (long long)ld1, (long long)ldo);
    CBRS_REQUIRE((x2 == nullptr) == (f2 == 0), CBRS_E_INVALID, "dense: second source and f2 disagree");
    CBRS_REQUIRE(!x2 || ld2 >= f2, CBRS_E_INVALID, "dense: ld2=%lld < f2=%d", (long long)ld2, f2);
    CBRS_REQUIRE(act >= CBRS_ACT_NONE && act <= CBRS_ACT_TANH, CBRS_E_INVALID, "dense: act=%d", act);
    CBRS_REQUIRE(rowop >= CBRS_ROWOP_NONE && rowop <= CBRS_ROWOP_ATTN, CBRS_E_INVALID, "dense: rowop=%d", rowop);
    CBRS_REQUIRE(rowop != CBRS_ROWOP_ATTN || (a_self && a_neigh && p_out && q_out && n <= 128), CBRS_E_INVALID,
                 "dense: attention row-op needs a_self/a_neigh/p_out/q_out and n <= 128");
    if (m == 0) return CBRS_OK;
    cudaStream_t s = (cudaStream_t)stream;
    DenseParams p{x1, ld1, idx1, f1, x2, ld2, idx2, f2, w, b, m, n, act, rowop, a_self, a_neigh, p_out, q_out, out, ldo};
    p.n_peer = n_peers;
    p.out_bf16 = out_dtype == CBRS_DTYPE_BF16;
    p.group_h = group_h;
    p.group_rows = group_rows;
    for (int q = 0; q < CBRS_MAX_PEERS - 1; ++q) {
        p.out_peer[q] = q < n_peers ? (float *)out_peers_host[q] : nullptr;
        p.q_peer[q] = (q < n_peers && q_peers_host) ? (float *)q_peers_host[q] : nullptr;
        if (q < n_peers)
            CBRS_REQUIRE(p.out_peer[q] && ((uintptr_t)p.out_peer[q] % 16) == ((uintptr_t)out % 16), CBRS_E_INVALID,
                         "dense: peer copy %d is null or aligned differently from the local output", q);
    }
    if (rowop == CBRS_ROWOP_L2NORM && n > 128) {  // row spans several column tiles: normalise in a second pass
        p.act = CBRS_ACT_NONE;
        p.rowop = CBRS_ROWOP_NONE;
        int rc = fast_ok(p) ? launch_dense_fast<128>(p, s) : launch_dense<128>(p, s);
        if (rc) return rc;
        row_l2norm_act_kernel<<<(unsigned)cdiv(m * 32, 256), 256, 0, s>>>(out, ldo, m, n, act);
        CBRS_CHECK_LAUNCH("row_l2norm_act");
        return CBRS_OK;
    }
    if (fast_ok(p)) {
        if (n <= 32) return launch_dense_fast<32>(p, s);
        if (n <= 64) return launch_dense_fast<64>(p, s);
        return launch_dense_fast<128>(p, s);
    }
    if (n <= 16) return launch_dense<16>(p, s);
    if (n <= 32) return launch_dense<32>(p, s);
    if (n <= 64) return launch_dense<64>(p, s);
    return launch_dense<128>(p, s);
}

extern "C" int cbrs_dense(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2, int64_t ld2,
                          const int64_t *idx2, int32_t f2, const float *w, const float *b, int64_t m, int32_t n,
                          int act, int rowop, const float *a_self, const float *a_neigh, float *p_out, float *q_out,
                          float *out, int64_t ldo, void *stream) {
    return dense_impl(x1, ld1, idx1, f1, x2, ld2, idx2, f2, w, b, m, n, act, rowop, a_self, a_neigh, p_out, q_out, out,
                      ldo, nullptr, nullptr, 0, CBRS_DTYPE_F32, stream);
}

extern "C" int cbrs_dense_bcast(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2,
                                int64_t ld2, const int64_t *idx2, int32_t f2, const float *w, const float *b, int64_t m,
                                int32_t n, int act, int rowop, const float *a_self, const float *a_neigh, float *p_out,
                                float *q_out, float *out, int64_t ldo, void *const *out_peers_host,
                                void *const *q_peers_host, int n_peers, void *stream) {
    return dense_impl(x1, ld1, idx1, f1, x2, ld2, idx2, f2, w, b, m, n, act, rowop, a_self, a_neigh, p_out, q_out, out,
                      ldo, out_peers_host, q_peers_host, n_peers, CBRS_DTYPE_F32, stream);
}

extern "C" int cbrs_dense_ex(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2, int64_t ld2,
                             const int64_t *idx2, int32_t f2, const float *w, const float *b, int64_t m, int32_t n, int act,
                             int rowop, const float *a_self, const float *a_neigh, float *p_out, float *q_out, void *out,
                             int64_t ldo, int out_dtype, void *const *out_peers_host, void *const *q_peers_host, int n_peers,
                             void *stream) {
    return dense_impl(x1, ld1, idx1, f1, x2, ld2, idx2, f2, w, b, m, n, act, rowop, a_self, a_neigh, p_out, q_out,
                      (float *)out, ldo, out_peers_host, q_peers_host, n_peers, out_dtype, stream);
}

extern "C" int cbrs_dense_grouped(const float *x, int64_t ldx, const float *w_cat, int64_t m, int32_t f, int32_t h,
                                  int32_t n_groups, void *out, int64_t ldo, int64_t group_rows, int out_dtype,
                                  void *const *out_peers_host, int n_peers, void *stream) {
    CBRS_REQUIRE(h > 0 && n_groups > 0, CBRS_E_INVALID, "dense_grouped: h=%d n_groups=%d", h, n_groups);
    return dense_impl(x, ldx, nullptr, f, nullptr, 0, nullptr, 0, w_cat, nullptr, m, h * n_groups, CBRS_ACT_NONE,
                      CBRS_ROWOP_NONE, nullptr, nullptr, nullptr, nullptr, (float *)out, ldo, out_peers_host, nullptr, n_peers,
                      out_dtype, stream, h, group_rows);
}
