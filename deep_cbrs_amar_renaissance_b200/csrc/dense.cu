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

namespace cbrs {

struct DenseParams {
    const float *x1; int64_t ld1; const int64_t *idx1; int32_t f1;
    const float *x2; int64_t ld2; const int64_t *idx2; int32_t f2;
    const float *w; const float *b;
    int64_t m; int32_t n;
    int act; int rowop;
    const float *a_self; const float *a_neigh; float *p_out; float *q_out;
    float *out; int64_t ldo;
};

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
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int ng = n0 + tx * TN + j;
            if (ng < p.n) p.out[m * p.ldo + ng] = apply_act(acc[i][j] * scale, p.act);
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

template <int BN>
static int launch_dense(const DenseParams &p, cudaStream_t s) {
    dim3 grid((unsigned)cdiv(p.m, kBM), (unsigned)cdiv(p.n, BN));
    dense_kernel<BN><<<grid, kDenseThreads, 0, s>>>(p);
    CBRS_CHECK_LAUNCH("dense");
    return CBRS_OK;
}

}  // namespace cbrs

using namespace cbrs;

extern "C" int cbrs_dense(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2, int64_t ld2,
                          const int64_t *idx2, int32_t f2, const float *w, const float *b, int64_t m, int32_t n,
                          int act, int rowop, const float *a_self, const float *a_neigh, float *p_out, float *q_out,
                          float *out, int64_t ldo, void *stream) {
    CBRS_REQUIRE(x1 && w && out, CBRS_E_INVALID, "dense: null argument");
    CBRS_REQUIRE(m >= 0 && n > 0 && f1 > 0 && f2 >= 0 && ld1 >= f1 && ldo >= n, CBRS_E_INVALID,
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
    if (rowop == CBRS_ROWOP_L2NORM && n > 128) {  // row spans several column tiles: normalise in a second pass
        p.act = CBRS_ACT_NONE;
        p.rowop = CBRS_ROWOP_NONE;
        int rc = launch_dense<128>(p, s);
        if (rc) return rc;
        row_l2norm_act_kernel<<<(unsigned)cdiv(m * 32, 256), 256, 0, s>>>(out, ldo, m, n, act);
        CBRS_CHECK_LAUNCH("row_l2norm_act");
        return CBRS_OK;
    }
    if (n <= 16) return launch_dense<16>(p, s);
    if (n <= 32) return launch_dense<32>(p, s);
    if (n <= 64) return launch_dense<64>(p, s);
    return launch_dense<128>(p, s);
}
