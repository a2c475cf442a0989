// Training-step kernels (scope row (f)-1): the backward of the forward hot path, the loss
// and the optimiser.  The reference trains through Keras `fit` with binary cross-entropy,
// an l2 penalty on the embeddings and GNN weights, and Adam
// (/root/reference/src/experiment.py:155-188, config.yaml:50-58, src/models/gnn.py:239-246);
// autograd does the differentiation there.  Here every derivative is an explicit kernel:
//
//   cbrs_act_grad        dPre = dOut * act'(out)                          (relu / sigmoid / tanh)
//   cbrs_dense_grad_w    dW = [X1[idx1] || X2[idx2]]^T dPre, db = colsum(dPre)
//                        fixed two-level reduction over row slabs => deterministic
//   cbrs_transpose_f32   W^T, so dX = dPre W^T runs on the forward dense kernel
//   cbrs_scatter_add_rows  dTable[idx[m]] += dRows[m] for the embedding lookups; stable sort
//                        of (idx, m), then each run is summed in ascending m => deterministic
//   cbrs_l2norm_relu_grad  backward of GraphSageConv's l2-normalise + relu
//   cbrs_scale_rows_inv_degree  rows / degree (backward of the mean aggregator: A^T (dAgg / deg))
//   cbrs_axpby2d         out = ca*a + cb*b on strided 2-D views (gradient accumulation)
//   cbrs_bce             Keras binary cross-entropy (probabilities clipped to [1e-7, 1-1e-7]) + dL/dp
//   cbrs_sum_squares     sum w^2 (the l2 penalty's value)
//   cbrs_adam_step       Keras Adam with the l2 gradient 2*l2*w folded in
//
// The sparse backward needs no new kernel: A_hat is symmetric, so dZ = A_hat dPre is
// cbrs_spmm_csr on the same CSR.
#include "common.cuh"

#include <string.h>

namespace cbrs {

// ------------------------------------------------------------------ activation backward
__global__ void act_grad_kernel(const float *__restrict__ dout, int64_t ldd, const float *__restrict__ out, int64_t ldo,
                                int64_t rows, int32_t d, int act, float *__restrict__ dpre, int64_t ldp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    const float g = dout[r * ldd + c];
    const float o = out[r * ldo + c];
    float v;
    switch (act) {
        case CBRS_ACT_RELU: v = o > 0.f ? g : 0.f; break;
        case CBRS_ACT_SIGMOID: v = g * (o * (1.f - o)); break;
        case CBRS_ACT_TANH: v = g * (1.f - o * o); break;
        default: v = g; break;
    }
    dpre[r * ldp + c] = v;
}

// ------------------------------------------------------------------ dW, db
// grid (ceil(K/32), ceil(N/32), slabs); each CTA reduces its slab of rows for a 32x32 tile of dW,
// 4 outputs per thread, and writes a partial; a second kernel adds the partials in slab order.
constexpr int kGwTile = 32, kGwRows = 32, kGwThreads = 256;

struct GradWParams {
    const float *x1; int64_t ld1; const int64_t *idx1; int32_t f1;
    const float *x2; int64_t ld2; const int64_t *idx2; int32_t f2;
    const float *dpre; int64_t ldd;
    int64_t m; int32_t n;
    int64_t rows_per_slab; int32_t slabs;
    float *partial;    // [slabs, K, n]
    float *partial_b;  // [slabs, n] or null
};

__global__ void __launch_bounds__(kGwThreads) grad_w_partial_kernel(const GradWParams p) {
    __shared__ float As[kGwRows][kGwTile + 1];
    __shared__ float Ds[kGwRows][kGwTile + 1];
    const int K = p.f1 + p.f2;
    const int k0 = blockIdx.x * kGwTile, n0 = blockIdx.y * kGwTile;
    const int64_t r0 = (int64_t)blockIdx.z * p.rows_per_slab;
    const int64_t r1 = r0 + p.rows_per_slab < p.m ? r0 + p.rows_per_slab : p.m;
    const int tid = threadIdx.x;
    const int tk = tid >> 3;        // 0..31: row of the dW tile (k)
    const int tn = (tid & 7) * 4;   // 4 consecutive columns (n)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float accb = 0.f;               // bias partial: threads 0..31 of the k-tile 0 CTAs own column tid
    for (int64_t rb = r0; rb < r1; rb += kGwRows) {
#pragma unroll
        for (int it = 0; it < (kGwRows * kGwTile) / kGwThreads; ++it) {
            const int e = tid + it * kGwThreads;
            const int rr = e / kGwTile, cc = e % kGwTile;
            const int64_t m = rb + rr;
            float a = 0.f, dv = 0.f;
            if (m < r1) {
                const int kg = k0 + cc;
                if (kg < K) {
                    if (kg < p.f1) a = p.x1[(p.idx1 ? p.idx1[m] : m) * p.ld1 + kg];
                    else a = p.x2[(p.idx2 ? p.idx2[m] : m) * p.ld2 + (kg - p.f1)];
                }
                const int ng = n0 + cc;
                if (ng < p.n) dv = p.dpre[m * p.ldd + ng];
            }
            As[rr][cc] = a;
            Ds[rr][cc] = dv;
        }
        __syncthreads();
#pragma unroll 8
        for (int rr = 0; rr < kGwRows; ++rr) {
            const float a = As[rr][tk];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(a, Ds[rr][tn + j], acc[j]);
        }
        if (p.partial_b && blockIdx.x == 0 && tid < kGwTile) {
#pragma unroll 8
            for (int rr = 0; rr < kGwRows; ++rr) accb += Ds[rr][tid];
        }
        __syncthreads();
    }
    const int kg = k0 + tk;
    if (kg < K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int ng = n0 + tn + j;
            if (ng < p.n) p.partial[((int64_t)blockIdx.z * K + kg) * p.n + ng] = acc[j];
        }
    }
    if (p.partial_b && blockIdx.x == 0 && tid < kGwTile && n0 + tid < p.n)
        p.partial_b[(int64_t)blockIdx.z * p.n + n0 + tid] = accb;
}

__global__ void grad_w_finish_kernel(const float *__restrict__ partial, int32_t slabs, int64_t elems,
                                     float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= elems) return;
    float acc = 0.f;
    for (int s = 0; s < slabs; ++s) acc += partial[(int64_t)s * elems + i];
    out[i] = acc;
}

__global__ void transpose_kernel(const float *__restrict__ src, int32_t rows, int32_t cols, float *__restrict__ dst) {
    __shared__ float tile[32][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int r = blockIdx.y * 32 + j;
        if (r < rows && c < cols) tile[j][threadIdx.x] = src[(int64_t)r * cols + c];
    }
    __syncthreads();
    const int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
        const int c2 = blockIdx.x * 32 + j;
        if (c2 < cols && r2 < rows) dst[(int64_t)c2 * rows + r2] = tile[threadIdx.x][j];
    }
}

// ------------------------------------------------------------------ scatter-add (embedding lookup backward)
__global__ void scatter_keys_kernel(const int64_t *__restrict__ idx, int64_t m, uint64_t *__restrict__ keys,
                                    uint32_t *__restrict__ payload) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    keys[i] = (uint64_t)idx[i];
    payload[i] = (uint32_t)i;
}

// one group of 32 lanes per sorted position that STARTS a run of equal indices; it adds the run's
// source rows in sorted (= ascending m, the sort is stable) order into the table row
__global__ void scatter_runs_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ payload, int64_t m,
                                    const float *__restrict__ src, int64_t lds, int32_t d, float *__restrict__ dst,
                                    int64_t ldd) {
    const int64_t pos = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (pos >= m) return;
    const uint64_t key = keys[pos];
    if (pos > 0 && keys[pos - 1] == key) return;
    for (int c = lane; c < d; c += 32) {
        float acc = dst[(int64_t)key * ldd + c];
        for (int64_t q = pos; q < m && keys[q] == key; ++q) acc += src[(int64_t)payload[q] * lds + c];
        dst[(int64_t)key * ldd + c] = acc;
    }
}

// ------------------------------------------------------------------ GraphSage l2-normalise + relu backward
// n = v / s, s = sqrt(max(sum v^2, 1e-12)), out = relu(n):  dn = dout * (n > 0);
// dv = (dn - n * (n . dn)) / s   (the floor branch, sum v^2 < 1e-12, is a plain scale: dv = dn / 1e-6)
__global__ void l2norm_relu_grad_kernel(const float *__restrict__ v, int64_t ldv, const float *__restrict__ dout,
                                        int64_t ldd, int64_t rows, int32_t d, int relu, float *__restrict__ dv,
                                        int64_t ldo) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float x = v[row * ldv + c]; ss = fmaf(x, x, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const bool floored = ss < 1e-12f;
    const float inv = 1.f / sqrtf(fmaxf(ss, 1e-12f));
    float dot = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float n = v[row * ldv + c] * inv;
        const float g = (!relu || n > 0.f) ? dout[row * ldd + c] : 0.f;
        dot = fmaf(n, g, dot);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    for (int c = lane; c < d; c += 32) {
        const float n = v[row * ldv + c] * inv;
        const float g = (!relu || n > 0.f) ? dout[row * ldd + c] : 0.f;
        dv[row * ldo + c] = floored ? g * inv : (g - n * dot) * inv;
    }
}

__global__ void l2norm_act_kernel(const float *__restrict__ v, int64_t ldv, int64_t rows, int32_t d, int relu,
                                  float *__restrict__ out, int64_t ldo) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float x = v[row * ldv + c]; ss = fmaf(x, x, ss); }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.f / sqrtf(fmaxf(ss, 1e-12f));
    for (int c = lane; c < d; c += 32) {
        const float n = v[row * ldv + c] * inv;
        out[row * ldo + c] = relu ? fmaxf(n, 0.f) : n;
    }
}

__global__ void scale_rows_inv_degree_kernel(const float *__restrict__ x, int64_t ldx, const int64_t *__restrict__ rowptr,
                                             int64_t rows, int32_t d, float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    const int64_t deg = rowptr[r + 1] - rowptr[r];
    out[r * ldo + c] = deg > 0 ? x[r * ldx + c] / (float)deg : 0.f;
}

__global__ void axpby2d_kernel(const float *__restrict__ a, int64_t lda, float ca, const float *__restrict__ b, int64_t ldb,
                               float cb, int64_t rows, int32_t d, float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    float v = ca * a[r * lda + c];
    if (b) v = fmaf(cb, b[r * ldb + c], v);
    out[r * ldo + c] = v;
}

// ------------------------------------------------------------------ loss
// One CTA; a fixed strided + tree reduction => the loss value is reproducible.
__global__ void __launch_bounds__(1024) bce_kernel(const float *__restrict__ p, const float *__restrict__ y, int64_t n,
                                                  float *__restrict__ loss_out, float *__restrict__ dp,
                                                  float *__restrict__ correct_out) {
    __shared__ float sh[1024], shc[1024];
    float acc = 0.f, cor = 0.f;
    const float inv_n = 1.f / (float)n;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const float pi = p[i], yi = y[i];
        const float pc = fminf(fmaxf(pi, 1e-7f), 1.f - 1e-7f);
        acc += -(yi * logf(pc) + (1.f - yi) * logf(1.f - pc));
        const bool inside = pi >= 1e-7f && pi <= 1.f - 1e-7f;  // clip_by_value passes no gradient outside
        if (dp) dp[i] = inside ? (-(yi / pc) + (1.f - yi) / (1.f - pc)) * inv_n : 0.f;
        cor += ((pi > 0.5f) == (yi > 0.5f)) ? 1.f : 0.f;
    }
    sh[threadIdx.x] = acc;
    shc[threadIdx.x] = cor;
    __syncthreads();
    for (int s = 512; s; s >>= 1) {
        if (threadIdx.x < s) { sh[threadIdx.x] += sh[threadIdx.x + s]; shc[threadIdx.x] += shc[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *loss_out = sh[0] * inv_n;
        if (correct_out) *correct_out = shc[0];
    }
}

// two levels, both in a fixed order => reproducible: each CTA reduces a contiguous slice (strided + tree), one CTA
// then adds the per-CTA partials in ascending order
constexpr int kSsqBlocks = 592;  // 148 SMs x 4
__global__ void __launch_bounds__(1024) sum_squares_partial_kernel(const float *__restrict__ w, int64_t n, int64_t per_block,
                                                                  float *__restrict__ partial) {
    __shared__ float sh[1024];
    const int64_t b = (int64_t)blockIdx.x * per_block;
    const int64_t e = b + per_block < n ? b + per_block : n;
    float acc = 0.f;
    for (int64_t i = b + threadIdx.x; i < e; i += 1024) acc = fmaf(w[i], w[i], acc);
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 512; s; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void sum_squares_finish_kernel(const float *__restrict__ partial, int blocks, float scale, float *__restrict__ out,
                                          int accumulate) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float acc = 0.f;
    for (int b = 0; b < blocks; ++b) acc += partial[b];
    *out = (accumulate ? *out : 0.f) + scale * acc;
}

// Keras Adam (optimizer_v2): lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t);  m, v moving averages;
// w -= lr_t * m / (sqrt(v) + eps).  The l2 regulariser's gradient 2*l2*w joins g here.
__global__ void adam_kernel(float *__restrict__ w, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                            int64_t n, float lr_t, const float *__restrict__ lr_t_dev, float b1, float b2, float eps,
                            float l2) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lr_t_dev) lr_t = *lr_t_dev;  // CUDA-graph replay: the step-dependent rate lives in device memory
    const float wi = w[i];
    const float gi = fmaf(2.f * l2, wi, g[i]);
    const float mi = fmaf(1.f - b1, gi - m[i], m[i]);
    const float vi = fmaf(1.f - b2, gi * gi - v[i], v[i]);
    m[i] = mi;
    v[i] = vi;
    w[i] = wi - lr_t * mi / (sqrtf(vi) + eps);
}

constexpr int kAdamMaxTensors = 48;
struct AdamMultiParams {
    float *w[kAdamMaxTensors];
    const float *g[kAdamMaxTensors];
    float *m[kAdamMaxTensors];
    float *v[kAdamMaxTensors];
    int64_t n[kAdamMaxTensors];
    float l2[kAdamMaxTensors];
    int32_t block_end[kAdamMaxTensors];  // exclusive prefix of 256-thread blocks
    int n_tensors;
    float lr_t;
    const float *lr_t_dev;
    float b1, b2, eps;
};

// one launch for all weight tensors of the model (a step updates ~40 small tensors)
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamMultiParams p) {
    int t = 0;
    while (t < p.n_tensors - 1 && (int)blockIdx.x >= p.block_end[t]) ++t;
    const int first = t == 0 ? 0 : p.block_end[t - 1];
    const int64_t i = (int64_t)(blockIdx.x - first) * 256 + threadIdx.x;
    if (i >= p.n[t]) return;
    const float lr_t = p.lr_t_dev ? *p.lr_t_dev : p.lr_t;
    const float wi = p.w[t][i];
    const float gi = fmaf(2.f * p.l2[t], wi, p.g[t][i]);
    const float mi = fmaf(1.f - p.b1, gi - p.m[t][i], p.m[t][i]);
    const float vi = fmaf(1.f - p.b2, gi * gi - p.v[t][i], p.v[t][i]);
    p.m[t][i] = mi;
    p.v[t][i] = vi;
    p.w[t][i] = wi - lr_t * mi / (sqrtf(vi) + p.eps);
}

}  // namespace cbrs

using namespace cbrs;

extern "C" int cbrs_adam_step_multi(int32_t n_tensors, void *const *w_host, const void *const *g_host, void *const *m_host,
                                    void *const *v_host, const int64_t *n_host, const float *l2_host, float lr_t,
                                    const float *lr_t_dev, float beta1, float beta2, float eps, void *stream) {
    CBRS_REQUIRE(n_tensors >= 0 && w_host && g_host && m_host && v_host && n_host, CBRS_E_INVALID,
                 "adam_step_multi: bad argument");
    for (int base = 0; base < n_tensors; base += kAdamMaxTensors) {
        AdamMultiParams p;
        memset(&p, 0, sizeof(p));
        const int cnt = n_tensors - base < kAdamMaxTensors ? n_tensors - base : kAdamMaxTensors;
        int64_t blocks = 0;
        for (int t = 0; t < cnt; ++t) {
            CBRS_REQUIRE(w_host[base + t] && g_host[base + t] && m_host[base + t] && v_host[base + t] && n_host[base + t] > 0,
                         CBRS_E_INVALID, "adam_step_multi: tensor %d", base + t);
            p.w[t] = (float *)w_host[base + t]; p.g[t] = (const float *)g_host[base + t];
            p.m[t] = (float *)m_host[base + t]; p.v[t] = (float *)v_host[base + t];
            p.n[t] = n_host[base + t]; p.l2[t] = l2_host ? l2_host[base + t] : 0.f;
            blocks += cdiv(p.n[t], 256);
            CBRS_REQUIRE(blocks < ((int64_t)1 << 31), CBRS_E_INVALID, "adam_step_multi: too many elements");
            p.block_end[t] = (int32_t)blocks;
        }
        p.n_tensors = cnt; p.lr_t = lr_t; p.lr_t_dev = lr_t_dev; p.b1 = beta1; p.b2 = beta2; p.eps = eps;
        adam_multi_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p);
        CBRS_CHECK_LAUNCH("adam_step_multi");
    }
    return CBRS_OK;
}

extern "C" int cbrs_act_grad(const float *dout, int64_t ldd, const float *out, int64_t ldo, int64_t rows, int32_t d,
                             int act, float *dpre, int64_t ldp, void *stream) {
    CBRS_REQUIRE(dout && out && dpre, CBRS_E_INVALID, "act_grad: null argument");
    CBRS_REQUIRE(rows >= 0 && d > 0 && ldd >= d && ldo >= d && ldp >= d, CBRS_E_INVALID, "act_grad: bad shape");
    CBRS_REQUIRE(act >= CBRS_ACT_NONE && act <= CBRS_ACT_TANH, CBRS_E_INVALID, "act_grad: act=%d", act);
    if (rows == 0) return CBRS_OK;
    act_grad_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(dout, ldd, out, ldo, rows, d, act, dpre,
                                                                                     ldp);
    CBRS_CHECK_LAUNCH("act_grad");
    return CBRS_OK;
}

// Row slabs: enough of them that tiles x slabs fills the chip twice (a batch-sized GEMM has only a few
// 32x32 output tiles), never fewer than 32 rows per slab, at most 256.  A pure function of the shape,
// so the reduction tree - and the bits of dW - are the same on every run.
static int grad_w_slabs(int64_t m, int32_t k, int32_t n) {
    const int64_t tiles = cdiv(k, kGwTile) * cdiv(n, kGwTile);
    int64_t want = cdiv(2 * kSMs, tiles);
    const int64_t most = cdiv(m, kGwRows);
    if (want > most) want = most;
    if (want > 256) want = 256;
    return (int)(want < 1 ? 1 : want);
}

extern "C" size_t cbrs_dense_grad_w_workspace_bytes(int64_t m, int32_t k, int32_t n) {
    const size_t slabs = (size_t)grad_w_slabs(m, k, n);
    return align_up(slabs * (size_t)k * (size_t)n * sizeof(float)) + align_up(slabs * (size_t)n * sizeof(float)) + 256;
}

extern "C" int cbrs_dense_grad_w(const float *x1, int64_t ld1, const int64_t *idx1, int32_t f1, const float *x2,
                                 int64_t ld2, const int64_t *idx2, int32_t f2, const float *dpre, int64_t ldd, int64_t m,
                                 int32_t n, float *dw, float *db, void *workspace, size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(x1 && dpre && dw, CBRS_E_INVALID, "dense_grad_w: null argument");
    CBRS_REQUIRE(m > 0 && n > 0 && f1 > 0 && f2 >= 0 && ld1 >= f1 && ldd >= n, CBRS_E_INVALID, "dense_grad_w: bad shape");
    CBRS_REQUIRE((x2 == nullptr) == (f2 == 0) && (!x2 || ld2 >= f2), CBRS_E_INVALID, "dense_grad_w: second source");
    const int K = f1 + f2;
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_dense_grad_w_workspace_bytes(m, K, n), CBRS_E_WORKSPACE,
                 "dense_grad_w: workspace too small");
    GradWParams p;
    p.x1 = x1; p.ld1 = ld1; p.idx1 = idx1; p.f1 = f1; p.x2 = x2; p.ld2 = ld2; p.idx2 = idx2; p.f2 = f2;
    p.dpre = dpre; p.ldd = ldd; p.m = m; p.n = n;
    p.slabs = grad_w_slabs(m, K, n);
    p.rows_per_slab = cdiv(cdiv(m, p.slabs), kGwRows) * kGwRows;
    Arena a(workspace, workspace_bytes);
    p.partial = a.take<float>((size_t)p.slabs * K * n);
    p.partial_b = db ? a.take<float>((size_t)p.slabs * n) : nullptr;
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((unsigned)cdiv(K, kGwTile), (unsigned)cdiv(n, kGwTile), (unsigned)p.slabs);
    grad_w_partial_kernel<<<grid, kGwThreads, 0, s>>>(p);
    CBRS_CHECK_LAUNCH("grad_w_partial");
    grad_w_finish_kernel<<<(unsigned)cdiv((int64_t)K * n, 256), 256, 0, s>>>(p.partial, p.slabs, (int64_t)K * n, dw);
    CBRS_CHECK_LAUNCH("grad_w_finish");
    if (db) {
        grad_w_finish_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(p.partial_b, p.slabs, n, db);
        CBRS_CHECK_LAUNCH("grad_b_finish");
    }
    return CBRS_OK;
}

extern "C" int cbrs_transpose_f32(const float *src, int32_t rows, int32_t cols, float *dst, void *stream) {
    CBRS_REQUIRE(src && dst && rows > 0 && cols > 0, CBRS_E_INVALID, "transpose: bad argument");
    dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, rows, cols, dst);
    CBRS_CHECK_LAUNCH("transpose");
    return CBRS_OK;
}

extern "C" size_t cbrs_scatter_add_rows_workspace_bytes(int64_t m) {
    return align_up((size_t)m * sizeof(uint64_t)) + align_up((size_t)m * sizeof(uint32_t)) + sort_workspace_bytes(m) + 512;
}

extern "C" int cbrs_scatter_add_rows(const float *src, int64_t lds, const int64_t *idx, int64_t m, int32_t d,
                                     int64_t n_rows, float *dst, int64_t ldd, void *workspace, size_t workspace_bytes,
                                     void *stream) {
    CBRS_REQUIRE(src && idx && dst, CBRS_E_INVALID, "scatter_add_rows: null argument");
    CBRS_REQUIRE(m >= 0 && d > 0 && lds >= d && ldd >= d && n_rows > 0 && m < ((int64_t)1 << 32), CBRS_E_INVALID,
                 "scatter_add_rows: bad shape");
    if (m == 0) return CBRS_OK;
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_scatter_add_rows_workspace_bytes(m), CBRS_E_WORKSPACE,
                 "scatter_add_rows: workspace too small");
    Arena a(workspace, workspace_bytes);
    uint64_t *keys = a.take<uint64_t>((size_t)m);
    uint32_t *payload = a.take<uint32_t>((size_t)m);
    void *sort_ws = a.base + a.off;
    const size_t sort_bytes = a.cap - a.off;
    cudaStream_t s = (cudaStream_t)stream;
    scatter_keys_kernel<<<(unsigned)cdiv(m, 256), 256, 0, s>>>(idx, m, keys, payload);
    CBRS_CHECK_LAUNCH("scatter_keys");
    int bits = 1;
    while (((int64_t)1 << bits) < n_rows) ++bits;
    int rc = sort_pairs_u64(keys, payload, m, bits, sort_ws, sort_bytes, s);
    if (rc) return rc;
    scatter_runs_kernel<<<(unsigned)cdiv(m * 32, 256), 256, 0, s>>>(keys, payload, m, src, lds, d, dst, ldd);
    CBRS_CHECK_LAUNCH("scatter_runs");
    return CBRS_OK;
}

extern "C" int cbrs_l2norm_relu_grad(const float *v, int64_t ldv, const float *dout, int64_t ldd, int64_t rows, int32_t d,
                                     int relu, float *dv, int64_t ldo, void *stream) {
    CBRS_REQUIRE(v && dout && dv, CBRS_E_INVALID, "l2norm_relu_grad: null argument");
    CBRS_REQUIRE(rows >= 0 && d > 0 && ldv >= d && ldd >= d && ldo >= d, CBRS_E_INVALID, "l2norm_relu_grad: bad shape");
    if (rows == 0) return CBRS_OK;
    l2norm_relu_grad_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(v, ldv, dout, ldd, rows, d,
                                                                                              relu, dv, ldo);
    CBRS_CHECK_LAUNCH("l2norm_relu_grad");
    return CBRS_OK;
}

extern "C" int cbrs_l2norm_act(const float *v, int64_t ldv, int64_t rows, int32_t d, int relu, float *out, int64_t ldo,
                               void *stream) {
    CBRS_REQUIRE(v && out, CBRS_E_INVALID, "l2norm_act: null argument");
    CBRS_REQUIRE(rows >= 0 && d > 0 && ldv >= d && ldo >= d, CBRS_E_INVALID, "l2norm_act: bad shape");
    if (rows == 0) return CBRS_OK;
    l2norm_act_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(v, ldv, rows, d, relu, out, ldo);
    CBRS_CHECK_LAUNCH("l2norm_act");
    return CBRS_OK;
}

extern "C" int cbrs_scale_rows_inv_degree(const float *x, int64_t ldx, const int64_t *rowptr, int64_t rows, int32_t d,
                                          float *out, int64_t ldo, void *stream) {
    CBRS_REQUIRE(x && rowptr && out, CBRS_E_INVALID, "scale_rows_inv_degree: null argument");
    CBRS_REQUIRE(rows >= 0 && d > 0 && ldx >= d && ldo >= d, CBRS_E_INVALID, "scale_rows_inv_degree: bad shape");
    if (rows == 0) return CBRS_OK;
    scale_rows_inv_degree_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, rowptr, rows, d,
                                                                                                 out, ldo);
    CBRS_CHECK_LAUNCH("scale_rows_inv_degree");
    return CBRS_OK;
}

extern "C" int cbrs_axpby2d(const float *a, int64_t lda, float ca, const float *b, int64_t ldb, float cb, int64_t rows,
                            int32_t d, float *out, int64_t ldo, void *stream) {
    CBRS_REQUIRE(a && out, CBRS_E_INVALID, "axpby2d: null argument");
    CBRS_REQUIRE(rows >= 0 && d > 0 && lda >= d && ldo >= d && (!b || ldb >= d), CBRS_E_INVALID, "axpby2d: bad shape");
    if (rows == 0) return CBRS_OK;
    axpby2d_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(a, lda, ca, b, ldb, cb, rows, d, out, ldo);
    CBRS_CHECK_LAUNCH("axpby2d");
    return CBRS_OK;
}

extern "C" int cbrs_bce(const float *p, const float *y, int64_t n, float *loss_out, float *dp_out, float *correct_out,
                        void *stream) {
    CBRS_REQUIRE(p && y && loss_out && n > 0, CBRS_E_INVALID, "bce: bad argument");
    bce_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p, y, n, loss_out, dp_out, correct_out);
    CBRS_CHECK_LAUNCH("bce");
    return CBRS_OK;
}

extern "C" size_t cbrs_sum_squares_workspace_bytes(void) { return kSsqBlocks * sizeof(float) + 256; }

extern "C" int cbrs_sum_squares(const float *w, int64_t n, float scale, float *out, int accumulate, void *workspace,
                                size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(w && out && n > 0, CBRS_E_INVALID, "sum_squares: bad argument");
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_sum_squares_workspace_bytes(), CBRS_E_WORKSPACE,
                 "sum_squares: workspace too small");
    int64_t per_block = cdiv(n, kSsqBlocks);
    per_block = cdiv(per_block, 1024) * 1024;
    const int blocks = (int)cdiv(n, per_block);
    cudaStream_t s = (cudaStream_t)stream;
    sum_squares_partial_kernel<<<blocks, 1024, 0, s>>>(w, n, per_block, (float *)workspace);
    CBRS_CHECK_LAUNCH("sum_squares_partial");
    sum_squares_finish_kernel<<<1, 32, 0, s>>>((const float *)workspace, blocks, scale, out, accumulate);
    CBRS_CHECK_LAUNCH("sum_squares_finish");
    return CBRS_OK;
}

extern "C" int cbrs_adam_step(float *w, const float *g, float *m, float *v, int64_t n, float lr_t,
                              const float *lr_t_dev, float beta1, float beta2, float eps, float l2, void *stream) {
    CBRS_REQUIRE(w && g && m && v && n > 0, CBRS_E_INVALID, "adam_step: bad argument");
    adam_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, g, m, v, n, lr_t, lr_t_dev, beta1, beta2, eps,
                                                                          l2);
    CBRS_CHECK_LAUNCH("adam_step");
    return CBRS_OK;
}

// ------------------------------------------------------------------ hybrid tweaks: attention fusion, residual add
// FusionLayer('attention') (/root/reference/src/layers/fusion.py:56-68): x = stack([a, b]); att = softmax_over_the_two(
// tanh(x @ W)); out = sum(att * x).  With ta = tanh(a W), tb = tanh(b W) (two cbrs_dense calls, shared W):
// s = sigmoid(ta - tb); out = s*a + (1-s)*b.
namespace cbrs {
__global__ void attn_fuse_kernel(const float *__restrict__ a, int64_t lda, const float *__restrict__ b, int64_t ldb,
                                 const float *__restrict__ ta, const float *__restrict__ tb, int64_t rows, int32_t d,
                                 float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    const float ea = expf(ta[i]), eb = expf(tb[i]);  // tanh output: |t| <= 1, no overflow; same form as tf.math.softmax
    const float inv = 1.f / (ea + eb);
    out[r * ldo + c] = (ea * inv) * a[r * lda + c] + (eb * inv) * b[r * ldb + c];
}
// da = g*s, db = g*(1-s) (direct paths); dta = g*(a-b)*s*(1-s), dtb = -dta (through the softmax)
__global__ void attn_fuse_grad_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ a, int64_t lda,
                                      const float *__restrict__ b, int64_t ldb, const float *__restrict__ ta,
                                      const float *__restrict__ tb, int64_t rows, int32_t d, float *__restrict__ da,
                                      float *__restrict__ db, float *__restrict__ dta, float *__restrict__ dtb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    const float ea = expf(ta[i]), eb = expf(tb[i]);
    const float s = ea / (ea + eb);
    const float gv = g[r * ldg + c];
    da[i] = gv * s;
    db[i] = gv * (1.f - s);
    const float t = gv * (a[r * lda + c] - b[r * ldb + c]) * s * (1.f - s);
    dta[i] = t;
    dtb[i] = -t;
}
// out = act(a + b + c): the residual classifier's activation(residual(x) + x1 + x2) (src/models/hybrid.py:89)
__global__ void add3_act_kernel(const float *__restrict__ a, int64_t lda, const float *__restrict__ b, int64_t ldb,
                                const float *__restrict__ c3, int64_t ldc, int64_t rows, int32_t d, int act,
                                float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    float v = (a[r * lda + c] + b[r * ldb + c]) + c3[r * ldc + c];
    switch (act) {
        case CBRS_ACT_RELU: v = fmaxf(v, 0.f); break;
        case CBRS_ACT_SIGMOID: v = 1.f / (1.f + expf(-v)); break;
        case CBRS_ACT_TANH: v = tanhf(v); break;
        default: break;
    }
    out[r * ldo + c] = v;
}
}  // namespace cbrs

extern "C" int cbrs_attn_fuse(const float *a, int64_t lda, const float *b, int64_t ldb, const float *ta, const float *tb,
                              int64_t rows, int32_t d, float *out, int64_t ldo, void *stream) {
    CBRS_REQUIRE(a && b && ta && tb && out && rows >= 0 && d > 0 && lda >= d && ldb >= d && ldo >= d, CBRS_E_INVALID,
                 "attn_fuse: bad argument");
    if (rows == 0) return CBRS_OK;
    attn_fuse_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, ta, tb, rows, d, out, ldo);
    CBRS_CHECK_LAUNCH("attn_fuse");
    return CBRS_OK;
}

extern "C" int cbrs_attn_fuse_grad(const float *g, int64_t ldg, const float *a, int64_t lda, const float *b, int64_t ldb,
                                   const float *ta, const float *tb, int64_t rows, int32_t d, float *da, float *db,
                                   float *dta, float *dtb, void *stream) {
    CBRS_REQUIRE(g && a && b && ta && tb && da && db && dta && dtb && rows >= 0 && d > 0 && ldg >= d && lda >= d && ldb >= d,
                 CBRS_E_INVALID, "attn_fuse_grad: bad argument");
    if (rows == 0) return CBRS_OK;
    attn_fuse_grad_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, a, lda, b, ldb, ta, tb, rows,
                                                                                          d, da, db, dta, dtb);
    CBRS_CHECK_LAUNCH("attn_fuse_grad");
    return CBRS_OK;
}

extern "C" int cbrs_add3_act(const float *a, int64_t lda, const float *b, int64_t ldb, const float *c, int64_t ldc,
                             int64_t rows, int32_t d, int act, float *out, int64_t ldo, void *stream) {
    CBRS_REQUIRE(a && b && c && out && rows >= 0 && d > 0 && lda >= d && ldb >= d && ldc >= d && ldo >= d, CBRS_E_INVALID,
                 "add3_act: bad argument");
    CBRS_REQUIRE(act >= CBRS_ACT_NONE && act <= CBRS_ACT_TANH, CBRS_E_INVALID, "add3_act: act=%d", act);
    if (rows == 0) return CBRS_OK;
    add3_act_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, c, ldc, rows, d, act, out, ldo);
    CBRS_CHECK_LAUNCH("add3_act");
    return CBRS_OK;
}
