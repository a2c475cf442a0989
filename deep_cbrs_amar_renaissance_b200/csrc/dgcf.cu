// DGCF operator build + layer pieces (scope row (f)-3).
//
// The reference builds the DGCF propagation matrix on the host with scipy
// (/root/reference/src/layers/dgcf_conv.py:38-80): crosshop = A.dot(A); both A and crosshop go
// through gcn_filter; crosshop is high-pass filtered (entries > eps, eps picked from
// [1e-1, 1e-2, 1e-3, 5e-4] so that the surviving edge count is closest in ratio to A_hat's);
// result = A_hat + crosshop_filtered + I.  The layer is out = M (x * sigmoid(w)), w [N,1]
// (dgcf_conv.py:32-36,101-102).
//
// Device form.  A.dot(A) is expanded into its elementary products (i, j, a_ik * a_kj) - one warp per
// row i walks its edges (i,k) and copies row k scaled by a_ik - and the COO of products goes through the
// SAME sort / duplicate-sum / normalise pipeline as every other graph (cbrs_graph_build_csr with
// DEDUP_SUM | ADD_SELF_LOOPS | SYM_NORM == gcn_filter(crosshop)).  cbrs_count_above and
// cbrs_csr_filter_above implement the threshold search and the filter; the final sum of three sparse
// matrices is one more cbrs_graph_build_csr(DEDUP_SUM) over their concatenated entries.
#include "common.cuh"

namespace cbrs {

__global__ void spgemm_row_counts_kernel(const int64_t *__restrict__ rowptr_l, const int32_t *__restrict__ colidx_l,
                                         const int64_t *__restrict__ rowptr_r, int64_t n_rows, int64_t *__restrict__ counts) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int64_t acc = 0;
    for (int64_t e = rowptr_l[row] + lane; e < rowptr_l[row + 1]; e += 32) {
        const int k = colidx_l[e];
        acc += rowptr_r[k + 1] - rowptr_r[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) counts[row] = acc;
}

__global__ void spgemm_expand_kernel(const int64_t *__restrict__ rowptr_l, const int32_t *__restrict__ colidx_l,
                                     const float *__restrict__ vals_l, const int64_t *__restrict__ rowptr_r,
                                     const int32_t *__restrict__ colidx_r, const float *__restrict__ vals_r, int64_t n_rows,
                                     const int64_t *__restrict__ row_offsets, int32_t *__restrict__ coo_row,
                                     int32_t *__restrict__ coo_col, float *__restrict__ coo_val) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    int64_t out = row_offsets[row];
    for (int64_t e = rowptr_l[row]; e < rowptr_l[row + 1]; ++e) {   // ascending k: products of a row are emitted in a fixed order
        const int k = colidx_l[e];
        const float a = vals_l ? vals_l[e] : 1.f;
        const int64_t b = rowptr_r[k], n = rowptr_r[k + 1] - b;
        for (int64_t t = lane; t < n; t += 32) {
            coo_row[out + t] = (int32_t)row;
            coo_col[out + t] = colidx_r[b + t];
            coo_val[out + t] = a * (vals_r ? vals_r[b + t] : 1.f);
        }
        out += n;
    }
}

constexpr int kMaxEps = 8;
struct EpsList { float eps[kMaxEps]; int n; };

__global__ void count_above_kernel(const float *__restrict__ vals, int64_t n, EpsList el, unsigned long long *__restrict__ counts) {
    __shared__ unsigned long long sh[kMaxEps];
    if (threadIdx.x < kMaxEps) sh[threadIdx.x] = 0ull;
    __syncthreads();
    unsigned int local[kMaxEps];
#pragma unroll
    for (int j = 0; j < kMaxEps; ++j) local[j] = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = vals[i];
#pragma unroll
        for (int j = 0; j < kMaxEps; ++j)
            if (j < el.n && v > el.eps[j]) ++local[j];
    }
#pragma unroll
    for (int j = 0; j < kMaxEps; ++j)
        if (j < el.n && local[j]) atomicAdd(&sh[j], (unsigned long long)local[j]);
    __syncthreads();
    if (threadIdx.x < el.n && sh[threadIdx.x]) atomicAdd(&counts[threadIdx.x], sh[threadIdx.x]);  // integer: order-free
}

__global__ void above_flags_kernel(const float *__restrict__ vals, int64_t n, float eps, uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = vals[i] > eps ? 1u : 0u;
}

// one warp per row: every kept entry goes to its scanned slot with its row id
__global__ void above_scatter_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                                     const float *__restrict__ vals, int64_t n_rows, float eps,
                                     const uint32_t *__restrict__ scan, int32_t *__restrict__ coo_row,
                                     int32_t *__restrict__ coo_col, float *__restrict__ coo_val) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    for (int64_t e = rowptr[row] + lane; e < rowptr[row + 1]; e += 32) {
        const float v = vals[e];
        if (v > eps) {
            const uint32_t o = scan[e];
            coo_row[o] = (int32_t)row;
            coo_col[o] = colidx[e];
            coo_val[o] = v;
        }
    }
}

__global__ void row_gate_kernel(const float *__restrict__ x, int64_t ldx, const float *__restrict__ w, int64_t rows,
                                int32_t d, float *__restrict__ out, int64_t ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * d) return;
    const int64_t r = i / d;
    const int c = (int)(i % d);
    out[r * ldo + c] = x[r * ldx + c] * (1.f / (1.f + expf(-w[r])));
}

// dx = g * sigmoid(w);  dw[r] = sigmoid'(w[r]) * sum_c g[r,c] x[r,c]
__global__ void row_gate_grad_kernel(const float *__restrict__ g, int64_t ldg, const float *__restrict__ x, int64_t ldx,
                                     const float *__restrict__ w, int64_t rows, int32_t d, float *__restrict__ dx,
                                     int64_t lddx, float *__restrict__ dw) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float s = 1.f / (1.f + expf(-w[row]));
    float dot = 0.f;
    for (int c = lane; c < d; c += 32) {
        const float gv = g[row * ldg + c];
        dot = fmaf(gv, x[row * ldx + c], dot);
        dx[row * lddx + c] = gv * s;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) dw[row] = dot * s * (1.f - s);
}

}  // namespace cbrs

using namespace cbrs;

extern "C" size_t cbrs_spgemm_workspace_bytes(int64_t n_rows) { return scan_i64_workspace_bytes(n_rows + 1) + 512; }

extern "C" int cbrs_spgemm_count(const cbrs_csr_t *left, const cbrs_csr_t *right, int64_t *row_offsets, int64_t *total_out,
                                 void *workspace, size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(left && right && row_offsets && total_out, CBRS_E_INVALID, "spgemm_count: null argument");
    CBRS_REQUIRE(left->n_rows > 0 && left->rowptr && right->rowptr && (left->nnz == 0 || left->colidx), CBRS_E_INVALID,
                 "spgemm_count: bad descriptor");
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_spgemm_workspace_bytes(left->n_rows), CBRS_E_WORKSPACE,
                 "spgemm_count: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    spgemm_row_counts_kernel<<<(unsigned)cdiv(left->n_rows * 32, 256), 256, 0, s>>>(left->rowptr, left->colidx, right->rowptr,
                                                                                   left->n_rows, row_offsets);
    CBRS_CHECK_LAUNCH("spgemm_row_counts");
    return scan_i64_exclusive(row_offsets, left->n_rows, total_out, workspace, workspace_bytes, s);
}

extern "C" int cbrs_spgemm_expand(const cbrs_csr_t *left, const cbrs_csr_t *right, const int64_t *row_offsets,
                                  int32_t *coo_row, int32_t *coo_col, float *coo_val, void *stream) {
    CBRS_REQUIRE(left && right && row_offsets && coo_row && coo_col && coo_val, CBRS_E_INVALID, "spgemm_expand: null argument");
    spgemm_expand_kernel<<<(unsigned)cdiv(left->n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        left->rowptr, left->colidx, left->vals, right->rowptr, right->colidx, right->vals, left->n_rows, row_offsets, coo_row,
        coo_col, coo_val);
    CBRS_CHECK_LAUNCH("spgemm_expand");
    return CBRS_OK;
}

extern "C" int cbrs_count_above(const float *vals, int64_t n, const float *eps_host, int32_t n_eps, uint64_t *counts,
                                void *stream) {
    CBRS_REQUIRE(vals && eps_host && counts && n >= 0 && n_eps > 0 && n_eps <= kMaxEps, CBRS_E_INVALID,
                 "count_above: bad argument (at most %d thresholds)", kMaxEps);
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(counts, 0, sizeof(uint64_t) * n_eps, s);
    CBRS_REQUIRE(e == cudaSuccess, CBRS_E_CUDA, "count_above: %s", cudaGetErrorString(e));
    if (n == 0) return CBRS_OK;
    EpsList el;
    el.n = n_eps;
    for (int j = 0; j < kMaxEps; ++j) el.eps[j] = j < n_eps ? eps_host[j] : 0.f;
    const int64_t want = cdiv(n, 256 * 8);
    count_above_kernel<<<(unsigned)(want < 4 * kSMs ? (want < 1 ? 1 : want) : 4 * kSMs), 256, 0, s>>>(
        vals, n, el, (unsigned long long *)counts);
    CBRS_CHECK_LAUNCH("count_above");
    return CBRS_OK;
}

extern "C" size_t cbrs_csr_filter_above_workspace_bytes(int64_t nnz) {
    return align_up((size_t)nnz * 4) + scan_u32_workspace_bytes(nnz) + 512;
}

extern "C" int cbrs_csr_filter_above(const cbrs_csr_t *g, float eps, int32_t *coo_row, int32_t *coo_col, float *coo_val,
                                     void *workspace, size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(g && g->rowptr && g->vals && coo_row && coo_col && coo_val, CBRS_E_INVALID, "csr_filter_above: null argument");
    CBRS_REQUIRE(g->nnz < (int64_t)0xffffffffll, CBRS_E_INVALID, "csr_filter_above: nnz out of range");
    if (g->nnz == 0) return CBRS_OK;
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_csr_filter_above_workspace_bytes(g->nnz), CBRS_E_WORKSPACE,
                 "csr_filter_above: workspace too small");
    Arena a(workspace, workspace_bytes);
    uint32_t *flags = a.take<uint32_t>((size_t)g->nnz);
    cudaStream_t s = (cudaStream_t)stream;
    above_flags_kernel<<<(unsigned)cdiv(g->nnz, 256), 256, 0, s>>>(g->vals, g->nnz, eps, flags);
    CBRS_CHECK_LAUNCH("above_flags");
    int rc = scan_u32_exclusive(flags, g->nnz, nullptr, a.base + a.off, a.cap - a.off, s);
    if (rc) return rc;
    above_scatter_kernel<<<(unsigned)cdiv(g->n_rows * 32, 256), 256, 0, s>>>(g->rowptr, g->colidx, g->vals, g->n_rows, eps, flags,
                                                                             coo_row, coo_col, coo_val);
    CBRS_CHECK_LAUNCH("above_scatter");
    return CBRS_OK;
}

extern "C" int cbrs_row_gate(const float *x, int64_t ldx, const float *w, int64_t rows, int32_t d, float *out, int64_t ldo,
                             void *stream) {
    CBRS_REQUIRE(x && w && out && rows >= 0 && d > 0 && ldx >= d && ldo >= d, CBRS_E_INVALID, "row_gate: bad argument");
    if (rows == 0) return CBRS_OK;
    row_gate_kernel<<<(unsigned)cdiv(rows * d, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, w, rows, d, out, ldo);
    CBRS_CHECK_LAUNCH("row_gate");
    return CBRS_OK;
}

extern "C" int cbrs_row_gate_grad(const float *g, int64_t ldg, const float *x, int64_t ldx, const float *w, int64_t rows,
                                  int32_t d, float *dx, int64_t lddx, float *dw, void *stream) {
    CBRS_REQUIRE(g && x && w && dx && dw && rows >= 0 && d > 0 && ldg >= d && ldx >= d && lddx >= d, CBRS_E_INVALID,
                 "row_gate_grad: bad argument");
    if (rows == 0) return CBRS_OK;
    row_gate_grad_kernel<<<(unsigned)cdiv(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(g, ldg, x, ldx, w, rows, d, dx, lddx, dw);
    CBRS_CHECK_LAUNCH("row_gate_grad");
    return CBRS_OK;
}
