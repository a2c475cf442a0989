// Fused GAT layer for sm_100a: edge score + edge softmax + aggregation in one pass.  Row P3.
//
// Replaces spektral GATConv._call_single as built at /root/reference/src/models/gnn.py:321-328
// (attn_heads=1, add_self_loops=True, dropout 0): 4 gathers + unsorted_segment_max/sum/sum +
// softmax, unfused, in the reference.  Restated semantics: SURVEY.md Appendix A.4.
//
// Per chunk of a row (G lanes): one pass - a batch of G edges gathers q[col] (4 B per edge,
// L2-resident), the running max is raised (online softmax: accumulator and sums rescaled when
// it moves), w = exp(e - max) one edge per lane, then the H-wide z rows stream exactly like in
// the SpMM kernel.  The edge set {A minus self loops} +
// {(i,i)} is formed on the fly: existing diagonal entries are masked and the first chunk
// of each row adds the self edge.  Heavy rows park (max, sum, acc) per chunk and a merge
// kernel rescales them in ascending chunk order (fixed tree => launch-shape independent).
#include "common.cuh"

namespace cbrs {

struct GatParams {
    const int64_t *rowptr;
    const int32_t *colidx;
    const int32_t *chunk_row;
    const int64_t *chunk_begin;
    const int32_t *chunk_slot;
    const int32_t *chunk_len;
    int64_t n_chunks;
    int32_t chunk_edges;
    int64_t row_offset;
    const float *z;
    int64_t ldz;
    const float *p;  // [N] global
    const float *q;  // [N] global
    float *y;
    int64_t ldy;
    int32_t h;
    const float *bias;
    int relu;
    float *part_acc;  // [n_slots, h]
    float *part_ms;   // [n_slots, 2]
    const int32_t *heavy_row;
    const int64_t *heavy_slot_ptr;
    int64_t n_heavy;
    float *y_peer[CBRS_MAX_PEERS - 1];  // multi-GPU: peer-mapped copies of y (same layout)
    int n_peer;
};

constexpr int kGatThreads = 256;

__device__ __forceinline__ float leaky02(float x) { return x > 0.f ? x : 0.2f * x; }

template <int G>
__device__ __forceinline__ float group_max(float v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(m, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o; o >>= 1) v += __shfl_xor_sync(m, v, o, G);
    return v;
}

// As for the SpMM kernel (profiles/r01_tune_spmm.log), occupancy beats per-warp memory-level parallelism for this
// gather: 4 row loads in flight at 64 registers (32 warps per SM) instead of 8 at 104 registers (16 warps).
template <int G, int VEC>
__global__ void __launch_bounds__(kGatThreads, 4) gat_chunk_kernel(const GatParams p) {
    constexpr int U = G < 4 ? G : 4;
    const int64_t gid = ((int64_t)blockIdx.x * kGatThreads + threadIdx.x) / G;
    if (gid >= p.n_chunks) return;
    const int lane = threadIdx.x & 31;
    const int lg = lane & (G - 1);
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (lane & ~(G - 1)));

    const int32_t row = p.chunk_row[gid];
    const int64_t grow = row + p.row_offset;
    const int64_t row_b = p.rowptr[row], row_e = p.rowptr[row + 1];
    const int64_t b = p.chunk_begin[gid];
    const int64_t e = p.chunk_len ? b + p.chunk_len[gid] : ((b + p.chunk_edges < row_e) ? b + p.chunk_edges : row_e);
    const int32_t slot = p.chunk_slot[gid];
    const bool first = (b == row_b);  // the first chunk of a row carries the self edge
    const float pi = __ldg(p.p + grow);

    // ONE pass over the chunk's edges (online softmax): the running maximum m is raised G edges at a time and the
    // accumulator, the lanes' weight sums and the self edge's weight are rescaled by exp(m_old - m_new) when it moves - a
    // group-uniform decision, 4 multiplies per lane and batch.  (Round 1 made a first pass over colidx and q[col] for
    // the maximum: one more index stream, one more 4-byte gather per edge and a serialised round trip before the row
    // gathers could start.)  The sequence of maxima depends on the chunk's edges only, so the result is the same under
    // any launch shape or row partition.
    float e_self = 0.f;
    if (first) e_self = leaky02(pi + __ldg(p.q + grow));
    float m = -INFINITY;

    float s_total = 0.f;
    for (int c0 = 0; c0 < p.h; c0 += G * VEC) {
        const int col = c0 + lg * VEC;
        const bool col_ok = col < p.h;
        const float *zcol = p.z + (col_ok ? col : 0);
        float acc[VEC];
#pragma unroll
        for (int t = 0; t < VEC; ++t) acc[t] = 0.f;
        float s_lane = 0.f;
        float w_self = first ? 1.f : 0.f;   // exp(e_self - m) with m = e_self
        m = first ? e_self : -INFINITY;
        if (first && col_ok) {
            if (VEC == 4) {
                const float4 zz = ldg4(zcol + grow * p.ldz);
                acc[0] = zz.x; acc[1 % VEC] = zz.y; acc[2 % VEC] = zz.z; acc[3 % VEC] = zz.w;
            } else {
                acc[0] = __ldg(zcol + grow * p.ldz);
            }
        }
        for (int64_t base = b; base < e; base += G) {
            const int64_t idx = base + lg;
            int c = 0;
            float ev = -INFINITY;
            if (idx < e) {
                c = ld_stream_i32(p.colidx + idx);
                if (c != grow) ev = leaky02(pi + __ldg(p.q + c));
            }
            const float bm = group_max<G>(ev, gmask);
            if (bm > m) {   // uniform over the group
                const float sc = (m == -INFINITY) ? 0.f : expf(m - bm);
#pragma unroll
                for (int t = 0; t < VEC; ++t) acc[t] *= sc;
                s_lane *= sc;
                w_self *= sc;
                m = bm;
            }
            const float w = (ev > -INFINITY) ? expf(ev - m) : 0.f;
            s_lane += w;
            const int cnt = (e - base < G) ? (int)(e - base) : G;
#pragma unroll
            for (int k0 = 0; k0 < G; k0 += U) {
                if (k0 >= cnt) break;
                int cc[U];
                float ww[U];
                float zr[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    cc[u] = __shfl_sync(gmask, c, k0 + u, G);
                    ww[u] = __shfl_sync(gmask, w, k0 + u, G);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (col_ok && k0 + u < cnt && ww[u] != 0.f) {
                        if (VEC == 4) {
                            const float4 zz = ldg4(zcol + (int64_t)cc[u] * p.ldz);
                            zr[u][0] = zz.x; zr[u][1 % VEC] = zz.y; zr[u][2 % VEC] = zz.z; zr[u][3 % VEC] = zz.w;
                        } else {
                            zr[u][0] = __ldg(zcol + (int64_t)cc[u] * p.ldz);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (col_ok && k0 + u < cnt && ww[u] != 0.f) {
#pragma unroll
                        for (int t = 0; t < VEC; ++t) acc[t] = fmaf(ww[u], zr[u][t], acc[t]);
                    }
            }
        }
        if (c0 == 0) s_total = group_sum<G>(s_lane, gmask) + w_self;
        if (!col_ok) continue;
        if (slot >= 0) {
#pragma unroll
            for (int t = 0; t < VEC; ++t) p.part_acc[(int64_t)slot * p.h + col + t] = acc[t];
        } else {
            const float inv = 1.f / (s_total + 1e-9f);
#pragma unroll
            for (int t = 0; t < VEC; ++t) {
                float o = acc[t] * inv;
                if (p.bias) o += __ldg(p.bias + col + t);
                if (p.relu) o = fmaxf(o, 0.f);
                p.y[(int64_t)row * p.ldy + col + t] = o;
                for (int r = 0; r < p.n_peer; ++r) p.y_peer[r][(int64_t)row * p.ldy + col + t] = o;
            }
        }
    }
    if (slot >= 0 && lg == 0) {
        p.part_ms[(int64_t)slot * 2 + 0] = m;
        p.part_ms[(int64_t)slot * 2 + 1] = s_total;
    }
}

// heavy rows: softmax-merge of the parked chunk partials, ascending chunk order
template <int G>
__global__ void __launch_bounds__(kGatThreads) gat_heavy_kernel(const GatParams p) {
    const int64_t gid = ((int64_t)blockIdx.x * kGatThreads + threadIdx.x) / G;
    if (gid >= p.n_heavy) return;
    const int lg = threadIdx.x & (G - 1);
    const int32_t row = p.heavy_row[gid];
    const int64_t s0 = p.heavy_slot_ptr[gid], s1 = p.heavy_slot_ptr[gid + 1];
    float mx = -INFINITY;
    for (int64_t s = s0; s < s1; ++s) mx = fmaxf(mx, p.part_ms[s * 2]);
    float tot = 0.f;
    for (int64_t s = s0; s < s1; ++s) tot += p.part_ms[s * 2 + 1] * expf(p.part_ms[s * 2] - mx);
    const float inv = 1.f / (tot + 1e-9f);
    for (int col = lg; col < p.h; col += G) {
        float acc = 0.f;
        for (int64_t s = s0; s < s1; ++s) acc = fmaf(p.part_acc[s * p.h + col], expf(p.part_ms[s * 2] - mx), acc);
        float o = acc * inv;
        if (p.bias) o += __ldg(p.bias + col);
        if (p.relu) o = fmaxf(o, 0.f);
        p.y[(int64_t)row * p.ldy + col] = o;
        for (int r = 0; r < p.n_peer; ++r) p.y_peer[r][(int64_t)row * p.ldy + col] = o;
    }
}

template <int G, int VEC>
static int launch(const GatParams &p, cudaStream_t s) {
    if (p.n_chunks > 0) {
        gat_chunk_kernel<G, VEC><<<(unsigned)cdiv(p.n_chunks * G, kGatThreads), kGatThreads, 0, s>>>(p);
        CBRS_CHECK_LAUNCH("gat_chunk");
    }
    if (p.n_heavy > 0) {
        gat_heavy_kernel<G><<<(unsigned)cdiv(p.n_heavy * G, kGatThreads), kGatThreads, 0, s>>>(p);
        CBRS_CHECK_LAUNCH("gat_heavy");
    }
    return CBRS_OK;
}

template <int VEC>
static int dispatch_g(int lanes_needed, const GatParams &p, cudaStream_t s) {
    if (lanes_needed <= 1) return launch<1, VEC>(p, s);
    if (lanes_needed <= 2) return launch<2, VEC>(p, s);
    if (lanes_needed <= 4) return launch<4, VEC>(p, s);
    if (lanes_needed <= 8) return launch<8, VEC>(p, s);
    if (lanes_needed <= 16) return launch<16, VEC>(p, s);
    return launch<32, VEC>(p, s);
}

}  // namespace cbrs

using namespace cbrs;

extern "C" size_t cbrs_gat_workspace_bytes(const cbrs_csr_t *g, int32_t h) {
    if (!g) return 0;
    return align_up((size_t)g->n_slots * (size_t)h * sizeof(float)) + align_up((size_t)g->n_slots * 2 * sizeof(float)) + 256;
}

static int gat_impl(const cbrs_csr_t *g, int64_t row_offset, const float *z, int64_t ldz, const float *pvec,
                    const float *qvec, float *y, int64_t ldy, int32_t h, const float *bias, int relu,
                    void *const *y_peers_host, int n_peers, void *workspace, size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(g && z && pvec && qvec && y, CBRS_E_INVALID, "gat: null argument");
    CBRS_REQUIRE(n_peers >= 0 && n_peers < CBRS_MAX_PEERS && (n_peers == 0 || y_peers_host), CBRS_E_INVALID,
                 "gat: n_peers=%d (at most %d peer copies)", n_peers, CBRS_MAX_PEERS - 1);
    CBRS_REQUIRE(h > 0 && ldz >= h && ldy >= h && row_offset >= 0, CBRS_E_INVALID, "gat: h=%d ldz=%lld ldy=%lld", h,
                 (long long)ldz, (long long)ldy);
    CBRS_REQUIRE(g->n_rows >= 0 && g->n_chunks >= 0 && g->chunk_edges > 0, CBRS_E_INVALID, "gat: bad graph descriptor");
    if (g->n_rows == 0) return CBRS_OK;
    CBRS_REQUIRE(g->rowptr && g->chunk_row && g->chunk_begin && g->chunk_slot && (g->nnz == 0 || g->colidx),
                 CBRS_E_INVALID, "gat: graph descriptor has null arrays");
    Arena a(workspace, workspace_bytes);
    GatParams p;
    p.part_acc = nullptr; p.part_ms = nullptr;
    if (g->n_slots > 0) {
        p.part_acc = a.take<float>((size_t)g->n_slots * h);
        p.part_ms = a.take<float>((size_t)g->n_slots * 2);
        CBRS_REQUIRE(workspace && p.part_acc && p.part_ms, CBRS_E_WORKSPACE, "gat: workspace too small");
    }
    p.rowptr = g->rowptr; p.colidx = g->colidx; p.chunk_row = g->chunk_row; p.chunk_begin = g->chunk_begin;
    p.chunk_slot = g->chunk_slot; p.chunk_len = g->chunk_len; p.n_chunks = g->n_chunks; p.chunk_edges = g->chunk_edges; p.row_offset = row_offset;
    p.z = z; p.ldz = ldz; p.p = pvec; p.q = qvec; p.y = y; p.ldy = ldy; p.h = h; p.bias = bias; p.relu = relu;
    p.heavy_row = g->heavy_row; p.heavy_slot_ptr = g->heavy_slot_ptr; p.n_heavy = g->n_heavy;
    p.n_peer = n_peers;
    for (int r = 0; r < CBRS_MAX_PEERS - 1; ++r) {
        p.y_peer[r] = r < n_peers ? (float *)y_peers_host[r] : nullptr;
        if (r < n_peers) CBRS_REQUIRE(p.y_peer[r], CBRS_E_INVALID, "gat: peer copy %d is null", r);
    }
    const bool vec4 = (h % 4 == 0) && (ldz % 4 == 0) && ((uintptr_t)z % 16 == 0);
    cudaStream_t s = (cudaStream_t)stream;
    return vec4 ? dispatch_g<4>(h / 4, p, s) : dispatch_g<1>(h, p, s);
}

extern "C" int cbrs_gat_csr(const cbrs_csr_t *g, int64_t row_offset, const float *z, int64_t ldz, const float *pvec,
                            const float *qvec, float *y, int64_t ldy, int32_t h, const float *bias, int relu,
                            void *workspace, size_t workspace_bytes, void *stream) {
    return gat_impl(g, row_offset, z, ldz, pvec, qvec, y, ldy, h, bias, relu, nullptr, 0, workspace, workspace_bytes,
                    stream);
}

extern "C" int cbrs_gat_csr_bcast(const cbrs_csr_t *g, int64_t row_offset, const float *z, int64_t ldz,
                                  const float *pvec, const float *qvec, float *y, int64_t ldy, int32_t h,
                                  const float *bias, int relu, void *const *y_peers_host, int n_peers, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    return gat_impl(g, row_offset, z, ldz, pvec, qvec, y, ldy, h, bias, relu, y_peers_host, n_peers, workspace,
                    workspace_bytes, stream);
}
