// Device scan and stable LSD radix sort, hand-written for sm_100a.
//
// These are the primitives under the graph build (row G2/G3: the device
// equivalent of scipy's COO->CSR + tf.sparse.reorder) and the pair-list top-k
// (row T: pandas' stable multi-key sort, /root/reference/src/utilities/metrics.py:27).
// HBM-bound integer work: coalesced tile loads, shared-memory ranking, one
// global histogram per pass.  Stability is what makes duplicate summation and
// tie order match the reference bit-for-bit.
#include "common.cuh"

namespace cbrs {

// ------------------------------------------------------------------ scan
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(T *data, int64_t n, T *tile_sums, T *total_out) {
    __shared__ T warp_sums[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T v[kScanItems];
    T sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? data[base + k] : (T)0;
        sum += v[k];
    }
    // inclusive scan of the per-thread sums inside the warp
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = (lane < kScanThreads / 32) ? warp_sums[lane] : (T)0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    T run = inc - sum + (warp ? warp_sums[warp - 1] : (T)0);  // exclusive prefix of this thread
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == kScanThreads - 1) {
        if (tile_sums) tile_sums[blockIdx.x] = run;
        if (total_out) *total_out = run;
    }
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(T *data, int64_t n, const T *tile_offsets) {
    const T off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) data[base + k] += off;
}

template <typename T>
static size_t scan_ws_bytes(int64_t n) {
    size_t total = 0;
    while (n > kScanTile) {
        n = cdiv(n, kScanTile);
        total += align_up((size_t)n * sizeof(T));
    }
    return total + 256;
}

template <typename T>
static int scan_rec(T *data, int64_t n, T *total_out, Arena &ws, cudaStream_t s) {
    if (n <= 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(T), s);
        return CBRS_OK;
    }
    const int64_t nb = cdiv(n, kScanTile);
    if (nb == 1) {
        scan_tile_kernel<T><<<1, kScanThreads, 0, s>>>(data, n, nullptr, total_out);
        CBRS_CHECK_LAUNCH("scan_tile");
        return CBRS_OK;
    }
    T *sums = ws.take<T>((size_t)nb);
    CBRS_REQUIRE(sums, CBRS_E_WORKSPACE, "scan: workspace too small");
    scan_tile_kernel<T><<<(unsigned)nb, kScanThreads, 0, s>>>(data, n, sums, nullptr);
    CBRS_CHECK_LAUNCH("scan_tile");
    int rc = scan_rec<T>(sums, nb, total_out, ws, s);
    if (rc) return rc;
    scan_add_kernel<T><<<(unsigned)nb, kScanThreads, 0, s>>>(data, n, sums);
    CBRS_CHECK_LAUNCH("scan_add");
    return CBRS_OK;
}

size_t scan_u32_workspace_bytes(int64_t n) { return scan_ws_bytes<uint32_t>(n); }
size_t scan_i64_workspace_bytes(int64_t n) { return scan_ws_bytes<int64_t>(n); }

int scan_u32_exclusive(uint32_t *data, int64_t n, uint32_t *total_out, void *ws, size_t ws_bytes, cudaStream_t s) {
    Arena a(ws, ws_bytes);
    return scan_rec<uint32_t>(data, n, total_out, a, s);
}
int scan_i64_exclusive(int64_t *data, int64_t n, int64_t *total_out, void *ws, size_t ws_bytes, cudaStream_t s) {
    Arena a(ws, ws_bytes);
    return scan_rec<int64_t>(data, n, total_out, a, s);
}

// ------------------------------------------------------------------ radix sort
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortItems = 16;                          // keys per thread
constexpr int kSortTile = kSortThreads * kSortItems;    // 4096 keys per CTA
constexpr int kSortSeg = 32 * kSortItems;               // contiguous keys owned by one warp
constexpr int kRadix = 256;

// tile histogram of one 8-bit digit -> hist[digit][tile]
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t *__restrict__ keys, int64_t n, int shift,
                                                                 uint32_t *__restrict__ hist, int64_t n_tiles) {
    __shared__ uint32_t h[kRadix];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int k = 0; k < kSortItems; ++k) {
        int64_t i = base + (int64_t)k * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&h[(unsigned)(keys[i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter.  Order inside a tile = (warp segment, iteration, lane), i.e. input order.
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t *__restrict__ kin, const uint32_t *__restrict__ pin,
                                                                    uint64_t *__restrict__ kout, uint32_t *__restrict__ pout,
                                                                    int64_t n, int shift, const uint32_t *__restrict__ offsets,
                                                                    int64_t n_tiles) {
    __shared__ uint32_t wcount[kSortWarps][kRadix];
    for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&wcount[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const int64_t seg = (int64_t)blockIdx.x * kSortTile + (int64_t)warp * kSortSeg;
    uint64_t key[kSortItems];
    uint16_t rank[kSortItems];
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int64_t i = seg + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? kin[i] : 0ull;
        const unsigned d = valid ? ((unsigned)(key[k] >> shift) & 0xffu) : 0xffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned r = __popc(peers & lt);
        unsigned before = 0;
        if (valid) before = wcount[warp][d];
        __syncwarp();
        if (valid && r == 0) wcount[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[k] = (uint16_t)(before + r);
    }
    __syncthreads();
    {   // counts -> global bases: tile offset of the digit + counts of the lower warps
        const int d = threadIdx.x;
        uint32_t run = offsets[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            uint32_t c = wcount[w][d];
            wcount[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        const int64_t i = seg + k * 32 + lane;
        if (i < n) {
            const unsigned d = (unsigned)(key[k] >> shift) & 0xffu;
            const uint32_t pos = wcount[warp][d] + rank[k];
            kout[pos] = key[k];
            pout[pos] = pin[i];
        }
    }
}

size_t sort_workspace_bytes(int64_t n) {
    const int64_t n_tiles = cdiv(n > 0 ? n : 1, kSortTile);
    const int64_t n_hist = n_tiles * kRadix;
    return align_up((size_t)n * 8) + align_up((size_t)n * 4) + align_up((size_t)n_hist * 4) +
           scan_u32_workspace_bytes(n_hist) + 1024;
}

int sort_pairs_u64(uint64_t *keys, uint32_t *payload, int64_t n, int key_bits, void *ws, size_t ws_bytes,
                   cudaStream_t s) {
    CBRS_REQUIRE(n >= 0 && n < (int64_t)0xffffffffll, CBRS_E_INVALID, "sort: n=%lld out of range", (long long)n);
    CBRS_REQUIRE(key_bits >= 0 && key_bits <= 64, CBRS_E_INVALID, "sort: key_bits=%d", key_bits);
    if (n <= 1 || key_bits == 0) return CBRS_OK;
    Arena a(ws, ws_bytes);
    const int64_t n_tiles = cdiv(n, kSortTile);
    const int64_t n_hist = n_tiles * kRadix;
    uint64_t *kalt = a.take<uint64_t>((size_t)n);
    uint32_t *palt = a.take<uint32_t>((size_t)n);
    uint32_t *hist = a.take<uint32_t>((size_t)n_hist);
    CBRS_REQUIRE(kalt && palt && hist, CBRS_E_WORKSPACE, "sort: workspace too small (%zu bytes)", ws_bytes);
    const size_t scan_off = a.off;
    uint64_t *kin = keys, *kout = kalt;
    uint32_t *pin = payload, *pout = palt;
    for (int shift = 0; shift < key_bits; shift += 8) {
        radix_hist_kernel<<<(unsigned)n_tiles, kSortThreads, 0, s>>>(kin, n, shift, hist, n_tiles);
        CBRS_CHECK_LAUNCH("radix_hist");
        a.off = scan_off;  // the scan scratch is reused every pass
        Arena sa((char *)ws + a.off, ws_bytes - a.off);
        int rc = scan_rec<uint32_t>(hist, n_hist, nullptr, sa, s);
        if (rc) return rc;
        radix_scatter_kernel<<<(unsigned)n_tiles, kSortThreads, 0, s>>>(kin, pin, kout, pout, n, shift, hist, n_tiles);
        CBRS_CHECK_LAUNCH("radix_scatter");
        uint64_t *tk = kin; kin = kout; kout = tk;
        uint32_t *tp = pin; pin = pout; pout = tp;
    }
    if (kin != keys) {
        cudaMemcpyAsync(keys, kin, (size_t)n * 8, cudaMemcpyDeviceToDevice, s);
        cudaMemcpyAsync(payload, pin, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
    }
    return CBRS_OK;
}

}  // namespace cbrs

extern "C" size_t cbrs_sort_workspace_bytes(int64_t n) { return cbrs::sort_workspace_bytes(n); }
extern "C" int cbrs_sort_pairs_u64(uint64_t *keys, uint32_t *payload, int64_t n, int key_bits, void *workspace,
                                   size_t workspace_bytes, void *stream) {
    return cbrs::sort_pairs_u64(keys, payload, n, key_bits, workspace, workspace_bytes, (cudaStream_t)stream);
}
