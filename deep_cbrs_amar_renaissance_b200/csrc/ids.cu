// Id compaction on device (scope row G0 / (f)-2): what the reference does on the host with
// np.unique(col, return_inverse=True) for the training ratings and an [n,1] == [U] broadcast +
// argwhere for the test ratings and property triples (/root/reference/src/data/loaders.py:47-54,64-66;
// the broadcast materialises n x U booleans, 1.1 GB at MovieLens-1M).
//   cbrs_compact_ids : sorted unique values + for every input its index among them (bit-exact ==
//                      np.unique(return_inverse=True)): stable radix sort of sign-flipped keys,
//                      head flags, scan
//   cbrs_lookup_ids  : index of every id in a sorted vocabulary by binary search, -1 when absent
#include "common.cuh"

namespace cbrs {

constexpr uint64_t kSignFlip = 0x8000000000000000ull;

__global__ void ids_keys_kernel(const int64_t *__restrict__ ids, int64_t n, uint64_t *__restrict__ keys,
                                uint32_t *__restrict__ payload) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = (uint64_t)ids[i] ^ kSignFlip;  // order-preserving map int64 -> uint64
    payload[i] = (uint32_t)i;
}

__global__ void ids_heads_kernel(const uint64_t *__restrict__ keys, int64_t n, uint32_t *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// flags hold the EXCLUSIVE scan of the head flags: rank of position i = scan[i] + head(i) - 1
__global__ void ids_emit_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ payload,
                                const uint32_t *__restrict__ scan, int64_t n, int64_t *__restrict__ uniques,
                                int64_t *__restrict__ inverse, int64_t *__restrict__ n_unique) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool head = (i == 0 || keys[i] != keys[i - 1]);
    const uint32_t rank = scan[i] + (head ? 1u : 0u) - 1u;
    if (head) uniques[rank] = (int64_t)(keys[i] ^ kSignFlip);
    inverse[payload[i]] = (int64_t)rank;
    if (i == n - 1 && n_unique) *n_unique = (int64_t)rank + 1;
}

__global__ void ids_lookup_kernel(const int64_t *__restrict__ vocab, int64_t n_vocab, const int64_t *__restrict__ ids,
                                  int64_t n, int64_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = ids[i];
    int64_t lo = 0, hi = n_vocab;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (vocab[mid] < v) lo = mid + 1; else hi = mid;
    }
    out[i] = (lo < n_vocab && vocab[lo] == v) ? lo : -1;
}

}  // namespace cbrs

using namespace cbrs;

extern "C" size_t cbrs_compact_ids_workspace_bytes(int64_t n) {
    return align_up((size_t)n * 8) + 2 * align_up((size_t)n * 4) + sort_workspace_bytes(n) + scan_u32_workspace_bytes(n) + 1024;
}

extern "C" int cbrs_compact_ids(const int64_t *ids, int64_t n, int64_t *uniques_out, int64_t *inverse_out,
                                int64_t *n_unique_out, void *workspace, size_t workspace_bytes, void *stream) {
    CBRS_REQUIRE(ids && uniques_out && inverse_out && n_unique_out, CBRS_E_INVALID, "compact_ids: null argument");
    CBRS_REQUIRE(n > 0 && n < (int64_t)0xffffffffll, CBRS_E_INVALID, "compact_ids: n=%lld", (long long)n);
    CBRS_REQUIRE(workspace && workspace_bytes >= cbrs_compact_ids_workspace_bytes(n), CBRS_E_WORKSPACE,
                 "compact_ids: workspace too small");
    Arena a(workspace, workspace_bytes);
    uint64_t *keys = a.take<uint64_t>((size_t)n);
    uint32_t *payload = a.take<uint32_t>((size_t)n);
    uint32_t *flags = a.take<uint32_t>((size_t)n);
    void *sort_ws = a.take<char>(sort_workspace_bytes(n));
    void *scan_ws = a.take<char>(scan_u32_workspace_bytes(n));
    CBRS_REQUIRE(keys && payload && flags && sort_ws && scan_ws, CBRS_E_WORKSPACE, "compact_ids: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)cdiv(n, 256);
    ids_keys_kernel<<<grid, 256, 0, s>>>(ids, n, keys, payload);
    CBRS_CHECK_LAUNCH("ids_keys");
    int rc = sort_pairs_u64(keys, payload, n, 64, sort_ws, sort_workspace_bytes(n), s);
    if (rc) return rc;
    ids_heads_kernel<<<grid, 256, 0, s>>>(keys, n, flags);
    CBRS_CHECK_LAUNCH("ids_heads");
    rc = scan_u32_exclusive(flags, n, nullptr, scan_ws, scan_u32_workspace_bytes(n), s);
    if (rc) return rc;
    ids_emit_kernel<<<grid, 256, 0, s>>>(keys, payload, flags, n, uniques_out, inverse_out, n_unique_out);
    CBRS_CHECK_LAUNCH("ids_emit");
    return CBRS_OK;
}

extern "C" int cbrs_lookup_ids(const int64_t *vocab_sorted, int64_t n_vocab, const int64_t *ids, int64_t n,
                               int64_t *index_out, void *stream) {
    CBRS_REQUIRE(vocab_sorted && ids && index_out && n_vocab > 0 && n >= 0, CBRS_E_INVALID, "lookup_ids: bad argument");
    if (n == 0) return CBRS_OK;
    ids_lookup_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(vocab_sorted, n_vocab, ids, n, index_out);
    CBRS_CHECK_LAUNCH("ids_lookup");
    return CBRS_OK;
}
