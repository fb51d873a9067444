// prep.cu -- sparse-matrix preparation on the device: what the reference does per
// epoch or per fit on the host -- `X.T.tocsr()` twice per epoch (cymf/wmf.pyx:112), the per-user positive sets
// (cymf/bpr.pyx:146-147: the sorted CSR row IS the set here) -- plus the nnz-balanced row deal of the sharded ALS.
// At C5 (1e9 nonzeros) host-side scipy transposes take minutes; these kernels take milliseconds and are exact and
// DETERMINISTIC (no atomics decide an order): everything is built from
//     exclusive_scan       three-phase scan (tile sums -> one-CTA scan of the sums -> tile scan with base)
//     radix_sort_pairs     stable LSD radix sort of (u32 key, u32 value), 8 bits per pass: tile histograms in
//                          digit-major order, scan, stable in-tile ranking with match.any + per-warp counters
// All are streaming passes over the nonzeros: HBM-bound, 16 B per element and pass.
#include "common.cuh"

namespace cymf {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TI> __device__ __forceinline__ uint64_t scan_load(const TI *p, int64_t idx, int64_t n) {
    return idx < n ? (uint64_t)p[idx] : 0ull;
}

// phase 1: sums[tile] = sum of the tile's elements
template <typename TI>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const TI *__restrict__ in, int64_t n,
                                                                      uint64_t *__restrict__ sums) {
    __shared__ uint64_t warp_sum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) s += scan_load(in, base + t * SCAN_THREADS + threadIdx.x, n);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += warp_sum[w];
        sums[blockIdx.x] = t;
    }
}

// block-wide exclusive scan of one value per thread (1024 threads max); returns the exclusive prefix, *total = sum
__device__ __forceinline__ uint64_t block_exclusive(uint64_t v, uint64_t *smem, uint64_t *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    uint64_t incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < nwarp ? smem[lane] : 0ull, wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, wi, off);
            if (lane >= off) wi += t;
        }
        smem[lane] = wi - w;                      // exclusive prefix of the warp totals
        if (lane == 31) smem[32] = wi;            // grand total
    }
    __syncthreads();
    const uint64_t out = smem[warp] + incl - v;
    *total = smem[32];
    __syncthreads();
    return out;
}

// phase 2: in-place exclusive scan of the tile sums by ONE CTA walking them in chunks with a running carry
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint64_t *__restrict__ sums, int64_t m) {
    __shared__ uint64_t smem[33];
    uint64_t carry = 0;
    for (int64_t base = 0; base < m; base += blockDim.x) {
        const int64_t idx = base + threadIdx.x;
        const uint64_t v = idx < m ? sums[idx] : 0ull;
        uint64_t total;
        const uint64_t ex = block_exclusive(v, smem, &total);
        if (idx < m) sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) sums[m] = carry;        // grand total after the last tile
}

// phase 3: out[i] = base[tile] + exclusive prefix inside the tile; thread t owns SCAN_ITEMS consecutive elements
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(const TI *in, int64_t n,     // in may alias out
                                                                  const uint64_t *__restrict__ sums, TO *out,
                                                                  int write_total) {
    __shared__ uint64_t smem[33];
    const int64_t first = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint64_t v[SCAN_ITEMS], mine = 0;
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) { v[t] = scan_load(in, first + t, n); mine += v[t]; }
    uint64_t total;
    uint64_t run = sums[blockIdx.x] + block_exclusive(mine, smem, &total);
#pragma unroll
    for (int t = 0; t < SCAN_ITEMS; ++t) {
        if (first + t < n) out[first + t] = (TO)run;
        run += v[t];
    }
    if (write_total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = (TO)sums[gridDim.x];
}

static inline int64_t scan_tiles(int64_t n) { return n > 0 ? (n + SCAN_TILE - 1) / SCAN_TILE : 1; }

template <typename TI, typename TO>
static int exclusive_scan(const TI *in, TO *out, int64_t n, int write_total, uint64_t *sums, cudaStream_t st) {
    const int64_t tiles = scan_tiles(n);
    scan_tile_sums_kernel<TI><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums);
    CYMF_LAUNCHED();
    scan_sums_kernel<<<1, 1024, 0, st>>>(sums, tiles);
    CYMF_LAUNCHED();
    scan_tiles_kernel<TI, TO><<<(unsigned)tiles, SCAN_THREADS, 0, st>>>(in, n, sums, out, write_total);
    CYMF_LAUNCHED();
    return 0;
}

// ---- stable LSD radix sort of (key, value) pairs, 8 bits per pass --------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_ROUNDS = 16;                              // elements per lane and tile
constexpr int SORT_TILE = SORT_THREADS * SORT_ROUNDS;        // 4096
constexpr int SORT_WARP_SPAN = 32 * SORT_ROUNDS;             // consecutive elements owned by one warp

// hist[d * tiles + tile] = number of keys of the tile whose digit is d
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const uint32_t *__restrict__ keys, int64_t n, int shift,
                                                                  uint32_t *__restrict__ hist, int64_t tiles) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
#pragma unroll
    for (int t = 0; t < SORT_ROUNDS; ++t) {
        const int64_t idx = base + t * SORT_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);        // counts only: order-free
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

// One tile: stable ranks inside each warp round from eight ballots (one per digit bit; cheaper than match.any),
// per-warp digit counters, exclusive prefixes over warps and over digits, then the tile is written to shared memory in
// digit-sorted order and leaves as contiguous runs -- every digit's elements of the tile are consecutive in the
// output, so consecutive threads store consecutive addresses.
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const uint32_t *__restrict__ keys,
                                                                     const uint32_t *__restrict__ vals, int64_t n,
                                                                     int shift, const uint32_t *__restrict__ offsets,
                                                                     int64_t tiles, uint32_t *__restrict__ out_keys,
                                                                     uint32_t *__restrict__ out_vals) {
    __shared__ uint32_t cnt[SORT_WARPS][256];
    __shared__ uint32_t gbase[256], dstart[256];
    __shared__ uint32_t skey[SORT_TILE], sval[SORT_TILE];
    __shared__ uint64_t scan_tmp[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t = threadIdx.x; t < SORT_WARPS * 256; t += SORT_THREADS) (&cnt[0][0])[t] = 0;
    gbase[threadIdx.x] = offsets[(int64_t)threadIdx.x * tiles + blockIdx.x];
    __syncthreads();
    const int64_t tile_base = (int64_t)blockIdx.x * SORT_TILE;
    const int64_t wbase = tile_base + (int64_t)warp * SORT_WARP_SPAN;
    uint32_t k[SORT_ROUNDS], v[SORT_ROUNDS];
    uint16_t off[SORT_ROUNDS];
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        const bool valid = idx < n;
        k[r] = valid ? keys[idx] : 0u;
        v[r] = (valid && vals) ? vals[idx] : 0u;
        const uint32_t d = (k[r] >> shift) & 255u;
        unsigned peers = __ballot_sync(0xffffffffu, valid);                  // lanes holding the same digit as mine
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? bal : ~bal;
        }
        uint32_t before = 0;
        if (valid) {
            const int leader = __ffs(peers) - 1;
            if (lane == leader) { before = cnt[warp][d]; cnt[warp][d] = before + __popc(peers); }
            before = __shfl_sync(peers, before, leader);
        }
        off[r] = (uint16_t)(before + __popc(peers & lt));                    // rank inside the warp's span, stable
        __syncwarp();
    }
    __syncthreads();
    uint32_t total = 0;
    {   // digit = threadIdx.x: exclusive prefix over the warps, then over the digits
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) { const uint32_t t = cnt[w][threadIdx.x]; cnt[w][threadIdx.x] = total; total += t; }
    }
    uint64_t tile_total;
    dstart[threadIdx.x] = (uint32_t)block_exclusive((uint64_t)total, scan_tmp, &tile_total);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t idx = wbase + r * 32 + lane;
        if (idx < n) {
            const uint32_t d = (k[r] >> shift) & 255u;
            const uint32_t lp = dstart[d] + cnt[warp][d] + off[r];
            skey[lp] = k[r];
            sval[lp] = v[r];
        }
    }
    __syncthreads();
    const int n_tile = (int)(n - tile_base < SORT_TILE ? n - tile_base : SORT_TILE);
    for (int s = threadIdx.x; s < n_tile; s += SORT_THREADS) {
        const uint32_t key = skey[s];
        const uint32_t d = (key >> shift) & 255u;
        const int64_t pos = (int64_t)gbase[d] + (s - dstart[d]);
        out_keys[pos] = key;
        if (out_vals) out_vals[pos] = sval[s];
    }
}

static inline int64_t sort_tiles(int64_t n) { return n > 0 ? (n + SORT_TILE - 1) / SORT_TILE : 1; }
static inline size_t align256(size_t b) { return (b + 255) / 256 * 256; }

// workspace layout: [alt keys n u32][alt vals n u32][hist 256*tiles u32][scan sums]
static size_t sort_workspace_bytes(int64_t n) {
    const int64_t tiles = sort_tiles(n);
    return 2 * align256((size_t)n * 4) + align256((size_t)256 * tiles * 4) +
           align256((size_t)(scan_tiles(256 * tiles) + 1) * 8);
}

// Sorts in place (keys, vals hold the result).  `bits` low bits of the key are significant.
static int radix_sort_pairs(uint32_t *keys, uint32_t *vals, int64_t n, int bits, void *workspace, cudaStream_t st) {
    if (n <= 1) return 0;
    const int64_t tiles = sort_tiles(n);
    char *ws = (char *)workspace;
    uint32_t *alt_k = (uint32_t *)ws; ws += align256((size_t)n * 4);
    uint32_t *alt_v = (uint32_t *)ws; ws += align256((size_t)n * 4);
    uint32_t *hist = (uint32_t *)ws; ws += align256((size_t)256 * tiles * 4);
    uint64_t *sums = (uint64_t *)ws;
    uint32_t *src_k = keys, *src_v = vals, *dst_k = alt_k, *dst_v = alt_v;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    for (int p = 0; p < passes; ++p) {
        sort_hist_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(src_k, n, 8 * p, hist, tiles);
        CYMF_LAUNCHED();
        CYMF_TRY((exclusive_scan<uint32_t, uint32_t>(hist, hist, 256 * tiles, 0, sums, st)));
        sort_scatter_kernel<<<(unsigned)tiles, SORT_THREADS, 0, st>>>(src_k, src_v, n, 8 * p, hist, tiles, dst_k, dst_v);
        CYMF_LAUNCHED();
        uint32_t *t = src_k; src_k = dst_k; dst_k = t;
        t = src_v; src_v = dst_v; dst_v = t;
    }
    if (src_k != keys) {
        CYMF_CUDA(cudaMemcpyAsync(keys, src_k, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        if (vals) CYMF_CUDA(cudaMemcpyAsync(vals, src_v, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    return 0;
}

// ---- CSR helpers --------------------------------------------------------------------------------------------------
// row_of[e] = r for every stored entry e of row r (one warp per row)
__global__ void expand_rows_kernel(const int64_t *__restrict__ indptr, int64_t rows, uint32_t *__restrict__ row_of) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
        const int64_t lo = indptr[r], hi = indptr[r + 1];
        for (int64_t e = lo + lane; e < hi; e += 32) row_of[e] = (uint32_t)r;
    }
}
__global__ void count_keys_kernel(const int32_t *__restrict__ keys, int64_t n, uint32_t *__restrict__ counts) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(counts + keys[t], 1u);                                     // integer counts: order-free
}
__global__ void row_degrees_kernel(const int64_t *__restrict__ indptr, int64_t rows, uint32_t *__restrict__ inv_deg,
                                   uint32_t *__restrict__ ids) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        inv_deg[r] = ~(uint32_t)(indptr[r + 1] - indptr[r]);                 // ascending ~deg == descending deg
        ids[r] = (uint32_t)r;
    }
}
// heaviest-first round-robin deal: the q-th heaviest row goes to rank q % world, position q / world
__global__ void deal_slots_kernel(const uint32_t *__restrict__ by_weight, int64_t rows, int32_t world, int64_t per_rank,
                                  int64_t *__restrict__ slot_row, int64_t *__restrict__ row_slot) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < rows; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t slot = (q % world) * per_rank + q / world;
        const int64_t row = by_weight[q];
        slot_row[slot] = row;
        row_slot[row] = slot;
    }
}
__global__ void fill_i64_kernel(int64_t *dst, int64_t n, int64_t value) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = value;
}
__global__ void block_lengths_kernel(const int64_t *__restrict__ indptr, const int64_t *__restrict__ slot_row,
                                     int64_t count, uint32_t *__restrict__ len) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < count; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = slot_row[q];
        len[q] = row < 0 ? 0u : (uint32_t)(indptr[row + 1] - indptr[row]);
    }
}
// block row q = source row slot_row[q] with its columns renamed by col_slot (one warp per row)
__global__ void block_copy_kernel(const int64_t *__restrict__ indptr, const int32_t *__restrict__ indices,
                                  const int64_t *__restrict__ slot_row, int64_t count,
                                  const int64_t *__restrict__ col_slot, const int64_t *__restrict__ blk_ptr,
                                  int32_t *__restrict__ blk_idx) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < count; q += warps) {
        const int64_t row = slot_row[q];
        if (row < 0) continue;
        const int64_t lo = indptr[row], n = indptr[row + 1] - lo, dst = blk_ptr[q];
        for (int64_t e = lane; e < n; e += 32) {
            const int32_t c = indices[lo + e];
            blk_idx[dst + e] = col_slot ? (int32_t)col_slot[c] : c;
        }
    }
}

static inline int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

}  // namespace cymf

using namespace cymf;

extern "C" int64_t cymf_scan_workspace_bytes(int64_t n) { return (int64_t)align256((size_t)(scan_tiles(n) + 1) * 8); }

extern "C" int cymf_exclusive_scan_u32_dev(const uint32_t *in, int64_t *out, int64_t n, void *workspace, void *stream) {
    CYMF_REQUIRE((in || n == 0) && out && workspace && n >= 0, "bad argument");
    return exclusive_scan<uint32_t, int64_t>(in, out, n, 1, (uint64_t *)workspace, (cudaStream_t)stream);
}

extern "C" int64_t cymf_sort_workspace_bytes(int64_t n) { return (int64_t)sort_workspace_bytes(n); }

extern "C" int cymf_sort_pairs_dev(uint32_t *keys, uint32_t *values, int64_t n, int32_t key_bits, void *workspace,
                                   void *stream) {
    CYMF_REQUIRE((keys || n == 0) && workspace && n >= 0 && key_bits >= 1 && key_bits <= 32, "bad argument");
    return radix_sort_pairs(keys, values, n, key_bits, workspace, (cudaStream_t)stream);
}

extern "C" int64_t cymf_csr_transpose_workspace_bytes(int64_t rows, int64_t cols, int64_t nnz) {
    (void)rows;
    return (int64_t)(align256((size_t)nnz * 4) + align256((size_t)(cols + 1) * 4) + sort_workspace_bytes(nnz) +
                     (size_t)cymf_scan_workspace_bytes(cols));
}

// CSR of X^T from the CSR of X.  Output rows are sorted by original row id (stable sort of the row-major entries
// by column), i.e. exactly scipy's X.T.tocsr() with sorted indices (cymf/wmf.pyx:112).
extern "C" int cymf_csr_transpose_dev(const int64_t *indptr, const int32_t *indices, int64_t rows, int64_t cols,
                                      int64_t nnz, int64_t *t_indptr, int32_t *t_indices, void *workspace,
                                      void *stream) {
    CYMF_REQUIRE(indptr && t_indptr && workspace && rows >= 0 && cols >= 0 && nnz >= 0, "bad argument");
    CYMF_REQUIRE(rows < ((int64_t)1 << 32) && cols < ((int64_t)1 << 31), "shape exceeds 32-bit row / column ids");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    uint32_t *keys = (uint32_t *)ws; ws += align256((size_t)nnz * 4);
    uint32_t *counts = (uint32_t *)ws; ws += align256((size_t)(cols + 1) * 4);
    void *sort_ws = ws; ws += sort_workspace_bytes(nnz);
    void *scan_ws = ws;
    CYMF_CUDA(cudaMemsetAsync(counts, 0, (size_t)(cols + 1) * 4, st));
    if (nnz > 0) {
        CYMF_REQUIRE(indices && t_indices, "null pointer");
        count_keys_kernel<<<flat_grid(nnz), 256, 0, st>>>(indices, nnz, counts);
        CYMF_LAUNCHED();
    }
    CYMF_TRY((exclusive_scan<uint32_t, int64_t>(counts, t_indptr, cols, 1, (uint64_t *)scan_ws, st)));
    if (nnz > 0) {
        CYMF_CUDA(cudaMemcpyAsync(keys, indices, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, st));
        expand_rows_kernel<<<flat_grid(rows * 32), 256, 0, st>>>(indptr, rows, (uint32_t *)t_indices);
        CYMF_LAUNCHED();
        CYMF_TRY(radix_sort_pairs(keys, (uint32_t *)t_indices, nnz, bits_for(cols), sort_ws, st));
    }
    return 0;
}

extern "C" int64_t cymf_deal_workspace_bytes(int64_t rows) {
    return (int64_t)(2 * align256((size_t)rows * 4) + sort_workspace_bytes(rows));
}

// Heaviest-first round-robin deal of the rows of a CSR over `world` ranks (cymf_b200's sharded ALS): rows sorted by
// decreasing degree (stable: ties keep ascending row id); the q-th goes to rank q % world, position q / world.
// slot_row[world * per_rank] (slot -> row id, -1 = phantom), row_slot[rows] (row id -> slot).
extern "C" int cymf_deal_rows_dev(const int64_t *indptr, int64_t rows, int32_t world, int64_t per_rank,
                                  int64_t *slot_row, int64_t *row_slot, void *workspace, void *stream) {
    CYMF_REQUIRE(indptr && slot_row && row_slot && workspace, "null pointer");
    CYMF_REQUIRE(rows >= 0 && world >= 1 && per_rank * world >= rows && rows < ((int64_t)1 << 32), "bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    uint32_t *keys = (uint32_t *)ws; ws += align256((size_t)rows * 4);
    uint32_t *ids = (uint32_t *)ws; ws += align256((size_t)rows * 4);
    fill_i64_kernel<<<flat_grid(per_rank * world), 256, 0, st>>>(slot_row, per_rank * world, -1);
    CYMF_LAUNCHED();
    if (rows == 0) return 0;
    row_degrees_kernel<<<flat_grid(rows), 256, 0, st>>>(indptr, rows, keys, ids);
    CYMF_LAUNCHED();
    CYMF_TRY(radix_sort_pairs(keys, ids, rows, 32, ws, st));
    deal_slots_kernel<<<flat_grid(rows), 256, 0, st>>>(ids, rows, world, per_rank, slot_row, row_slot);
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int64_t cymf_csr_block_workspace_bytes(int64_t count) {
    return (int64_t)(align256((size_t)(count + 1) * 4) + (size_t)cymf_scan_workspace_bytes(count));
}

// Row block of a CSR in dealt order: block row q = source row slot_row[q] (empty when -1), columns renamed through
// col_slot (NULL = keep).  Two calls: blk_indices == NULL computes blk_indptr[count + 1] only (its last entry is the
// block's nnz, which sizes blk_indices); the second call copies the entries.
extern "C" int cymf_csr_block_dev(const int64_t *indptr, const int32_t *indices, const int64_t *slot_row, int64_t count,
                                  const int64_t *col_slot, int64_t *blk_indptr, int32_t *blk_indices, void *workspace,
                                  void *stream) {
    CYMF_REQUIRE(indptr && slot_row && blk_indptr && workspace && count >= 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (!blk_indices) {
        char *ws = (char *)workspace;
        uint32_t *len = (uint32_t *)ws; ws += align256((size_t)(count + 1) * 4);
        block_lengths_kernel<<<flat_grid(count), 256, 0, st>>>(indptr, slot_row, count, len);
        CYMF_LAUNCHED();
        return exclusive_scan<uint32_t, int64_t>(len, blk_indptr, count, 1, (uint64_t *)ws, st);
    }
    CYMF_REQUIRE(indices, "null pointer");
    if (count == 0) return 0;
    block_copy_kernel<<<flat_grid(count * 32), 256, 0, st>>>(indptr, indices, slot_row, count, col_slot, blk_indptr,
                                                           blk_indices);
    CYMF_LAUNCHED();
    return 0;
}

// ---- co-occurrence counting: the loop of read_text (cymf/glove.pyx:218-221) ----------------------------------------
//     for every kept token j of a line and every earlier kept token k of the same line with j - k <= window:
//         M[x_j, x_k] += 1.0 / (j - k)                        (an unordered_map<long, double> in the reference)
// Sort-based and exact: pair slot p = j * window + (window - d), d = j - k, enumerates the reference's updates in ITS
// order (j ascending, k ascending); two stable radix sorts (by column, then by row) group the slots of a cell while
// keeping that order, and each cell is summed sequentially in f64 -- the same additions in the same order, so the
// counts are bit-identical to the reference's.  Output triplets are sorted by (row, col).
namespace cymf {

struct CoocView {
    const int32_t *tokens, *pos;
    int64_t n_pairs;
    int32_t window, vocab;
    __device__ __forceinline__ bool decode(uint32_t p, int32_t *row, int32_t *col, int32_t *dist) const {
        const int64_t j = p / (uint32_t)window;
        const int32_t d = window - (int32_t)(p % (uint32_t)window);
        *dist = d;
        if (d > pos[j]) return false;                       // would reach before the start of the line
        *row = tokens[j];
        *col = tokens[j - d];
        return true;
    }
};

__global__ void cooc_col_keys_kernel(const CoocView v, uint32_t *__restrict__ keys, uint32_t *__restrict__ ids) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < v.n_pairs; p += (int64_t)gridDim.x * blockDim.x) {
        int32_t r, c, d;
        keys[p] = v.decode((uint32_t)p, &r, &c, &d) ? (uint32_t)c : (uint32_t)v.vocab;   // invalid slots sort last
        ids[p] = (uint32_t)p;
    }
}
__global__ void cooc_row_keys_kernel(const CoocView v, const uint32_t *__restrict__ ids, uint32_t *__restrict__ keys) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < v.n_pairs; q += (int64_t)gridDim.x * blockDim.x) {
        int32_t r, c, d;
        keys[q] = v.decode(ids[q], &r, &c, &d) ? (uint32_t)r : (uint32_t)v.vocab;
    }
}
// head[q] = 1 when sorted slot q starts a new (row, col) cell; *n_valid = number of valid slots (they sort first)
__global__ void cooc_heads_kernel(const CoocView v, const uint32_t *__restrict__ ids, uint32_t *__restrict__ head,
                                  int64_t *__restrict__ n_valid) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < v.n_pairs; q += (int64_t)gridDim.x * blockDim.x) {
        int32_t r, c, d, r0, c0, d0;
        uint32_t h = 0;
        if (v.decode(ids[q], &r, &c, &d)) {
            h = (q == 0 || !v.decode(ids[q - 1], &r0, &c0, &d0) || r0 != r || c0 != c) ? 1u : 0u;
            if (q + 1 == v.n_pairs || !v.decode(ids[q + 1], &r0, &c0, &d0)) *n_valid = q + 1;   // exactly one such q
        }
        head[q] = h;
    }
}
// start[c] = first sorted slot of cell c (cells are numbered by the scan of the head flags); start[nnz] = n_valid
__global__ void cooc_cell_starts_kernel(const uint32_t *__restrict__ head, const int64_t *__restrict__ cell_of,
                                        int64_t n_pairs, const int64_t *__restrict__ n_valid, uint32_t *__restrict__ start) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n_pairs; q += (int64_t)gridDim.x * blockDim.x)
        if (head[q]) start[cell_of[q]] = (uint32_t)q;
    if (blockIdx.x == 0 && threadIdx.x == 0) start[cell_of[n_pairs]] = (uint32_t)*n_valid;
}
// A cell's slots are contiguous and in corpus order; 1.0 / distance is added one by one (glove.pyx:221) -- f64 addition
// is not associative, so the order IS the result.  One thread per cell for the short ones (3.5 slots on average);
// cells of >= 64 slots go to a work list for cooc_sum_long_kernel.
constexpr int COOC_LONG = 64;
__global__ void cooc_sum_short_kernel(const CoocView v, const uint32_t *__restrict__ ids,
                                      const uint32_t *__restrict__ start, const int64_t *__restrict__ nnz_ptr,
                                      int64_t capacity, int32_t *__restrict__ rows, int32_t *__restrict__ cols,
                                      double *__restrict__ vals, uint32_t *__restrict__ work,
                                      unsigned long long *__restrict__ n_work) {
    extern __shared__ double recip[];                       // recip[s] = 1.0 / (window - s), s = slot % window
    for (int s = threadIdx.x; s < v.window; s += blockDim.x) recip[s] = __ddiv_rn(1.0, (double)(v.window - s));
    __syncthreads();
    const int64_t cells = *nnz_ptr < capacity ? *nnz_ptr : capacity;
    const uint32_t window = (uint32_t)v.window;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t lo = start[c], hi = start[c + 1];
        int32_t r, cc, d;
        v.decode(ids[lo], &r, &cc, &d);
        rows[c] = r;
        cols[c] = cc;
        if (hi - lo >= (uint32_t)COOC_LONG) { work[atomicAdd(n_work, 1ull)] = (uint32_t)c; continue; }
        double sum = 0.0;
        for (uint32_t t = lo; t < hi; ++t) sum = __dadd_rn(sum, recip[ids[t] % window]);
        vals[c] = sum;
    }
}
// One warp per long cell (the most frequent pair of a Zipf corpus holds ~1 % of all updates): 32 addends are fetched
// with one coalesced load and handed to the running sum in slot order through shuffles, so the serial chain carries
// only the additions themselves.  Slots past the end contribute +0.0, which leaves any sum unchanged.
__global__ void __launch_bounds__(128) cooc_sum_long_kernel(const CoocView v, const uint32_t *__restrict__ ids,
                                                            const uint32_t *__restrict__ start,
                                                            const uint32_t *__restrict__ work,
                                                            const unsigned long long *__restrict__ n_work,
                                                            double *__restrict__ vals) {
    extern __shared__ double recip[];
    for (int s = threadIdx.x; s < v.window; s += blockDim.x) recip[s] = __ddiv_rn(1.0, (double)(v.window - s));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t window = (uint32_t)v.window;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n = (int64_t)*n_work;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += warps) {
        const uint32_t c = work[w];
        const uint32_t lo = start[c], hi = start[c + 1];
        double sum = 0.0;
        uint32_t t = lo + lane;
        double next = t < hi ? recip[ids[t] % window] : 0.0;
        for (uint32_t base = lo; base < hi; base += 32) {
            const double a = next;
            t = base + 32 + lane;
            next = (base + 32 < hi && t < hi) ? recip[ids[t] % window] : 0.0;        // next chunk already in flight
#pragma unroll
            for (int j = 0; j < 32; ++j) sum = __dadd_rn(sum, __shfl_sync(0xffffffffu, a, j));
        }
        if (lane == 0) vals[c] = sum;
    }
}

}  // namespace cymf

extern "C" int64_t cymf_cooc_workspace_bytes(int64_t n_tokens, int32_t window) {
    const int64_t P = n_tokens * (int64_t)window;
    return (int64_t)(3 * align256((size_t)(P + 1) * 4) + align256((size_t)(P + 1) * 8) + sort_workspace_bytes(P) +
                     (size_t)cymf_scan_workspace_bytes(P) + 256);
}

// tokens[n_tokens]: kept-word ids of the corpus, lines concatenated; pos_in_line[n_tokens]: index of the token within
// its line.  rows / cols / vals (capacity entries, <= n_tokens * window needed) receive the cells sorted by
// (row, col); *nnz_out (DEVICE int64) the number of cells (entries beyond `capacity` are dropped, nnz_out still counts).
extern "C" int cymf_cooc_count_dev(const int32_t *tokens, const int32_t *pos_in_line, int64_t n_tokens, int32_t vocab,
                                   int32_t window, int32_t *rows, int32_t *cols, double *vals, int64_t capacity,
                                   int64_t *nnz_out, void *workspace, void *stream) {
    CYMF_REQUIRE(nnz_out && workspace && n_tokens >= 0 && vocab > 0 && window > 0 && capacity >= 0, "bad argument");
    const int64_t P = n_tokens * (int64_t)window;
    CYMF_REQUIRE(P < ((int64_t)1 << 32), "n_tokens * window_size must stay below 2^32");
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 0) { CYMF_CUDA(cudaMemsetAsync(nnz_out, 0, 8, st)); return 0; }
    CYMF_REQUIRE(tokens && pos_in_line && rows && cols && vals, "null pointer");
    char *ws = (char *)workspace;
    uint32_t *keys = (uint32_t *)ws; ws += align256((size_t)(P + 1) * 4);      // sort keys, then the cells' first slots
    uint32_t *ids = (uint32_t *)ws; ws += align256((size_t)(P + 1) * 4);
    uint32_t *head = (uint32_t *)ws; ws += align256((size_t)(P + 1) * 4);      // head flags, then the long-cell work list
    int64_t *cell_of = (int64_t *)ws; ws += align256((size_t)(P + 1) * 8);
    void *sort_ws = ws; ws += sort_workspace_bytes(P);
    void *scan_ws = ws; ws += (size_t)cymf_scan_workspace_bytes(P);
    int64_t *n_valid = (int64_t *)ws;
    unsigned long long *n_work = (unsigned long long *)(ws + 8);
    CYMF_CUDA(cudaMemsetAsync(n_valid, 0, 16, st));
    const CoocView v{tokens, pos_in_line, P, window, vocab};
    const int bits = bits_for((int64_t)vocab + 1);
    cooc_col_keys_kernel<<<flat_grid(P), 256, 0, st>>>(v, keys, ids);
    CYMF_LAUNCHED();
    CYMF_TRY(radix_sort_pairs(keys, ids, P, bits, sort_ws, st));
    cooc_row_keys_kernel<<<flat_grid(P), 256, 0, st>>>(v, ids, keys);
    CYMF_LAUNCHED();
    CYMF_TRY(radix_sort_pairs(keys, ids, P, bits, sort_ws, st));
    cooc_heads_kernel<<<flat_grid(P), 256, 0, st>>>(v, ids, head, n_valid);
    CYMF_LAUNCHED();
    CYMF_TRY((exclusive_scan<uint32_t, int64_t>(head, cell_of, P, 1, (uint64_t *)scan_ws, st)));
    CYMF_CUDA(cudaMemcpyAsync(nnz_out, cell_of + P, 8, cudaMemcpyDeviceToDevice, st));
    uint32_t *start = keys, *work = head;
    cooc_cell_starts_kernel<<<flat_grid(P), 256, 0, st>>>(head, cell_of, P, n_valid, start);
    CYMF_LAUNCHED();
    cooc_sum_short_kernel<<<flat_grid(P), 256, (size_t)window * 8, st>>>(v, ids, start, nnz_out, capacity, rows, cols, vals,
                                                                      work, n_work);
    CYMF_LAUNCHED();
    cooc_sum_long_kernel<<<(unsigned)sm_count() * 8, 128, (size_t)window * 8, st>>>(v, ids, start, work, n_work, vals);
    CYMF_LAUNCHED();
    return 0;
}
