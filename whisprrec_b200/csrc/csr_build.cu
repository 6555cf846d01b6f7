// Graph construction on the device (reference models/general/LightGCN.py:54-88):
// (user, item) edge list -> CSR structure of the bipartite adjacency [[0, R], [R^T, 0]], rows and columns ascending,
// duplicate pairs dropped (the reference builds R from per-user SETS).  Everything is hand-written here: a stable LSD
// radix sort of packed 64-bit keys (8-bit digits; per-tile histograms -> multi-level exclusive scan -> stable scatter
// with warp match-any ranking), adjacent-unique compaction, and row pointers filled from the key boundaries (no atomics,
// so the result is deterministic).  The fp32 weights are NOT computed here: d^-1/2 comes from the caller's NumPy call
// (the reference's own, LightGCN.py:89-93, which is not correctly rounded) and wr_csr_norm_weights applies it.
#include "common.cuh"

namespace wr {

constexpr int RS_THREADS = 256, RS_KEYS = 16, RS_TILE = RS_THREADS * RS_KEYS;      // 4096 keys per CTA
constexpr int SC_TILE = 4096;                                                      // scan: elements per CTA
constexpr uint64_t KEY_INVALID = ~0ull;

// ---- keys ----
__global__ void __launch_bounds__(256) pack_keys_kernel(const int64_t *__restrict__ hi_ids, const int64_t *__restrict__ lo_ids,
                                                         int64_t n, int64_t hi_bound, int64_t lo_bound, uint64_t *keys,
                                                         WrWorkspace *ws) {
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t h = hi_ids[k], l = lo_ids[k];
        const bool ok = (uint64_t)h < (uint64_t)hi_bound && (uint64_t)l < (uint64_t)lo_bound;
        if (!ok) atomicOr(&ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
        keys[k] = ok ? ((uint64_t)h << 32) | (uint64_t)l : KEY_INVALID;       // invalid pairs sort to the end and are dropped
    }
}

// (hi << 32 | lo) -> (lo << 32 | hi) for the first *n_dev keys
__global__ void __launch_bounds__(256) swap_halves_kernel(const uint64_t *__restrict__ in, uint64_t *out, const int64_t *n_dev) {
    const int64_t n = *n_dev;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t x = in[k];
        out[k] = (x << 32) | (x >> 32);
    }
}

// ---- radix sort pass: per-tile digit histogram, hist[digit * n_tiles + tile] ----
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const uint64_t *__restrict__ keys, const int64_t *n_dev, int shift,
                                                                uint32_t *hist, int n_tiles) {
    __shared__ uint32_t h[256];
    const int64_t n = *n_dev;
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k) {
        const int64_t idx = base + k * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// ---- radix sort pass: stable scatter.  Warp w of a tile owns keys [w * 32 * RS_KEYS, (w + 1) * 32 * RS_KEYS) of it,
//      taken 32 at a time; equal digits are ranked by (iteration, lane) with match-any, then by warp, then by tile. ----
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *out, const int64_t *n_dev,
                                                                   int shift, const uint32_t *__restrict__ offs, int n_tiles) {
    __shared__ uint32_t cnt[RS_THREADS / 32][256];
    __shared__ uint32_t gbase[256];
    const int64_t n = *n_dev;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)w * 32 * RS_KEYS;
    uint64_t key[RS_KEYS];
    uint32_t rank[RS_KEYS];
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k) {
        const int64_t idx = base + k * 32 + lane;
        const bool valid = idx < n;
        key[k] = valid ? in[idx] : 0;
        const uint32_t d = valid ? (uint32_t)((key[k] >> shift) & 255u) : 256u;
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        const uint32_t lt = __popc(mask & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid) prev = cnt[w][d];
        __syncwarp();
        if (valid && lt == 0) cnt[w][d] = prev + __popc(mask);
        __syncwarp();
        rank[k] = prev + lt;
    }
    __syncthreads();
    {   // thread d: exclusive prefix of digit d over the warps, and the tile's global base for that digit
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_THREADS / 32; ++ww) {
            const uint32_t t = cnt[ww][d];
            cnt[ww][d] = run;
            run += t;
        }
        gbase[d] = offs[(size_t)d * n_tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k) {
        const int64_t idx = base + k * 32 + lane;
        if (idx < n) {
            const uint32_t d = (uint32_t)((key[k] >> shift) & 255u);
            out[(size_t)gbase[d] + cnt[w][d] + rank[k]] = key[k];
        }
    }
}

// ---- exclusive scan of uint32 (multi-level: tile sums -> scan of the sums -> apply) ----
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *smem /*[32]*/, uint32_t *total) {
    // exclusive prefix of v over the 256 threads of the CTA; *total = sum
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        smem[lane] = s;         // inclusive over warps
    }
    __syncthreads();
    const uint32_t warp_off = w ? smem[w - 1] : 0;
    *total = smem[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_off + x - v;
}

__global__ void __launch_bounds__(256) scan_tile_sums_kernel(const uint32_t *__restrict__ data, int64_t n, uint32_t *sums) {
    __shared__ uint32_t sm[32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * (SC_TILE / 256);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SC_TILE / 256; ++k)
        if (base + k < n) s += data[base + k];
    uint32_t total;
    block_exclusive_scan(s, sm, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

// in place: data[i] <- tile_off[tile] + exclusive prefix inside the tile
__global__ void __launch_bounds__(256) scan_apply_kernel(uint32_t *data, int64_t n, const uint32_t *__restrict__ tile_off) {
    __shared__ uint32_t sm[32];
    const int64_t base = (int64_t)blockIdx.x * SC_TILE + (int64_t)threadIdx.x * (SC_TILE / 256);
    uint32_t v[SC_TILE / 256];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < SC_TILE / 256; ++k) {
        v[k] = base + k < n ? data[base + k] : 0;
        s += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(s, sm, &total) + (tile_off ? tile_off[blockIdx.x] : 0u);
#pragma unroll
    for (int k = 0; k < SC_TILE / 256; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
}

// levels: data (n) -> sums1 (ceil(n / 4096)) -> sums2 -> ...; tmp must hold the sums of all levels
static int exclusive_scan_u32(uint32_t *data, int64_t n, uint32_t *tmp, cudaStream_t st) {
    if (n <= 0) return WR_OK;
    const int64_t tiles = (n + SC_TILE - 1) / SC_TILE;
    if (tiles == 1) {
        scan_apply_kernel<<<1, 256, 0, st>>>(data, n, nullptr);
        WR_CHECK_LAUNCH();
        return WR_OK;
    }
    scan_tile_sums_kernel<<<(unsigned)tiles, 256, 0, st>>>(data, n, tmp);
    WR_CHECK_LAUNCH();
    const int rc = exclusive_scan_u32(tmp, tiles, tmp + tiles, st);
    if (rc) return rc;
    scan_apply_kernel<<<(unsigned)tiles, 256, 0, st>>>(data, n, tmp);
    WR_CHECK_LAUNCH();
    return WR_OK;
}

static int64_t scan_tmp_elems(int64_t n) {
    int64_t total = 0;
    while (n > SC_TILE) {
        n = (n + SC_TILE - 1) / SC_TILE;
        total += n;
    }
    return total + 1;
}

// ---- adjacent-unique compaction of sorted keys ----
__global__ void __launch_bounds__(256) unique_count_kernel(const uint64_t *__restrict__ keys, const int64_t *n_dev, uint32_t *tile_cnt) {
    __shared__ uint32_t sm[32];
    const int64_t n = *n_dev;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_KEYS;
    uint32_t c = 0;
    uint64_t prev = base > 0 && base - 1 < n ? keys[base - 1] : KEY_INVALID;
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k) {
        const uint64_t x = base + k < n ? keys[base + k] : KEY_INVALID;
        c += (x != KEY_INVALID && (x != prev || base + k == 0)) ? 1u : 0u;
        prev = x;
    }
    uint32_t total;
    block_exclusive_scan(c, sm, &total);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) unique_compact_kernel(const uint64_t *__restrict__ keys, const int64_t *n_dev,
                                                              const uint32_t *__restrict__ tile_off, uint64_t *out) {
    __shared__ uint32_t sm[32];
    const int64_t n = *n_dev;
    const int64_t base = (int64_t)blockIdx.x * RS_TILE + (int64_t)threadIdx.x * RS_KEYS;
    uint64_t x[RS_KEYS];
    bool keep[RS_KEYS];
    uint32_t c = 0;
    uint64_t prev = base > 0 && base - 1 < n ? keys[base - 1] : KEY_INVALID;
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k) {
        x[k] = base + k < n ? keys[base + k] : KEY_INVALID;
        keep[k] = x[k] != KEY_INVALID && (x[k] != prev || base + k == 0);
        c += keep[k] ? 1u : 0u;
        prev = x[k];
    }
    uint32_t total;
    uint32_t pos = block_exclusive_scan(c, sm, &total) + tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < RS_KEYS; ++k)
        if (keep[k]) out[pos++] = x[k];
}

// n_out = tile_off[last] + tile_cnt of the last tile: computed before the scan overwrote tile_cnt, so the caller passes
// the exclusive offsets and the count of the last tile separately
__global__ void set_count_kernel(const uint32_t *tile_off_last, const uint32_t *last_cnt, int64_t *n_out) {
    *n_out = (int64_t)*tile_off_last + (int64_t)*last_cnt;
}
__global__ void set_scalar_kernel(int64_t *p, int64_t v) { *p = v; }
__global__ void save_u32_kernel(const uint32_t *src, uint32_t *dst) { *dst = *src; }
__global__ void save_i64_kernel(const int64_t *src, int64_t *dst) { *dst = *src; }

// keys of the kept edges of a CSR matrix: edge e of the main graph sits in row upper_bound(rowptr, e) - 1
__global__ void __launch_bounds__(256) subgraph_keys_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                             int64_t n_rows, const int64_t *__restrict__ keep, int64_t K,
                                                             int transpose, uint64_t *keys, WrWorkspace *ws) {
    const int64_t nnz = rowptr[n_rows];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < K; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = keep[k];
        if ((uint64_t)e >= (uint64_t)nnz) {
            atomicOr(&ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
            keys[k] = KEY_INVALID;
            continue;
        }
        int64_t lo = 0, hi = n_rows;                 // first row whose rowptr[row + 1] > e
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (rowptr[mid + 1] > e) hi = mid; else lo = mid + 1;
        }
        const uint64_t r = (uint64_t)lo, c = (uint64_t)(uint32_t)col[e];
        keys[k] = transpose ? (c << 32) | r : (r << 32) | c;
    }
}

// ---- CSR rows from sorted distinct keys (hi = row inside the block of rows, lo = column inside the other block) ----
// rowptr[row_base + r] = edge_base + (first key whose hi >= r); col[edge_base + k] = col_base + lo_k.
__global__ void __launch_bounds__(256) rows_from_keys_kernel(const uint64_t *__restrict__ keys, const int64_t *n_dev, int64_t n_rows,
                                                              int64_t row_base, int64_t col_base, int64_t edge_base_factor,
                                                              int64_t *rowptr, int32_t *col, int write_end) {
    const int64_t n = *n_dev;
    const int64_t edge_base = edge_base_factor * n;        // 0 for the user rows, n for the item rows
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n == 0) {
        for (int64_t r = t0; r < n_rows; r += stride) rowptr[row_base + r] = edge_base;
        if (write_end && t0 == 0) rowptr[row_base + n_rows] = edge_base;
        return;
    }
    for (int64_t k = t0; k < n; k += stride) {
        const uint64_t x = keys[k];
        const int64_t r = (int64_t)(x >> 32);
        col[edge_base + k] = (int32_t)(col_base + (int64_t)(x & 0xffffffffu));
        const int64_t rp = k ? (int64_t)(keys[k - 1] >> 32) : -1;
        for (int64_t q = rp + 1; q <= r; ++q) rowptr[row_base + q] = edge_base + k;       // rows rp+1 .. r start at edge k
        if (k == n - 1) {
            for (int64_t q = r + 1; q < n_rows; ++q) rowptr[row_base + q] = edge_base + n; // trailing empty rows
            if (write_end) rowptr[row_base + n_rows] = edge_base + n;
        }
    }
}

__global__ void finish_nnz_kernel(const int64_t *n_dev, int64_t *nnz_out) { *nnz_out = 2 * *n_dev; }

struct BuildScratch {
    uint64_t *keysA, *keysB;
    uint32_t *hist, *scan_tmp, *tile_cnt, *last_cnt;
    int64_t *n_dev;
    int n_tiles;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t carve(BuildScratch *b, char *base, int64_t E) {
    const int64_t n_tiles = (E + RS_TILE - 1) / RS_TILE;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += align256(bytes); return p; };
    uint64_t *a = (uint64_t *)take((size_t)E * 8), *bb = (uint64_t *)take((size_t)E * 8);
    uint32_t *hist = (uint32_t *)take((size_t)256 * n_tiles * 4);
    uint32_t *tmp = (uint32_t *)take((size_t)scan_tmp_elems(256 * n_tiles) * 4);
    uint32_t *tc = (uint32_t *)take((size_t)(n_tiles + 1) * 4);
    uint32_t *lc = (uint32_t *)take(256);
    int64_t *nd = (int64_t *)take(256);
    if (b) *b = BuildScratch{a, bb, hist, tmp, tc, lc, nd, (int)n_tiles};
    return off;
}

static int bits_for(int64_t bound) {      // bits needed for values in [0, bound)
    int b = 0;
    while (b < 32 && ((int64_t)1 << b) < bound) ++b;
    return b < 1 ? 1 : b;
}

// LSD passes over the digits that can be non-zero; the sorted keys end up in *cur (keysA or keysB)
static int radix_sort(BuildScratch &s, uint64_t **cur, uint64_t **alt, int lo_bits, int hi_bits, cudaStream_t st) {
    int shifts[8], ns = 0;
    for (int sh = 0; sh < lo_bits; sh += 8) shifts[ns++] = sh;
    for (int sh = 0; sh < hi_bits; sh += 8) shifts[ns++] = 32 + sh;
    for (int i = 0; i < ns; ++i) {
        radix_hist_kernel<<<s.n_tiles, RS_THREADS, 0, st>>>(*cur, s.n_dev, shifts[i], s.hist, s.n_tiles);
        WR_CHECK_LAUNCH();
        const int rc = exclusive_scan_u32(s.hist, (int64_t)256 * s.n_tiles, s.scan_tmp, st);
        if (rc) return rc;
        radix_scatter_kernel<<<s.n_tiles, RS_THREADS, 0, st>>>(*cur, *alt, s.n_dev, shifts[i], s.hist, s.n_tiles);
        WR_CHECK_LAUNCH();
        uint64_t *t = *cur;
        *cur = *alt;
        *alt = t;
    }
    return WR_OK;
}

}  // namespace wr

using namespace wr;

extern "C" size_t wr_csr_build_scratch_bytes(int64_t E) {
    if (E <= 0) return 0;
    return carve(nullptr, nullptr, E) + 256;
}

extern "C" int wr_csr_build(const int64_t *edge_u, const int64_t *edge_i, int64_t E, int64_t n_users, int64_t n_items,
                            int64_t *rowptr, int32_t *col, int64_t *nnz_out, void *scratch, size_t scratch_bytes, void *ws,
                            void *stream) {
    if (!edge_u || !edge_i || !rowptr || !col || !nnz_out || !scratch || !ws) return WR_E_NULL;
    if (E <= 0 || E >= ((int64_t)1 << 32) - RS_TILE || n_users <= 0 || n_items <= 0 || n_users >= INT32_MAX ||
        n_items >= INT32_MAX || n_users + n_items >= INT32_MAX)
        return WR_E_SIZE;
    if (scratch_bytes < wr_csr_build_scratch_bytes(E)) return WR_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    char *base = (char *)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    carve(&s, base, E);
    const int ubits = bits_for(n_users), ibits = bits_for(n_items);
    const int g = (int)((E + 255) / 256 < 16 * (int64_t)kSMs ? (E + 255) / 256 : 16 * (int64_t)kSMs);
    set_scalar_kernel<<<1, 1, 0, st>>>(s.n_dev, E);
    WR_CHECK_LAUNCH();
    // 1. user-major keys (u << 32 | i), sorted
    pack_keys_kernel<<<g, 256, 0, st>>>(edge_u, edge_i, E, n_users, n_items, s.keysA, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    uint64_t *cur = s.keysA, *alt = s.keysB;
    // invalid keys (all ones) need every digit looked at to stay last; they only exist after an id error, which the
    // status word reports -- the passes below are sized for valid keys and leave such keys wherever their low digits put
    // them, which is why unique_* test for KEY_INVALID explicitly
    int rc = radix_sort(s, &cur, &alt, ibits, ubits, st);
    if (rc) return rc;
    // 2. drop duplicate pairs: the distinct keys go to `alt`, their number to n_dev
    unique_count_kernel<<<s.n_tiles, 256, 0, st>>>(cur, s.n_dev, s.tile_cnt);
    WR_CHECK_LAUNCH();
    save_u32_kernel<<<1, 1, 0, st>>>(s.tile_cnt + s.n_tiles - 1, s.last_cnt);
    WR_CHECK_LAUNCH();
    rc = exclusive_scan_u32(s.tile_cnt, s.n_tiles, s.scan_tmp, st);
    if (rc) return rc;
    unique_compact_kernel<<<s.n_tiles, 256, 0, st>>>(cur, s.n_dev, s.tile_cnt, alt);
    WR_CHECK_LAUNCH();
    set_count_kernel<<<1, 1, 0, st>>>(s.tile_cnt + s.n_tiles - 1, s.last_cnt, s.n_dev);
    WR_CHECK_LAUNCH();
    {
        uint64_t *t = cur;
        cur = alt;
        alt = t;
    }
    // 3. user rows: rowptr[0 .. U), col[0 .. E') = U + item
    rows_from_keys_kernel<<<g, 256, 0, st>>>(cur, s.n_dev, n_users, 0, n_users, 0, rowptr, col, 0);
    WR_CHECK_LAUNCH();
    // 4. item-major keys (i << 32 | u) of the distinct pairs, sorted; item rows: rowptr[U .. U + I], col[E' .. 2E') = user
    swap_halves_kernel<<<g, 256, 0, st>>>(cur, alt, s.n_dev);
    WR_CHECK_LAUNCH();
    {
        uint64_t *t = cur;
        cur = alt;
        alt = t;
    }
    rc = radix_sort(s, &cur, &alt, ubits, ibits, st);
    if (rc) return rc;
    rows_from_keys_kernel<<<g, 256, 0, st>>>(cur, s.n_dev, n_items, n_users, 0, 1, rowptr, col, 1);
    WR_CHECK_LAUNCH();
    finish_nnz_kernel<<<1, 1, 0, st>>>(s.n_dev, nnz_out);
    WR_CHECK_LAUNCH();
    return WR_OK;
}

// wr_subgraph_csr: CSR structure (or the structure of the transpose) of the sub-graph that keeps edges keep[0..K) of a
// square CSR matrix -- SGL's edge-dropout views (utils/augmentor.py:77-111; the kept edge numbers come from Python's
// random stream, wr_pyrandom_sample).  Same machinery as wr_csr_build: packed keys, radix sort, row pointers.
extern "C" size_t wr_subgraph_csr_scratch_bytes(int64_t K) { return wr_csr_build_scratch_bytes(K); }

extern "C" int wr_subgraph_csr(const int64_t *rowptr, const int32_t *col, int64_t n_rows, const int64_t *keep, int64_t K,
                               int transpose, int64_t *out_rowptr, int32_t *out_col, int64_t *nnz_out, void *scratch,
                               size_t scratch_bytes, void *ws, void *stream) {
    if (!rowptr || !col || !keep || !out_rowptr || !out_col || !nnz_out || !scratch || !ws) return WR_E_NULL;
    if (K <= 0 || K >= ((int64_t)1 << 32) - RS_TILE || n_rows <= 0 || n_rows >= INT32_MAX) return WR_E_SIZE;
    if (scratch_bytes < wr_csr_build_scratch_bytes(K)) return WR_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    BuildScratch s;
    char *base = (char *)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    carve(&s, base, K);
    const int bits = bits_for(n_rows);
    const int g = (int)((K + 255) / 256 < 16 * (int64_t)kSMs ? (K + 255) / 256 : 16 * (int64_t)kSMs);
    set_scalar_kernel<<<1, 1, 0, st>>>(s.n_dev, K);
    WR_CHECK_LAUNCH();
    subgraph_keys_kernel<<<g, 256, 0, st>>>(rowptr, col, n_rows, keep, K, transpose, s.keysA, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    uint64_t *cur = s.keysA, *alt = s.keysB;
    int rc = radix_sort(s, &cur, &alt, bits, bits, st);
    if (rc) return rc;
    // the kept edge numbers are distinct, but a repeated one must not become a repeated entry
    unique_count_kernel<<<s.n_tiles, 256, 0, st>>>(cur, s.n_dev, s.tile_cnt);
    WR_CHECK_LAUNCH();
    save_u32_kernel<<<1, 1, 0, st>>>(s.tile_cnt + s.n_tiles - 1, s.last_cnt);
    WR_CHECK_LAUNCH();
    rc = exclusive_scan_u32(s.tile_cnt, s.n_tiles, s.scan_tmp, st);
    if (rc) return rc;
    unique_compact_kernel<<<s.n_tiles, 256, 0, st>>>(cur, s.n_dev, s.tile_cnt, alt);
    WR_CHECK_LAUNCH();
    set_count_kernel<<<1, 1, 0, st>>>(s.tile_cnt + s.n_tiles - 1, s.last_cnt, s.n_dev);
    WR_CHECK_LAUNCH();
    rows_from_keys_kernel<<<g, 256, 0, st>>>(alt, s.n_dev, n_rows, 0, 0, 0, out_rowptr, out_col, 1);
    WR_CHECK_LAUNCH();
    save_i64_kernel<<<1, 1, 0, st>>>(s.n_dev, nnz_out);
    WR_CHECK_LAUNCH();
    return WR_OK;
}
