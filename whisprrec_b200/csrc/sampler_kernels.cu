// Negative sampling on the device, bit-exact with the reference's NumPy stream (models/BaseModel.py:167-177):
//
//     neg = np.random.randint(1, n_items, size=(N, 1))              # bulk draw from the global MT19937
//     for i, u in enumerate(user_id):                               # then, row by row, in order
//         while neg[i] in train_clicked_set[u]: neg[i] = np.random.randint(1, n_items)
//
// NumPy's legacy randint for a range that fits 32 bits serves every attempt with ONE raw MT19937 output r:
// v = r & mask, accepted iff v <= rng (rng = n_items - 2, mask = 2^k - 1 >= rng).  So the accepted values form one
// stream A[0], A[1], ...; the bulk draw takes A[0..N) and the redraws of the rejected rows take the following
// entries in row order.  The two recurrences (across 624-word blocks, across rejected rows) run in one CTA each,
// parallel inside; everything else uses the whole grid:
//   mt_stream_kernel   twists the state block by block (three dependency-free phases per block), tempers, filters
//                      and compacts the accepted values in order;
//   neg_bulk_kernel    (whole grid) the bulk draw: membership tests (binary search in the sorted train CSR) for every
//                      row in parallel; chunk_scan_kernel / bad_list_kernel compact the rejected rows in row order;
//   neg_redraw_kernel  hands the rest of the stream out to the rejected rows speculatively -- every rejected row assumes
//                      its first redraw succeeds; the first one whose does not is resolved sequentially and the rest
//                      re-speculate from there.
#include "common.cuh"

namespace wr {

constexpr int MT_N = 624, MT_M = 397, MT_THREADS = 640;

__device__ __forceinline__ uint32_t mt_mix(uint32_t hi, uint32_t lo, uint32_t src) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return src ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

// state_blocks[b * 624 + t]: untempered state of block b (block 0 = the caller's key, of which `pos0` outputs are
// already consumed).  acc_val / acc_raw: accepted values (low + v) and the raw index (counted from the start of block
// 0) of each, in stream order; n_acc: how many.
__global__ void __launch_bounds__(MT_THREADS) mt_stream_kernel(const uint32_t *key0, int pos0, int nblocks,
                                                               uint32_t mask, uint32_t rng, int32_t low,
                                                               uint32_t *state_blocks, int32_t *acc_val,
                                                               uint32_t *acc_raw, int64_t *n_acc) {
    __shared__ uint32_t mt[MT_N];
    __shared__ int warp_cnt[MT_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t < MT_N) mt[t] = key0[t];
    __syncthreads();
    int64_t base = 0;
    for (int b = 0; b < nblocks; ++b) {
        if (b > 0) {
            // the recurrence reaches back M = 397 words, so 227 words at a time have no dependency among themselves
            uint32_t nv = 0;
            if (t < 227) nv = mt_mix(mt[t], mt[t + 1], mt[t + MT_M]);
            __syncthreads();
            if (t < 227) mt[t] = nv;
            __syncthreads();
            if (t >= 227 && t < 454) nv = mt_mix(mt[t], mt[t + 1], mt[t - 227]);
            __syncthreads();
            if (t >= 227 && t < 454) mt[t] = nv;
            __syncthreads();
            if (t >= 454 && t < 623) nv = mt_mix(mt[t], mt[t + 1], mt[t - 227]);
            if (t == 623) nv = mt_mix(mt[623], mt[0], mt[MT_M - 1]);
            __syncthreads();
            if (t >= 454 && t < MT_N) mt[t] = nv;
            __syncthreads();
        }
        bool ok = false;
        uint32_t v = 0;
        if (t < MT_N) {
            const uint32_t s = mt[t];
            state_blocks[(int64_t)b * MT_N + t] = s;
            v = mt_temper(s) & mask;
            ok = v <= rng && (b > 0 || t >= pos0);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int i = 0; i < MT_THREADS / 32; ++i) {
            const int c = warp_cnt[i];
            if (i < w) before += c;
            total += c;
        }
        if (ok) {
            const int64_t o = base + before + __popc(bal & ((1u << lane) - 1u));
            acc_val[o] = low + (int32_t)v;
            acc_raw[o] = (uint32_t)((int64_t)b * MT_N + t);
        }
        base += total;
        __syncthreads();
    }
    if (t == 0) *n_acc = base;
}

__device__ __forceinline__ bool clicked(const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, int64_t u,
                                        int32_t item) {
    int64_t lo = ptr[u], hi = ptr[u + 1];
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int32_t x = __ldg(idx + mid);
        if (x == item) return true;
        if (x < item) lo = mid + 1; else hi = mid;
    }
    return false;
}

constexpr int NA_THREADS = 1024;

// (1) the bulk draw, every row in parallel: neg[i] = A[i]; flag the rows whose candidate is in the user's train set and
// count them per 1024-row chunk
__global__ void __launch_bounds__(NA_THREADS) neg_bulk_kernel(const int64_t *__restrict__ user, int64_t N, int64_t n_users,
                                                              const int64_t *__restrict__ train_ptr,
                                                              const int32_t *__restrict__ train_idx,
                                                              const int32_t *__restrict__ acc_val,
                                                              const int64_t *__restrict__ n_acc_p, int64_t *neg_out,
                                                              uint8_t *flag, int32_t *chunk_cnt, WrWorkspace *ws) {
    __shared__ int warp_cnt[NA_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t i = (int64_t)blockIdx.x * NA_THREADS + t;
    bool bad = false;
    if (i < N && i < *n_acc_p) {
        const int64_t u = user[i];
        const int32_t v = acc_val[i];
        neg_out[i] = v;
        if ((uint64_t)u < (uint64_t)n_users) bad = clicked(train_ptr, train_idx, u, v);
        else atomicOr(&ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
    }
    if (i < N) flag[i] = bad ? 1 : 0;
    const uint32_t bal = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    if (t == 0) {
        int m = 0;
        for (int k = 0; k < NA_THREADS / 32; ++k) m += warp_cnt[k];
        chunk_cnt[blockIdx.x] = m;
    }
}

// (2) exclusive scan of the chunk counts (one CTA; the counts are few: N / 1024)
__global__ void __launch_bounds__(NA_THREADS) chunk_scan_kernel(const int32_t *chunk_cnt, int64_t n_chunks, int64_t *chunk_off,
                                                                int64_t *n_bad) {
    __shared__ long long part[NA_THREADS];
    __shared__ long long carry;
    const int t = threadIdx.x;
    if (t == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_chunks; base += NA_THREADS) {
        const int64_t c = base + t;
        const long long v = c < n_chunks ? chunk_cnt[c] : 0;
        part[t] = v;
        __syncthreads();
        for (int o = 1; o < NA_THREADS; o <<= 1) {          // Hillis-Steele inclusive scan
            const long long add = t >= o ? part[t - o] : 0;
            __syncthreads();
            part[t] += add;
            __syncthreads();
        }
        if (c < n_chunks) chunk_off[c] = carry + part[t] - v;
        __syncthreads();
        if (t == NA_THREADS - 1) carry += part[t];
        __syncthreads();
    }
    if (t == 0) *n_bad = carry;
}

// (3) the rejected rows, in row order, as one compact list
__global__ void __launch_bounds__(NA_THREADS) bad_list_kernel(const uint8_t *__restrict__ flag, int64_t N,
                                                              const int64_t *__restrict__ chunk_off, int32_t *bad_rows) {
    __shared__ int warp_cnt[NA_THREADS / 32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t i = (int64_t)blockIdx.x * NA_THREADS + t;
    const bool bad = i < N && flag[i];
    const uint32_t bal = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) warp_cnt[w] = __popc(bal);
    __syncthreads();
    int rank = __popc(bal & ((1u << lane) - 1u));
    for (int k = 0; k < w; ++k) rank += warp_cnt[k];
    if (bad) bad_rows[chunk_off[blockIdx.x] + rank] = (int32_t)i;
}

// (4) the redraws: one CTA hands the stream A[N..) out to the rejected rows in order.  If row j of a window starts
// its draws at stream position q + j + d (d = how many extra draws the rows before it needed), it takes the first
// position from there on whose value is not in its train set.  A round evaluates the membership table
// miss[j][d], j < 64 rows, d < 16 shifts, with all 1024 threads at once; one thread then walks the table (row by row:
// advance d while miss[j][d]) and resolves as many rows as the table covers -- usually the whole window, so the dependent
// chain costs one membership latency per ~64 rows instead of one per rejected redraw.
// out_info[0] = accepted values consumed in total, out_info[1] = 1 if the stream ran dry (caller retries with more)
constexpr int RD_ROWS = 64, RD_SHIFTS = 16;
__global__ void __launch_bounds__(NA_THREADS) neg_redraw_kernel(const int64_t *__restrict__ user, int64_t N,
                                                                const int64_t *__restrict__ train_ptr,
                                                                const int32_t *__restrict__ train_idx,
                                                                const int32_t *__restrict__ acc_val,
                                                                const int64_t *__restrict__ n_acc_p,
                                                                const int32_t *__restrict__ bad_rows,
                                                                const int64_t *__restrict__ n_bad_p, int64_t *neg_out,
                                                                int64_t *out_info) {
    static_assert(RD_ROWS * RD_SHIFTS == NA_THREADS, "one (row, shift) pair per thread");
    __shared__ uint32_t s_miss[RD_ROWS];
    __shared__ int s_shift[RD_ROWS];
    __shared__ int s_resolved;
    __shared__ long long s_q, s_done;
    __shared__ int s_dry;
    const int t = threadIdx.x, j = t >> 4, d = t & 15, lane = t & 31;
    const int64_t n_acc = *n_acc_p, n_bad = *n_bad_p;
    if (t == 0) {
        s_q = N;          // A[0..N) was the bulk draw; redraws start here
        s_done = 0;
        s_dry = n_acc < N ? 1 : 0;
    }
    __syncthreads();
    while (true) {
        const long long q = s_q, done = s_done;
        if (s_dry || done >= n_bad) break;
        if (q + RD_ROWS + RD_SHIFTS >= n_acc) {         // not enough stream left to speculate safely
            __syncthreads();
            if (t == 0) s_dry = 1;
            __syncthreads();
            break;
        }
        const int m = (int)min((long long)RD_ROWS, n_bad - done);
        bool miss = false;
        int64_t row = 0;
        if (j < m) {
            row = bad_rows[done + j];
            miss = clicked(train_ptr, train_idx, user[row], acc_val[q + j + d]);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, miss);       // a warp holds rows 2w (low half) and 2w + 1
        if ((lane & 15) == 0) s_miss[j] = (lane < 16 ? bal : (bal >> 16)) & 0xffffu;
        __syncthreads();
        if (t == 0) {
            int dd = 0, r = 0;
            for (; r < m; ++r) {
                const uint32_t mb = s_miss[r];
                while (dd < RD_SHIFTS && ((mb >> dd) & 1u)) ++dd;
                if (dd >= RD_SHIFTS) break;              // row r needs positions beyond the table: next round
                s_shift[r] = dd;
            }
            s_resolved = r;
            // resolved rows consumed positions up to q + (r - 1) + dd; an unresolved row r has used up q + r + 15
            s_q = r < m ? q + r + RD_SHIFTS : q + m + dd;
            s_done = done + r;
        }
        __syncthreads();
        if (j < s_resolved && d == 0) neg_out[row] = acc_val[q + j + s_shift[j]];
        __syncthreads();
    }
    if (t == 0) {
        out_info[0] = s_q;
        out_info[1] = s_dry;
    }
}

}  // namespace wr

using namespace wr;

extern "C" size_t wr_neg_sample_scratch_bytes(int64_t N, int64_t n_items) {
    if (N <= 0 || n_items < 3) return 0;
    const uint64_t rng = (uint64_t)n_items - 2;
    uint64_t mask = rng;
    for (int sh = 1; sh <= 32; sh <<= 1) mask |= mask >> sh;
    const double p = (double)(rng + 1) / (double)(mask + 1);
    const int64_t raw = (int64_t)((double)N * 1.15 / p) + 65536;       // bulk + ~15 % redraws and slack
    const int64_t nblocks = (raw + MT_N - 1) / MT_N + 1;
    const int64_t n_chunks = (N + 1023) / 1024;
    // counters | state blocks (u32) + accepted values (i32) + accepted raw indices (u32) | per-row flag (u8) and
    // rejected-row list (i32) | per-chunk count (i32) and offset (i64)
    return 1024 + (size_t)nblocks * MT_N * 12 + ((size_t)N * 5 + 64) + ((size_t)n_chunks * 12 + 64);
}

extern "C" int wr_neg_sample_mt19937(const uint32_t *host_key, int pos, int64_t N, const int64_t *user, int64_t n_users,
                                     int64_t n_items, const int64_t *train_ptr, const int32_t *train_idx,
                                     int64_t *neg_out, uint32_t *host_key_out, int *host_pos_out, void *scratch,
                                     size_t scratch_bytes, void *ws, void *stream) {
    if (!host_key || !user || !train_ptr || !train_idx || !neg_out || !host_key_out || !host_pos_out || !scratch || !ws)
        return WR_E_NULL;
    if (N <= 0 || n_users <= 0 || n_items < 3 || n_items - 2 > 0xfffffffell || pos < 0 || pos > MT_N) return WR_E_SIZE;
    const size_t need = wr_neg_sample_scratch_bytes(N, n_items);
    if (scratch_bytes < need) return WR_E_SIZE;
    if ((reinterpret_cast<uintptr_t>(scratch) & 15u) != 0) return WR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t rng = (uint32_t)(n_items - 2);
    uint32_t mask = rng;
    for (int sh = 1; sh <= 16; sh <<= 1) mask |= mask >> sh;
    const int64_t n_chunks = (N + 1023) / 1024;
    const size_t tail = ((size_t)N * 5 + 64) + ((size_t)n_chunks * 12 + 64);
    const int64_t nblocks = (int64_t)((scratch_bytes - 1024 - tail) / ((size_t)MT_N * 12));
    if (nblocks > INT32_MAX || nblocks < 2) return WR_E_SIZE;
    uint8_t *base = static_cast<uint8_t *>(scratch);
    // layout: [0, 1024) counters | state blocks | accepted values | accepted raw indices; the caller's key is block 0
    int64_t *counters = reinterpret_cast<int64_t *>(base);                         // [0] n_acc, [1..2] out_info
    uint32_t *state_blocks = reinterpret_cast<uint32_t *>(base + 1024);
    int32_t *acc_val = reinterpret_cast<int32_t *>(state_blocks + nblocks * MT_N);
    uint32_t *acc_raw = reinterpret_cast<uint32_t *>(acc_val + nblocks * MT_N);
    int32_t *bad_rows = reinterpret_cast<int32_t *>(acc_raw + nblocks * MT_N);
    uint8_t *flag = reinterpret_cast<uint8_t *>(bad_rows + N);
    int64_t *chunk_off = reinterpret_cast<int64_t *>((reinterpret_cast<uintptr_t>(flag + N) + 15) & ~(uintptr_t)15);
    int32_t *chunk_cnt = reinterpret_cast<int32_t *>(chunk_off + n_chunks);
    cudaError_t e = cudaMemcpyAsync(state_blocks, host_key, MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    // block 0 is read from state_blocks (its own output buffer): the kernel loads it to shared memory first
    mt_stream_kernel<<<1, MT_THREADS, 0, st>>>(state_blocks, pos, (int)nblocks, mask, rng, 1, state_blocks, acc_val,
                                               acc_raw, counters);
    WR_CHECK_LAUNCH();
    if (n_chunks > INT32_MAX) return WR_E_SIZE;
    neg_bulk_kernel<<<(int)n_chunks, NA_THREADS, 0, st>>>(user, N, n_users, train_ptr, train_idx, acc_val, counters,
                                                          neg_out, flag, chunk_cnt, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    chunk_scan_kernel<<<1, NA_THREADS, 0, st>>>(chunk_cnt, n_chunks, chunk_off, counters + 3);
    WR_CHECK_LAUNCH();
    bad_list_kernel<<<(int)n_chunks, NA_THREADS, 0, st>>>(flag, N, chunk_off, bad_rows);
    WR_CHECK_LAUNCH();
    neg_redraw_kernel<<<1, NA_THREADS, 0, st>>>(user, N, train_ptr, train_idx, acc_val, counters, bad_rows, counters + 3,
                                                neg_out, counters + 1);
    WR_CHECK_LAUNCH();
    int64_t info[3];
    e = cudaMemcpyAsync(info, counters, sizeof(info), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    if (info[2] != 0 || info[1] < N) return WR_E_SIZE;          // stream ran dry: call again with a larger scratch
    // NumPy's state after consuming everything up to the last accepted value that was handed out
    uint32_t last_raw = 0;
    e = cudaMemcpyAsync(&last_raw, acc_raw + (info[1] - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    const int64_t consumed = (int64_t)last_raw + 1;             // raw outputs read, counted from the start of block 0
    int64_t blk = consumed / MT_N;
    int p = (int)(consumed % MT_N);
    if (p == 0) { blk -= 1; p = MT_N; }                        // the twist is deferred until the next draw
    e = cudaMemcpyAsync(host_key_out, state_blocks + blk * MT_N, MT_N * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    *host_pos_out = p;
    return WR_OK;
}
