// Negative sampling on the device, bit-exact with the reference's NumPy stream (models/BaseModel.py:167-177):
//
//     neg = np.random.randint(1, n_items, size=(N, 1))              # bulk draw from the global MT19937
//     for i, u in enumerate(user_id):                               # then, row by row, in order
//         while neg[i] in train_clicked_set[u]: neg[i] = np.random.randint(1, n_items)
//
// NumPy's legacy randint for a range that fits 32 bits serves every attempt with ONE raw MT19937 output r:
// v = r & mask, accepted iff v <= rng (rng = n_items - 2, mask = 2^k - 1 >= rng).  So the accepted values form one
// stream A[0], A[1], ...; the bulk draw takes A[0..N) and the redraws of the rejected rows take the following
// entries in row order.  Two single-CTA kernels (the recurrences are sequential across 624-word blocks and across
// rejected rows, parallel inside):
//   mt_stream_kernel   twists the state block by block (three dependency-free phases per block), tempers, filters
//                      and compacts the accepted values in order;
//   neg_assign_kernel  walks the rows 1024 at a time: membership tests (binary search in the sorted train CSR) in
//                      parallel, then hands the stream out to the rejected rows speculatively -- every rejected row
//                      assumes its first redraw succeeds; the first one whose does not is resolved sequentially and the
//                      rest re-speculate from there.
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

// out_info[0] = accepted values consumed in total, out_info[1] = 1 if the stream ran dry (caller retries with more)
__global__ void __launch_bounds__(NA_THREADS) neg_assign_kernel(const int64_t *__restrict__ user, int64_t N, int64_t n_users,
                                                                const int64_t *__restrict__ train_ptr,
                                                                const int32_t *__restrict__ train_idx,
                                                                const int32_t *__restrict__ acc_val,
                                                                const int64_t *__restrict__ n_acc_p, int64_t *neg_out,
                                                                int64_t *out_info, WrWorkspace *ws) {
    __shared__ int warp_cnt[NA_THREADS / 32];
    __shared__ int s_first;
    __shared__ long long s_q;
    __shared__ volatile int s_dry;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t n_acc = *n_acc_p;
    if (t == 0) {
        s_q = N;          // A[0..N) is the bulk draw; redraws start here
        s_dry = n_acc < N ? 1 : 0;
    }
    __syncthreads();
    if (s_dry) {
        if (t == 0) { out_info[0] = 0; out_info[1] = 1; }
        return;
    }
    for (int64_t base = 0; base < N; base += NA_THREADS) {
        const int64_t i = base + t;
        int64_t u = 0;
        bool bad = false;
        if (i < N) {
            u = user[i];
            const int32_t v = acc_val[i];
            neg_out[i] = v;
            if ((uint64_t)u < (uint64_t)n_users) bad = clicked(train_ptr, train_idx, u, v);
            else atomicOr(&ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
        }
        // rank of this row among the chunk's rejected rows, in row order
        const uint32_t bal = __ballot_sync(0xffffffffu, bad);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int rank = __popc(bal & ((1u << lane) - 1u)), m = 0;
        for (int k = 0; k < NA_THREADS / 32; ++k) {
            const int c = warp_cnt[k];
            if (k < w) rank += c;
            m += c;
        }
        int done = 0;               // ranks [0, done) have their final negative
        while (done < m) {
            if (t == 0) s_first = m;
            __syncthreads();
            const long long qb = s_q;               // stream index that rank `done` draws next
            int32_t v = 0;
            long long c = 0;
            bool miss = false;
            const bool mine = bad && rank >= done;
            if (mine) {
                c = qb + (rank - done);
                if (c >= n_acc) { s_dry = 1; miss = true; }
                else { v = acc_val[c]; miss = clicked(train_ptr, train_idx, u, v); }
                if (miss) atomicMin(&s_first, rank);
            }
            __syncthreads();
            const int first = s_first;
            if (mine && rank < first) neg_out[i] = v;          // speculation held for these
            if (mine && rank == first) {                        // this row keeps drawing until it is clear
                while (!s_dry) {
                    ++c;
                    if (c >= n_acc) { s_dry = 1; break; }
                    v = acc_val[c];
                    if (!clicked(train_ptr, train_idx, u, v)) break;
                }
                neg_out[i] = v;
                s_q = c + 1;
            }
            if (first == m && t == 0) s_q = qb + (m - done);
            __syncthreads();
            if (s_dry) break;
            done = first == m ? m : first + 1;
        }
        __syncthreads();
        if (s_dry) break;
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
    // state blocks (u32) + accepted values (i32) + accepted raw indices (u32) + counters
    return (size_t)nblocks * MT_N * 12 + 1024;
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
    const int64_t nblocks = (int64_t)((scratch_bytes - 1024) / ((size_t)MT_N * 12));
    if (nblocks > INT32_MAX) return WR_E_SIZE;
    uint8_t *base = static_cast<uint8_t *>(scratch);
    // layout: [0, 1024) counters | state blocks | accepted values | accepted raw indices; the caller's key is block 0
    int64_t *counters = reinterpret_cast<int64_t *>(base);                         // [0] n_acc, [1..2] out_info
    uint32_t *state_blocks = reinterpret_cast<uint32_t *>(base + 1024);
    int32_t *acc_val = reinterpret_cast<int32_t *>(state_blocks + nblocks * MT_N);
    uint32_t *acc_raw = reinterpret_cast<uint32_t *>(acc_val + nblocks * MT_N);
    cudaError_t e = cudaMemcpyAsync(state_blocks, host_key, MT_N * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return (int)e;
    // block 0 is read from state_blocks (its own output buffer): the kernel loads it to shared memory first
    mt_stream_kernel<<<1, MT_THREADS, 0, st>>>(state_blocks, pos, (int)nblocks, mask, rng, 1, state_blocks, acc_val,
                                               acc_raw, counters);
    WR_CHECK_LAUNCH();
    neg_assign_kernel<<<1, NA_THREADS, 0, st>>>(user, N, n_users, train_ptr, train_idx, acc_val, counters, neg_out,
                                                counters + 1, (WrWorkspace *)ws);
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
