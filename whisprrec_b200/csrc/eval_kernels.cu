// Full-ranking evaluation, fp32-exact variant: a register-tiled score kernel whose epilogue applies the
// per-user history mask (bitmask built from the sorted history CSR), counts items that beat the target and
// maintains a per-row top-k -- the [rows, n_items] score matrix of BaseRunner.interface is never written.
#include "common.cuh"

namespace wr {

constexpr int EV_TR = 64;   // eval rows per CTA
constexpr int EV_TI = 64;   // items per tile
constexpr int EV_KMAX = 32;

struct EvalParams {
    const float *Uemb, *Iemb;
    const int64_t *user, *pos;
    int64_t R, n_users, n_items;
    const int64_t *hist_ptr;
    const int32_t *hist_idx;
    int k;
    int32_t *topk_idx;
    float *topk_val;
    int32_t *rank;
    float *target;
    float *scores;  // optional dense [R, n_items] output (unmasked), the reference's full_predict
    int splits, tiles_per_split, n_tiles;
    WrWorkspace *ws;
    // item-shard mode (wr_eval_rank_topk_shard): Uemb holds the R gathered user rows, the target scores are an
    // input (the target item may live on another shard: pos = -1) and `target` is not written
    const float *target_in;
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Thread (ty, tx) of the 16x16 grid owns rows ty+16i and columns tx+16j (i, j < 4): with a row pitch of D+4
// floats the 128-bit shared loads of a quarter warp then fall into 8 distinct 4-bank groups (no conflicts).
template <int D, bool TOPK>
__global__ void __launch_bounds__(256) eval_rank_kernel(EvalParams p) {
    constexpr int LD = D + 4;
    constexpr int D4 = D / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *As = reinterpret_cast<float *>(smem_raw);            // [EV_TR][LD]
    float *Bs = As + EV_TR * LD;                                 // [2][EV_TI][LD]
    uint32_t *mask = reinterpret_cast<uint32_t *>(Bs + 2 * EV_TI * LD);  // [2][EV_TR][2]
    float *st_s = reinterpret_cast<float *>(mask + 2 * EV_TR * 2);       // [EV_TR]
    float *S = st_s + EV_TR;                                     // [EV_TR][EV_TI+1]     (TOPK only)
    float *tkv = S + EV_TR * (EV_TI + 1);                        // [EV_TR][EV_KMAX]     (TOPK only)
    int32_t *tki = reinterpret_cast<int32_t *>(tkv + EV_TR * EV_KMAX);

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4, lane = tid & 31, warp = tid >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * EV_TR;
    const int t0 = blockIdx.y * p.tiles_per_split;
    const int t1 = min(p.n_tiles, t0 + p.tiles_per_split);

    // ---- user rows of this CTA -> shared (zero rows for r >= R or an out-of-range id) ----
    for (int c = tid; c < EV_TR * D4; c += 256) {
        const int i = c / D4, v = c - i * D4;
        const int64_t r = row0 + i;
        float4 x = f4_zero();
        if (r < p.R) {
            const int64_t u = p.user[r];
            if ((uint64_t)u < (uint64_t)p.n_users) x = ldg4(p.Uemb + (p.target_in ? r : u) * D + 4 * v);
        }
        *reinterpret_cast<float4 *>(As + i * LD + 4 * v) = x;
    }
    auto load_tile = [&](int t, int buf) {
        const int64_t j0 = (int64_t)t * EV_TI;
        float *dst = Bs + buf * EV_TI * LD;
        for (int c = tid; c < EV_TI * D4; c += 256) {
            const int i = c / D4, v = c - i * D4;
            if (j0 + i < p.n_items) cp_async16(dst + i * LD + 4 * v, p.Iemb + (j0 + i) * D + 4 * v);
        }
        cp_async_commit();
    };
    if (t0 < t1) load_tile(t0, 0);
    if (TOPK) {
        for (int c = tid; c < EV_TR * EV_KMAX; c += 256) {
            tkv[c] = -INFINITY;
            tki[c] = -1;
        }
    }
    __syncthreads();

    // ---- per-row state held by threads 0..63: target score, history cursor ----
    int64_t cur = 0, hend = 0;
    bool row_live = false;
    if (tid < EV_TR) {
        const int64_t r = row0 + tid;
        float st = 0.f;
        if (r < p.R) {
            const int64_t u = p.user[r], it = p.pos[r];
            if ((uint64_t)u < (uint64_t)p.n_users && (p.target_in || (uint64_t)it < (uint64_t)p.n_items)) {
                row_live = true;
                if (p.target_in) {
                    st = p.target_in[r];
                } else {
                    const float *a = As + tid * LD;
                    const float *b = p.Iemb + it * D;
#pragma unroll 4
                    for (int v = 0; v < D4; ++v) {
                        const float4 x = *reinterpret_cast<const float4 *>(a + 4 * v), y = ldg4(b + 4 * v);
                        st = fmaf(x.x, y.x, st);
                        st = fmaf(x.y, y.y, st);
                        st = fmaf(x.z, y.z, st);
                        st = fmaf(x.w, y.w, st);
                    }
                }
                cur = p.hist_ptr[u];
                hend = p.hist_ptr[u + 1];
                // first history entry at or after this CTA's first item
                const int32_t first = (int32_t)min((int64_t)t0 * EV_TI, (int64_t)INT32_MAX);
                int64_t lo = cur, hi = hend;
                while (lo < hi) {
                    const int64_t mid = (lo + hi) >> 1;
                    if (p.hist_idx[mid] < first) lo = mid + 1; else hi = mid;
                }
                cur = lo;
                if (blockIdx.y == 0 && !p.target_in) p.target[r] = st;
            } else {
                atomicOr(&p.ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
            }
        }
        st_s[tid] = st;
    }

    int cnt[4] = {0, 0, 0, 0};
    for (int t = t0; t < t1; ++t) {
        const int buf = (t - t0) & 1;
        const int64_t j0 = (int64_t)t * EV_TI;
        if (t + 1 < t1) load_tile(t + 1, buf ^ 1);
        if (tid < EV_TR) {
            uint32_t m0 = 0, m1 = 0;
            if (!row_live) {
                m0 = m1 = 0xffffffffu;
            } else {
                while (cur < hend) {
                    const int64_t h = (int64_t)p.hist_idx[cur] - j0;
                    if (h >= EV_TI) break;
                    if (h >= 32) m1 |= 1u << (h - 32); else if (h >= 0) m0 |= 1u << h;
                    ++cur;
                }
                const int64_t live = p.n_items - j0;  // columns past the table are masked
                if (live < 32) { m0 |= ~0u << (live < 0 ? 0 : live); m1 = ~0u; }
                else if (live < 64) m1 |= ~0u << (live - 32);
            }
            mask[(buf * EV_TR + tid) * 2 + 0] = m0;
            mask[(buf * EV_TR + tid) * 2 + 1] = m1;
        }
        if (t + 1 < t1) cp_async_wait<1>(); else cp_async_wait<0>();
        __syncthreads();

        const float *Bt = Bs + buf * EV_TI * LD;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 2
        for (int v = 0; v < D4; ++v) {
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(As + (ty + 16 * i) * LD + 4 * v);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(Bt + (tx + 16 * j) * LD + 4 * v);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // the same d = 0..D-1 FMA chain as the target score above: comparisons are consistent
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int rl = ty + 16 * i;
            const float st = st_s[rl];
            const uint32_t m0 = mask[(buf * EV_TR + rl) * 2 + 0], m1 = mask[(buf * EV_TR + rl) * 2 + 1];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t word = j < 2 ? m0 : m1;
                const bool masked = (word >> (tx + 16 * (j & 1))) & 1u;
                cnt[i] += (!masked && acc[i][j] > st) ? 1 : 0;
                if (TOPK) S[rl * (EV_TI + 1) + tx + 16 * j] = masked ? -INFINITY : acc[i][j];
                if (p.scores) {
                    const int64_t r = row0 + rl, c = j0 + tx + 16 * j;
                    if (r < p.R && c < p.n_items) p.scores[r * p.n_items + c] = acc[i][j];
                }
            }
        }
        if (TOPK) {
            __syncthreads();
            // warp w merges the tile's scores of rows 8w..8w+7 into their sorted lists (lane = list slot)
            const uint32_t kmask = p.k >= 32 ? 0xffffffffu : ((1u << p.k) - 1u);
            for (int q = 0; q < 8; ++q) {
                const int rl = warp * 8 + q;
                float tv = tkv[rl * EV_KMAX + lane];
                int32_t ti = tki[rl * EV_KMAX + lane];
                float tau = __shfl_sync(0xffffffffu, tv, p.k - 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float s = S[rl * (EV_TI + 1) + h * 32 + lane];
                    const int32_t id = (int32_t)(j0 + h * 32 + lane);
                    // items arrive in ascending id, so a tie with the k-th best never displaces it
                    uint32_t pend = __ballot_sync(0xffffffffu, s > tau);
                    while (pend) {
                        const int src = __ffs(pend) - 1;
                        pend &= pend - 1;
                        const float v = __shfl_sync(0xffffffffu, s, src);
                        const int32_t vid = __shfl_sync(0xffffffffu, id, src);
                        if (!(v > tau)) continue;
                        const int ins = __popc(__ballot_sync(0xffffffffu, tv >= v) & kmask);
                        const float up_v = __shfl_up_sync(0xffffffffu, tv, 1);
                        const int32_t up_i = __shfl_up_sync(0xffffffffu, ti, 1);
                        if (lane == ins) { tv = v; ti = vid; }
                        else if (lane > ins) { tv = up_v; ti = up_i; }
                        tau = __shfl_sync(0xffffffffu, tv, p.k - 1);
                    }
                }
                tkv[rl * EV_KMAX + lane] = tv;
                tki[rl * EV_KMAX + lane] = ti;
            }
        }
        __syncthreads();
    }

    // ---- ranks: fold the 16 column-threads of each row ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int c = cnt[i];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        const int64_t r = row0 + ty + 16 * i;
        if (tx == 0 && r < p.R) {
            if (p.splits == 1) p.rank[r] = 1 + c;
            else atomicAdd(&p.rank[r], c);
        }
    }
    if (TOPK) {
        for (int c = tid; c < EV_TR * p.k; c += 256) {
            const int i = c / p.k, s = c - i * p.k;
            const int64_t r = row0 + i;
            if (r < p.R) {
                p.topk_idx[r * p.k + s] = tki[i * EV_KMAX + s];
                p.topk_val[r * p.k + s] = tkv[i * EV_KMAX + s];
            }
        }
    }
}

__global__ void fill_i32_kernel(int32_t *x, int64_t n, int32_t v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}

// HR@k / NDCG@k sums in float64; per-block partials are added in block order by the last block.
__global__ void __launch_bounds__(256) metrics_kernel(const int32_t *__restrict__ rank, int64_t R, int nk, int k0,
                                                       int k1, int k2, int k3, int k4, int k5, int k6, int k7,
                                                       double *out, WrWorkspace *ws) {
    const int ks[8] = {k0, k1, k2, k3, k4, k5, k6, k7};
    __shared__ double red[8][16];
    __shared__ bool flag;
    double hr[8], nd[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) hr[q] = nd[q] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < R; i += (int64_t)gridDim.x * blockDim.x) {
        const int rk = rank[i];
        const double gain = 1.0 / log2((double)rk + 1.0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < nk && rk <= ks[q]) { hr[q] += 1.0; nd[q] += gain; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            hr[q] += __shfl_xor_sync(0xffffffffu, hr[q], o);
            nd[q] += __shfl_xor_sync(0xffffffffu, nd[q], o);
        }
        if (lane == 0) { red[w][q] = hr[q]; red[w][8 + q] = nd[q]; }
    }
    __syncthreads();
    if (threadIdx.x < 16) {
        double t = 0.0;
        for (int ww = 0; ww < 8; ++ww) t += red[ww][threadIdx.x];
        ws->dpartial[blockIdx.x * 16 + threadIdx.x] = t;
    }
    if (last_block_arrives(&ws->ticket[2], &flag)) {
        if (threadIdx.x < 16) {
            double t = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) t += __ldcg(&ws->dpartial[b * 16 + threadIdx.x]);
            const int q = threadIdx.x & 7;
            if (q < nk) out[(threadIdx.x >> 3) * nk + q] = t / (double)R;
        }
    }
}

template <int D, bool TOPK>
static int launch_eval(EvalParams &p, cudaStream_t st, int row_tiles) {
    constexpr int LD = D + 4;
    size_t smem = (size_t)(EV_TR * LD + 2 * EV_TI * LD) * 4 + 2 * EV_TR * 2 * 4 + EV_TR * 4;
    if (TOPK) smem += (size_t)EV_TR * (EV_TI + 1) * 4 + (size_t)EV_TR * EV_KMAX * 8;
    cudaError_t e = cudaFuncSetAttribute(eval_rank_kernel<D, TOPK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(row_tiles, p.splits);
    eval_rank_kernel<D, TOPK><<<grid, 256, smem, st>>>(p);
    e = cudaGetLastError();
    return (int)e;
}

}  // namespace wr

using namespace wr;

// eval_tcgen05.cu
int wr_eval_rank_tc(const float *Uemb, const float *Iemb, const int64_t *user, const int64_t *pos, int64_t R,
                    int64_t n_users, int64_t n_items, int D, const int64_t *hist_ptr, const int32_t *hist_idx,
                    int32_t *rank, float *target, const float *target_in, float *scores_out, int k, int32_t *topk_idx,
                    float *topk_val, void *scratch, WrWorkspace *ws, cudaStream_t st, int precision);

static int eval_rank_topk_impl(const float *Uemb, const float *Iemb, const int64_t *user, const int64_t *pos,
                               int64_t R, int64_t n_users, int64_t n_items, int D, const int64_t *hist_ptr,
                               const int32_t *hist_idx, int k, int precision, int32_t *topk_idx, float *topk_val,
                               int32_t *rank, float *target, const float *target_in, float *scores_out,
                               void *scratch, void *ws, void *stream) {
    if (!Uemb || !Iemb || !user || !pos || !hist_ptr || !hist_idx || !rank || !ws) return WR_E_NULL;
    if (!target && !target_in) return WR_E_NULL;
    if ((topk_idx == nullptr) != (topk_val == nullptr)) return WR_E_NULL;
    if (R <= 0 || n_users <= 0 || n_items <= 0 || n_items > INT32_MAX - 1024) return WR_E_SIZE;
    if (D <= 0 || (D & 3)) return WR_E_DIM;
    if (topk_idx && (k < 1 || k > EV_KMAX)) return WR_E_TOPK;
    if (!wr_aligned16(Uemb) || !wr_aligned16(Iemb)) return WR_E_ALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    const bool topk = topk_idx != nullptr;
    if (precision == 1 || precision == 2) {
        if (precision == 2 && topk) return WR_E_TOPK;
        return wr_eval_rank_tc(Uemb, Iemb, user, pos, R, n_users, n_items, D, hist_ptr, hist_idx, rank, target,
                               target_in, scores_out, topk ? k : 0, topk_idx, topk_val, scratch, (WrWorkspace *)ws, st,
                               precision);
    }
    if (precision != 0) return WR_E_PRECISION;
    EvalParams p{Uemb, Iemb, user, pos, R, n_users, n_items, hist_ptr, hist_idx, topk ? k : 1,
                 topk_idx, topk_val, rank, target, scores_out, 1, 0, 0, (WrWorkspace *)ws, target_in};
    p.n_tiles = (int)((n_items + EV_TI - 1) / EV_TI);
    const int64_t row_tiles64 = (R + EV_TR - 1) / EV_TR;
    if (row_tiles64 > INT32_MAX) return WR_E_SIZE;
    const int row_tiles = (int)row_tiles64;
    // few eval rows: split the item range over CTAs (rank counts merge with integer atomics -> still exact)
    int splits = 1;
    if (!topk && row_tiles < 2 * kSMs) splits = (2 * kSMs + row_tiles - 1) / row_tiles;
    if (splits > p.n_tiles) splits = p.n_tiles;
    if (splits > 65535) splits = 65535;
    p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
    p.splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    if (p.splits > 1) {
        fill_i32_kernel<<<(int)min((int64_t)kSMs * 4, (R + 255) / 256), 256, 0, st>>>(rank, R, 1);
        WR_CHECK_LAUNCH();
    }
    int rc;
#define WR_EVAL_CALL(DD) rc = topk ? launch_eval<DD, true>(p, st, row_tiles) : launch_eval<DD, false>(p, st, row_tiles)
    switch (D) {
        case 16: WR_EVAL_CALL(16); break;
        case 32: WR_EVAL_CALL(32); break;
        case 64: WR_EVAL_CALL(64); break;
        case 128: WR_EVAL_CALL(128); break;
        default: return WR_E_DIM;
    }
#undef WR_EVAL_CALL
    return rc;
}

extern "C" int wr_eval_rank_topk(const float *Uemb, const float *Iemb, const int64_t *user, const int64_t *pos,
                                 int64_t R, int64_t n_users, int64_t n_items, int D, const int64_t *hist_ptr,
                                 const int32_t *hist_idx, int k, int precision, int32_t *topk_idx, float *topk_val,
                                 int32_t *rank, float *target, float *scores_out, void *scratch, void *ws,
                                 void *stream) {
    if (!target) return WR_E_NULL;
    return eval_rank_topk_impl(Uemb, Iemb, user, pos, R, n_users, n_items, D, hist_ptr, hist_idx, k, precision,
                               topk_idx, topk_val, rank, target, nullptr, scores_out, scratch, ws, stream);
}

extern "C" int wr_eval_rank_topk_shard(const float *Urows, const float *Iemb, const int64_t *user,
                                       const int64_t *pos_local, int64_t R, int64_t n_users, int64_t n_items_local,
                                       int D, const int64_t *hist_ptr, const int32_t *hist_idx, int k, int precision,
                                       const float *target, int32_t *topk_idx, float *topk_val, int32_t *rank,
                                       void *scratch, void *ws, void *stream) {
    if (!target) return WR_E_NULL;
    return eval_rank_topk_impl(Urows, Iemb, user, pos_local, R, n_users, n_items_local, D, hist_ptr, hist_idx, k,
                               precision, topk_idx, topk_val, rank, nullptr, target, nullptr, scratch, ws, stream);
}

extern "C" int wr_metrics(const int32_t *rank, int64_t R, const int *host_ks, int nk, double *out, void *ws,
                          void *stream) {
    if (!rank || !host_ks || !out || !ws) return WR_E_NULL;
    if (R <= 0 || nk < 1 || nk > 8) return WR_E_SIZE;
    int ks[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < nk; ++i) ks[i] = host_ks[i];
    int64_t g = (R + 255) / 256;
    if (g > 64) g = 64;
    metrics_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(rank, R, nk, ks[0], ks[1], ks[2], ks[3], ks[4], ks[5],
                                                              ks[6], ks[7], out, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    return WR_OK;
}
