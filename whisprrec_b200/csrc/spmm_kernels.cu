// LightGCN propagation: CSR SpMM over the symmetric-normalised bipartite adjacency with the layer-mean
// (and, in the backward direction, the pooled-gradient addend) fused into the epilogue.
#include "common.cuh"

namespace wr {

// where the rows of X live: this GPU, or the row shards of all GPUs (the layer output is never all-gathered: a
// neighbour's row is read from its owner over NVLink inside the SpMM)
struct LocalX {
    const float *X;
    __device__ __forceinline__ const float *row(int c, int D) const { return X + (int64_t)c * D; }
};
struct ShardX {
    wr_shards s;
    __device__ __forceinline__ const float *row(int c, int D) const { return shard_node_row(s, c, D); }
};

template <class XACC>
struct SpmmParamsT {
    const int64_t *rowptr;
    const int32_t *col;
    const float *val;
    int64_t N;
    XACC X;
    float *Y;
    float *add;
    int zero_add;
    const float *acc_in;
    float *acc_out;
    float acc_div;
    wr_spmm_plan plan;       // by value; long_threshold = INT64_MAX when there is no plan
    int row_blocks;          // CTAs [0, row_blocks) walk rows, the next chunk_blocks walk the chunks of split rows
    int chunk_blocks;
    // copy-engine push (wr_csr_spmm_sharded_dma): blocks of rows have to complete in a steady stream from the start of the
    // kernel, so the two kinds of CTA are interleaved in launch order in proportion to their work (both then sweep the
    // row range for the whole duration) and the row walk starts at row_rot (the item rows: their split rows are the last
    // the chunk walk reaches, so their short rows go first)
    int interleave;
    int64_t row_rot;
    // fused all-gather of the output (multi-GPU): every finished row of Y is also stored into each peer's copy of this
    // rank's shard (posted NVLink writes behind the arithmetic), so the next layer starts without a gather pass
    float *push[WR_MAX_WORLD];
    int n_push;
    // progress words for a copy-engine push beside the kernel (wr_csr_spmm_sharded_dma): rows are counted per block of
    // 2^prog_shift rows; whoever finishes a block's last row publishes prog_flag[block] = prog_epoch
    int32_t *prog_count;
    uint32_t *prog_flag;
    int prog_shift;
    uint32_t prog_epoch;
};
using SpmmParams = SpmmParamsT<LocalX>;

// acc += sum_{e in [beg, end)} val[e] * X[col[e]] for the lane's slice of the row (partial over lane groups).
template <int LPR, int VPL, int UNROLL, class XACC>
__device__ __forceinline__ void spmm_accumulate(const SpmmParamsT<XACC> &p, int64_t beg, int64_t end, int lane, int sub,
                                                int grp, float4 (&acc)[VPL]) {
    using RG = RowGroup<LPR, VPL>;
    constexpr int D = RG::D;
    constexpr int EPS = RG::GROUPS;  // edges per step
    // Tables far larger than L2 (power-law graphs at scale): the rows of the highest-degree nodes are re-read thousands
    // of times per pass, the rest a few dozen times with reuse distances of gigabytes.  plan.hot_bits marks the former;
    // they are loaded with an L2 evict_last policy (they stay), the index / weight streams with evict_first.
    const uint32_t *x_rows = p.plan.x_rows;
    const uint32_t *hot_bits = x_rows ? nullptr : p.plan.hot_bits;
    uint64_t pol_keep = 0, pol_stream = 0;
    if (hot_bits) {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    }
    for (int64_t base = beg; base < end; base += 32) {
        const int cnt = (int)min((int64_t)32, end - base);
        int c = 0;
        float w = 0.f;
        if (lane < cnt) {
            if (hot_bits) {
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(c) : "l"(p.col + base + lane), "l"(pol_stream));
                asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(w) : "l"(p.val + base + lane), "l"(pol_stream));
                if ((__ldg(hot_bits + (c >> 5)) >> (c & 31)) & 1u) c |= (int)0x80000000u;      // carried in the sign bit
            } else {
                c = __ldg(p.col + base + lane);
                w = __ldg(p.val + base + lane);
                // sparse X (the pooled gradient of a batch): a row whose bit is clear is zero -- it is not fetched
                if (x_rows && !((__ldg(x_rows + (c >> 5)) >> (c & 31)) & 1u)) c = -1;
            }
        }
        if (x_rows && !__any_sync(0xffffffffu, lane < cnt && c >= 0)) continue;
        for (int j = 0; j < cnt; j += EPS * UNROLL) {
            float4 x[UNROLL][VPL];
            float ww[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int e = j + u * EPS + grp;
                int cc = __shfl_sync(0xffffffffu, c, e & 31);
                ww[u] = __shfl_sync(0xffffffffu, w, e & 31);
                if (e < cnt) {
                    if (hot_bits) {
                        const bool hot = cc < 0;
                        cc &= 0x7fffffff;
                        const float *row = p.X.row(cc, D);
                        if (hot) {
#pragma unroll
                            for (int v = 0; v < VPL; ++v)
                                asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                                             : "=f"(x[u][v].x), "=f"(x[u][v].y), "=f"(x[u][v].z), "=f"(x[u][v].w)
                                             : "l"(row + 4 * (sub + v * LPR)), "l"(pol_keep));
                        } else {
                            RG::load(row, sub, x[u]);
                        }
                    } else if (cc >= 0) {
                        RG::load(p.X.row(cc, D), sub, x[u]);
                    } else {
                        RG::zero(x[u]);
                    }
                } else {
                    RG::zero(x[u]);
                    ww[u] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
#pragma unroll
                for (int v = 0; v < VPL; ++v) acc[v] = fma4(ww[u], x[u][v], acc[v]);
        }
    }
    // fold the 32/LPR partial rows held by the lane groups
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o);
            acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
            acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o);
            acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
        }
}

// y (+ add) -> Y, running layer sum / mean.  Called by the lanes of group 0 with their slice of the finished row.
template <class XACC>
__device__ __forceinline__ void spmm_row_epilogue(const SpmmParamsT<XACC> &p, int64_t off, float4 y) {
    if (p.add) {
        float4 *ap = reinterpret_cast<float4 *>(p.add + off);
        y = add4(y, *ap);
        if (p.zero_add) *ap = f4_zero();
    }
    if (p.Y) {
        *reinterpret_cast<float4 *>(p.Y + off) = y;
        for (int g = 0; g < p.n_push; ++g) *reinterpret_cast<float4 *>(p.push[g] + off) = y;
    }
    if (p.acc_out) {
        const float4 a = add4(*reinterpret_cast<const float4 *>(p.acc_in + off), y);
        // ATen's mean is sum().div_(count): a true division, not a multiply by the reciprocal
        *reinterpret_cast<float4 *>(p.acc_out + off) =
            p.acc_div == 1.0f ? a : make_float4(a.x / p.acc_div, a.y / p.acc_div, a.z / p.acc_div, a.w / p.acc_div);
    }
}

// Lane 0, after __syncwarp(): row `row` of Y is complete (all lanes' stores happen-before through the warp barrier).
template <class XACC>
__device__ __forceinline__ void spmm_row_done(const SpmmParamsT<XACC> &p, int64_t row) {
    const int64_t k = row >> p.prog_shift;
    const int64_t first = k << p.prog_shift;
    const int rows_here = (int)min((int64_t)1 << p.prog_shift, p.N - first);
    __threadfence();
    if (atomicAdd(p.prog_count + k, 1) == rows_here - 1) {
        p.prog_count[k] = 0;                   // ready for the next call
        __threadfence_system();                // the block's rows, observed through the counter, before the flag
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.prog_flag + k), "r"(p.prog_epoch) : "memory");
    }
}

// A warp owns a row (or, for rows longer than plan.long_threshold, one <= chunk-sized slice of it).  The row of
// D = 4*LPR*VPL floats is covered by LPR lanes, so 32/LPR neighbour rows are fetched per step (one 128-bit load per
// lane each), UNROLL steps in flight.  Column ids / weights are read 32 at a time, coalesced, and handed round
// with shuffles.  Slices of a split row are reduced into plan.slot_partial with 128-bit REDs; the warp that
// arrives last owns the row's epilogue, so one launch covers everything and power-law rows cannot become the tail.
template <int LPR, int VPL, int UNROLL, class XACC>
__global__ void __launch_bounds__(256) csr_spmm_kernel(SpmmParamsT<XACC> p) {
    using RG = RowGroup<LPR, VPL>;
    constexpr int D = RG::D;
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    bool is_chunk = (int)blockIdx.x >= p.row_blocks;
    int kind_idx = is_chunk ? (int)blockIdx.x - p.row_blocks : (int)blockIdx.x;
    if (p.interleave) {
        const int64_t total = p.row_blocks + p.chunk_blocks;
        const int c0 = (int)((int64_t)blockIdx.x * p.chunk_blocks / total);
        const int c1 = (int)(((int64_t)blockIdx.x + 1) * p.chunk_blocks / total);
        is_chunk = c1 > c0;
        kind_idx = is_chunk ? c0 : (int)blockIdx.x - c0;
    }
    if (!is_chunk) {
        const int64_t warp = (int64_t)kind_idx * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int64_t nwarps = (int64_t)p.row_blocks * (blockDim.x >> 5);
        for (int64_t it = warp; it < p.N; it += nwarps) {
            int64_t row = it + p.row_rot;
            if (row >= p.N) row -= p.N;
            const int64_t beg = __ldg(p.rowptr + row), end = __ldg(p.rowptr + row + 1);
            if (end - beg > p.plan.long_threshold) continue;      // split rows are handled below
            float4 acc[VPL];
            RG::zero(acc);
            spmm_accumulate<LPR, VPL, UNROLL, XACC>(p, beg, end, lane, sub, grp, acc);
            if (grp == 0) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) spmm_row_epilogue(p, row * D + 4 * (sub + v * LPR), acc[v]);
            }
            if (p.prog_count) {
                __syncwarp();
                if (lane == 0) spmm_row_done(p, row);
            }
        }
    } else {
        const int64_t warp = (int64_t)kind_idx * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const int64_t nwarps = (int64_t)p.chunk_blocks * (blockDim.x >> 5);
        for (int64_t ch = warp; ch < p.plan.n_chunks; ch += nwarps) {
            const int64_t beg = __ldg(p.plan.chunk_beg + ch);
            const int64_t end = beg + __ldg(p.plan.chunk_len + ch);
            const int slot = __ldg(p.plan.chunk_slot + ch);
            float4 acc[VPL];
            RG::zero(acc);
            spmm_accumulate<LPR, VPL, UNROLL, XACC>(p, beg, end, lane, sub, grp, acc);
            float *part = p.plan.slot_partial + (int64_t)slot * D;
            if (grp == 0) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) red_add_v4(part + 4 * (sub + v * LPR), acc[v]);
            }
            // last slice to arrive finishes the row
            __threadfence();
            __syncwarp();
            int prev = 0;
            if (lane == 0) prev = atomicAdd(p.plan.slot_arrivals + slot, 1);
            prev = __shfl_sync(0xffffffffu, prev, 0);
            if (prev == __ldg(p.plan.slot_chunks + slot) - 1) {
                __threadfence();
                const int64_t row = __ldg(p.plan.chunk_row + ch);
                if (grp == 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        float4 *pp = reinterpret_cast<float4 *>(part + 4 * (sub + v * LPR));
                        const float4 y = __ldcg(pp);
                        __stcg(pp, f4_zero());                   // scratch and counter are left ready for reuse
                        spmm_row_epilogue(p, row * D + 4 * (sub + v * LPR), y);
                    }
                }
                if (lane == 0) p.plan.slot_arrivals[slot] = 0;
                if (p.prog_count) {
                    __syncwarp();
                    if (lane == 0) spmm_row_done(p, row);
                }
            }
        }
    }
}

// Any D % 4 == 0.
template <class XACC>
__global__ void __launch_bounds__(256) csr_spmm_generic_kernel(SpmmParamsT<XACC> p, int D) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    const int D4 = D >> 2;
    for (int64_t row = warp; row < p.N; row += nwarps) {
        const int64_t beg = p.rowptr[row], end = p.rowptr[row + 1];
        for (int v = lane; v < D4; v += 32) {
            float4 acc = f4_zero();
            for (int64_t e = beg; e < end; ++e) {
                const int c = __ldg(p.col + e);
                if (p.plan.x_rows && !((__ldg(p.plan.x_rows + (c >> 5)) >> (c & 31)) & 1u)) continue;
                acc = fma4(__ldg(p.val + e), ldg4(p.X.row(c, D) + 4 * v), acc);
            }
            spmm_row_epilogue(p, row * D + 4 * v, acc);
        }
    }
}

__global__ void __launch_bounds__(256) csr_norm_weights_kernel(const int64_t *__restrict__ rowptr,
                                                                const int32_t *__restrict__ col,
                                                                const float *__restrict__ dinv, int64_t N,
                                                                float *__restrict__ val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = warp; row < N; row += nwarps) {
        const int64_t beg = rowptr[row], end = rowptr[row + 1];
        // LightGCN.py:95-97: (D^-1/2 A) D^-1/2 on a binary A: two roundings, row factor first
        const float dr = __fmul_rn(dinv[row], 1.0f);
        for (int64_t e = beg + lane; e < end; e += 32) val[e] = __fmul_rn(dr, dinv[col[e]]);
    }
}

}  // namespace wr

using namespace wr;

// CTAs of csr_spmm_kernel<...> that are resident at once on the current device (cached per instantiation).
template <int LPR, int VPL, int UNROLL, class XACC>
static int spmm_resident_ctas() {
    static int cached = 0;
    if (!cached) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, csr_spmm_kernel<LPR, VPL, UNROLL, XACC>, 256, 0) != cudaSuccess ||
            per_sm < 1)
            per_sm = 4;
        cached = per_sm * kSMs;
    }
    return cached;
}

template <class XACC>
static int spmm_resident_for(int D) {
    switch (D) {
        case 16: return spmm_resident_ctas<4, 1, 2, XACC>();
        case 32: return spmm_resident_ctas<8, 1, 2, XACC>();
        case 64: return spmm_resident_ctas<16, 1, 8, XACC>();
        case 128: return spmm_resident_ctas<32, 1, 4, XACC>();
        default: return spmm_resident_ctas<32, 2, 2, XACC>();
    }
}

template <class XACC>
static int spmm_launch(SpmmParamsT<XACC> &p, int D, const wr_spmm_plan *host_plan, cudaStream_t st, int64_t nnz_hint = 0) {
    p.plan = wr_spmm_plan{};
    p.plan.long_threshold = INT64_MAX;
    const bool fast = D == 16 || D == 32 || D == 64 || D == 128 || D == 256;
    if (host_plan && fast && host_plan->n_chunks > 0) {
        const wr_spmm_plan &q = *host_plan;
        if (!q.chunk_row || !q.chunk_beg || !q.chunk_len || !q.chunk_slot || !q.slot_chunks || !q.slot_arrivals ||
            !q.slot_partial)
            return WR_E_NULL;
        if (q.long_threshold < 1 || q.n_long < 1 || !wr_aligned16(q.slot_partial)) return WR_E_SIZE;
        p.plan = q;
    }
    if (host_plan && fast) p.plan.hot_bits = host_plan->hot_bits;
    if (host_plan) p.plan.x_rows = host_plan->x_rows;
    int64_t g = (p.N + 7) / 8;
    if (g > 16 * kSMs) g = 16 * kSMs;
    p.row_blocks = (int)g;
    int64_t gc = (p.plan.n_chunks + 7) / 8;
    if (gc > 16 * kSMs) gc = 16 * kSMs;
    if (p.interleave) {
        // Exactly the CTAs that are resident at once (a second wave would sweep the row range again after the first has
        // finished it: no block of rows would be complete before the end), split in proportion to the edges each kind
        // walks (slices average ~7/8 of the 128-edge limit)
        const int64_t total = spmm_resident_for<XACC>(D);
        double share = p.plan.n_chunks > 0 ? (nnz_hint > 0 ? (double)p.plan.n_chunks * 112.0 / (double)nnz_hint : 0.5) : 0.0;
        share = share < 0.02 ? 0.02 : (share > 0.9 ? 0.9 : share);
        int64_t c = (int64_t)(share * total + 0.5);
        if (c > (p.plan.n_chunks + 7) / 8) c = (p.plan.n_chunks + 7) / 8;
        if (c < 1 && p.plan.n_chunks > 0) c = 1;
        int64_t r = total - c;
        if (r > (p.N + 7) / 8) r = (p.N + 7) / 8;
        if (r < 1) r = 1;
        g = r;
        gc = c;
    }
    p.row_blocks = (int)g;
    p.chunk_blocks = (int)gc;
    const int grid = (int)(g + gc);
    switch (D) {
        case 16: csr_spmm_kernel<4, 1, 2, XACC><<<grid, 256, 0, st>>>(p); break;
        case 32: csr_spmm_kernel<8, 1, 2, XACC><<<grid, 256, 0, st>>>(p); break;
        case 64: csr_spmm_kernel<16, 1, 8, XACC><<<grid, 256, 0, st>>>(p); break;
        case 128: csr_spmm_kernel<32, 1, 4, XACC><<<grid, 256, 0, st>>>(p); break;
        case 256: csr_spmm_kernel<32, 2, 2, XACC><<<grid, 256, 0, st>>>(p); break;
        default: csr_spmm_generic_kernel<XACC><<<p.row_blocks, 256, 0, st>>>(p, D);
    }
    WR_CHECK_LAUNCH();
    return WR_OK;
}

static int spmm_check(const int64_t *rowptr, const int32_t *col, const float *val, int64_t N, int D, float *Y,
                      float *add, const float *acc_in, float *acc_out) {
    if (!rowptr || !col || !val) return WR_E_NULL;
    if (!Y && !acc_out) return WR_E_NULL;
    if (acc_out && !acc_in) return WR_E_NULL;
    if (N <= 0) return WR_E_SIZE;
    if (D <= 0 || (D & 3)) return WR_E_DIM;
    if ((Y && !wr_aligned16(Y)) || (add && !wr_aligned16(add)) || (acc_in && !wr_aligned16(acc_in)) ||
        (acc_out && !wr_aligned16(acc_out)))
        return WR_E_ALIGN;
    return WR_OK;
}

extern "C" int wr_csr_spmm(const int64_t *rowptr, const int32_t *col, const float *val, int64_t N, int D,
                           const float *X, float *Y, float *add, int zero_add, const float *acc_in,
                           float *acc_out, float acc_div, const wr_spmm_plan *host_plan, void *stream) {
    if (!X) return WR_E_NULL;
    const int rc = spmm_check(rowptr, col, val, N, D, Y, add, acc_in, acc_out);
    if (rc) return rc;
    if (!wr_aligned16(X)) return WR_E_ALIGN;
    if (X == Y || X == acc_out) return WR_E_SIZE;
    SpmmParams p{rowptr, col, val, N, {X}, Y, add, zero_add, acc_in, acc_out, acc_div, {}, 0, 0, 0, 0, {}, 0, nullptr, nullptr, 0, 0};
    return spmm_launch(p, D, host_plan, (cudaStream_t)stream);
}

extern "C" int wr_csr_spmm_sharded(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n_local, int D,
                                   const wr_shards *host_X, float *Y, float *add, int zero_add, const float *acc_in,
                                   float *acc_out, float acc_div, const wr_spmm_plan *host_plan,
                                   float *const host_push[WR_MAX_WORLD], void *stream) {
    if (!host_X) return WR_E_NULL;
    int rc = wr_check_shards(host_X);
    if (rc) return rc;
    rc = spmm_check(rowptr, col, val, n_local, D, Y, add, acc_in, acc_out);
    if (rc) return rc;
    if (n_local != host_X->rows_u_local + host_X->rows_i_local) return WR_E_SIZE;
    if (host_X->base[host_X->rank] == Y || host_X->base[host_X->rank] == acc_out) return WR_E_SIZE;
    SpmmParamsT<ShardX> p{rowptr, col, val, n_local, {*host_X}, Y, add, zero_add, acc_in, acc_out, acc_div, {}, 0, 0, 0, 0, {}, 0,
                          nullptr, nullptr, 0, 0};
    if (host_push) {
        if (!Y) return WR_E_NULL;
        for (int g = 0; g < host_X->world; ++g) {
            if (g == host_X->rank || !host_push[g]) continue;
            if (!wr_aligned16(host_push[g])) return WR_E_ALIGN;
            p.push[p.n_push++] = host_push[g];
        }
    }
    return spmm_launch(p, D, host_plan, (cudaStream_t)stream);
}

// cuStreamWaitValue32 through the runtime's driver entry point table (the library does not link libcuda).
typedef int (*wr_wait_value32_fn)(void *stream, unsigned long long addr, uint32_t value, unsigned int flags);
static wr_wait_value32_fn wait_value32() {
    static wr_wait_value32_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (wr_wait_value32_fn)sym;
    }
    return fn;
}

extern "C" int wr_csr_spmm_sharded_dma(const int64_t *rowptr, const int32_t *col, const float *val, int64_t n_local, int D,
                                       const wr_shards *host_X, float *Y, float *add, int zero_add, const float *acc_in,
                                       float *acc_out, float acc_div, const wr_spmm_plan *host_plan,
                                       float *const host_push[WR_MAX_WORLD], int32_t *progress, int64_t progress_words,
                                       int64_t block_rows, uint32_t epoch, int64_t nnz, void *side_stream, void *stream) {
    if (!host_X || !host_push || !progress || !Y || !side_stream) return WR_E_NULL;
    int rc = wr_check_shards(host_X);
    if (rc) return rc;
    rc = spmm_check(rowptr, col, val, n_local, D, Y, add, acc_in, acc_out);
    if (rc) return rc;
    if (!(D == 16 || D == 32 || D == 64 || D == 128 || D == 256)) return WR_E_DIM;
    if (n_local != host_X->rows_u_local + host_X->rows_i_local) return WR_E_SIZE;
    if (host_X->base[host_X->rank] == Y || host_X->base[host_X->rank] == acc_out) return WR_E_SIZE;
    if (block_rows < 32 || (block_rows & (block_rows - 1)) || epoch == 0) return WR_E_SIZE;
    const int64_t nblk = (n_local + block_rows - 1) / block_rows;
    if (progress_words < 2 * nblk) return WR_E_SIZE;
    wr_wait_value32_fn wait = wait_value32();
    if (!wait) return (int)cudaErrorNotSupported;
    SpmmParamsT<ShardX> p{rowptr, col, val, n_local, {*host_X}, Y, add, zero_add, acc_in, acc_out, acc_div, {}, 0, 0, 1, host_X->rows_u_local, {}, 0,
                          progress, (uint32_t *)progress + nblk, 0, epoch};
    while (((int64_t)1 << p.prog_shift) < block_rows) ++p.prog_shift;
    cudaStream_t st = (cudaStream_t)stream, sd = (cudaStream_t)side_stream;
    rc = spmm_launch(p, D, host_plan, st, nnz);
    if (rc) return rc;
    const int world = host_X->world, rank = host_X->rank;
    for (int64_t k = 0; k < nblk; ++k) {
        // CU_STREAM_WAIT_VALUE_GEQ = 0: (int32_t)(*addr - value) >= 0
        const int wrc = wait(sd, (unsigned long long)(uintptr_t)(p.prog_flag + k), epoch, 0u);
        if (wrc) return wrc;
        const int64_t first = k * block_rows;
        const size_t bytes = (size_t)(min(block_rows, n_local - first) * D) * sizeof(float);
        for (int r0 = 1; r0 < world; ++r0) {
            const int g = (rank + r0) % world;               // peers in a rotated order: links evenly loaded
            if (!host_push[g]) continue;
            const cudaError_t e = cudaMemcpyAsync(host_push[g] + first * D, Y + first * D, bytes, cudaMemcpyDefault, sd);
            if (e != cudaSuccess) return (int)e;
        }
    }
    cudaEvent_t ev;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev, sd);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ev, 0);
    cudaEventDestroy(ev);
    return e == cudaSuccess ? WR_OK : (int)e;
}

// All-gather of one shard by the copy engines: this rank's rows -> every peer's copy (no SM work, ~20 % faster than the
// pull kernel on 8 B200s).  The caller's barrier after it makes the copies visible to their readers.
extern "C" int wr_push_shard_dma(const float *src, int64_t n_floats, int world, int rank,
                                 float *const host_push[WR_MAX_WORLD], void *stream) {
    if (!src || !host_push) return WR_E_NULL;
    if (world < 1 || world > WR_MAX_WORLD || rank < 0 || rank >= world || n_floats <= 0) return WR_E_SIZE;
    for (int r0 = 1; r0 < world; ++r0) {
        const int g = (rank + r0) % world;
        if (!host_push[g]) continue;
        const cudaError_t e = cudaMemcpyAsync(host_push[g], src, (size_t)n_floats * sizeof(float), cudaMemcpyDefault,
                                              (cudaStream_t)stream);
        if (e != cudaSuccess) return (int)e;
    }
    return WR_OK;
}

extern "C" int wr_csr_norm_weights(const int64_t *rowptr, const int32_t *col, const float *dinv, int64_t N,
                                   float *val, void *stream) {
    if (!rowptr || !col || !dinv || !val) return WR_E_NULL;
    if (N <= 0) return WR_E_SIZE;
    int64_t g = (N + 7) / 8;
    if (g > 16 * kSMs) g = 16 * kSMs;
    csr_norm_weights_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(rowptr, col, dinv, N, val);
    WR_CHECK_LAUNCH();
    return WR_OK;
}
