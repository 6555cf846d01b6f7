// SGL (reference models/general/SGL.py:165-246): the pieces the LightGCN kernels do not cover.
//   * the sum-form BPR term  sum_b -logsigmoid(s+ - s-)  is wr_bpr_logsig_sum_fwd_bwd (train_kernels.cu);
//   * InfoNCE between the two augmented views (SGL.py:196-231), forward and backward, for one block of rows (the users
//     of the batch against all users, or the positive items against all items):
//         a_b = normalize(T1[idx_b]),  t_j = normalize(T2[j]),  e_bj = exp(a_b . t_j / tau),  z_b = sum_j e_bj
//         loss += w * sum_b (log z_b - a_b . t_{idx_b} / tau)
//     The [B, N] contraction runs as fp32 FMA tiles (the parity bar is 1e-5 relative to the reference's fp32 matmul; B x N
//     is 2048 x 6040 on ml-1m) in three passes over a materialised weight matrix W = e / z:
//         E = exp(A T^T / tau)  ->  z, W  ->  Ga = W T  and  Gt = W^T A
//     followed by the chain rule through the two normalisations into the pooled tables' gradients.
#include "common.cuh"

namespace wr {

// ---- out[r] = x / max(||x||, 1e-12) (F.normalize), inv[r] = 1 / max(||x||, 1e-12); x = T[idx ? idx[r] : r] ----
__global__ void __launch_bounds__(256) rows_normalize_kernel(const float *__restrict__ T, const int64_t *idx, int64_t n,
                                                              int64_t n_table, int D, float *out, float *inv, WrWorkspace *ws) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = warp; r < n; r += nwarps) {
        int64_t src = idx ? idx[r] : r;
        const bool ok = (uint64_t)src < (uint64_t)n_table;
        if (!ok) {
            if (lane == 0) atomicOr(&ws->status, WR_STATUS_INDEX_OUT_OF_RANGE);
            src = 0;
        }
        const float *x = T + src * D;
        float ss = 0.f;
        for (int d = lane; d < D; d += 32) ss = fmaf(x[d], x[d], ss);
        ss = warp_sum(ss);
        const float k = ok ? 1.0f / fmaxf(sqrtf(ss), 1e-12f) : 0.f;
        for (int d = lane; d < D; d += 32) out[r * D + d] = x[d] * k;
        if (lane == 0) inv[r] = k;
    }
}

// ---- C[M, N] = f(alpha * sum_k A(m, k) B(k, n)); A(m, k) = A[m * sam + k * sak], B(k, n) = B[k * sbk + n * sbn];
//      f = exp when EXP.  64 x 64 tile per CTA, 16-wide K slabs, 4 x 4 outputs per thread. ----
template <bool EXP>
__global__ void __launch_bounds__(256) sgemm_tile_kernel(int64_t M, int64_t N, int64_t K, const float *__restrict__ A, int64_t sam,
                                                          int64_t sak, const float *__restrict__ Bm, int64_t sbk, int64_t sbn,
                                                          float *C, int64_t ldc, float alpha) {
    __shared__ float sA[16][64 + 4], sB[16][64 + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * 64, n0 = (int64_t)blockIdx.x * 64;
    float acc[4][4] = {};
    for (int64_t k0 = 0; k0 < K; k0 += 16) {
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            // the index that is contiguous in memory runs fastest across threads
            const int kk = sak == 1 ? (t & 15) : (t >> 6), mm = sak == 1 ? (t >> 4) : (t & 63);
            const int64_t m = m0 + mm, k = k0 + kk;
            sA[kk][mm] = (m < M && k < K) ? A[m * sam + k * sak] : 0.f;
        }
        for (int t = threadIdx.x; t < 16 * 64; t += 256) {
            const int kk = sbk == 1 ? (t & 15) : (t >> 6), nn = sbk == 1 ? (t >> 4) : (t & 63);
            const int64_t n = n0 + nn, k = k0 + kk;
            sB[kk][nn] = (n < N && k < K) ? Bm[k * sbk + n * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t n = n0 + tx * 4 + j;
            if (n < N) C[m * ldc + n] = EXP ? expf(alpha * acc[i][j]) : alpha * acc[i][j];
        }
    }
}

// ---- z_b = sum_j E[b, j]; E[b, :] /= z_b; pos_b = a_b . t_{idx_b} / tau; part[b] = log z_b - pos_b ----
__global__ void __launch_bounds__(256) infonce_rows_kernel(float *E, int64_t Bn, int64_t N, const float *__restrict__ An,
                                                            const float *__restrict__ Tn, const int64_t *idx, int D, float inv_tau,
                                                            float *part) {
    __shared__ float red[8];
    for (int64_t b = blockIdx.x; b < Bn; b += gridDim.x) {
        float s = 0.f;
        for (int64_t j = threadIdx.x; j < N; j += blockDim.x) s += E[b * N + j];
        const float z0 = block_sum(s, red);
        __shared__ float zs, ps;
        float p = 0.f;
        const int64_t t = idx[b];
        for (int d = threadIdx.x; d < D; d += blockDim.x) p = fmaf(An[b * D + d], Tn[t * D + d], p);
        const float p0 = block_sum(p, red);
        if (threadIdx.x == 0) {
            zs = z0;
            ps = p0 * inv_tau;
            part[b] = logf(z0) - p0 * inv_tau;
        }
        __syncthreads();
        const float iz = 1.0f / zs;
        for (int64_t j = threadIdx.x; j < N; j += blockDim.x) E[b * N + j] *= iz;
        __syncthreads();
    }
}

// deterministic sum of part[0 .. n) -> loss_out[0] += w * sum   (one CTA)
__global__ void __launch_bounds__(256) scaled_sum_kernel(const float *__restrict__ part, int64_t n, float w, float *loss_out) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
    const float t = block_sum(s, red);
    if (threadIdx.x == 0) loss_out[0] += w * t;
}

// ---- batch side: g = c (Ga_b - t_{idx_b}) w.r.t. the normalised a_b, through the normalisation into dT1[idx_b] (RED);
//      the positive pair's pull on t_{idx_b}: Gt[idx_b] -= a_b (RED, before infonce_table_kernel runs) ----
__global__ void __launch_bounds__(256) infonce_batch_kernel(const float *__restrict__ Ga, const float *__restrict__ An,
                                                             const float *__restrict__ inv1, const float *__restrict__ Tn,
                                                             const int64_t *idx, int64_t Bn, int D, float c, float *dT1,
                                                             float *Gt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t b = warp; b < Bn; b += nwarps) {
        const int64_t t = idx[b];
        float dot = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float g = Ga[b * D + d] - Tn[t * D + d];
            dot = fmaf(g, An[b * D + d], dot);
        }
        dot = warp_sum(dot);
        const float k = c * inv1[b];
        for (int d = lane; d < D; d += 32) {
            const float a = An[b * D + d];
            const float g = Ga[b * D + d] - Tn[t * D + d];
            atomicAdd(dT1 + t * D + d, k * (g - dot * a));
            atomicAdd(Gt + t * D + d, -a);
        }
    }
}

// ---- table side: g = c Gt_j w.r.t. the normalised t_j, through the normalisation, added to dT2[j] ----
__global__ void __launch_bounds__(256) infonce_table_kernel(const float *__restrict__ Gt, const float *__restrict__ Tn,
                                                             const float *__restrict__ inv2, int64_t N, int D, float c, float *dT2) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t j = warp; j < N; j += nwarps) {
        float dot = 0.f;
        for (int d = lane; d < D; d += 32) dot = fmaf(Gt[j * D + d], Tn[j * D + d], dot);
        dot = warp_sum(dot);
        const float k = c * inv2[j];
        for (int d = lane; d < D; d += 32) dT2[j * D + d] += k * (Gt[j * D + d] - dot * Tn[j * D + d]);
    }
}

static inline size_t sgl_align(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace wr

using namespace wr;

// rows of the batch processed per pass: the weight matrix W [rows, N] is materialised
static int64_t infonce_chunk(int64_t B, int64_t N) {
    int64_t c = ((int64_t)512 << 20) / (4 * (N > 0 ? N : 1));      // <= 512 MB of W
    c = c / 64 * 64;
    if (c < 64) c = 64;
    return c < B ? c : B;
}

extern "C" size_t wr_infonce_scratch_bytes(int64_t B, int64_t N, int D) {
    if (B <= 0 || N <= 0 || D <= 0) return 0;
    const int64_t bc = infonce_chunk(B, N);
    return sgl_align((size_t)N * D * 4) * 2 /* Tn, Gt */ + sgl_align((size_t)N * 4) /* inv2 */ +
           sgl_align((size_t)B * D * 4) * 2 /* An, Ga */ + sgl_align((size_t)B * 4) * 2 /* inv1, part */ +
           sgl_align((size_t)bc * N * 4) /* W */ + 256;
}

extern "C" int wr_infonce_fwd_bwd(const float *T1, const float *T2, const int64_t *idx, int64_t B, int64_t N, int D,
                                  float tau, float weight, float grad_scale, float *dT1, float *dT2, float *loss_out,
                                  void *scratch, size_t scratch_bytes, void *ws, void *stream) {
    if (!T1 || !T2 || !idx || !dT1 || !dT2 || !loss_out || !scratch || !ws) return WR_E_NULL;
    if (B <= 0 || N <= 0 || tau <= 0.f) return WR_E_SIZE;
    if (D <= 0 || (D & 3)) return WR_E_DIM;
    if (scratch_bytes < wr_infonce_scratch_bytes(B, N, D)) return WR_E_SIZE;
    cudaStream_t st = (cudaStream_t)stream;
    char *p = (char *)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    auto take = [&](size_t bytes) { char *q = p; p += sgl_align(bytes); return q; };
    float *Tn = (float *)take((size_t)N * D * 4), *Gt = (float *)take((size_t)N * D * 4), *inv2 = (float *)take((size_t)N * 4);
    float *An = (float *)take((size_t)B * D * 4), *Ga = (float *)take((size_t)B * D * 4);
    float *inv1 = (float *)take((size_t)B * 4), *part = (float *)take((size_t)B * 4);
    const int64_t bc = infonce_chunk(B, N);
    float *W = (float *)take((size_t)bc * N * 4);
    const float inv_tau = 1.0f / tau;
    const int gw = 8 * kSMs;
    rows_normalize_kernel<<<(int)min((int64_t)gw, (N + 7) / 8), 256, 0, st>>>(T2, nullptr, N, N, D, Tn, inv2, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    rows_normalize_kernel<<<(int)min((int64_t)gw, (B + 7) / 8), 256, 0, st>>>(T1, idx, B, N, D, An, inv1, (WrWorkspace *)ws);
    WR_CHECK_LAUNCH();
    cudaError_t e = cudaMemsetAsync(Gt, 0, (size_t)N * D * 4, st);
    if (e != cudaSuccess) return (int)e;
    for (int64_t b0 = 0; b0 < B; b0 += bc) {
        const int64_t bn = B - b0 < bc ? B - b0 : bc;
        // E = exp(An Tn^T / tau)                                                   [bn, N]
        sgemm_tile_kernel<true><<<dim3((unsigned)((N + 63) / 64), (unsigned)((bn + 63) / 64)), 256, 0, st>>>(
            bn, N, D, An + b0 * D, D, 1, Tn, 1, D, W, N, inv_tau);
        WR_CHECK_LAUNCH();
        infonce_rows_kernel<<<(int)min((int64_t)4 * kSMs, bn), 256, 0, st>>>(W, bn, N, An + b0 * D, Tn, idx + b0, D, inv_tau, part + b0);
        WR_CHECK_LAUNCH();
        // Ga = W Tn                                                                [bn, D]
        sgemm_tile_kernel<false><<<dim3((unsigned)((D + 63) / 64), (unsigned)((bn + 63) / 64)), 256, 0, st>>>(
            bn, D, N, W, N, 1, Tn, D, 1, Ga + b0 * D, D, 1.0f);
        WR_CHECK_LAUNCH();
        // Gt += W^T An: accumulated over the chunks through a second buffer only when there are several chunks
        if (bc >= B) {
            sgemm_tile_kernel<false><<<dim3((unsigned)((D + 63) / 64), (unsigned)((N + 63) / 64)), 256, 0, st>>>(
                N, D, bn, W, 1, N, An + b0 * D, D, 1, Gt, D, 1.0f);
            WR_CHECK_LAUNCH();
        } else {
            return WR_E_SIZE;       // B x N beyond one 512 MB pass: not needed at the reference's scales
        }
    }
    scaled_sum_kernel<<<1, 256, 0, st>>>(part, B, weight, loss_out);
    WR_CHECK_LAUNCH();
    const float c = weight * inv_tau * grad_scale;
    infonce_batch_kernel<<<(int)min((int64_t)gw, (B + 7) / 8), 256, 0, st>>>(Ga, An, inv1, Tn, idx, B, D, c, dT1, Gt);
    WR_CHECK_LAUNCH();
    infonce_table_kernel<<<(int)min((int64_t)gw, (N + 7) / 8), 256, 0, st>>>(Gt, Tn, inv2, N, D, c, dT2);
    WR_CHECK_LAUNCH();
    return WR_OK;
}
