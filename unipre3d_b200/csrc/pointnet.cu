// pointnet.cu -- the per-group mini-PointNet of the tokenizer (openpoints/models/backbone/transformer.py:210-243
// `Encoder`) around its four dense GEMMs:
//
//     x (B,G,K,3) -Conv1d(3,128)-> BN -> ReLU -Conv1d(128,256)-> f ; fg = max_K f ; [fg || f] -Conv1d(512,512)-> BN -> ReLU
//       -Conv1d(512,C)-> max_K -> tokens (B,G,C)
//
// The dense 1x1 convolutions stay library GEMMs over R = B*G*K rows.  Everything else -- the K=3 first layer, both
// train-mode BatchNorms (statistics, normalise+ReLU, and their backward reductions), the per-group max-pools and their
// scatter backward, and the "global || local" concat (never materialised: W3 [fg || f] = W3g fg + W3l f, so the global
// half is a (B*G)-row GEMM whose result is broadcast over the K rows inside the BatchNorm passes) -- are the fused
// kernels below.  BatchNorm statistics are per-CTA partial sums reduced in fp64 by a finalize kernel (deterministic,
// no atomics); the conv biases in front of a train-mode BatchNorm have a mathematically zero gradient.
//
// Rows are (group, k): r = g*K + k, g < Gt = B*G.  Activations between GEMMs are AT = float or __nv_bfloat16.
#include <cuda_bf16.h>

#include <atomic>
#include <stdlib.h>

#include "common.cuh"

namespace up3d {

template <typename AT> struct PVec4;
template <> struct PVec4<float> {
    using Raw = float4;               // what a load leaves in registers until the value is used
    static __device__ __forceinline__ Raw load_raw(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ float4 cvt(Raw r) { return r; }
    static __device__ __forceinline__ float4 load(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void store(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
};
template <> struct PVec4<__nv_bfloat16> {
    using Raw = uint2;                // 4 packed bf16: batches of in-flight loads cost half the registers of float4
    static __device__ __forceinline__ Raw load_raw(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint2 *>(p); }
    static __device__ __forceinline__ float4 cvt(Raw r) {
        const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.x));
        const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&r.y));
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ float4 load(const __nv_bfloat16 *p) { return cvt(load_raw(p)); }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned *>(&a);
        r.y = *reinterpret_cast<const unsigned *>(&b);
        *reinterpret_cast<uint2 *>(p) = r;
    }
};

// ------------------------------------------------------------------------------------------ first layer (K = 3)
// nb is the tokenizer's (B,3,G,K) neighbourhood tensor: row r = b*GK + j reads nb[(b*3 + k)*GK + j].
constexpr int PN_TILE = 128;      // rows per CTA tile of the first-layer kernels
constexpr int PN_C1 = 128;        // channels of the first layer (thread = channel)

__device__ __forceinline__ void pn_load_x_tile(float (*xs)[PN_TILE], const float *__restrict__ nb, int R, int GK, int r0) {
    for (int i = threadIdx.x; i < 3 * PN_TILE; i += blockDim.x) {
        const int k = i / PN_TILE, rr = i - k * PN_TILE, r = r0 + rr;
        float v = 0.f;
        if (r < R) {
            const int b = r / GK, j = r - b * GK;
            v = nb[((size_t)b * 3 + k) * GK + j];
        }
        xs[k][rr] = v;
    }
}

// partials (gridDim.x, 3, C1): per-CTA [shift, sum (z - shift), sum (z - shift)^2] of z = W1 x + b1 over the CTA's ONE
// row tile (gridDim.x == ceil(R / PN_TILE)); shift = z of the tile's first row, so the sums carry no cancellation even
// when |mean| >> std (merged in fp64 by bn_reduce_finalize_kernel).
__global__ void __launch_bounds__(PN_C1)
pn_conv1_stats_kernel(int R, int GK, const float *__restrict__ nb, const float *__restrict__ W1, const float *__restrict__ b1,
                      float *__restrict__ partials) {
    __shared__ float xs[3][PN_TILE];
    const int c = threadIdx.x;
    const float w0 = W1[3 * c], w1 = W1[3 * c + 1], w2 = W1[3 * c + 2], bb = b1[c];
    const int r0 = blockIdx.x * PN_TILE;
    pn_load_x_tile(xs, nb, R, GK, r0);
    __syncthreads();
    const int n = min(PN_TILE, R - r0);
    const float shift = fmaf(w0, xs[0][0], fmaf(w1, xs[1][0], fmaf(w2, xs[2][0], bb)));
    float s = 0.f, q = 0.f;
    for (int rr = 0; rr < n; ++rr) {
        const float z = fmaf(w0, xs[0][rr], fmaf(w1, xs[1][rr], fmaf(w2, xs[2][rr], bb))) - shift;
        s += z;
        q = fmaf(z, z, q);
    }
    partials[((size_t)blockIdx.x * 3 + 0) * PN_C1 + c] = shift;
    partials[((size_t)blockIdx.x * 3 + 1) * PN_C1 + c] = s;
    partials[((size_t)blockIdx.x * 3 + 2) * PN_C1 + c] = q;
}

// y1[r, c] = relu(a[c] * (W1 x_r + b1)[c] + d[c]);  stats (4, C1): mean, rstd, a = gamma*rstd, d = beta - mean*a
template <typename AT>
__global__ void __launch_bounds__(256)
pn_conv1_bn_relu_kernel(int R, int GK, const float *__restrict__ nb, const float *__restrict__ W1, const float *__restrict__ b1,
                        const float *__restrict__ stats, AT *__restrict__ y1) {
    __shared__ float xs[3][PN_TILE];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c0 = 4 * lane;
    float w[4][3], bb[4], a[4], d[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        w[i][0] = W1[3 * (c0 + i)]; w[i][1] = W1[3 * (c0 + i) + 1]; w[i][2] = W1[3 * (c0 + i) + 2];
        bb[i] = b1[c0 + i];
        a[i] = stats[2 * PN_C1 + c0 + i];
        d[i] = stats[3 * PN_C1 + c0 + i];
    }
    for (int r0 = blockIdx.x * PN_TILE; r0 < R; r0 += gridDim.x * PN_TILE) {
        __syncthreads();
        pn_load_x_tile(xs, nb, R, GK, r0);
        __syncthreads();
        const int n = min(PN_TILE, R - r0);
        for (int rr = warp; rr < n; rr += 8) {
            const float x0 = xs[0][rr], x1 = xs[1][rr], x2 = xs[2][rr];
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float z = fmaf(w[i][0], x0, fmaf(w[i][1], x1, fmaf(w[i][2], x2, bb[i])));
                o[i] = fmaxf(fmaf(a[i], z, d[i]), 0.f);
            }
            PVec4<AT>::store(y1 + (size_t)(r0 + rr) * PN_C1 + c0, make_float4(o[0], o[1], o[2], o[3]));
        }
    }
}

// backward of ReLU + BatchNorm + first layer.  PASS 0: partials (gridDim.x, 2, C1) of sum dy, sum dy*xhat with
// dy = dy1 * [a z + d > 0].  PASS 1: dz = a (dy - m1 - xhat m2) with m = sums / R; accumulates gW1 (C1,3) += dz x^T,
// gb1 (C1) += dz (float atomics, one per CTA and entry).
template <typename AT, int PASS>
__global__ void __launch_bounds__(PN_C1)
pn_conv1_bwd_kernel(int R, int GK, const float *__restrict__ nb, const float *__restrict__ W1, const float *__restrict__ b1,
                    const float *__restrict__ stats, const AT *__restrict__ dy1, const float *__restrict__ sums,
                    float inv_count, float *__restrict__ partials, float *__restrict__ gW1, float *__restrict__ gb1) {
    __shared__ float xs[3][PN_TILE];
    const int c = threadIdx.x;
    const float w0 = W1[3 * c], w1 = W1[3 * c + 1], w2 = W1[3 * c + 2], bb = b1[c];
    const float mean = stats[c], rstd = stats[PN_C1 + c], a = stats[2 * PN_C1 + c], d = stats[3 * PN_C1 + c];
    float m1 = 0.f, m2 = 0.f;
    if (PASS == 1) { m1 = sums[c] * inv_count; m2 = sums[PN_C1 + c] * inv_count; }
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    for (int r0 = blockIdx.x * PN_TILE; r0 < R; r0 += gridDim.x * PN_TILE) {
        __syncthreads();
        pn_load_x_tile(xs, nb, R, GK, r0);
        __syncthreads();
        const int n = min(PN_TILE, R - r0);
#pragma unroll 8
        for (int rr = 0; rr < n; ++rr) {
            const float x0 = xs[0][rr], x1 = xs[1][rr], x2 = xs[2][rr];
            const float z = fmaf(w0, x0, fmaf(w1, x1, fmaf(w2, x2, bb)));
            const float g = (float)dy1[(size_t)(r0 + rr) * PN_C1 + c];
            const float dy = fmaf(a, z, d) > 0.f ? g : 0.f;
            const float xh = (z - mean) * rstd;
            if (PASS == 0) {
                acc0 += dy;
                acc1 = fmaf(dy, xh, acc1);
            } else {
                const float dz = a * (dy - m1 - xh * m2);
                acc0 = fmaf(dz, x0, acc0); acc1 = fmaf(dz, x1, acc1); acc2 = fmaf(dz, x2, acc2);
                acc3 += dz;
            }
        }
    }
    if (PASS == 0) {
        partials[((size_t)blockIdx.x * 2 + 0) * PN_C1 + c] = acc0;
        partials[((size_t)blockIdx.x * 2 + 1) * PN_C1 + c] = acc1;
    } else {
        // partials (gridDim.x, 2, 2*C1): row 0 = [dW[:,0] | dW[:,1]], row 1 = [dW[:,2] | db]  (same-address atomics from
        // every CTA serialised this kernel: 53 -> ~15 us)
        float *pr = partials + (size_t)blockIdx.x * 4 * PN_C1;
        pr[c] = acc0; pr[PN_C1 + c] = acc1; pr[2 * PN_C1 + c] = acc2; pr[3 * PN_C1 + c] = acc3;
    }
}

// ------------------------------------------------------------------------------------------ partial-sum reduction / BN finalize
// grid = ceil(C / 32), block = (32 columns, FIN_LANES lanes over the partials).
// MODE 0 (backward sums): partials (nPart, 2, C) -> sums (2, C), fp64 accumulation.
// MODE 1 (forward statistics): partials (nPart, 3, C) = [shift, sum (z-shift), sum (z-shift)^2] over
//   n_p = min(rows_per_part, total_rows - p*rows_per_part) rows each.  Two passes in fp64, no cancellation:
//     mean = sum_p (n_p shift_p + s_p) / N ;  M2 = sum_p [ q_p + 2 (shift_p - mean) s_p + n_p (shift_p - mean)^2 ]
//   -> optional triple_out (3, C) = [mean, 0, M2] (the same format, for a second-level merge across ranks) and/or
//   stats (4, C) = mean, rstd, a = gamma*rstd, d = beta - mean*a plus nn.BatchNorm1d's running-statistics update
//   (momentum, unbiased variance).
constexpr int FIN_LANES = 32;

__device__ __forceinline__ double fin_lane_sum(double v, double (*sh)[32]) {
    __syncthreads();
    sh[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll 8
    for (int i = 0; i < FIN_LANES; ++i) t += sh[i][threadIdx.x];
    return t;                                   // every lane gets the column total
}

template <int MODE>
__global__ void __launch_bounds__(32 * FIN_LANES)
bn_reduce_finalize_kernel(int nPart, int C, const float *__restrict__ partials, float *__restrict__ sums, int rows_per_part,
                          long long total_rows, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                          float momentum, float *__restrict__ running_mean, float *__restrict__ running_var,
                          long long *__restrict__ nbt, float *__restrict__ triple_out, float *__restrict__ stats) {
    __shared__ double sh[FIN_LANES][32];
    const int c = blockIdx.x * 32 + threadIdx.x, ly = threadIdx.y;
    const bool ok = c < C;
    if (MODE == 0) {
        const int nrows = rows_per_part;        // MODE 0 reuses this argument: rows (1 or 2) of each partial
        double s = 0.0, q = 0.0;
        if (ok) {
#pragma unroll 4
            for (int p = ly; p < nPart; p += FIN_LANES) {
                s += (double)partials[((size_t)p * nrows + 0) * C + c];
                if (nrows == 2) q += (double)partials[((size_t)p * nrows + 1) * C + c];
            }
        }
        s = fin_lane_sum(s, sh);
        q = fin_lane_sum(q, sh);
        if (ly == 0 && ok) {
            sums[c] = (float)s;
            if (nrows == 2) sums[C + c] = (float)q;
        }
        return;
    }
    const double N = (double)total_rows;
    double m = 0.0;
    if (ok) {
#pragma unroll 4
        for (int p = ly; p < nPart; p += FIN_LANES) {
            const long long left = total_rows - (long long)p * rows_per_part;
            const double n = (double)(left < rows_per_part ? (left > 0 ? left : 0) : rows_per_part);
            m += n * (double)partials[((size_t)p * 3 + 0) * C + c] + (double)partials[((size_t)p * 3 + 1) * C + c];
        }
    }
    const double mean = fin_lane_sum(m, sh) / N;
    double m2 = 0.0;
    if (ok) {
#pragma unroll 4
        for (int p = ly; p < nPart; p += FIN_LANES) {
            const long long left = total_rows - (long long)p * rows_per_part;
            const double n = (double)(left < rows_per_part ? (left > 0 ? left : 0) : rows_per_part);
            const double dlt = (double)partials[((size_t)p * 3 + 0) * C + c] - mean;
            m2 += (double)partials[((size_t)p * 3 + 2) * C + c] + 2.0 * dlt * (double)partials[((size_t)p * 3 + 1) * C + c] +
                  n * dlt * dlt;
        }
    }
    m2 = fin_lane_sum(m2, sh);
    if (m2 < 0.0) m2 = 0.0;
    if (ly == 0 && ok) {
        if (triple_out) {
            triple_out[c] = (float)mean;
            triple_out[C + c] = 0.f;
            triple_out[2 * C + c] = (float)m2;
        }
        if (stats) {
            const double var = m2 / N;
            const float rstd = (float)(1.0 / sqrt(var + (double)eps));
            const float a = gamma[c] * rstd;
            stats[c] = (float)mean;
            stats[C + c] = rstd;
            stats[2 * C + c] = a;
            stats[3 * C + c] = beta[c] - (float)mean * a;
            if (running_mean) {
                const double unbiased = N > 1.0 ? m2 / (N - 1.0) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
            }
        }
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0 && ly == 0) nbt[0] += 1;
}

// ------------------------------------------------------------------------------------------ group-tile kernels over (R, C)
// CTA = `gpc` consecutive groups x one column tile; blockDim = (CT column threads, RL row lanes); a column thread owns
// 4 consecutive columns; row lane l handles rows k = l, l+RL, ... of each group.
constexpr int GT_RL = 4;
constexpr int GT_MAXK = 64;
#ifndef UP3D_GT_BATCH
#define UP3D_GT_BATCH 8
#endif
#ifndef UP3D_GT_MIN_CTAS
#define UP3D_GT_MIN_CTAS 1
#endif
constexpr int GT_BATCH = UP3D_GT_BATCH;        // rows of a group handled per lane are unrolled in batches of GT_BATCH
constexpr int GT_MIN_CTAS = UP3D_GT_MIN_CTAS;  // resident CTAs per SM the register budget is sized for

struct GroupTileArgs {
    int Gt, K, C, gpc;
    float inv_count;              // 1 / R (backward)
    const void *in0;              // MODE-dependent main (R, C) input, AT
    const void *in1;              // second (R, C) input, AT
    const float *gpart;           // (Gt, C) fp32 broadcast term (global half of the concat GEMM), may be NULL
    const float *bias;            // (C) conv bias added to in0 + gpart, may be NULL
    const float *stats;           // (4, C): mean, rstd, a, d
    const float *sums;            // (2, C): backward sums
    const int *arg;               // (Gt, C) arg-max row of each group
    const void *small_in;         // (Gt, C) per-group input (max_scatter: d tokens; combine: dfg), AT
    void *out;                    // (R, C) AT
    void *small_out;              // (Gt, C): group_max values (AT) / bn_bwd_apply group sums (AT)
    int *arg_out;                 // (Gt, C)
    float *partials;              // (gridDim.x, 2, C)
    float *colsum;                // (C) float atomics
};

enum { GT_STATS = 0, GT_APPLY = 1, GT_BWD_REDUCE = 2, GT_BWD_APPLY = 3, GT_MAX = 4, GT_SCATTER = 5, GT_COMBINE = 6 };

// ---- bulk-async staging (STAGED variant): a group's K x C tile is ONE contiguous block of memory when the CTA spans all
// columns, so the CTA streams whole group tiles into a ring of shared-memory stages with cp.async.bulk (one elected
// thread, completion on an mbarrier) and computes from shared memory.  GT_STAGES tiles (per input) are in flight per SM
// at all times, without holding them in registers: the direct-load variant ran load -> wait -> compute -> barrier per
// group and reached ~1.5 TB/s (45 us for the 67 MB of the backward-reduce pass).
constexpr int GT_STAGES = 3;
constexpr int GT_BAR_BYTES = 128;

__device__ __forceinline__ uint32_t gt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gt_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gt_smem_u32(dst)), "l"(src), "r"(bytes), "r"(gt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void gt_mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(gt_smem_u32(bar)), "r"(parity) : "memory");
    }
}

template <int MODE> struct GtInputs { static constexpr int N = (MODE == 2 || MODE == 3) ? 2 : (MODE == 5 ? 0 : 1); };

template <typename AT, int MODE, bool STAGED>
__global__ void __launch_bounds__(512, GT_MIN_CTAS)
group_tile_kernel(GroupTileArgs A) {
    __shared__ float4 sh0[GT_RL][128];
    __shared__ float4 sh1[GT_RL][128];
    extern __shared__ __align__(128) uint8_t gt_dyn[];
    const int ct = threadIdx.x, lane_r = threadIdx.y;
    const int col = (blockIdx.y * blockDim.x + ct) * 4;
    const bool col_ok = col < A.C;
    const int C = A.C, K = A.K;
    const AT *in0 = reinterpret_cast<const AT *>(A.in0);
    const AT *in1 = reinterpret_cast<const AT *>(A.in1);
    AT *out = reinterpret_cast<AT *>(A.out);
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), mean = bias, rstd = bias, a = bias, d = bias, m1 = bias, m2 = bias;
    if (col_ok) {
        if (A.bias) bias = *reinterpret_cast<const float4 *>(A.bias + col);
        if (A.stats) {
            mean = *reinterpret_cast<const float4 *>(A.stats + col);
            rstd = *reinterpret_cast<const float4 *>(A.stats + C + col);
            a = *reinterpret_cast<const float4 *>(A.stats + 2 * C + col);
            d = *reinterpret_cast<const float4 *>(A.stats + 3 * C + col);
        }
        if (MODE == GT_BWD_APPLY) {
            m1 = *reinterpret_cast<const float4 *>(A.sums + col);
            m2 = *reinterpret_cast<const float4 *>(A.sums + C + col);
            m1.x *= A.inv_count; m1.y *= A.inv_count; m1.z *= A.inv_count; m1.w *= A.inv_count;
            m2.x *= A.inv_count; m2.y *= A.inv_count; m2.z *= A.inv_count; m2.w *= A.inv_count;
        }
    }
    float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;       // per-CTA partial sums (STATS / BWD_REDUCE / COMBINE)
    const int g_begin = blockIdx.x * A.gpc, g_end = min(A.Gt, g_begin + A.gpc);
    float4 shift = make_float4(0.f, 0.f, 0.f, 0.f);             // STATS: z of the CTA's first row (same for all row lanes)
    if (MODE == GT_STATS && col_ok && !STAGED) {
        shift = PVec4<AT>::load(in0 + (size_t)g_begin * K * C + col);
        shift.x += bias.x; shift.y += bias.y; shift.z += bias.z; shift.w += bias.w;
        if (A.gpart) {
            const float4 t = *reinterpret_cast<const float4 *>(A.gpart + (size_t)g_begin * C + col);
            shift.x += t.x; shift.y += t.y; shift.z += t.z; shift.w += t.w;
        }
    }
    constexpr int NIN = GtInputs<MODE>::N;
    const uint32_t tile_bytes = (uint32_t)K * (uint32_t)C * (uint32_t)sizeof(AT);
    uint64_t *full = reinterpret_cast<uint64_t *>(gt_dyn);
    uint8_t *ring = gt_dyn + GT_BAR_BYTES;
    const bool elected = STAGED && threadIdx.x == 0 && threadIdx.y == 0;
    auto issue = [&](int g, int st) {                               // elected thread: group g's tile(s) -> stage st
        gt_mbar_expect_tx(&full[st], NIN * tile_bytes);
        gt_bulk_g2s(ring + (size_t)st * NIN * tile_bytes, in0 + (size_t)g * K * C, tile_bytes, &full[st]);
        if (NIN == 2) gt_bulk_g2s(ring + ((size_t)st * NIN + 1) * tile_bytes, in1 + (size_t)g * K * C, tile_bytes, &full[st]);
    };
    if (STAGED) {
        if (elected) {
#pragma unroll
            for (int st = 0; st < GT_STAGES; ++st) gt_mbar_init(&full[st], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (elected)
            for (int st = 0; st < GT_STAGES && g_begin + st < g_end; ++st) issue(g_begin + st, st);
    }
    if (MODE == GT_STATS && col_ok && STAGED) {                     // (the unstaged variant read it from global above)
        gt_mbar_wait(&full[0], 0);
        shift = PVec4<AT>::load(reinterpret_cast<const AT *>(ring) + col);
        shift.x += bias.x; shift.y += bias.y; shift.z += bias.z; shift.w += bias.w;
        if (A.gpart) {
            const float4 t = *reinterpret_cast<const float4 *>(A.gpart + (size_t)g_begin * C + col);
            shift.x += t.x; shift.y += t.y; shift.z += t.z; shift.w += t.w;
        }
    }
    for (int g = g_begin; g < g_end; ++g) {
        const int it = g - g_begin, stg = it % GT_STAGES;
        const AT *s0 = reinterpret_cast<const AT *>(ring + (size_t)stg * NIN * tile_bytes);
        const AT *s1 = reinterpret_cast<const AT *>(ring + ((size_t)stg * NIN + 1) * tile_bytes);
        if (STAGED) gt_mbar_wait(&full[stg], (uint32_t)((it / GT_STAGES) & 1));
        float4 gp = bias;                                         // broadcast term of this group: gpart + bias
        if (col_ok && A.gpart) {
            const float4 t = *reinterpret_cast<const float4 *>(A.gpart + (size_t)g * C + col);
            gp.x += t.x; gp.y += t.y; gp.z += t.z; gp.w += t.w;
        }
        float4 small = make_float4(0.f, 0.f, 0.f, 0.f);
        int4 am = make_int4(-1, -1, -1, -1);
        if (col_ok && (MODE == GT_SCATTER || MODE == GT_COMBINE)) {
            small = PVec4<AT>::load(reinterpret_cast<const AT *>(A.small_in) + (size_t)g * C + col);
            am = *reinterpret_cast<const int4 *>(A.arg + (size_t)g * C + col);
        }
        float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        int4 bidx = make_int4(0, 0, 0, 0);
        float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);             // group sum (BWD_APPLY)
        constexpr int NB = STAGED ? 4 : GT_BATCH;     // rows in flight per lane (shared-memory reads need few)
        for (int kb = lane_r; kb < K; kb += GT_RL * NB) {
            float4 u[NB], v[NB];
            if (MODE != GT_SCATTER) {
#pragma unroll
                for (int i = 0; i < NB; ++i) {
                    const int k = kb + i * GT_RL;
                    if (k < K && col_ok) {
                        if (STAGED) {
                            const size_t so = (size_t)k * C + col;
                            u[i] = PVec4<AT>::load(s0 + so);
                            if (MODE == GT_BWD_REDUCE || MODE == GT_BWD_APPLY) v[i] = PVec4<AT>::load(s1 + so);
                        } else {
                            const size_t o = ((size_t)g * K + k) * C + col;
                            u[i] = PVec4<AT>::load(in0 + o);
                            if (MODE == GT_BWD_REDUCE || MODE == GT_BWD_APPLY) v[i] = PVec4<AT>::load(in1 + o);
                        }
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NB; ++i) {
                const int k = kb + i * GT_RL;
                if (k >= K || !col_ok) continue;
                const size_t o = ((size_t)g * K + k) * C + col;
                const float4 ui = u[i], vi = v[i];
                if (MODE == GT_STATS) {
                    const float z0 = ui.x + gp.x - shift.x, z1 = ui.y + gp.y - shift.y;
                    const float z2 = ui.z + gp.z - shift.z, z3 = ui.w + gp.w - shift.w;
                    p0.x += z0; p0.y += z1; p0.z += z2; p0.w += z3;
                    p1.x = fmaf(z0, z0, p1.x); p1.y = fmaf(z1, z1, p1.y); p1.z = fmaf(z2, z2, p1.z); p1.w = fmaf(z3, z3, p1.w);
                } else if (MODE == GT_APPLY) {
                    float4 y;
                    y.x = fmaxf(fmaf(a.x, ui.x + gp.x, d.x), 0.f);
                    y.y = fmaxf(fmaf(a.y, ui.y + gp.y, d.y), 0.f);
                    y.z = fmaxf(fmaf(a.z, ui.z + gp.z, d.z), 0.f);
                    y.w = fmaxf(fmaf(a.w, ui.w + gp.w, d.w), 0.f);
                    PVec4<AT>::store(out + o, y);
                } else if (MODE == GT_BWD_REDUCE || MODE == GT_BWD_APPLY) {
                    // u = dy (gradient of the ReLU output), v = zl (pre-BN local GEMM output)
                    const float z0 = vi.x + gp.x, z1 = vi.y + gp.y, z2 = vi.z + gp.z, z3 = vi.w + gp.w;
                    const float g0 = fmaf(a.x, z0, d.x) > 0.f ? ui.x : 0.f, g1 = fmaf(a.y, z1, d.y) > 0.f ? ui.y : 0.f;
                    const float g2 = fmaf(a.z, z2, d.z) > 0.f ? ui.z : 0.f, g3 = fmaf(a.w, z3, d.w) > 0.f ? ui.w : 0.f;
                    const float h0 = (z0 - mean.x) * rstd.x, h1 = (z1 - mean.y) * rstd.y;
                    const float h2 = (z2 - mean.z) * rstd.z, h3 = (z3 - mean.w) * rstd.w;
                    if (MODE == GT_BWD_REDUCE) {
                        p0.x += g0; p0.y += g1; p0.z += g2; p0.w += g3;
                        p1.x = fmaf(g0, h0, p1.x); p1.y = fmaf(g1, h1, p1.y); p1.z = fmaf(g2, h2, p1.z); p1.w = fmaf(g3, h3, p1.w);
                    } else {
                        float4 dz;
                        dz.x = a.x * (g0 - m1.x - h0 * m2.x); dz.y = a.y * (g1 - m1.y - h1 * m2.y);
                        dz.z = a.z * (g2 - m1.z - h2 * m2.z); dz.w = a.w * (g3 - m1.w - h3 * m2.w);
                        PVec4<AT>::store(out + o, dz);
                        gs.x += dz.x; gs.y += dz.y; gs.z += dz.z; gs.w += dz.w;
                    }
                } else if (MODE == GT_MAX) {
                    if (ui.x > best.x) { best.x = ui.x; bidx.x = k; }
                    if (ui.y > best.y) { best.y = ui.y; bidx.y = k; }
                    if (ui.z > best.z) { best.z = ui.z; bidx.z = k; }
                    if (ui.w > best.w) { best.w = ui.w; bidx.w = k; }
                } else if (MODE == GT_SCATTER) {
                    PVec4<AT>::store(out + o, make_float4(am.x == k ? small.x : 0.f, am.y == k ? small.y : 0.f,
                                                          am.z == k ? small.z : 0.f, am.w == k ? small.w : 0.f));
                } else if (MODE == GT_COMBINE) {
                    float4 y = ui;
                    if (am.x == k) y.x += small.x;
                    if (am.y == k) y.y += small.y;
                    if (am.z == k) y.z += small.z;
                    if (am.w == k) y.w += small.w;
                    PVec4<AT>::store(out + o, y);
                    p0.x += y.x; p0.y += y.y; p0.z += y.z; p0.w += y.w;
                }
            }
        }
        if (MODE == GT_BWD_APPLY || MODE == GT_MAX) {          // combine the row lanes of this group
            __syncthreads();
            sh0[lane_r][ct] = (MODE == GT_MAX) ? best : gs;
            if (MODE == GT_MAX) sh1[lane_r][ct] = make_float4(__int_as_float(bidx.x), __int_as_float(bidx.y),
                                                              __int_as_float(bidx.z), __int_as_float(bidx.w));
            __syncthreads();
            if (lane_r == 0 && col_ok) {
                if (MODE == GT_BWD_APPLY) {
#pragma unroll
                    for (int l = 1; l < GT_RL; ++l) {
                        const float4 t = sh0[l][ct];
                        gs.x += t.x; gs.y += t.y; gs.z += t.z; gs.w += t.w;
                    }
                    PVec4<AT>::store(reinterpret_cast<AT *>(A.small_out) + (size_t)g * C + col, gs);
                } else {
#pragma unroll
                    for (int l = 1; l < GT_RL; ++l) {
                        const float4 t = sh0[l][ct];
                        const float4 ti = sh1[l][ct];
                        const int i0 = __float_as_int(ti.x), i1 = __float_as_int(ti.y), i2 = __float_as_int(ti.z),
                                  i3 = __float_as_int(ti.w);
                        // first occurrence wins ties (rows are visited in increasing k within a lane)
                        if (t.x > best.x || (t.x == best.x && i0 < bidx.x)) { best.x = t.x; bidx.x = i0; }
                        if (t.y > best.y || (t.y == best.y && i1 < bidx.y)) { best.y = t.y; bidx.y = i1; }
                        if (t.z > best.z || (t.z == best.z && i2 < bidx.z)) { best.z = t.z; bidx.z = i2; }
                        if (t.w > best.w || (t.w == best.w && i3 < bidx.w)) { best.w = t.w; bidx.w = i3; }
                    }
                    PVec4<AT>::store(reinterpret_cast<AT *>(A.small_out) + (size_t)g * C + col, best);
                    *reinterpret_cast<int4 *>(A.arg_out + (size_t)g * C + col) = bidx;
                }
            }
        }
        if (STAGED && g + GT_STAGES < g_end) {                   // every thread is done with this stage: refill it
            __syncthreads();
            if (elected) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(g + GT_STAGES, stg);
            }
        }
    }
    if (MODE == GT_STATS || MODE == GT_BWD_REDUCE || MODE == GT_COMBINE) {
        __syncthreads();
        sh0[lane_r][ct] = p0;
        sh1[lane_r][ct] = p1;
        __syncthreads();
        if (lane_r == 0 && col_ok) {
#pragma unroll
            for (int l = 1; l < GT_RL; ++l) {
                const float4 t = sh0[l][ct], t1 = sh1[l][ct];
                p0.x += t.x; p0.y += t.y; p0.z += t.z; p0.w += t.w;
                p1.x += t1.x; p1.y += t1.y; p1.z += t1.z; p1.w += t1.w;
            }
            if (MODE == GT_COMBINE) {
                if (A.partials) *reinterpret_cast<float4 *>(A.partials + (size_t)blockIdx.x * C + col) = p0;
            } else if (MODE == GT_STATS) {
                *reinterpret_cast<float4 *>(A.partials + ((size_t)blockIdx.x * 3 + 0) * C + col) = shift;
                *reinterpret_cast<float4 *>(A.partials + ((size_t)blockIdx.x * 3 + 1) * C + col) = p0;
                *reinterpret_cast<float4 *>(A.partials + ((size_t)blockIdx.x * 3 + 2) * C + col) = p1;
            } else {
                *reinterpret_cast<float4 *>(A.partials + ((size_t)blockIdx.x * 2 + 0) * C + col) = p0;
                *reinterpret_cast<float4 *>(A.partials + ((size_t)blockIdx.x * 2 + 1) * C + col) = p1;
            }
        }
    }
}

// Staged (bulk-async ring) when the CTA spans all columns (a group tile is then contiguous), the tile is 16-byte granular,
// the ring fits and a CTA has several groups to pipeline; UP3D_GT_STAGED=0 keeps the direct-load variant.
static std::atomic<int> g_gt_staged{-1};         // -1: not decided yet (environment), 0 / 1: off / on
static bool gt_staged_enabled() {
    int v = g_gt_staged.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("UP3D_GT_STAGED");
        v = (e && e[0] == '0') ? 0 : 1;
        g_gt_staged.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}

template <typename AT, int MODE>
static int launch_group_tile(const GroupTileArgs &A, cudaStream_t st) {
    const int cthreads = A.C / 4;
    int ct = cthreads <= 128 ? cthreads : 128;
    if (cthreads > 128) {
        for (int t = 128; t >= 32; t -= 32)
            if (cthreads % t == 0) { ct = t; break; }
    }
    const dim3 block(ct, GT_RL), grid(div_up(A.Gt, A.gpc), div_up(cthreads, ct));
    constexpr int NIN = GtInputs<MODE>::N;
    const size_t tile_bytes = (size_t)A.K * A.C * sizeof(AT);
    const size_t dyn = GT_BAR_BYTES + (size_t)GT_STAGES * NIN * tile_bytes;
    if (NIN > 0 && gt_staged_enabled() && grid.y == 1 && A.gpc >= 2 && tile_bytes % 16 == 0 && dyn <= 200 * 1024) {
        static std::atomic<size_t> configured{0};
        if (dyn > configured.load(std::memory_order_acquire)) {
            UP3D_CUDA_OK(cudaFuncSetAttribute((const void *)group_tile_kernel<AT, MODE, true>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            configured.store(dyn, std::memory_order_release);
        }
        group_tile_kernel<AT, MODE, true><<<grid, block, dyn, st>>>(A);
    } else {
        group_tile_kernel<AT, MODE, false><<<grid, block, 0, st>>>(A);
    }
    return 0;
}

}  // namespace up3d

using namespace up3d;

static bool pn_aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

/* ---- first layer ---- */
extern "C" int up3d_pn_stats_tile_rows(void) { return PN_TILE; }

extern "C" int up3d_pn_conv1_stats(int R, int GK, const float *nb, const float *W1, const float *b1, float *partials,
                                   up3d_stream_t stream) {
    UP3D_CHECK_ARG(R > 0 && GK > 0 && R % GK == 0, "up3d_pn_conv1_stats: R must be a positive multiple of G*K");
    UP3D_CHECK_ARG(nb && W1 && b1 && partials, "up3d_pn_conv1_stats: NULL pointer");
    pn_conv1_stats_kernel<<<div_up(R, PN_TILE), PN_C1, 0, (cudaStream_t)stream>>>(R, GK, nb, W1, b1, partials);
    UP3D_LAUNCH_OK("pn_conv1_stats_kernel");
    return 0;
}

extern "C" int up3d_pn_conv1_bn_relu(int act_bf16, int R, int GK, const float *nb, const float *W1, const float *b1,
                                     const float *stats, void *y1, up3d_stream_t stream) {
    UP3D_CHECK_ARG(R > 0 && GK > 0 && R % GK == 0, "up3d_pn_conv1_bn_relu: R must be a positive multiple of G*K");
    UP3D_CHECK_ARG(nb && W1 && b1 && stats && y1 && pn_aligned16(y1), "up3d_pn_conv1_bn_relu: NULL or misaligned pointer");
    const int grid = min(div_up(R, PN_TILE), UP3D_NUM_SMS * 4);
    if (act_bf16)
        pn_conv1_bn_relu_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(R, GK, nb, W1, b1, stats, (__nv_bfloat16 *)y1);
    else
        pn_conv1_bn_relu_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(R, GK, nb, W1, b1, stats, (float *)y1);
    UP3D_LAUNCH_OK("pn_conv1_bn_relu_kernel");
    return 0;
}

extern "C" int up3d_pn_conv1_bwd(int act_bf16, int pass, int R, int GK, const float *nb, const float *W1, const float *b1,
                                 const float *stats, const void *dy1, const float *sums, double count, float *partials,
                                 int n_partials, float *gW1, float *gb1, up3d_stream_t stream) {
    UP3D_CHECK_ARG(R > 0 && GK > 0 && R % GK == 0, "up3d_pn_conv1_bwd: R must be a positive multiple of G*K");
    UP3D_CHECK_ARG(nb && W1 && b1 && stats && dy1 && n_partials > 0, "up3d_pn_conv1_bwd: NULL pointer");
    UP3D_CHECK_ARG(partials != nullptr && (pass == 0 || (sums && count > 0)), "up3d_pn_conv1_bwd: missing arguments for this pass");
    const float inv_count = pass == 0 ? 0.f : (float)(1.0 / count);
    cudaStream_t st = (cudaStream_t)stream;
#define PN_BWD(AT, P) pn_conv1_bwd_kernel<AT, P><<<n_partials, PN_C1, 0, st>>>(R, GK, nb, W1, b1, stats, (const AT *)dy1, sums, inv_count, partials, gW1, gb1)
    if (act_bf16) { if (pass == 0) PN_BWD(__nv_bfloat16, 0); else PN_BWD(__nv_bfloat16, 1); }
    else { if (pass == 0) PN_BWD(float, 0); else PN_BWD(float, 1); }
#undef PN_BWD
    UP3D_LAUNCH_OK("pn_conv1_bwd_kernel");
    return 0;
}

extern "C" int up3d_bn_reduce_sums(int n_partials, int n_rows, int C, const float *partials, float *sums, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n_partials > 0 && C > 0 && partials && sums && (n_rows == 1 || n_rows == 2), "up3d_bn_reduce_sums: bad arguments");
    bn_reduce_finalize_kernel<0><<<div_up(C, 32), dim3(32, FIN_LANES), 0, (cudaStream_t)stream>>>(
        n_partials, C, partials, sums, n_rows, 0, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, nullptr, nullptr, nullptr);
    UP3D_LAUNCH_OK("bn_reduce_finalize_kernel<sums>");
    return 0;
}

extern "C" int up3d_bn_reduce_finalize(int n_partials, int C, const float *partials, int rows_per_partial, int64_t total_rows,
                                       const float *gamma, const float *beta, float eps, float momentum, float *running_mean,
                                       float *running_var, int64_t *num_batches_tracked, float *triple_out, float *stats,
                                       up3d_stream_t stream) {
    UP3D_CHECK_ARG(n_partials > 0 && C > 0 && partials && rows_per_partial > 0 && total_rows > 0,
                   "up3d_bn_reduce_finalize: bad arguments");
    UP3D_CHECK_ARG(triple_out || stats, "up3d_bn_reduce_finalize: nothing to write");
    UP3D_CHECK_ARG(!stats || (gamma && beta), "up3d_bn_reduce_finalize: stats need gamma/beta");
    UP3D_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "up3d_bn_reduce_finalize: running stats come in pairs");
    bn_reduce_finalize_kernel<1><<<div_up(C, 32), dim3(32, FIN_LANES), 0, (cudaStream_t)stream>>>(
        n_partials, C, partials, nullptr, rows_per_partial, (long long)total_rows, gamma, beta, eps, momentum, running_mean,
        running_var, (long long *)num_batches_tracked, triple_out, stats);
    UP3D_LAUNCH_OK("bn_reduce_finalize_kernel<stats>");
    return 0;
}

/* ---- group-tile passes ---- */
extern "C" int up3d_set_group_tile_staging(int on) {
    const bool prev = gt_staged_enabled();
    g_gt_staged.store(on ? 1 : 0, std::memory_order_relaxed);
    return prev ? 1 : 0;
}

static int gt_common(const char *name, int Gt, int K, int C, int gpc) {
    UP3D_CHECK_ARG(Gt > 0 && K > 0 && C > 0 && C % 4 == 0 && gpc > 0, "%s: bad sizes Gt=%d K=%d C=%d", name, Gt, K, C);
    UP3D_CHECK_ARG(C / 4 <= 128 || (C / 4) % 32 == 0, "%s: C=%d not supported", name, C);
    return 0;
}

#define GT_DISPATCH(MODE)                                         \
    do {                                                          \
        int rc_ = act_bf16 ? launch_group_tile<__nv_bfloat16, MODE>(A, (cudaStream_t)stream) \
                           : launch_group_tile<float, MODE>(A, (cudaStream_t)stream);       \
        if (rc_) return rc_;                                      \
    } while (0)

extern "C" int up3d_gbn_stats(int act_bf16, int Gt, int K, int C, int gpc, const void *zl, const float *gpart, const float *bias,
                              float *partials, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_gbn_stats", Gt, K, C, gpc)) return rc;
    UP3D_CHECK_ARG(zl && partials && pn_aligned16(zl) && pn_aligned16(gpart) && pn_aligned16(bias) && pn_aligned16(partials),
                   "up3d_gbn_stats: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = gpc; A.in0 = zl; A.gpart = gpart; A.bias = bias; A.partials = partials;
    GT_DISPATCH(GT_STATS);
    UP3D_LAUNCH_OK("group_tile_kernel<stats>");
    return 0;
}

extern "C" int up3d_gbn_apply_relu(int act_bf16, int Gt, int K, int C, int gpc, const void *zl, const float *gpart,
                                   const float *bias, const float *stats, void *y, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_gbn_apply_relu", Gt, K, C, gpc)) return rc;
    UP3D_CHECK_ARG(zl && stats && y && pn_aligned16(zl) && pn_aligned16(gpart) && pn_aligned16(bias) && pn_aligned16(stats) &&
                   pn_aligned16(y), "up3d_gbn_apply_relu: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = gpc; A.in0 = zl; A.gpart = gpart; A.bias = bias; A.stats = stats; A.out = y;
    GT_DISPATCH(GT_APPLY);
    UP3D_LAUNCH_OK("group_tile_kernel<apply>");
    return 0;
}

extern "C" int up3d_gbn_bwd_reduce(int act_bf16, int Gt, int K, int C, int gpc, const void *dy, const void *zl, const float *gpart,
                                   const float *bias, const float *stats, float *partials, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_gbn_bwd_reduce", Gt, K, C, gpc)) return rc;
    UP3D_CHECK_ARG(dy && zl && stats && partials && pn_aligned16(dy) && pn_aligned16(zl) && pn_aligned16(gpart) &&
                   pn_aligned16(bias) && pn_aligned16(stats) && pn_aligned16(partials), "up3d_gbn_bwd_reduce: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = gpc; A.in0 = dy; A.in1 = zl; A.gpart = gpart; A.bias = bias; A.stats = stats;
    A.partials = partials;
    GT_DISPATCH(GT_BWD_REDUCE);
    UP3D_LAUNCH_OK("group_tile_kernel<bwd_reduce>");
    return 0;
}

extern "C" int up3d_gbn_bwd_apply(int act_bf16, int Gt, int K, int C, int gpc, const void *dy, const void *zl, const float *gpart,
                                  const float *bias, const float *stats, const float *sums, double count, void *dz,
                                  void *dgroup, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_gbn_bwd_apply", Gt, K, C, gpc)) return rc;
    UP3D_CHECK_ARG(count > 0, "up3d_gbn_bwd_apply: count must be positive");
    UP3D_CHECK_ARG(dy && zl && stats && sums && dz && dgroup && pn_aligned16(dy) && pn_aligned16(zl) && pn_aligned16(gpart) &&
                   pn_aligned16(bias) && pn_aligned16(stats) && pn_aligned16(sums) && pn_aligned16(dz) && pn_aligned16(dgroup),
                   "up3d_gbn_bwd_apply: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = gpc; A.in0 = dy; A.in1 = zl; A.gpart = gpart; A.bias = bias; A.stats = stats;
    A.sums = sums; A.inv_count = (float)(1.0 / count); A.out = dz; A.small_out = dgroup;
    GT_DISPATCH(GT_BWD_APPLY);
    UP3D_LAUNCH_OK("group_tile_kernel<bwd_apply>");
    return 0;
}

extern "C" int up3d_group_max(int act_bf16, int Gt, int K, int C, const void *x, void *out, int32_t *arg, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_group_max", Gt, K, C, 1)) return rc;
    UP3D_CHECK_ARG(x && out && arg && pn_aligned16(x) && pn_aligned16(out) && pn_aligned16(arg), "up3d_group_max: NULL or misaligned pointer");
    GroupTileArgs A{};
    // no per-CTA partials here, so the groups-per-CTA split is free: one wave of CTAs, each pipelining its groups
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = act_bf16 ? (Gt + UP3D_NUM_SMS - 1) / UP3D_NUM_SMS : 1; A.in0 = x; A.small_out = out; A.arg_out = arg;
    GT_DISPATCH(GT_MAX);
    UP3D_LAUNCH_OK("group_tile_kernel<max>");
    return 0;
}

extern "C" int up3d_group_max_scatter(int act_bf16, int Gt, int K, int C, const void *dpooled, const int32_t *arg, void *dx,
                                      up3d_stream_t stream) {
    if (int rc = gt_common("up3d_group_max_scatter", Gt, K, C, 1)) return rc;
    UP3D_CHECK_ARG(dpooled && arg && dx && pn_aligned16(dpooled) && pn_aligned16(arg) && pn_aligned16(dx),
                   "up3d_group_max_scatter: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = 1; A.small_in = dpooled; A.arg = arg; A.out = dx;
    GT_DISPATCH(GT_SCATTER);
    UP3D_LAUNCH_OK("group_tile_kernel<scatter>");
    return 0;
}

extern "C" int up3d_group_combine(int act_bf16, int Gt, int K, int C, int gpc, const void *dlocal, const void *dpooled,
                                  const int32_t *arg, void *dx, float *colsum_partials, up3d_stream_t stream) {
    if (int rc = gt_common("up3d_group_combine", Gt, K, C, gpc)) return rc;
    UP3D_CHECK_ARG(dlocal && dpooled && arg && dx && pn_aligned16(dlocal) && pn_aligned16(dpooled) && pn_aligned16(arg) &&
                   pn_aligned16(dx), "up3d_group_combine: NULL or misaligned pointer");
    GroupTileArgs A{};
    A.Gt = Gt; A.K = K; A.C = C; A.gpc = gpc; A.in0 = dlocal; A.small_in = dpooled; A.arg = arg; A.out = dx; A.partials = colsum_partials;
    GT_DISPATCH(GT_COMBINE);
    UP3D_LAUNCH_OK("group_tile_kernel<combine>");
    return 0;
}
