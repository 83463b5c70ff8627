// attention.cu -- multi-head self-attention of the point-Transformer Blocks
// (/root/reference/openpoints/models/backbone/transformer.py:36-77: softmax(q k^T * scale) v, no mask, no dropout)
// for the token counts of this path: L = G+1 = 129 tokens, head_dim 64, read straight from the fused
// qkv GEMM output (T, 3C) and written straight into the layouts the neighbouring GEMMs consume:
//
//   forward   o (T, C)  <- qkv          (+ log-sum-exp per (b, h, token) for the backward)
//   backward  dqkv (T, 3C) <- qkv, o, lse, do      in ONE launch: no dq/dk/dv staging tensors, no concat, no
//             fp32->bf16 dq conversion pass, no separate dO.O reduction (the library path needed 5 launches here)
//
// One CTA per (48-row tile, head, sample): K and V (and, backward, Q and dO) of the whole (b, h) live in shared
// memory (L <= 192), so there is no online-softmax rescaling: S = Q K^T is formed once per tile, soft-maxed in
// shared memory, and fed back as the A operand of the second GEMM.  Tensor-core math through the portable warp-level
// MMA API (bf16 inputs, fp32 accumulation) -- these are 129x129x64 problems, far below the size where tcgen05/TMEM
// staging pays off; the win is launch count and layout, not peak FLOP/s.
//
// Backward, per CTA with tile index t (rows t0..t0+47):
//   role A (queries i in tile):  S = Q_t K^T, dP = dO_t V^T, dS = P o (dP - delta_i) * scale,  dQ_t = dS K
//   role B (keys j in tile):     S^T = K_t Q^T, dP^T = V_t dO^T, P^T, dS^T likewise,            dV_t = P^T dO, dK_t = dS^T Q
// Every output row is owned by exactly one CTA: no atomics, deterministic.
#include <cuda_bf16.h>
#include <mma.h>
#include <stdlib.h>

#include "common.cuh"

namespace up3d {

using namespace nvcuda;
typedef __nv_bfloat16 bf16;

constexpr int AT_D = 64;          // head dim
constexpr int AT_QT = 48;         // rows per CTA tile (3 row strips of 16)
constexpr int AT_LD = 72;         // bf16 row stride of the (rows, 64) operand tiles (64 + 8 pad: 144 B, 32 B-aligned tiles)
constexpr int AT_OLD = 68;        // fp32 row stride of the (48, 64) output staging tiles
constexpr int AT_THREADS = 384;   // 12 warps: (row strip wr = warp % 3, column quarter wq = warp / 3); the kernels are
                                  // latency-bound chains of fragment loads and MMAs, so warps per SM matter more than tile reuse
constexpr int AT_MAX_L = 192;
constexpr int AT_NP = AT_MAX_L / 64;   // column pairs per lane in the row-wise passes (c = 2*lane + 64*j)

__host__ __device__ inline int at_round16(int x) { return (x + 15) & ~15; }
__host__ __device__ inline size_t at_align128(size_t x) { return (x + 127) & ~(size_t)127; }
// fp32 staging regions hold the (48 x Lp) score tile (ld Lp+4) and later the (48 x 64) output tile (ld AT_OLD)
__host__ __device__ inline int at_stage_ld(int Lp) { return Lp + 4 > 68 ? Lp + 4 : 68; }
// rows of the full-(b,h) operand tiles in shared memory: every 48-row tile must be addressable (zero-filled past L)
__host__ __device__ inline int at_rows(int L) { const int a = at_round16(L), b = (L + AT_QT - 1) / AT_QT * AT_QT; return a > b ? a : b; }

// (rows x 64) bf16 tile of a (T, row_stride) matrix -> shared [n_rows][AT_LD]; rows >= n_valid are zero-filled.
// Asynchronous 16-byte copies (cp.async, no register staging): a thread issues all of its copies back to back, so the
// whole tile is in flight at once; the caller waits with at_async_wait() + __syncthreads().
__device__ __forceinline__ void at_load_rows(bf16 *dst, const bf16 *src, size_t row_stride, int row0, int n_valid, int n_rows) {
    for (int i = threadIdx.x; i < n_rows * 8; i += AT_THREADS) {
        const int r = i >> 3, ch = i & 7;
        const bool ok = r < n_valid;
        const bf16 *g = ok ? src + (size_t)(row0 + r) * row_stride + ch * 8 : src;
        const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + r * AT_LD + ch * 8);
        const int nbytes = ok ? 16 : 0;              // src-size 0: the 16 destination bytes are zero-filled
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(g), "r"(nbytes) : "memory");
    }
}
__device__ __forceinline__ void at_async_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// out (48 x Lp, fp32, ld = ldS) = X (48 x 64 rows of Xs starting at x_row0) * Y^T (Y: Lp x 64), both [.][AT_LD] bf16;
// this warp: row strip wr, column tiles ct = wq, wq+4, ...
__device__ __forceinline__ void at_gemm_xyT(float *out, int ldS, const bf16 *Xs, int x_row0, const bf16 *Ys, int Lp, int wr, int wq) {
    wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> a[4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) wmma::load_matrix_sync(a[ks], Xs + (x_row0 + wr * 16) * AT_LD + ks * 16, AT_LD);
    const int nct = Lp / 16;
    for (int ct = wq; ct < nct; ct += 4) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::col_major> b[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) wmma::load_matrix_sync(b[ks], Ys + ct * 16 * AT_LD + ks * 16, AT_LD);
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> c;
        wmma::fill_fragment(c, 0.f);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) wmma::mma_sync(c, a[ks], b[ks], c);
        wmma::store_matrix_sync(out + wr * 16 * ldS + ct * 16, c, ldS, wmma::mem_row_major);
    }
}

// out (48 x 64, fp32, ld = AT_OLD) = P (48 x Lp bf16, ld = ldP) * Y (Lp x 64, [.][AT_LD] bf16); this warp: row strip wr,
// column tile wq; two independent accumulators (even / odd k-steps) shorten the MMA dependency chain
__device__ __forceinline__ void at_gemm_py(float *out, const bf16 *Ps, int ldP, const bf16 *Ys, int Lp, int wr, int wq) {
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> c0, c1;
    wmma::fill_fragment(c0, 0.f);
    wmma::fill_fragment(c1, 0.f);
    const int nks = Lp / 16;
    for (int ks = 0; ks < nks; ks += 2) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> a0, a1;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> b0, b1;
        wmma::load_matrix_sync(a0, Ps + wr * 16 * ldP + ks * 16, ldP);
        wmma::load_matrix_sync(b0, Ys + ks * 16 * AT_LD + wq * 16, AT_LD);
        if (ks + 1 < nks) {
            wmma::load_matrix_sync(a1, Ps + wr * 16 * ldP + (ks + 1) * 16, ldP);
            wmma::load_matrix_sync(b1, Ys + (ks + 1) * 16 * AT_LD + wq * 16, AT_LD);
        }
        wmma::mma_sync(c0, a0, b0, c0);
        if (ks + 1 < nks) wmma::mma_sync(c1, a1, b1, c1);
    }
#pragma unroll
    for (int i = 0; i < c0.num_elements; ++i) c0.x[i] += c1.x[i];
    wmma::store_matrix_sync(out + wr * 16 * AT_OLD + wq * 16, c0, AT_OLD, wmma::mem_row_major);
}

// (48 x 64) fp32 staging tile -> bf16 rows of a (T, row_stride) matrix, optionally scaled per row
__device__ __forceinline__ void at_store_rows(bf16 *dst, size_t row_stride, int row0, int n_valid, const float *stage,
                                              const float *row_scale) {
    for (int i = threadIdx.x; i < AT_QT * 8; i += AT_THREADS) {
        const int r = i >> 3, ch = i & 7;
        if (r >= n_valid) continue;
        const float s = row_scale ? row_scale[r] : 1.f;
        const float *p = stage + r * AT_OLD + ch * 8;
        __nv_bfloat162 h[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(p[2 * k] * s, p[2 * k + 1] * s);
        *reinterpret_cast<uint4 *>(dst + (size_t)(row0 + r) * row_stride + ch * 8) = *reinterpret_cast<const uint4 *>(h);
    }
}

struct AttnSmemFwd {
    size_t k, v, q, s, p, rowinv, total;
    __host__ __device__ explicit AttnSmemFwd(int Lp) {
        size_t o = 0;
        k = o; o = at_align128(o + (size_t)Lp * AT_LD * 2);
        v = o; o = at_align128(o + (size_t)Lp * AT_LD * 2);
        q = o; o = at_align128(o + (size_t)AT_QT * AT_LD * 2);
        s = o; o = at_align128(o + (size_t)AT_QT * at_stage_ld(Lp) * 4);
        p = o; o = at_align128(o + (size_t)AT_QT * (Lp + 8) * 2);
        rowinv = o; o = at_align128(o + AT_QT * 4);
        total = o;
    }
};

__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(int L, int H, float scale, const bf16 *__restrict__ qkv, bf16 *__restrict__ o, float *__restrict__ lse) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int Lp = at_round16(L), C = H * AT_D;
    const AttnSmemFwd lay(Lp);
    bf16 *Ks = reinterpret_cast<bf16 *>(smem + lay.k), *Vs = reinterpret_cast<bf16 *>(smem + lay.v);
    bf16 *Qs = reinterpret_cast<bf16 *>(smem + lay.q), *Pb = reinterpret_cast<bf16 *>(smem + lay.p);
    float *Sf = reinterpret_cast<float *>(smem + lay.s), *rowinv = reinterpret_cast<float *>(smem + lay.rowinv);
    const int t0 = blockIdx.x * AT_QT, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp % 3, wq = warp / 3;
    const int ldS = Lp + 4, ldP = Lp + 8;
    const size_t rs = (size_t)3 * C;
    const bf16 *base = qkv + (size_t)b * L * rs + h * AT_D;
    const int n_q = min(AT_QT, L - t0);

    at_load_rows(Ks, base + C, rs, 0, L, Lp);
    at_load_rows(Vs, base + 2 * C, rs, 0, L, Lp);
    at_load_rows(Qs, base, rs, t0, n_q, AT_QT);
    at_async_wait();
    __syncthreads();
    at_gemm_xyT(Sf, ldS, Qs, 0, Ks, Lp, wr, wq);
    __syncthreads();
    // softmax over the L valid keys; 4 rows per warp, a lane owns the column pairs c = 2*lane + 64*j (kept in
    // registers between the max and the exp pass: one shared-memory read, one packed bf16x2 write per pair)
#pragma unroll
    for (int rr = 0; rr < AT_QT / (AT_THREADS / 32); ++rr) {
        const int r = warp * (AT_QT / (AT_THREADS / 32)) + rr;
        const float *srow = Sf + r * ldS;
        float2 v[AT_NP];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < AT_NP; ++j) {
            const int c = 2 * lane + 64 * j;
            v[j] = make_float2(-INFINITY, -INFINITY);
            if (c < Lp) {
                const float2 t = *reinterpret_cast<const float2 *>(srow + c);
                if (c < L) v[j].x = t.x * scale;
                if (c + 1 < L) v[j].y = t.y * scale;
            }
            m = fmaxf(m, fmaxf(v[j].x, v[j].y));
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < AT_NP; ++j) {
            const int c = 2 * lane + 64 * j;
            if (c < Lp) {
                const float e0 = __expf(v[j].x - m), e1 = __expf(v[j].y - m);      // exp(-inf) = 0 for masked keys
                sum += e0 + e1;
                *reinterpret_cast<__nv_bfloat162 *>(Pb + r * ldP + c) = __floats2bfloat162_rn(e0, e1);
            }
        }
        sum = warp_sum(sum);
        if (lane == 0) {
            rowinv[r] = 1.f / sum;
            if (r < n_q) lse[((size_t)b * H + h) * L + t0 + r] = m + __logf(sum);
        }
    }
    __syncthreads();
    at_gemm_py(Sf, Pb, ldP, Vs, Lp, wr, wq);          // S is dead: its region stages the (48 x 64) output
    __syncthreads();
    at_store_rows(o + (size_t)b * L * C + h * AT_D, C, t0, n_q, Sf, rowinv);
}

struct AttnSmemBwd {
    size_t q, k, v, dO, lse, delta, s, dp, p, ds, total;     // the forward output O is staged in the (not yet used) P/dS region
    __host__ __device__ AttnSmemBwd(int Lp, int Lr) {
        size_t o = 0;
        q = o; o = at_align128(o + (size_t)Lr * AT_LD * 2);
        k = o; o = at_align128(o + (size_t)Lr * AT_LD * 2);
        v = o; o = at_align128(o + (size_t)Lr * AT_LD * 2);
        dO = o; o = at_align128(o + (size_t)Lr * AT_LD * 2);
        lse = o; o = at_align128(o + (size_t)Lp * 4);
        delta = o; o = at_align128(o + (size_t)Lp * 4);
        s = o; o = at_align128(o + (size_t)AT_QT * at_stage_ld(Lp) * 4);
        dp = o; o = at_align128(o + (size_t)AT_QT * at_stage_ld(Lp) * 4);
        p = o; o = at_align128(o + (size_t)AT_QT * (Lp + 8) * 2);
        ds = o; o = at_align128(o + (size_t)AT_QT * (Lp + 8) * 2);
        total = o;
    }
};

__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_kernel(int L, int H, float scale, const bf16 *__restrict__ qkv, const bf16 *__restrict__ o,
                const float *__restrict__ lse, const bf16 *__restrict__ dout, bf16 *__restrict__ dqkv) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int Lp = at_round16(L), Lr = at_rows(L), C = H * AT_D;
    const AttnSmemBwd lay(Lp, Lr);
    bf16 *Qs = reinterpret_cast<bf16 *>(smem + lay.q), *Ks = reinterpret_cast<bf16 *>(smem + lay.k);
    bf16 *Vs = reinterpret_cast<bf16 *>(smem + lay.v), *dOs = reinterpret_cast<bf16 *>(smem + lay.dO);
    float *lse_s = reinterpret_cast<float *>(smem + lay.lse), *delta_s = reinterpret_cast<float *>(smem + lay.delta);
    float *Sf = reinterpret_cast<float *>(smem + lay.s), *dPf = reinterpret_cast<float *>(smem + lay.dp);
    bf16 *Pb = reinterpret_cast<bf16 *>(smem + lay.p), *dSb = reinterpret_cast<bf16 *>(smem + lay.ds);
    const int t0 = blockIdx.x * AT_QT, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wr = warp % 3, wq = warp / 3;
    const int ldS = Lp + 4, ldP = Lp + 8;
    const size_t rs = (size_t)3 * C;
    const bf16 *base = qkv + (size_t)b * L * rs + h * AT_D;
    const bf16 *obase = o + (size_t)b * L * C + h * AT_D, *dobase = dout + (size_t)b * L * C + h * AT_D;
    bf16 *dbase = dqkv + (size_t)b * L * rs + h * AT_D;
    const int n_t = min(AT_QT, L - t0);

    at_load_rows(Qs, base, rs, 0, L, Lr);
    at_load_rows(Ks, base + C, rs, 0, L, Lr);
    at_load_rows(Vs, base + 2 * C, rs, 0, L, Lr);
    at_load_rows(dOs, dobase, C, 0, L, Lr);
    for (int r = threadIdx.x; r < Lp; r += AT_THREADS) lse_s[r] = r < L ? lse[((size_t)b * H + h) * L + r] : 0.f;
    // delta_i = sum_d dO[i][d] * O[i][d]: 8 threads per row (one 16-byte chunk each), O straight from global
    {
        uint4 ov[(AT_MAX_L * 8 + AT_THREADS - 1) / AT_THREADS];
#pragma unroll
        for (int j = 0; j < (AT_MAX_L * 8 + AT_THREADS - 1) / AT_THREADS; ++j) {
            const int i = threadIdx.x + j * AT_THREADS, r = i >> 3, ch = i & 7;
            ov[j] = make_uint4(0u, 0u, 0u, 0u);
            if (i < Lp * 8 && r < L) ov[j] = *reinterpret_cast<const uint4 *>(obase + (size_t)r * C + ch * 8);
        }
        at_async_wait();
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (AT_MAX_L * 8 + AT_THREADS - 1) / AT_THREADS; ++j) {
            const int i = threadIdx.x + j * AT_THREADS, r = i >> 3, ch = i & 7;
            float v = 0.f;
            if (i < Lp * 8) {
                const uint4 dv = *reinterpret_cast<const uint4 *>(dOs + r * AT_LD + ch * 8);
                const __nv_bfloat162 *oh = reinterpret_cast<const __nv_bfloat162 *>(&ov[j]);
                const __nv_bfloat162 *dh = reinterpret_cast<const __nv_bfloat162 *>(&dv);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float2 of = __bfloat1622float2(oh[k]), df = __bfloat1622float2(dh[k]);
                    v = fmaf(of.x, df.x, fmaf(of.y, df.y, v));
                }
            }
            // (Lp*8 is a multiple of 32 and AT_THREADS a multiple of 8: the 8 lanes of a row are in one warp, all active)
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            if (i < Lp * 8 && ch == 0) delta_s[r] = v;
        }
    }
    __syncthreads();

    // ---------------- role A: queries of this tile -> dQ
    at_gemm_xyT(Sf, ldS, Qs, t0, Ks, Lp, wr, wq);
    at_gemm_xyT(dPf, ldS, dOs, t0, Vs, Lp, wr, wq);
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < AT_QT / (AT_THREADS / 32); ++rr) {
        const int r = warp * (AT_QT / (AT_THREADS / 32)) + rr;
        const bool row_ok = r < n_t;
        const float l = row_ok ? lse_s[t0 + r] : 0.f, dl = row_ok ? delta_s[t0 + r] : 0.f;
#pragma unroll
        for (int j = 0; j < AT_NP; ++j) {
            const int c = 2 * lane + 64 * j;
            if (c < Lp) {
                const float2 s2 = *reinterpret_cast<const float2 *>(Sf + r * ldS + c);
                const float2 d2 = *reinterpret_cast<const float2 *>(dPf + r * ldS + c);
                const float x = (row_ok && c < L) ? __expf(s2.x * scale - l) * (d2.x - dl) * scale : 0.f;
                const float y = (row_ok && c + 1 < L) ? __expf(s2.y * scale - l) * (d2.y - dl) * scale : 0.f;
                *reinterpret_cast<__nv_bfloat162 *>(dSb + r * ldP + c) = __floats2bfloat162_rn(x, y);
            }
        }
    }
    __syncthreads();
    at_gemm_py(Sf, dSb, ldP, Ks, Lp, wr, wq);
    __syncthreads();
    at_store_rows(dbase, rs, t0, n_t, Sf, nullptr);
    __syncthreads();

    // ---------------- role B: keys of this tile -> dK, dV  (transposed problem: rows = keys, columns = queries)
    at_gemm_xyT(Sf, ldS, Ks, t0, Qs, Lp, wr, wq);
    at_gemm_xyT(dPf, ldS, Vs, t0, dOs, Lp, wr, wq);
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < AT_QT / (AT_THREADS / 32); ++rr) {
        const int r = warp * (AT_QT / (AT_THREADS / 32)) + rr;
        const bool row_ok = r < n_t;
#pragma unroll
        for (int j = 0; j < AT_NP; ++j) {
            const int c = 2 * lane + 64 * j;
            if (c < Lp) {
                const float2 s2 = *reinterpret_cast<const float2 *>(Sf + r * ldS + c);
                const float2 d2 = *reinterpret_cast<const float2 *>(dPf + r * ldS + c);
                const float2 l2 = *reinterpret_cast<const float2 *>(lse_s + c);
                const float2 dl2 = *reinterpret_cast<const float2 *>(delta_s + c);
                const float p0 = (row_ok && c < L) ? __expf(s2.x * scale - l2.x) : 0.f;
                const float p1 = (row_ok && c + 1 < L) ? __expf(s2.y * scale - l2.y) : 0.f;
                *reinterpret_cast<__nv_bfloat162 *>(Pb + r * ldP + c) = __floats2bfloat162_rn(p0, p1);
                *reinterpret_cast<__nv_bfloat162 *>(dSb + r * ldP + c) =
                    __floats2bfloat162_rn(p0 * (d2.x - dl2.x) * scale, p1 * (d2.y - dl2.y) * scale);
            }
        }
    }
    __syncthreads();
    at_gemm_py(Sf, Pb, ldP, dOs, Lp, wr, wq);          // dV_t = P^T dO
    at_gemm_py(dPf, dSb, ldP, Qs, Lp, wr, wq);         // dK_t = dS^T Q
    __syncthreads();
    at_store_rows(dbase + 2 * C, rs, t0, n_t, Sf, nullptr);
    at_store_rows(dbase + C, rs, t0, n_t, dPf, nullptr);
}

}  // namespace up3d

namespace up3d {
// register-resident mma.sync variant (attention_mma.cu): the default; UP3D_ATTN_WMMA=1 selects the kernels above
int attn_fwd_mma_launch(int B, int L, int H, float scale, const void *qkv, void *o, float *lse, cudaStream_t st);
int attn_bwd_mma_launch(int B, int L, int H, float scale, const void *qkv, const void *o, const float *lse, const void *dout,
                        void *dqkv, cudaStream_t st);
static bool attn_use_mma() {
    static const bool v = []() { const char *e = getenv("UP3D_ATTN_WMMA"); return !(e && e[0] == '1'); }();
    return v;
}
}  // namespace up3d

using namespace up3d;

extern "C" int up3d_attn_max_len(void) { return AT_MAX_L; }

static int attn_check(const char *name, int B, int L, int H, int D) {
    UP3D_CHECK_ARG(B >= 0 && H > 0 && L > 0, "%s: bad sizes", name);
    UP3D_CHECK_ARG(D == AT_D, "%s: head_dim must be %d (got %d)", name, AT_D, D);
    UP3D_CHECK_ARG(L <= AT_MAX_L, "%s: at most %d tokens (got %d)", name, AT_MAX_L, L);
    return 0;
}

extern "C" int up3d_attn_fwd(int B, int L, int H, int D, float scale, const void *qkv, void *o, float *lse,
                             up3d_stream_t stream) {
    if (int rc = attn_check("up3d_attn_fwd", B, L, H, D)) return rc;
    if (B == 0) return 0;
    UP3D_CHECK_ARG(qkv && o && lse, "up3d_attn_fwd: NULL pointer");
    UP3D_CHECK_ARG(((((uintptr_t)qkv) | ((uintptr_t)o)) & 15) == 0, "up3d_attn_fwd: pointers must be 16-byte aligned");
    if (attn_use_mma()) return attn_fwd_mma_launch(B, L, H, scale, qkv, o, lse, (cudaStream_t)stream);
    const AttnSmemFwd lay(at_round16(L));
    static size_t configured = 0;
    if (lay.total > configured) {
        UP3D_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
        configured = lay.total;
    }
    attn_fwd_kernel<<<dim3(div_up(L, AT_QT), H, B), AT_THREADS, lay.total, (cudaStream_t)stream>>>(
        L, H, scale, (const bf16 *)qkv, (bf16 *)o, lse);
    UP3D_LAUNCH_OK("attn_fwd_kernel");
    return 0;
}

extern "C" int up3d_attn_bwd(int B, int L, int H, int D, float scale, const void *qkv, const void *o, const float *lse,
                             const void *dout, void *dqkv, up3d_stream_t stream) {
    if (int rc = attn_check("up3d_attn_bwd", B, L, H, D)) return rc;
    if (B == 0) return 0;
    UP3D_CHECK_ARG(qkv && o && lse && dout && dqkv, "up3d_attn_bwd: NULL pointer");
    UP3D_CHECK_ARG(((((uintptr_t)qkv) | ((uintptr_t)o) | ((uintptr_t)dout) | ((uintptr_t)dqkv)) & 15) == 0,
                   "up3d_attn_bwd: pointers must be 16-byte aligned");
    if (attn_use_mma()) return attn_bwd_mma_launch(B, L, H, scale, qkv, o, lse, dout, dqkv, (cudaStream_t)stream);
    const AttnSmemBwd lay(at_round16(L), at_rows(L));
    static size_t configured = 0;
    if (lay.total > configured) {
        UP3D_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lay.total));
        configured = lay.total;
    }
    attn_bwd_kernel<<<dim3(div_up(L, AT_QT), H, B), AT_THREADS, lay.total, (cudaStream_t)stream>>>(
        L, H, scale, (const bf16 *)qkv, (const bf16 *)o, lse, (const bf16 *)dout, (bf16 *)dqkv);
    UP3D_LAUNCH_OK("attn_bwd_kernel");
    return 0;
}
