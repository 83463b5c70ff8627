// backbone.cu -- the memory-bound glue of the point-Transformer encoder and the optimizer, fused.
//
// The transformer of the render-loss path (/root/reference/openpoints/models/backbone/transformer.py:89-207) runs
// 16 Blocks over B*(G+1) = 8*129 tokens of width 384: every tensor between two GEMMs is ~1.5 MB, so the step is bound
// by the NUMBER of passes over those tensors (the eager graph spends ~60 launches per block on adds, casts, LayerNorm,
// DropPath masks, bias-gradient reductions ...).  These kernels collapse each inter-GEMM stretch into ONE pass:
//
//   ln_fwd      xs = x + scale[b]*delta + pos ; y = LayerNorm(xs)          (Block.forward: x + drop_path(f(norm(x))),
//                                                                            TransformerEncoder.forward: block(x + pos))
//   ln_bwd      dx = g_res + LN'(dy) ; dgamma,dbeta += ; dpos += dx ; dY_prev = scale[b]*dx (cast) ; dbias_prev += colsum
//   gelu_fwd / gelu_bwd(+ bias-gradient column sums)                        (Mlp.forward, transformer.py:27-33, nn.GELU erf)
//   scale_cast_colsum                                                        (DropPath backward + bias gradient)
//   adamw: multi-tensor sum of squares -> clip coefficient -> AdamW update + bf16 shadow refresh in one HBM pass
//          (train_network.py:156-159 AdamW(eps=1e-15), 368-390 clip_grad_norm_(1.0) + skip-on-NaN)
//
// Activations between GEMMs are `AT` = float (reference precision) or __nv_bfloat16 (tensor-core GEMM operands);
// the residual stream, statistics and all parameter gradients stay fp32.
#include <cuda_bf16.h>

#include "common.cuh"

namespace up3d {

// ------------------------------------------------------------------------------------------ vector helpers
template <typename AT> struct Vec4;
template <> struct Vec4<float> {
    static __device__ __forceinline__ float4 load(const float *p) { return *reinterpret_cast<const float4 *>(p); }
    static __device__ __forceinline__ void store(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
    static __device__ __forceinline__ float4 round(float4 v) { return v; }
};
template <> struct Vec4<__nv_bfloat16> {
    static __device__ __forceinline__ float4 load(const __nv_bfloat16 *p) {
        const uint2 r = *reinterpret_cast<const uint2 *>(p);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&r.x);
        const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162 *>(&r.y);
        const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        uint2 r;
        r.x = *reinterpret_cast<const unsigned *>(&a);
        r.y = *reinterpret_cast<const unsigned *>(&b);
        *reinterpret_cast<uint2 *>(p) = r;
    }
    // value as the consumer GEMM will see it (bias gradients are summed over the ROUNDED dY, as torch does)
    static __device__ __forceinline__ float4 round(float4 v) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
        const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    }
};

constexpr int LN_WARPS = 8;

// ------------------------------------------------------------------------------------------ LayerNorm forward
// One warp per row; lane owns columns {128*j + 4*lane .. +3}, j < NV (C = 128*NV).
template <typename AT, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_fwd_kernel(int T, int L, const float *__restrict__ x, const AT *__restrict__ delta, const float *__restrict__ scale,
              const float *__restrict__ pos, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
              float *__restrict__ xs_out, AT *__restrict__ y, float *__restrict__ mean_out, float *__restrict__ rstd_out) {
    constexpr int C = 128 * NV;
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    if (row >= T) return;
    const size_t base = (size_t)row * C;
    const float s = (delta && scale) ? scale[row / L] : 1.f;
    float4 v[NV];
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int col = 128 * j + 4 * lane;
        v[j] = Vec4<float>::load(x + base + col);
        if (delta) {
            const float4 d = Vec4<AT>::load(delta + base + col);
            v[j].x += s * d.x; v[j].y += s * d.y; v[j].z += s * d.z; v[j].w += s * d.w;
        }
        if (pos) {
            const float4 p = Vec4<float>::load(pos + base + col);
            v[j].x += p.x; v[j].y += p.y; v[j].z += p.z; v[j].w += p.w;
        }
        if (xs_out) Vec4<float>::store(xs_out + base + col, v[j]);
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    if (!y) return;
    const float mean = warp_sum(sum) * (1.f / C);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
    if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int col = 128 * j + 4 * lane;
        const float4 g = Vec4<float>::load(gamma + col), b = Vec4<float>::load(beta + col);
        float4 o;
        o.x = (v[j].x - mean) * rstd * g.x + b.x;
        o.y = (v[j].y - mean) * rstd * g.y + b.y;
        o.z = (v[j].z - mean) * rstd * g.z + b.z;
        o.w = (v[j].w - mean) * rstd * g.w + b.w;
        Vec4<AT>::store(y + base + col, o);
    }
}

// ------------------------------------------------------------------------------------------ LayerNorm backward
// dx = g_res + rstd * (dy*gamma - mean_C(dy*gamma) - xhat * mean_C(dy*gamma*xhat))
// plus, in the same pass: dgamma += dy*xhat, dbeta += dy (column sums, one atomic per column per CTA),
// dpos += dx, dscaled = scale[b]*dx cast to AT (the dY of the Linear that produced the residual branch) and its
// column sum into dbias (that Linear's bias gradient).
template <typename AT, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
ln_bwd_kernel(int T, int L, const AT *__restrict__ dy, const float *__restrict__ xs, const float *__restrict__ mean,
              const float *__restrict__ rstd, const float *__restrict__ gamma, const float *__restrict__ g_res,
              const float *__restrict__ scale, float *__restrict__ dx, float *__restrict__ dpos,
              AT *__restrict__ dscaled, float *__restrict__ dgamma, float *__restrict__ dbeta,
              float *__restrict__ dbias) {
    constexpr int C = 128 * NV;
    __shared__ float s_red[LN_WARPS][C];
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 acc_g[NV], acc_b[NV], acc_s[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc_g[j] = acc_b[j] = acc_s[j] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int row = blockIdx.x * LN_WARPS + warp; row < T; row += gridDim.x * LN_WARPS) {
        const size_t base = (size_t)row * C;
        const float mu = mean[row], rs = rstd[row];
        float4 d[NV], xh[NV];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int col = 128 * j + 4 * lane;
            d[j] = Vec4<AT>::load(dy + base + col);
            const float4 xv = Vec4<float>::load(xs + base + col);
            const float4 g = Vec4<float>::load(gamma + col);
            xh[j] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
            acc_g[j].x += d[j].x * xh[j].x; acc_g[j].y += d[j].y * xh[j].y;
            acc_g[j].z += d[j].z * xh[j].z; acc_g[j].w += d[j].w * xh[j].w;
            acc_b[j].x += d[j].x; acc_b[j].y += d[j].y; acc_b[j].z += d[j].z; acc_b[j].w += d[j].w;
            d[j].x *= g.x; d[j].y *= g.y; d[j].z *= g.z; d[j].w *= g.w;          // dy * gamma
            c1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
            c2 += (d[j].x * xh[j].x + d[j].y * xh[j].y) + (d[j].z * xh[j].z + d[j].w * xh[j].w);
        }
        c1 = warp_sum(c1) * (1.f / C);
        c2 = warp_sum(c2) * (1.f / C);
        const float s = (dscaled && scale) ? scale[row / L] : 1.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int col = 128 * j + 4 * lane;
            float4 o;
            o.x = (d[j].x - c1 - xh[j].x * c2) * rs;
            o.y = (d[j].y - c1 - xh[j].y * c2) * rs;
            o.z = (d[j].z - c1 - xh[j].z * c2) * rs;
            o.w = (d[j].w - c1 - xh[j].w * c2) * rs;
            if (g_res) {
                const float4 r = Vec4<float>::load(g_res + base + col);
                o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            Vec4<float>::store(dx + base + col, o);
            if (dpos) {
                float4 p = Vec4<float>::load(dpos + base + col);
                p.x += o.x; p.y += o.y; p.z += o.z; p.w += o.w;
                Vec4<float>::store(dpos + base + col, p);
            }
            if (dscaled) {
                float4 q = make_float4(s * o.x, s * o.y, s * o.z, s * o.w);
                Vec4<AT>::store(dscaled + base + col, q);
                q = Vec4<AT>::round(q);
                acc_s[j].x += q.x; acc_s[j].y += q.y; acc_s[j].z += q.z; acc_s[j].w += q.w;
            }
        }
    }
    // column sums: warps -> shared -> one atomic per column per CTA, one array at a time
#pragma unroll 1
    for (int which = 0; which < 3; ++which) {
        float *dst = which == 0 ? dgamma : which == 1 ? dbeta : dbias;
        if (!dst || (which == 2 && !dscaled)) continue;        // uniform over the CTA
        __syncthreads();
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float4 a = which == 0 ? acc_g[j] : which == 1 ? acc_b[j] : acc_s[j];
            Vec4<float>::store(&s_red[warp][128 * j + 4 * lane], a);
        }
        __syncthreads();
        for (int col = threadIdx.x; col < C; col += LN_WARPS * 32) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < LN_WARPS; ++w) t += s_red[w][col];
            atomicAdd(dst + col, t);
        }
    }
}

// ------------------------------------------------------------------------------------------ GELU (erf form, nn.GELU default)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}

template <typename AT>
__global__ void __launch_bounds__(256) gelu_fwd_kernel(size_t n4, const AT *__restrict__ x, AT *__restrict__ y) {
    pdl_trigger();
    pdl_wait();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = Vec4<AT>::load(x + 4 * i);
        Vec4<AT>::store(y + 4 * i, make_float4(gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w)));
    }
}

// Column-threaded tile kernels: blockDim = (CT column threads, CL_LANES row lanes); a column thread owns 4 consecutive
// columns, the CTA covers CT*4 columns x CL_LANES*CL_U rows; row lane l handles rows r0 + l + CL_LANES*u.  Column sums
// are combined over the row lanes in shared memory, then ONE atomic per column per CTA (the atomics, not the data, bound
// these kernels: T*C/(rows per CTA) of them).
//   MODE 0: out = dy * gelu'(pre)                 (dy, pre: AT)          colsum(out) -> dbias
//   MODE 1: out = scale[row / L] * g  (g: fp32)                           colsum(out) -> dbias
constexpr int CL_LANES = 8, CL_U = 4, CL_MAX_CT = 64;
template <typename AT, int MODE>
__global__ void __launch_bounds__(CL_MAX_CT * CL_LANES)
col_tile_kernel(int T, int L, int Ccols, const void *__restrict__ in0_, const AT *__restrict__ pre,
                const float *__restrict__ scale, AT *__restrict__ out, float *__restrict__ dbias) {
    __shared__ float4 s_acc[CL_LANES][CL_MAX_CT];
    pdl_trigger();
    pdl_wait();
    const int ct = threadIdx.x, lane_r = threadIdx.y;
    const int col = (blockIdx.y * blockDim.x + ct) * 4;
    const bool col_ok = col < Ccols;
    const int r0 = blockIdx.x * (CL_LANES * CL_U);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col_ok) {
        float4 a[CL_U], p[CL_U];
#pragma unroll
        for (int u = 0; u < CL_U; ++u) {
            const int row = r0 + lane_r + CL_LANES * u;
            if (row < T) {
                const size_t o = (size_t)row * Ccols + col;
                if (MODE == 0) {
                    a[u] = Vec4<AT>::load(reinterpret_cast<const AT *>(in0_) + o);
                    p[u] = Vec4<AT>::load(pre + o);
                } else {
                    a[u] = Vec4<float>::load(reinterpret_cast<const float *>(in0_) + o);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CL_U; ++u) {
            const int row = r0 + lane_r + CL_LANES * u;
            if (row >= T) continue;
            float4 q;
            if (MODE == 0) {
                q = make_float4(a[u].x * gelu_grad_f(p[u].x), a[u].y * gelu_grad_f(p[u].y), a[u].z * gelu_grad_f(p[u].z),
                                a[u].w * gelu_grad_f(p[u].w));
            } else {
                const float s = scale ? scale[row / L] : 1.f;
                q = make_float4(s * a[u].x, s * a[u].y, s * a[u].z, s * a[u].w);
            }
            Vec4<AT>::store(out + (size_t)row * Ccols + col, q);
            q = Vec4<AT>::round(q);
            acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
        }
    }
    if (!dbias) return;                        // uniform over the CTA
    s_acc[lane_r][ct] = acc;
    __syncthreads();
    if (lane_r == 0 && col_ok) {
#pragma unroll
        for (int l = 1; l < CL_LANES; ++l) {
            const float4 t = s_acc[l][ct];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        atomicAdd(dbias + col, acc.x); atomicAdd(dbias + col + 1, acc.y);
        atomicAdd(dbias + col + 2, acc.z); atomicAdd(dbias + col + 3, acc.w);
    }
}

// ------------------------------------------------------------------------------------------ optimizer
// Multi-tensor layout: tensor i has numel[i] elements; the work list is cut into chunks of ADAM_CHUNK elements,
// chunk c covers tensor chunk_tensor[c] from element chunk_start[c].  Pointer tables live in device memory.
constexpr int ADAM_CHUNK = 4096;     // elements per chunk (16 KB fp32): 256 threads x 4 float4

__global__ void __launch_bounds__(256)
grad_sumsq_kernel(int n_chunks, const int *__restrict__ chunk_tensor, const int *__restrict__ chunk_start,
                  const long long *__restrict__ numel, const float *const *__restrict__ grads, float grad_scale,
                  double *__restrict__ acc) {
    __shared__ float s_part[8];
    float local = 0.f;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int t = chunk_tensor[c];
        const long long start = chunk_start[c], n = numel[t];
        const float *g = grads[t];
        const long long end = min(n, start + (long long)ADAM_CHUNK);
        if ((((uintptr_t)g) & 15) == 0 && (start & 3) == 0) {
            for (long long i = start + 4 * threadIdx.x; i < end; i += 4 * 256) {
                if (i + 3 < end) {
                    const float4 v = *reinterpret_cast<const float4 *>(g + i);
                    local += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
                } else {
                    for (long long k = i; k < end; ++k) local += g[k] * g[k];
                }
            }
        } else {
            for (long long i = start + threadIdx.x; i < end; i += 256) local += g[i] * g[i];
        }
    }
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 8) {
        float v = s_part[threadIdx.x];
        v += __shfl_xor_sync(0xffu, v, 4);
        v += __shfl_xor_sync(0xffu, v, 2);
        v += __shfl_xor_sync(0xffu, v, 1);
        if (threadIdx.x == 0) atomicAdd(acc, (double)v * (double)grad_scale * (double)grad_scale);
    }
}

// state (device, 32 bytes):  fp64 sum-of-squares accumulator | float step, total_norm, clip_coef, found_inf | u32 ticket | pad
struct AdamHyper { float beta1, beta2, eps, weight_decay, max_norm, grad_scale; };

__device__ __forceinline__ void adam_elem(float &p, float &m, float &v, float g, float lr, float wd, float b1, float b2,
                                          float eps, float step_size, float inv_bc2_sqrt) {
    p -= lr * wd * p;                                  // decoupled weight decay (AdamW)
    m = m + (1.f - b1) * (g - m);                      // lerp(m, g, 1 - beta1)
    v = b2 * v + (1.f - b2) * g * g;
    const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(256)
adamw_kernel(int n_chunks, const int *__restrict__ chunk_tensor, const int *__restrict__ chunk_start,
             const long long *__restrict__ numel, float *const *__restrict__ params, const float *const *__restrict__ grads,
             float *const *__restrict__ exp_avg, float *const *__restrict__ exp_avg_sq,
             __nv_bfloat16 *const *__restrict__ shadows, const int *__restrict__ group, const float *__restrict__ lrs,
             double *__restrict__ acc, float *__restrict__ state, AdamHyper h) {
    // every CTA derives the same scalars from the accumulator and the step counter (no grid sync needed); the LAST
    // CTA to finish -- by then every CTA has read them -- publishes the diagnostics, advances the step counter and
    // clears the accumulator for the next step (ticket in state[4])
    const float total = (float)sqrt(acc[0]);
    const bool bad = !isfinite(total);
    float coef = h.max_norm > 0.f ? h.max_norm / (total + 1e-6f) : 1.f;   // clip_grad_norm_: clamp(max_norm/(norm+1e-6), max=1)
    coef = fminf(coef, 1.f) * h.grad_scale;            // gradients enter as grad_scale * g (1/world under data parallelism)
    const float t = state[0] + 1.f;
    const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
    const float inv_bc2_sqrt = rsqrtf(bc2);
    // bad: train_network.py:336-340 -- skip the step, parameters / moments / step counter untouched
    for (int c = blockIdx.x; c < n_chunks && !bad; c += gridDim.x) {
        const int ti = chunk_tensor[c];
        const long long start = chunk_start[c], n = numel[ti];
        const long long end = min(n, start + (long long)ADAM_CHUNK);
        float *p = params[ti], *m = exp_avg[ti], *v = exp_avg_sq[ti];
        const float *g = grads[ti];
        __nv_bfloat16 *sh = shadows ? shadows[ti] : nullptr;
        const float lr = lrs[group[ti]];
        const float step_size = lr / bc1;
        const bool vec = ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v | (uintptr_t)g) & 15) == 0) && ((start & 3) == 0) &&
                         (sh == nullptr || (((uintptr_t)sh) & 7) == 0);
        if (vec) {
            for (long long i = start + 4 * threadIdx.x; i < end; i += 4 * 256) {
                if (i + 3 < end) {
                    float4 pp = *reinterpret_cast<float4 *>(p + i), mm = *reinterpret_cast<float4 *>(m + i);
                    float4 vv = *reinterpret_cast<float4 *>(v + i);
                    const float4 gg = *reinterpret_cast<const float4 *>(g + i);
                    adam_elem(pp.x, mm.x, vv.x, gg.x * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                    adam_elem(pp.y, mm.y, vv.y, gg.y * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                    adam_elem(pp.z, mm.z, vv.z, gg.z * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                    adam_elem(pp.w, mm.w, vv.w, gg.w * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                    *reinterpret_cast<float4 *>(p + i) = pp;
                    *reinterpret_cast<float4 *>(m + i) = mm;
                    *reinterpret_cast<float4 *>(v + i) = vv;
                    if (sh) Vec4<__nv_bfloat16>::store(sh + i, pp);
                } else {
                    for (long long k = i; k < end; ++k) {
                        float pp = p[k], mm = m[k], vv = v[k];
                        adam_elem(pp, mm, vv, g[k] * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                        p[k] = pp; m[k] = mm; v[k] = vv;
                        if (sh) sh[k] = __float2bfloat16_rn(pp);
                    }
                }
            }
        } else {
            for (long long k = start + threadIdx.x; k < end; k += 256) {
                float pp = p[k], mm = m[k], vv = v[k];
                adam_elem(pp, mm, vv, g[k] * coef, lr, h.weight_decay, h.beta1, h.beta2, h.eps, step_size, inv_bc2_sqrt);
                p[k] = pp; m[k] = mm; v[k] = vv;
                if (sh) sh[k] = __float2bfloat16_rn(pp);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned *ticket = reinterpret_cast<unsigned *>(state + 4);
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            state[1] = total;
            state[2] = bad ? 0.f : coef / h.grad_scale;
            state[3] = bad ? 1.f : 0.f;
            if (!bad) state[0] = t;            // the step counter only advances on applied steps
            acc[0] = 0.0;                      // ready for the next step's sum of squares
            *ticket = 0u;
            __threadfence();
        }
    }
}

template <typename AT>
static int launch_ln_fwd(int NV, int T, int L, const float *x, const void *delta, const float *scale, const float *pos,
                         const float *gamma, const float *beta, float eps, float *xs, void *y, float *mean, float *rstd,
                         cudaStream_t st) {
    const int grid = div_up(T, LN_WARPS);
    cudaError_t e = cudaSuccess;
#define LN_FWD_CASE(N)                                                                                           \
    case N:                                                                                                      \
        e = launch_pdl(ln_fwd_kernel<AT, N>, dim3(grid), dim3(LN_WARPS * 32), 0, st, T, L, x, (const AT *)delta, scale, pos, \
                       gamma, beta, eps, xs, (AT *)y, mean, rstd);                                               \
        break;
    switch (NV) {
        LN_FWD_CASE(1) LN_FWD_CASE(2) LN_FWD_CASE(3) LN_FWD_CASE(4) LN_FWD_CASE(6) LN_FWD_CASE(8)
        default: return set_error("up3d_ln_fwd: width %d not supported (128 x {1,2,3,4,6,8})", NV * 128);
    }
#undef LN_FWD_CASE
    if (e != cudaSuccess) return set_error("launch of ln_fwd_kernel failed: %s", cudaGetErrorString(e));
    return 0;
}

template <typename AT>
static int launch_ln_bwd(int NV, int T, int L, const void *dy, const float *xs, const float *mean, const float *rstd,
                         const float *gamma, const float *g_res, const float *scale, float *dx, float *dpos, void *dscaled,
                         float *dgamma, float *dbeta, float *dbias, cudaStream_t st) {
    const int grid = min(div_up(T, LN_WARPS), UP3D_NUM_SMS);     // one row per warp (measured faster than two)
    cudaError_t e = cudaSuccess;
#define LN_BWD_CASE(N)                                                                                              \
    case N:                                                                                                         \
        e = launch_pdl(ln_bwd_kernel<AT, N>, dim3(grid), dim3(LN_WARPS * 32), 0, st, T, L, (const AT *)dy, xs, mean, rstd,  \
                       gamma, g_res, scale, dx, dpos, (AT *)dscaled, dgamma, dbeta, dbias);                         \
        break;
    switch (NV) {
        LN_BWD_CASE(1) LN_BWD_CASE(2) LN_BWD_CASE(3) LN_BWD_CASE(4) LN_BWD_CASE(6) LN_BWD_CASE(8)
        default: return set_error("up3d_ln_bwd: width %d not supported (128 x {1,2,3,4,6,8})", NV * 128);
    }
#undef LN_BWD_CASE
    if (e != cudaSuccess) return set_error("launch of ln_bwd_kernel failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace up3d

using namespace up3d;

static bool aligned16(const void *p) { return (((uintptr_t)p) & 15) == 0; }

extern "C" int up3d_ln_fwd(int act_bf16, int T, int L, int C, const float *x, const void *delta, const float *scale,
                           const float *pos, const float *gamma, const float *beta, float eps, float *xs_out, void *y,
                           float *mean, float *rstd, up3d_stream_t stream) {
    UP3D_CHECK_ARG(T >= 0 && L > 0 && C > 0 && C % 128 == 0, "up3d_ln_fwd: bad sizes T=%d L=%d C=%d", T, L, C);
    if (T == 0) return 0;
    UP3D_CHECK_ARG(x != nullptr, "up3d_ln_fwd: x is NULL");
    UP3D_CHECK_ARG(y == nullptr || (gamma && beta && mean && rstd), "up3d_ln_fwd: y needs gamma/beta/mean/rstd");
    UP3D_CHECK_ARG(y != nullptr || xs_out != nullptr, "up3d_ln_fwd: nothing to write");
    UP3D_CHECK_ARG(aligned16(x) && aligned16(delta) && aligned16(pos) && aligned16(gamma) && aligned16(beta) &&
                   aligned16(xs_out) && aligned16(y), "up3d_ln_fwd: pointers must be 16-byte aligned");
    int rc = act_bf16 ? launch_ln_fwd<__nv_bfloat16>(C / 128, T, L, x, delta, scale, pos, gamma, beta, eps, xs_out, y, mean,
                                                     rstd, (cudaStream_t)stream)
                      : launch_ln_fwd<float>(C / 128, T, L, x, delta, scale, pos, gamma, beta, eps, xs_out, y, mean, rstd,
                                             (cudaStream_t)stream);
    if (rc) return rc;
    UP3D_LAUNCH_OK("ln_fwd_kernel");
    return 0;
}

extern "C" int up3d_ln_bwd(int act_bf16, int T, int L, int C, const void *dy, const float *xs, const float *mean,
                           const float *rstd, const float *gamma, const float *g_res, const float *scale, float *dx,
                           float *dpos, void *dscaled, float *dgamma, float *dbeta, float *dbias, up3d_stream_t stream) {
    UP3D_CHECK_ARG(T >= 0 && L > 0 && C > 0 && C % 128 == 0, "up3d_ln_bwd: bad sizes T=%d L=%d C=%d", T, L, C);
    if (T == 0) return 0;
    UP3D_CHECK_ARG(dy && xs && mean && rstd && gamma && dx, "up3d_ln_bwd: NULL pointer");
    UP3D_CHECK_ARG(aligned16(dy) && aligned16(xs) && aligned16(gamma) && aligned16(g_res) && aligned16(dx) &&
                   aligned16(dpos) && aligned16(dscaled), "up3d_ln_bwd: pointers must be 16-byte aligned");
    int rc = act_bf16 ? launch_ln_bwd<__nv_bfloat16>(C / 128, T, L, dy, xs, mean, rstd, gamma, g_res, scale, dx, dpos, dscaled,
                                                     dgamma, dbeta, dbias, (cudaStream_t)stream)
                      : launch_ln_bwd<float>(C / 128, T, L, dy, xs, mean, rstd, gamma, g_res, scale, dx, dpos, dscaled, dgamma,
                                             dbeta, dbias, (cudaStream_t)stream);
    if (rc) return rc;
    UP3D_LAUNCH_OK("ln_bwd_kernel");
    return 0;
}

extern "C" int up3d_gelu_fwd(int act_bf16, int64_t n, const void *x, void *y, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n >= 0 && n % 4 == 0, "up3d_gelu_fwd: n must be a multiple of 4");
    if (n == 0) return 0;
    UP3D_CHECK_ARG(x && y && aligned16(x) && aligned16(y), "up3d_gelu_fwd: NULL or misaligned pointer");
    const size_t n4 = (size_t)n / 4;
    const int grid = (int)min((size_t)UP3D_NUM_SMS * 8, (n4 + 255) / 256);
    if (act_bf16)
        UP3D_CUDA_OK(launch_pdl(gelu_fwd_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, n4,
                                (const __nv_bfloat16 *)x, (__nv_bfloat16 *)y));
    else
        UP3D_CUDA_OK(launch_pdl(gelu_fwd_kernel<float>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, n4, (const float *)x,
                                (float *)y));
    UP3D_LAUNCH_OK("gelu_fwd_kernel");
    return 0;
}

template <int MODE>
static int launch_col_tile(int act_bf16, int T, int L, int C, const void *in0, const void *pre, const float *scale, void *out,
                           float *dbias, cudaStream_t st) {
    int ct = 32;                             // largest warp multiple <= CL_MAX_CT dividing the C/4 column groups
    for (int t = CL_MAX_CT; t >= 32; t -= 32)
        if ((C / 4) % t == 0) { ct = t; break; }
    const dim3 block(ct, CL_LANES), grid(div_up(T, CL_LANES * CL_U), div_up(C / 4, ct));
    if (act_bf16)
        UP3D_CUDA_OK(launch_pdl(col_tile_kernel<__nv_bfloat16, MODE>, grid, block, 0, st, T, L, C, in0,
                                (const __nv_bfloat16 *)pre, scale, (__nv_bfloat16 *)out, dbias));
    else
        UP3D_CUDA_OK(launch_pdl(col_tile_kernel<float, MODE>, grid, block, 0, st, T, L, C, in0, (const float *)pre, scale,
                                (float *)out, dbias));
    return 0;
}

extern "C" int up3d_gelu_bwd(int act_bf16, int T, int C, const void *dy, const void *pre, void *dx, float *dbias,
                             up3d_stream_t stream) {
    UP3D_CHECK_ARG(T >= 0 && C > 0 && C % 4 == 0, "up3d_gelu_bwd: bad sizes");
    if (T == 0) return 0;
    UP3D_CHECK_ARG(dy && pre && dx && aligned16(dy) && aligned16(pre) && aligned16(dx), "up3d_gelu_bwd: NULL or misaligned pointer");
    UP3D_CHECK_ARG(!act_bf16 || C % 8 == 0, "up3d_gelu_bwd: bf16 rows must be a multiple of 8 wide");
    launch_col_tile<0>(act_bf16, T, 1, C, dy, pre, nullptr, dx, dbias, (cudaStream_t)stream);
    UP3D_LAUNCH_OK("col_tile_kernel<gelu_bwd>");
    return 0;
}

extern "C" int up3d_scale_cast_colsum(int act_bf16, int T, int L, int C, const float *g, const float *scale, void *out,
                                      float *dbias, up3d_stream_t stream) {
    UP3D_CHECK_ARG(T >= 0 && L > 0 && C > 0 && C % 4 == 0, "up3d_scale_cast_colsum: bad sizes");
    if (T == 0) return 0;
    UP3D_CHECK_ARG(g && out && aligned16(g) && aligned16(out), "up3d_scale_cast_colsum: NULL or misaligned pointer");
    UP3D_CHECK_ARG(!act_bf16 || C % 8 == 0, "up3d_scale_cast_colsum: bf16 rows must be a multiple of 8 wide");
    launch_col_tile<1>(act_bf16, T, L, C, g, nullptr, scale, out, dbias, (cudaStream_t)stream);
    UP3D_LAUNCH_OK("col_tile_kernel<scale_cast>");
    return 0;
}

extern "C" int up3d_adamw_chunk_elems(void) { return ADAM_CHUNK; }

static int adamw_check(int n_tensors, int n_chunks, const void *a, const void *b, const void *c, const void *state) {
    UP3D_CHECK_ARG(n_tensors >= 0 && n_chunks >= 0, "up3d_adamw: bad sizes");
    UP3D_CHECK_ARG(n_tensors == 0 || n_chunks == 0 || (a && b && c && state), "up3d_adamw: NULL pointer");
    UP3D_CHECK_ARG((((uintptr_t)state) & 7) == 0, "up3d_adamw: state must be 8-byte aligned");
    return 0;
}

extern "C" int up3d_grad_sumsq(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                               const int64_t *numel, const float *const *grads, float grad_scale, void *state,
                               up3d_stream_t stream) {
    if (int rc = adamw_check(n_tensors, n_chunks, chunk_tensor, chunk_start, numel, state)) return rc;
    if (n_tensors == 0 || n_chunks == 0) return 0;
    UP3D_CHECK_ARG(grads != nullptr, "up3d_grad_sumsq: NULL pointer");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    const int grid = min(n_chunks, UP3D_NUM_SMS * 8);
    grad_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_chunks, chunk_tensor, chunk_start, (const long long *)numel,
                                                             grads, grad_scale, (double *)state);
    UP3D_LAUNCH_OK("grad_sumsq_kernel");
    return 0;
}

extern "C" int up3d_adamw_apply(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                                const int64_t *numel, float *const *params, const float *const *grads, float *const *exp_avg,
                                float *const *exp_avg_sq, void *const *bf16_shadows, const int32_t *group, const float *lrs,
                                float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                                void *state, up3d_stream_t stream) {
    if (int rc = adamw_check(n_tensors, n_chunks, chunk_tensor, chunk_start, numel, state)) return rc;
    if (n_tensors == 0 || n_chunks == 0) return 0;
    UP3D_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && group && lrs, "up3d_adamw_apply: NULL pointer");
    UP3D_CHECK_ARG(grad_scale > 0.f, "up3d_adamw_apply: grad_scale must be positive");
    const int grid = min(n_chunks, UP3D_NUM_SMS * 8);
    AdamHyper h{beta1, beta2, eps, weight_decay, max_norm, grad_scale};
    adamw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n_chunks, chunk_tensor, chunk_start, (const long long *)numel, params,
                                                        grads, exp_avg, exp_avg_sq, (__nv_bfloat16 *const *)bf16_shadows,
                                                        group, lrs, (double *)state, (float *)state + 2, h);
    UP3D_LAUNCH_OK("adamw_kernel");
    return 0;
}

extern "C" int up3d_adamw_step(int n_tensors, int n_chunks, const int32_t *chunk_tensor, const int32_t *chunk_start,
                               const int64_t *numel, float *const *params, const float *const *grads, float *const *exp_avg,
                               float *const *exp_avg_sq, void *const *bf16_shadows, const int32_t *group, const float *lrs,
                               float beta1, float beta2, float eps, float weight_decay, float max_norm, float grad_scale,
                               void *state, up3d_stream_t stream) {
    if (int rc = up3d_grad_sumsq(n_tensors, n_chunks, chunk_tensor, chunk_start, numel, grads, grad_scale, state, stream)) return rc;
    return up3d_adamw_apply(n_tensors, n_chunks, chunk_tensor, chunk_start, numel, params, grads, exp_avg, exp_avg_sq,
                            bf16_shadows, group, lrs, beta1, beta2, eps, weight_decay, max_norm, grad_scale, state, stream);
}
