// head.cu -- the small-tensor ends of the model around the backbone, each as one launch instead of ~30-70 eager ops:
//
//  * splat head: GaussianSplatPredictor._process_network_output (/root/reference/model/gaussian_predictor.py:279-328,
//    activations 249-254) + the concatenation render_predicted does (gaussian_renderer/__init__.py:66-69):
//    raw (B,P,14+3(M-1)) head output + centres (B,P,3) -> the rasterizer's inputs, forward and backward.
//  * fusion projection: FeatureFusion (fusion/feat_fusion.py:23-56, 88-131): project the group centres into the
//    source view, pixel rounding, in-image test, per-pixel nearest-depth test, and the image feature at the kept pixels
//    -- for the analytic stem field -- normalised with image_conv's GroupNorm statistics.  No gradient flows here.
//
// Both work on B x 128 points: launch-latency, not bandwidth, is what they cost; one CTA per object.
#include "common.cuh"

namespace up3d {

constexpr int HEAD_THREADS = 256;

__device__ __forceinline__ float block_sum_256(float v, float *sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < HEAD_THREADS / 32; ++i) t += sh[i];
    return t;
}

// raw row layout (gaussian_predictor.py:141-143 split_dimensions): xyz 0..2 | opacity 3 | scaling 4..6 | rotation 7..10 |
// features_dc 11..13 | features_rest 14.. ((M-1)*3 values, (k, c) row-major).
// Quirks reproduced: rotation is F.normalize(x, dim=-1, eps=1e-6) on the (B,4,P) tensor, i.e. each of the 4
// quaternion components is normalised over the P POINTS; scaling = exp(clamp(x, -1, 20)); isotropic broadcasts channel 4.
__global__ void __launch_bounds__(HEAD_THREADS)
splat_head_fwd_kernel(int P, int M, int Cr, const float *__restrict__ raw, const float *__restrict__ center, float offset_scale,
                      int isotropic, float *__restrict__ xyz, float *__restrict__ opacity, float *__restrict__ scaling,
                      float *__restrict__ rotation, float *__restrict__ shs, float *__restrict__ rot_norm) {
    __shared__ float sh[HEAD_THREADS / 32];
    __shared__ float s_inv[4];
    const int b = blockIdx.x;
    const float *r0 = raw + (size_t)b * P * Cr;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p = threadIdx.x; p < P; p += HEAD_THREADS) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float v = r0[(size_t)p * Cr + 7 + j]; acc[j] = fmaf(v, v, acc[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = block_sum_256(acc[j], sh);
        if (threadIdx.x == 0) {
            const float n = fmaxf(sqrtf(t), 1e-6f);          // F.normalize: x / max(||x||, eps)
            rot_norm[b * 4 + j] = n;
            s_inv[j] = n;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += HEAD_THREADS) {
        const float *r = r0 + (size_t)p * Cr;
        const size_t g = (size_t)b * P + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) xyz[g * 3 + c] = tanhf(r[c]) * offset_scale + center[g * 3 + c];
        opacity[g] = 1.f / (1.f + expf(-r[3]));
#pragma unroll
        for (int c = 0; c < 3; ++c) scaling[g * 3 + c] = expf(fminf(fmaxf(r[4 + (isotropic ? 0 : c)], -1.f), 20.f));
#pragma unroll
        for (int j = 0; j < 4; ++j) rotation[g * 4 + j] = r[7 + j] / s_inv[j];
        for (int k = 0; k < M * 3; ++k) shs[g * (size_t)(M * 3) + k] = r[11 + k];
    }
}

__global__ void __launch_bounds__(HEAD_THREADS)
splat_head_bwd_kernel(int P, int M, int Cr, const float *__restrict__ raw, const float *__restrict__ rot_norm, float offset_scale,
                      int isotropic, const float *__restrict__ d_xyz, const float *__restrict__ d_opacity,
                      const float *__restrict__ d_scaling, const float *__restrict__ d_rotation, const float *__restrict__ d_shs,
                      float *__restrict__ d_raw) {
    __shared__ float sh[HEAD_THREADS / 32];
    __shared__ float s_dot[4];
    const int b = blockIdx.x;
    const float *r0 = raw + (size_t)b * P * Cr;
    float n[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) n[j] = rot_norm[b * 4 + j];
    // normalize backward over the point axis: dx_p = (g_p - y_p * sum_q g_q y_q) / n   (n > eps; else dx = g / eps)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (d_rotation) {
        for (int p = threadIdx.x; p < P; p += HEAD_THREADS) {
            const size_t g = (size_t)b * P + p;
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(d_rotation[g * 4 + j], r0[(size_t)p * Cr + 7 + j] / n[j], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float t = block_sum_256(acc[j], sh);
        if (threadIdx.x == 0) s_dot[j] = n[j] > 1e-6f ? t : 0.f;     // clamped norm: no dependence on x through n
    }
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += HEAD_THREADS) {
        const float *r = r0 + (size_t)p * Cr;
        float *d = d_raw + ((size_t)b * P + p) * Cr;
        const size_t g = (size_t)b * P + p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float t = tanhf(r[c]);
            d[c] = d_xyz ? d_xyz[g * 3 + c] * offset_scale * (1.f - t * t) : 0.f;
        }
        const float s = 1.f / (1.f + expf(-r[3]));
        d[3] = d_opacity ? d_opacity[g] * s * (1.f - s) : 0.f;
        float dsc[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = r[4 + (isotropic ? 0 : c)];
            const bool pass = x >= -1.f && x <= 20.f;            // torch.clamp passes the gradient on [min, max]
            const float gsc = d_scaling ? d_scaling[g * 3 + c] : 0.f;
            dsc[c] = pass ? gsc * expf(x) : 0.f;
        }
        if (isotropic) { d[4] = dsc[0] + dsc[1] + dsc[2]; d[5] = 0.f; d[6] = 0.f; }
        else { d[4] = dsc[0]; d[5] = dsc[1]; d[6] = dsc[2]; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gr = d_rotation ? d_rotation[g * 4 + j] : 0.f;
            d[7 + j] = (gr - (r[7 + j] / n[j]) * s_dot[j]) / n[j];
        }
        for (int k = 0; k < M * 3; ++k) d[11 + k] = d_shs ? d_shs[g * (size_t)(M * 3) + k] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------ fusion projection
constexpr int FUS_MAX_N = 1024;     // group centres per object held in shared memory
constexpr int FUS_MAX_G = 64;

// grid (objects, slices): blockIdx.x = object b.  center (B,N,3); w2c (B,16) row-major world-to-camera (the inverse the host computed);
// image (B,3,H,W) is the source view of object b; the analytic stem field f[c] = sin(proj[c,:].rgb + shift[c]) is
// normalised with image_conv's GroupNorm statistics (sums (B,G,2) fp64 = [sum f, sum f^2] per group).
// Outputs: keep (B,N) uint8, pix (B,N,2) int32 (ix, iy; 0 where the point is outside), xhat (B,N,Cc) fp32 =
// GroupNorm-normalised field at image[b, :, ix, iy]  (feat_fusion.py's own `[b, :, pixel_x, pixel_y]` indexing).
__global__ void __launch_bounds__(HEAD_THREADS)
fusion_project_kernel(int N, int H, int W, int Cc, int G, float fx, float fy, float cx, float cy, float eps,
                      const float *__restrict__ center, const float *__restrict__ w2c, const float *__restrict__ image,
                      const float *__restrict__ proj, const float *__restrict__ shift, const double *__restrict__ sums,
                      unsigned char *__restrict__ keep, int *__restrict__ pix, float *__restrict__ xhat) {
    __shared__ int s_cell[FUS_MAX_N];
    __shared__ float s_depth[FUS_MAX_N];
    __shared__ float s_rgb[FUS_MAX_N][3];
    __shared__ float s_mean[FUS_MAX_G], s_rstd[FUS_MAX_G];
    const int b = blockIdx.x;
    const float *m = w2c + b * 16;
    for (int p = threadIdx.x; p < N; p += HEAD_THREADS) {
        const float x = center[((size_t)b * N + p) * 3], y = center[((size_t)b * N + p) * 3 + 1], z = center[((size_t)b * N + p) * 3 + 2];
        float cam[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {                           // (w2c @ [x y z 1]^T)_i, sequential accumulation over k
            float a = __fmul_rn(m[4 * i], x);
            a = __fmaf_rn(m[4 * i + 1], y, a);
            a = __fmaf_rn(m[4 * i + 2], z, a);
            cam[i] = __fmaf_rn(m[4 * i + 3], 1.f, a);
        }
        const float px = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cam[0], fx), cam[2]), cx));    // torch.round: half to even
        const float py = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cam[1], fy), cam[2]), cy));
        const bool inside = (px >= 0.f) && (py >= 0.f) && (px < (float)H) && (py < (float)W) && (cam[2] >= 0.f);
        const int ix = inside ? (int)px : 0, iy = inside ? (int)py : 0;
        s_cell[p] = inside ? iy * H + ix : -1;
        s_depth[p] = cam[2];
        if (blockIdx.y == 0) {
            pix[((size_t)b * N + p) * 2] = ix;
            pix[((size_t)b * N + p) * 2 + 1] = iy;
        }
        const float *img = image + (size_t)b * 3 * H * W + (size_t)ix * W + iy;             // image[b, :, ix, iy]
        s_rgb[p][0] = img[0]; s_rgb[p][1] = img[(size_t)H * W]; s_rgb[p][2] = img[(size_t)2 * H * W];
    }
    const int cpg = Cc / G;
    for (int g = threadIdx.x; g < G; g += HEAD_THREADS) {
        const double cnt = (double)cpg * H * W;
        const double mean = sums[((size_t)b * G + g) * 2] / cnt;
        double var = sums[((size_t)b * G + g) * 2 + 1] / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[g] = (float)mean;
        s_rstd[g] = rsqrtf(__fadd_rn((float)var, eps));
    }
    __syncthreads();
    for (int p = threadIdx.x; p < N; p += HEAD_THREADS) {
        const int c = s_cell[p];
        bool k = c >= 0;
        if (k) {
            float dmin = s_depth[p];
            for (int q = 0; q < N; ++q)
                if (s_cell[q] == c) dmin = fminf(dmin, s_depth[q]);
            k = s_depth[p] == dmin;                             // nearest point of the pixel (ties keep all, as the reference)
        }
        if (blockIdx.y == 0) keep[(size_t)b * N + p] = k ? 1 : 0;
    }
    // the N x C field evaluation (one sinf per element) is spread over gridDim.y CTAs per object; each of them redoes the
    // cheap projection above (with one CTA per object this loop alone was 20 of the kernel's 25 us)
    for (int i = blockIdx.y * HEAD_THREADS + threadIdx.x; i < N * Cc; i += HEAD_THREADS * gridDim.y) {
        const int p = i / Cc, c = i - p * Cc, g = c / cpg;
        float a = __fmul_rn(s_rgb[p][0], proj[3 * c]);
        a = __fmaf_rn(s_rgb[p][1], proj[3 * c + 1], a);
        a = __fmaf_rn(s_rgb[p][2], proj[3 * c + 2], a);
        const float v = sinf(__fadd_rn(a, shift[c]));
        xhat[((size_t)b * N + p) * Cc + c] = __fmul_rn(__fsub_rn(v, s_mean[g]), s_rstd[g]);
    }
}

}  // namespace up3d

using namespace up3d;

extern "C" int up3d_splat_head_fwd(int B, int P, int M, const float *raw, const float *center, float offset_scale,
                                   int isotropic, float *xyz, float *opacity, float *scaling, float *rotation, float *shs,
                                   float *rot_norm, up3d_stream_t stream) {
    UP3D_CHECK_ARG(B >= 0 && P > 0 && M >= 1, "up3d_splat_head_fwd: bad sizes");
    if (B == 0) return 0;
    UP3D_CHECK_ARG(raw && center && xyz && opacity && scaling && rotation && shs && rot_norm, "up3d_splat_head_fwd: NULL pointer");
    splat_head_fwd_kernel<<<B, HEAD_THREADS, 0, (cudaStream_t)stream>>>(P, M, 11 + 3 * M, raw, center, offset_scale, isotropic, xyz,
                                                                         opacity, scaling, rotation, shs, rot_norm);
    UP3D_LAUNCH_OK("splat_head_fwd_kernel");
    return 0;
}

extern "C" int up3d_splat_head_bwd(int B, int P, int M, const float *raw, const float *rot_norm, float offset_scale,
                                   int isotropic, const float *d_xyz, const float *d_opacity, const float *d_scaling,
                                   const float *d_rotation, const float *d_shs, float *d_raw, up3d_stream_t stream) {
    UP3D_CHECK_ARG(B >= 0 && P > 0 && M >= 1, "up3d_splat_head_bwd: bad sizes");
    if (B == 0) return 0;
    UP3D_CHECK_ARG(raw && rot_norm && d_raw, "up3d_splat_head_bwd: NULL pointer");
    splat_head_bwd_kernel<<<B, HEAD_THREADS, 0, (cudaStream_t)stream>>>(P, M, 11 + 3 * M, raw, rot_norm, offset_scale, isotropic,
                                                                         d_xyz, d_opacity, d_scaling, d_rotation, d_shs, d_raw);
    UP3D_LAUNCH_OK("splat_head_bwd_kernel");
    return 0;
}

extern "C" int up3d_fusion_project(int B, int N, int H, int W, int C, int G, float fx, float fy, float cx, float cy, float eps,
                                   const float *center, const float *w2c, const float *image, const float *proj,
                                   const float *shift, const double *sums, unsigned char *keep, int32_t *pix, float *xhat,
                                   up3d_stream_t stream) {
    UP3D_CHECK_ARG(B >= 0 && N > 0 && H > 0 && W > 0, "up3d_fusion_project: bad sizes");
    UP3D_CHECK_ARG(N <= FUS_MAX_N, "up3d_fusion_project: at most %d centres per object (got %d)", FUS_MAX_N, N);
    UP3D_CHECK_ARG(C > 0 && G > 0 && G <= FUS_MAX_G && C % G == 0, "up3d_fusion_project: need C %% G == 0, G <= %d", FUS_MAX_G);
    if (B == 0) return 0;
    UP3D_CHECK_ARG(center && w2c && image && proj && shift && sums && keep && pix && xhat, "up3d_fusion_project: NULL pointer");
    const int slices = max(1, min(16, div_up(N * C, HEAD_THREADS * 8)));
    fusion_project_kernel<<<dim3(B, slices), HEAD_THREADS, 0, (cudaStream_t)stream>>>(N, H, W, C, G, fx, fy, cx, cy, eps, center, w2c,
                                                                                       image, proj, shift, sums, keep, pix, xhat);
    UP3D_LAUNCH_OK("fusion_project_kernel");
    return 0;
}
