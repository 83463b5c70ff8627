// stem.cu -- GroupNorm statistics of the frozen image stem's feature field without materialising it.
//
// The reference feeds `decoder_block_3` of a frozen SD-VAE (model/image_predictor.py:56-81; weights not shipped) through
// the trainable image_conv = GroupNorm(32, 128, eps=1e-6) + Conv2d 1x1 (model/gaussian_predictor.py:61-66, 139).  The
// repo's weight-free stand-in (gaussian_predictor.FrozenImageStem) defines the 128-channel field analytically:
//     f[n, c, y, x] = sin(proj[c,:] . image[n, :, y, x] + shift[c]).
// FeatureFusion reads 128 pixels per object from Conv1x1(GroupNorm(f)) (fusion/feat_fusion.py:121-131), so the only
// full-image quantity is GroupNorm's per-(image, group) mean / variance.  This kernel produces those sums straight from
// the 3-channel image (6 MB at 8 x 256^2) instead of writing, re-reading and reducing the 268 MB dense field.
#include "common.cuh"

namespace up3d {

constexpr int STEM_MAX_C = 256;

// grid (ceil(HW / (256*STEM_PIX)), n), block 256: a thread accumulates STEM_PIX pixels (stride 256: coalesced) before
// the warp reduction, so the shuffles / shared atomics are amortised.  sums (n, G, 2) fp64: [sum f, sum f^2] over the
// group's channels and all pixels.
constexpr int STEM_PIX = 8;
__global__ void __launch_bounds__(256)
stem_group_stats_kernel(int HW, int Cc, int G, const float *__restrict__ image, const float *__restrict__ proj,
                        const float *__restrict__ shift, double *__restrict__ sums) {
    __shared__ float4 s_w[STEM_MAX_C];          // proj row + shift
    __shared__ float s_acc[2 * STEM_MAX_C];     // per group: sum, sum of squares (G <= C)
    const int n = blockIdx.y, pix0 = blockIdx.x * (256 * STEM_PIX) + threadIdx.x;
    for (int c = threadIdx.x; c < Cc; c += 256) s_w[c] = make_float4(proj[3 * c], proj[3 * c + 1], proj[3 * c + 2], shift[c]);
    for (int i = threadIdx.x; i < 2 * G; i += 256) s_acc[i] = 0.f;
    __syncthreads();
    const float *img = image + (size_t)n * 3 * HW;
    float x0[STEM_PIX], x1[STEM_PIX], x2[STEM_PIX];
    int n_valid = 0;
#pragma unroll
    for (int j = 0; j < STEM_PIX; ++j) {
        const int pix = pix0 + 256 * j;
        const bool valid = pix < HW;
        x0[j] = valid ? img[pix] : 0.f; x1[j] = valid ? img[HW + pix] : 0.f; x2[j] = valid ? img[2 * HW + pix] : 0.f;
        n_valid += valid ? 1 : 0;               // valid pixels are a prefix of j
    }
    const int cpg = Cc / G;
    for (int g = 0; g < G; ++g) {
        float s = 0.f, q = 0.f;
        for (int k = 0; k < cpg; ++k) {
            const float4 w = s_w[g * cpg + k];
#pragma unroll
            for (int j = 0; j < STEM_PIX; ++j) {
                // MUFU sine (|arg| is O(1) here: abs error ~1e-6, and averaged over the 2.6e5 samples of a group it is
                // far below the statistics' own fp32 rounding); an explicit rintf range reduction would double the
                // special-function-unit work that bounds this kernel
                const float a = fmaf(w.x, x0[j], fmaf(w.y, x1[j], fmaf(w.z, x2[j], w.w)));
                const float v = j < n_valid ? __sinf(a) : 0.f;
                s += v;
                q = fmaf(v, v, q);
            }
        }
        s = warp_sum(s);
        q = warp_sum(q);
        if ((threadIdx.x & 31) == 0) {
            atomicAdd(&s_acc[2 * g], s);
            atomicAdd(&s_acc[2 * g + 1], q);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * G; i += 256) atomicAdd(&sums[(size_t)n * 2 * G + i], (double)s_acc[i]);
}

}  // namespace up3d

using namespace up3d;

extern "C" int up3d_stem_group_stats(int n_images, int H, int W, int C, int G, const float *image, const float *proj,
                                     const float *shift, double *sums, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n_images >= 0 && H > 0 && W > 0, "up3d_stem_group_stats: bad image size");
    UP3D_CHECK_ARG(C > 0 && C <= STEM_MAX_C && G > 0 && C % G == 0, "up3d_stem_group_stats: need 0 < C <= %d, C %% G == 0",
                   STEM_MAX_C);
    if (n_images == 0) return 0;
    UP3D_CHECK_ARG(image && proj && shift && sums, "up3d_stem_group_stats: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    UP3D_CUDA_OK(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * G * (size_t)n_images, st));
    const int HW = H * W;
    stem_group_stats_kernel<<<dim3(div_up(HW, 256 * STEM_PIX), n_images), 256, 0, st>>>(HW, C, G, image, proj, shift, sums);
    UP3D_LAUNCH_OK("stem_group_stats_kernel");
    return 0;
}
