// loss.cu -- fused focal-L2 image loss (value + gradient in one pass).
//
// Restates /root/reference/utils/loss_utils.py:23-45 (focal_l2_loss) as called from
// /root/reference/train_network.py:276-283: per pixel the weight is normed_bg_rate when ALL three gt
// channels are `isclose` to the background colour (atol 1e-6, torch's default rtol 1e-5), else
// normed_non_bg_rate; loss = mean(w * (render - gt)^2) over (n,3,H,W).
#include "common.cuh"

namespace up3d {

__device__ __forceinline__ bool isclose_bg(float g, float bg) {
    // torch.isclose: |a - b| <= atol + rtol * |b| with a = gt, b = bg_color
    return fabsf(g - bg) <= 1e-6f + 1e-5f * fabsf(bg);
}

// gt may be the 8-bit image as decoded from the dataset (value / 255, the division the reference's loader does on the
// host) and may be a strided view: image n = (object b = n / views_per_obj, view v = n % views_per_obj) lives at
// gt + b * gt_obj_stride + v * 3*HW  (the trainer's gt_images[:, input_images:] slice, never copied).
template <typename GT> __device__ __forceinline__ float gt_value(GT v);
template <> __device__ __forceinline__ float gt_value<float>(float v) { return v; }
template <> __device__ __forceinline__ float gt_value<unsigned char>(unsigned char v) { return __fdiv_rn((float)v, 255.f); }

template <typename GT>
__global__ void __launch_bounds__(256)
focal_l2_kernel(long long n_pix_total, long long HW, const float *__restrict__ rendered, const GT *__restrict__ gt,
                long long views_per_obj, long long gt_obj_stride, const float *__restrict__ bg, float w_bg, float w_fg,
                float inv_count, double *__restrict__ loss_acc, float *__restrict__ dL) {
    __shared__ float s_part[8];
    const float b0 = bg[0], b1 = bg[1], b2 = bg[2];
    float local = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pix_total;
         p += (long long)gridDim.x * blockDim.x) {
        const long long img = p / HW, pix = p - img * HW;
        const size_t o = (size_t)img * 3 * HW + pix;
        const long long ob = img / views_per_obj, ov = img - ob * views_per_obj;
        const size_t go = (size_t)ob * gt_obj_stride + (size_t)ov * 3 * HW + pix;
        const float g0 = gt_value<GT>(gt[go]), g1 = gt_value<GT>(gt[go + HW]), g2 = gt_value<GT>(gt[go + 2 * HW]);
        const float r0 = rendered[o], r1 = rendered[o + HW], r2 = rendered[o + 2 * HW];
        const float w = (isclose_bg(g0, b0) && isclose_bg(g1, b1) && isclose_bg(g2, b2)) ? w_bg : w_fg;
        const float d0 = r0 - g0, d1 = r1 - g1, d2 = r2 - g2;
        local += w * (d0 * d0 + d1 * d1 + d2 * d2);
        if (dL) {
            const float s = 2.f * w * inv_count;
            dL[o] = s * d0; dL[o + HW] = s * d1; dL[o + 2 * HW] = s * d2;
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
        if (threadIdx.x == 0) atomicAdd(loss_acc, (double)v);
    }
}

__global__ void focal_l2_finish_kernel(const double *loss_acc, float inv_count, float *loss_out) {
    loss_out[0] = (float)(loss_acc[0] * (double)inv_count);
}

}  // namespace up3d

using namespace up3d;

extern "C" int up3d_focal_l2_loss_strided(int64_t n_images, int H, int W, const float *rendered, const void *gt, int gt_is_u8,
                                          int64_t views_per_object, int64_t gt_object_stride, const float *bg,
                                          float non_bg_rate, float bg_rate, float *loss_out, float *dL_drendered,
                                          up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(n_images >= 0 && H > 0 && W > 0, "up3d_focal_l2_loss: bad sizes");
    UP3D_CHECK_ARG(views_per_object > 0 && gt_object_stride >= 0, "up3d_focal_l2_loss: bad gt layout");
    UP3D_CHECK_ARG(loss_out != nullptr, "up3d_focal_l2_loss: loss_out must hold 4 floats (1 result + 8-byte accumulator)");
    // loss_out[0] = result; loss_out[2..3] (8-byte aligned) = double accumulator scratch
    UP3D_CHECK_ARG(((uintptr_t)loss_out & 7) == 0, "up3d_focal_l2_loss: loss_out must be 8-byte aligned");
    double *acc = (double *)(loss_out + 2);
    UP3D_CUDA_OK(cudaMemsetAsync(loss_out, 0, 4 * sizeof(float), stream));
    if (n_images == 0) return 0;
    UP3D_CHECK_ARG(rendered && gt && bg, "up3d_focal_l2_loss: null pointer");
    const long long HW = (long long)H * W, total = HW * n_images;
    const float wfg = 2.f * non_bg_rate / (bg_rate + non_bg_rate), wbg = 2.f * bg_rate / (bg_rate + non_bg_rate);
    const float inv_count = 1.0f / (float)(total * 3);
    long long blocks = (total + 255) / 256;
    if (blocks > UP3D_NUM_SMS * 8) blocks = UP3D_NUM_SMS * 8;
    if (gt_is_u8)
        focal_l2_kernel<unsigned char><<<(int)blocks, 256, 0, stream>>>(total, HW, rendered, (const unsigned char *)gt,
                                                                       views_per_object, gt_object_stride, bg, wbg, wfg,
                                                                       inv_count, acc, dL_drendered);
    else
        focal_l2_kernel<float><<<(int)blocks, 256, 0, stream>>>(total, HW, rendered, (const float *)gt, views_per_object,
                                                               gt_object_stride, bg, wbg, wfg, inv_count, acc, dL_drendered);
    UP3D_LAUNCH_OK("focal_l2_kernel");
    focal_l2_finish_kernel<<<1, 1, 0, stream>>>(acc, inv_count, loss_out);
    UP3D_LAUNCH_OK("focal_l2_finish_kernel");
    return 0;
}

extern "C" int up3d_focal_l2_loss(int64_t n_images, int H, int W, const float *rendered, const float *gt, const float *bg,
                                  float non_bg_rate, float bg_rate, float *loss_out, float *dL_drendered,
                                  up3d_stream_t stream_) {
    return up3d_focal_l2_loss_strided(n_images, H, W, rendered, gt, 0, n_images > 0 ? n_images : 1, 0, bg, non_bg_rate, bg_rate,
                                      loss_out, dL_drendered, stream_);
}
