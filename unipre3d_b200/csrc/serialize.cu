// serialize.cu -- space-filling-curve keys of voxel coordinates: the first stage of PTv3's `Point.serialization`
// (/root/reference/pointcept/models/utils/structure.py:47-107), which the reference computes with three table look-ups
// per coordinate and byte (serialization/z_order.py:62-96, on top of serialization/default.py:8-25).
//
// z-order ("z", and "z-trans" = x and y swapped): bit i of x, y, z goes to bits 3i+2, 3i+1, 3i of the key
// (z_order.py:41-51); the batch index is OR-ed in above bit 3*depth (default.py:22-24).  Pure integer work, one thread per
// point, 12 B in / 8 B out: HBM-bound by construction; the bit interleave is the classic 3-way "part by 2" magic-number spread.
#include "common.cuh"

namespace up3d {

// spread the low 21 bits of v so that bit i lands on bit 3i
__device__ __forceinline__ unsigned long long part1by2(unsigned long long v) {
    v &= 0x1fffffULL;
    v = (v | (v << 32)) & 0x1f00000000ffffULL;
    v = (v | (v << 16)) & 0x1f0000ff0000ffULL;
    v = (v | (v << 8)) & 0x100f00f00f00f00fULL;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ULL;
    v = (v | (v << 2)) & 0x1249249249249249ULL;
    return v;
}

__global__ void __launch_bounds__(256)
zorder_keys_kernel(long long n, int depth, int swap_xy, const int *__restrict__ grid_coord, const long long *__restrict__ batch,
                   long long *__restrict__ code) {
    const unsigned long long mask = depth >= 21 ? 0x1fffffULL : ((1ULL << depth) - 1ULL);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)(long long)grid_coord[3 * i], y = (unsigned long long)(long long)grid_coord[3 * i + 1];
        const unsigned long long z = (unsigned long long)(long long)grid_coord[3 * i + 2];
        if (swap_xy) { const unsigned long long t = x; x = y; y = t; }
        unsigned long long key = (part1by2(x & mask) << 2) | (part1by2(y & mask) << 1) | part1by2(z & mask);
        if (batch) key |= (unsigned long long)batch[i] << (3 * depth);
        code[i] = (long long)key;
    }
}

// Hilbert order ("hilbert", and "hilbert-trans" = x and y swapped): Skilling's transpose algorithm exactly as
// serialization/hilbert.py:96-191 runs it on bit planes -- for every bit from the top, for every dimension: if the bit is
// set invert the lower bits of dimension 0, else exchange the differing lower bits of dimension 0 and this dimension;
// then interleave (dimension 0 most significant within a triplet) and Gray-decode the 3*depth-bit word.
__global__ void __launch_bounds__(256)
hilbert_keys_kernel(long long n, int depth, int swap_xy, const int *__restrict__ grid_coord, const long long *__restrict__ batch,
                    long long *__restrict__ code) {
    const unsigned mask = (1u << depth) - 1u;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned X[3] = {(unsigned)grid_coord[3 * i] & mask, (unsigned)grid_coord[3 * i + 1] & mask,
                         (unsigned)grid_coord[3 * i + 2] & mask};
        if (swap_xy) { const unsigned t = X[0]; X[0] = X[1]; X[1] = t; }
        for (unsigned Q = 1u << (depth - 1); Q > 0; Q >>= 1) {
            const unsigned P = Q - 1;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (X[d] & Q) {
                    X[0] ^= P;
                } else {
                    const unsigned t = (X[0] ^ X[d]) & P;
                    X[0] ^= t;
                    X[d] ^= t;
                }
            }
        }
        unsigned long long g = (part1by2(X[0]) << 2) | (part1by2(X[1]) << 1) | part1by2(X[2]);
        g ^= g >> 1; g ^= g >> 2; g ^= g >> 4; g ^= g >> 8; g ^= g >> 16; g ^= g >> 32;        // Gray -> binary
        if (batch) g |= (unsigned long long)batch[i] << (3 * depth);
        code[i] = (long long)g;
    }
}

}  // namespace up3d

using namespace up3d;

extern "C" int up3d_hilbert_keys(int64_t n, int depth, int swap_xy, const int32_t *grid_coord, const int64_t *batch, int64_t *code,
                                 up3d_stream_t stream) {
    UP3D_CHECK_ARG(n >= 0, "up3d_hilbert_keys: negative size");
    UP3D_CHECK_ARG(depth >= 1 && depth <= 16, "up3d_hilbert_keys: depth must be in [1,16] (got %d)", depth);
    if (n == 0) return 0;
    UP3D_CHECK_ARG(grid_coord && code, "up3d_hilbert_keys: NULL pointer");
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (long long)UP3D_NUM_SMS * 16 ? blocks : (long long)UP3D_NUM_SMS * 16);
    hilbert_keys_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((long long)n, depth, swap_xy, grid_coord, (const long long *)batch,
                                                               (long long *)code);
    UP3D_LAUNCH_OK("hilbert_keys_kernel");
    return 0;
}

extern "C" int up3d_zorder_keys(int64_t n, int depth, int swap_xy, const int32_t *grid_coord, const int64_t *batch, int64_t *code,
                                up3d_stream_t stream) {
    UP3D_CHECK_ARG(n >= 0, "up3d_zorder_keys: negative size");
    UP3D_CHECK_ARG(depth >= 1 && depth <= 16, "up3d_zorder_keys: depth must be in [1,16] (got %d; the reference asserts <= 16)", depth);
    if (n == 0) return 0;
    UP3D_CHECK_ARG(grid_coord && code, "up3d_zorder_keys: NULL pointer");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    const long long blocks = (n + 255) / 256;
    const int grid = (int)(blocks < (long long)UP3D_NUM_SMS * 16 ? blocks : (long long)UP3D_NUM_SMS * 16);
    zorder_keys_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((long long)n, depth, swap_xy, grid_coord, (const long long *)batch,
                                                              (long long *)code);
    UP3D_LAUNCH_OK("zorder_keys_kernel");
    return 0;
}
