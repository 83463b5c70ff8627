// sparse_conv.cu -- submanifold / strided / inverse sparse 3-D convolution for the scene-level backbones.
//
// Replaces the arithmetic the reference takes from the un-vendored spconv v2 package: spconv.SubMConv3d (k = 1, 3, 5),
// spconv.SparseConv3d (k = 2, s = 2) and spconv.SparseInverseConv3d (k = 2) as used by
// /root/reference/pointcept/models/sparse_unet/spconv_unet_v1m1_base.py:57-83, 153-160, 209-216, 247-253 and by the
// xCPE of point_transformer_v3m1_base.py:281-287.  Convention = torch.nn.functional.conv3d's cross-correlation on the
// densified grid (what tests/test_sparse_conv_gpu.py checks against).
//
// Design: output-stationary implicit GEMM.  A "rulebook" nbr[o][i] (int32, -1 = no input) names, for kernel offset o and
// output row i, the input row that offset reads; all three layer types are the same kernel with different rulebooks:
//     SubM:      nbr[o][i] = row of voxel (coord_i + offset_o)        (built here by binary search in the sorted keys)
//     strided:   nbr[o][p] = the child of coarse voxel p at parity o   (from parent / parity arrays)
//     inverse:   nbr[o][i] = parent(i) if parity(i) == o else -1
// and the input gradient is the same kernel again with the transposed rulebook and per-offset transposed weights.
// One CTA owns 64 output rows; per offset it gathers the 64 x C_in input rows (fp32 in HBM -> bf16 in shared memory,
// zero rows where nbr < 0, offsets nobody in the tile uses are skipped) and multiplies them with W[o] (C_in x C_out,
// bf16, L2-resident) on the tensor cores (mma.sync through the WMMA API, fp32 accumulators kept in registers across
// all offsets).  No atomics in the forward / input-gradient direction: results are deterministic.
// The weight gradient reduces over voxels with fp32 atomics (as spconv does).
#include <cuda_bf16.h>
#include <mma.h>

#include "common.cuh"

namespace up3d {
namespace sp {

using namespace nvcuda;
typedef __nv_bfloat16 bf16;

constexpr int ROWS = 64;          // output rows per CTA (4 warps x 16)
constexpr int THREADS = 128;
constexpr int PAD = 8;            // bf16 elements of padding per shared-memory row (bank spread for ldmatrix)

// key = batch:16 | c0:16 | c1:16 | c2:16  (all coordinates in [0, 65535])
__device__ __forceinline__ long long pack_key(int b, int x, int y, int z) {
    return ((long long)b << 48) | ((long long)x << 32) | ((long long)y << 16) | (long long)z;
}

// nbr[o][i] for the K x K x K neighbourhood of every voxel; keys sorted ascending, coords (n,4) = (b, c0, c1, c2) in the
// same order.  Offset index o = (a * K + b) * K + c reads voxel coord + (a - r, b - r, c - r), r = K / 2.
__global__ void __launch_bounds__(256) subm_rulebook_kernel(int n, int K, const long long *__restrict__ keys,
                                                            const int *__restrict__ coords, int *__restrict__ nbr) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int KV = K * K * K;
    if (t >= (long long)n * KV) return;
    const int o = (int)(t / n), i = (int)(t % n);
    const int r = K / 2;
    const int a = o / (K * K) - r, b = (o / K) % K - r, c = o % K - r;
    const int bb = coords[4 * i], x = coords[4 * i + 1] + a, y = coords[4 * i + 2] + b, z = coords[4 * i + 3] + c;
    int res = -1;
    if (a == 0 && b == 0 && c == 0) {
        res = i;
    } else if (x >= 0 && y >= 0 && z >= 0 && x < 65536 && y < 65536 && z < 65536) {
        const long long key = pack_key(bb, x, y, z);
        int lo = 0, hi = n - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const long long km = keys[mid];
            if (km == key) { res = mid; break; }
            if (km < key) lo = mid + 1; else hi = mid - 1;
        }
    }
    nbr[(size_t)o * n + i] = res;
}

// ------------------------------------------------------------------------------------------------ forward / dX
// grid (ceil(n_out / 64), C_out / (16 * NT)); NT = 16-column accumulator fragments per warp.
// Per kernel offset and per 64-channel K chunk: the weight slice W[o][k0:k0+64, col0:col0+16 NT] is copied to shared
// memory with cp.async (coalesced 16-byte pieces, in flight while the input rows are gathered), the 64 gathered input rows
// are converted fp32 -> bf16 into shared memory, then both MMA operands come from shared memory.  (The first version read
// the B fragments straight from global memory: 36 dependent L2 round trips per offset made it 3.5x slower.)
constexpr int KC = 64;            // K chunk (input channels per staging round)

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

template <int NT>
__global__ void __launch_bounds__(THREADS)
conv_kernel(int n_out, int Cin, int Cout, int KV, const int *__restrict__ nbr, const float *__restrict__ in,
            const bf16 *__restrict__ w, float *__restrict__ out) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    constexpr int LDA = KC + PAD, LDB = 16 * NT + PAD;
    bf16 *As = reinterpret_cast<bf16 *>(smem_raw);                 // [ROWS][LDA]
    bf16 *Bs = As + ROWS * LDA;                                    // [KC][LDB]
    __shared__ int s_idx[ROWS];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int row0 = blockIdx.x * ROWS, col0 = blockIdx.y * (16 * NT);

    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) wmma::fill_fragment(acc[j], 0.f);

    for (int o = 0; o < KV; ++o) {
        int my = -1;
        if (tid < ROWS) {
            const int r = row0 + tid;
            my = r < n_out ? nbr[(size_t)o * n_out + r] : -1;
            s_idx[tid] = my;
        }
        // offsets that no row of this tile uses cost one barrier and no memory traffic
        if (!__syncthreads_or(my >= 0)) continue;
        for (int k0 = 0; k0 < Cin; k0 += KC) {
            const int kw = min(KC, Cin - k0);                      // multiple of 16
            // B slice: kw rows x 16 NT columns, 2 NT 16-byte pieces per row
            const bf16 *wsrc = w + ((size_t)o * Cin + k0) * Cout + col0;
            for (int e = tid; e < kw * (2 * NT); e += THREADS) {
                const int r = e / (2 * NT), p8 = e % (2 * NT);
                cp_async16(Bs + r * LDB + 8 * p8, wsrc + (size_t)r * Cout + 8 * p8);
            }
            // A chunk: gathered rows, fp32 -> bf16
            const int c4n = kw / 4;
            for (int e = tid; e < ROWS * c4n; e += THREADS) {
                const int r = e / c4n, c4 = e % c4n;
                const int src = s_idx[r];
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (src >= 0) v = *reinterpret_cast<const float4 *>(in + (size_t)src * Cin + k0 + 4 * c4);
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const unsigned *>(&lo);
                pk.y = *reinterpret_cast<const unsigned *>(&hi);
                *reinterpret_cast<uint2 *>(As + r * LDA + 4 * c4) = pk;
            }
            cp_async_wait_all();
            __syncthreads();
            for (int kc = 0; kc < kw; kc += 16) {
                wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> af;
                wmma::load_matrix_sync(af, As + (warp * 16) * LDA + kc, LDA);
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> bfr;
                    wmma::load_matrix_sync(bfr, Bs + kc * LDB + 16 * j, LDB);
                    wmma::mma_sync(acc[j], af, bfr, acc[j]);
                }
            }
            __syncthreads();                                       // As / Bs are rewritten by the next chunk
        }
    }
    // epilogue: fragments -> shared (fp32) -> coalesced rows
    float *Cs = reinterpret_cast<float *>(smem_raw);               // [ROWS][16 * NT + 4]
    const int ldc = 16 * NT + 4;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NT; ++j) wmma::store_matrix_sync(Cs + (warp * 16) * ldc + 16 * j, acc[j], ldc, wmma::mem_row_major);
    __syncthreads();
    const int q4 = 4 * NT;                                         // float4 chunks per output row slice
    for (int e = tid; e < ROWS * q4; e += THREADS) {
        const int r = e / q4, c4 = e % q4;
        if (row0 + r < n_out)
            *reinterpret_cast<float4 *>(out + (size_t)(row0 + r) * Cout + col0 + 4 * c4) =
                *reinterpret_cast<const float4 *>(Cs + r * ldc + 4 * c4);
    }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[o] (Cin x Cout) += sum_i in[nbr[o][i]]^T dout[i].  grid (voxel chunks, KV, (Cin/CTi) * (Cout/CTo)) with tiles of up to
// 128 x 128 channels (one tile for the 32..128-channel layers: the gathered inputs and the output gradients are then read
// once per (offset, row tile), not once per channel tile).  8 warps: warp w owns 16 input channels of the tile and all its
// output columns; a CTA accumulates its tile over its chunk of `rows_per_cta` output rows in registers and adds it once
// with fp32 atomics.
constexpr int WG_MAX = 128;
constexpr int WG_THREADS = 256;
__host__ __device__ inline int wg_tile(int c) {             // largest multiple of 16 <= 128 that divides c
    for (int t = WG_MAX; t >= 16; t -= 16)
        if (c % t == 0) return t;
    return 16;
}

__global__ void __launch_bounds__(WG_THREADS)
wgrad_kernel(int n_out, int Cin, int Cout, int rows_per_cta, const int *__restrict__ nbr, const float *__restrict__ in,
             const float *__restrict__ dout, float *__restrict__ dw) {
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const int ci_t = wg_tile(Cin), co_t = wg_tile(Cout);
    const int lda = ci_t + PAD, ldd = co_t + PAD;
    bf16 *As = reinterpret_cast<bf16 *>(smem_raw);               // [ROWS][lda]  gathered inputs  (voxel, cin)
    bf16 *Ds = As + ROWS * lda;                                  // [ROWS][ldd]  output gradients (voxel, cout)
    __shared__ int s_idx[ROWS];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int o = blockIdx.y;
    const int tiles_co = Cout / co_t;
    const int ci0 = (blockIdx.z / tiles_co) * ci_t, co0 = (blockIdx.z % tiles_co) * co_t;
    const int nfr = co_t / 16;                                   // accumulator fragments per warp (<= 8)
    const bool warp_on = warp * 16 < ci_t;
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) wmma::fill_fragment(acc[j], 0.f);

    const int r_begin = blockIdx.x * rows_per_cta, r_end = min(n_out, r_begin + rows_per_cta);
    for (int row0 = r_begin; row0 < r_end; row0 += ROWS) {
        int my = -1;
        if (tid < ROWS) {
            const int r = row0 + tid;
            my = r < r_end ? nbr[(size_t)o * n_out + r] : -1;
            s_idx[tid] = my;
        }
        if (!__syncthreads_or(my >= 0)) continue;
        for (int e = tid; e < ROWS * (ci_t / 4); e += WG_THREADS) {
            const int r = e / (ci_t / 4), c4 = e % (ci_t / 4);
            const int src = s_idx[r];
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (src >= 0) v = *reinterpret_cast<const float4 *>(in + (size_t)src * Cin + ci0 + 4 * c4);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned *>(&lo);
            pk.y = *reinterpret_cast<const unsigned *>(&hi);
            *reinterpret_cast<uint2 *>(As + r * lda + 4 * c4) = pk;
        }
        for (int e = tid; e < ROWS * (co_t / 4); e += WG_THREADS) {
            const int r = e / (co_t / 4), c4 = e % (co_t / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s_idx[r] >= 0) v = *reinterpret_cast<const float4 *>(dout + (size_t)(row0 + r) * Cout + co0 + 4 * c4);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned *>(&lo);
            pk.y = *reinterpret_cast<const unsigned *>(&hi);
            *reinterpret_cast<uint2 *>(Ds + r * ldd + 4 * c4) = pk;
        }
        __syncthreads();
        if (warp_on) {
            for (int kv = 0; kv < ROWS; kv += 16) {                // the reduction dimension is the voxel axis
                wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::col_major> af;      // (cin, voxel) = As^T
                wmma::load_matrix_sync(af, As + kv * lda + warp * 16, lda);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < nfr) {
                        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> bfr;
                        wmma::load_matrix_sync(bfr, Ds + kv * ldd + 16 * j, ldd);
                        wmma::mma_sync(acc[j], af, bfr, acc[j]);
                    }
                }
            }
        }
        __syncthreads();
    }
    // tile -> shared fp32 (aliases the staging buffers) -> atomics
    float *Ts = reinterpret_cast<float *>(smem_raw);               // [ci_t][co_t + 4]
    const int ldt = co_t + 4;
    __syncthreads();
    if (warp_on) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < nfr) wmma::store_matrix_sync(Ts + (warp * 16) * ldt + 16 * j, acc[j], ldt, wmma::mem_row_major);
    }
    __syncthreads();
    float *dst = dw + (size_t)o * Cin * Cout;
    for (int e = tid; e < ci_t * co_t; e += WG_THREADS) {
        const int r = e / co_t, c = e % co_t;
        const float v = Ts[r * ldt + c];
        if (v != 0.f) atomicAdd(dst + (size_t)(ci0 + r) * Cout + co0 + c, v);
    }
}

template <int NT>
static int launch_conv(int n_out, int Cin, int Cout, int KV, const int *nbr, const float *in, const bf16 *w, float *out,
                       cudaStream_t st) {
    const size_t smem = max((size_t)(ROWS * (KC + PAD) + KC * (16 * NT + PAD)) * sizeof(bf16),
                            (size_t)ROWS * (16 * NT + 4) * sizeof(float));
    static size_t configured = 0;
    if (smem > configured) {
        UP3D_CUDA_OK(cudaFuncSetAttribute(conv_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    const dim3 grid(div_up(n_out, ROWS), Cout / (16 * NT));
    conv_kernel<NT><<<grid, THREADS, smem, st>>>(n_out, Cin, Cout, KV, nbr, in, w, out);
    UP3D_LAUNCH_OK("sp::conv_kernel");
    return 0;
}

}  // namespace sp
}  // namespace up3d

using namespace up3d;

extern "C" int up3d_sparse_subm_rulebook(int n, int kernel_size, const int64_t *keys_sorted, const int32_t *coords,
                                         int32_t *nbr, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n >= 0 && (kernel_size == 1 || kernel_size == 3 || kernel_size == 5),
                   "up3d_sparse_subm_rulebook: kernel_size must be 1, 3 or 5 (got %d)", kernel_size);
    if (n == 0) return 0;
    UP3D_CHECK_ARG(keys_sorted && coords && nbr, "up3d_sparse_subm_rulebook: NULL pointer");
    const long long total = (long long)n * kernel_size * kernel_size * kernel_size;
    sp::subm_rulebook_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        n, kernel_size, (const long long *)keys_sorted, coords, nbr);
    UP3D_LAUNCH_OK("subm_rulebook_kernel");
    return 0;
}

extern "C" int up3d_sparse_conv(int n_out, int c_in, int c_out, int kernel_volume, const int32_t *nbr, const float *in,
                                const void *weight_bf16, float *out, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n_out >= 0 && c_in > 0 && c_out > 0 && kernel_volume > 0, "up3d_sparse_conv: bad sizes");
    UP3D_CHECK_ARG(c_in % 16 == 0 && c_out % 16 == 0 && c_in <= 1024,
                   "up3d_sparse_conv: channel counts must be multiples of 16 (pad), C_in <= 1024 (got %d -> %d)", c_in, c_out);
    if (n_out == 0) return 0;
    UP3D_CHECK_ARG(nbr && in && weight_bf16 && out, "up3d_sparse_conv: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const sp::bf16 *w = (const sp::bf16 *)weight_bf16;
    const int nf = c_out / 16;
    // widest accumulator strip that divides C_out (<= 16 fragments = 128 accumulator registers per thread)
    if (nf % 16 == 0) return sp::launch_conv<16>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
    if (nf % 8 == 0) return sp::launch_conv<8>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
    if (nf % 6 == 0) return sp::launch_conv<6>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
    if (nf % 4 == 0) return sp::launch_conv<4>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
    if (nf % 2 == 0) return sp::launch_conv<2>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
    return sp::launch_conv<1>(n_out, c_in, c_out, kernel_volume, nbr, in, w, out, st);
}

extern "C" int up3d_sparse_conv_wgrad(int n_out, int c_in, int c_out, int kernel_volume, const int32_t *nbr, const float *in,
                                      const float *dout, float *dweight, up3d_stream_t stream) {
    UP3D_CHECK_ARG(n_out >= 0 && c_in > 0 && c_out > 0 && kernel_volume > 0, "up3d_sparse_conv_wgrad: bad sizes");
    UP3D_CHECK_ARG(c_in % 16 == 0 && c_out % 16 == 0, "up3d_sparse_conv_wgrad: channel counts must be multiples of 16");
    const int ci_t = sp::wg_tile(c_in), co_t = sp::wg_tile(c_out);
    if (n_out == 0) return 0;
    UP3D_CHECK_ARG(nbr && in && dout && dweight, "up3d_sparse_conv_wgrad: NULL pointer");
    const size_t smem = max((size_t)sp::ROWS * (ci_t + sp::PAD + co_t + sp::PAD) * sizeof(sp::bf16),
                            (size_t)ci_t * (co_t + 4) * sizeof(float));
    static size_t configured = 0;
    if (smem > configured) {
        UP3D_CUDA_OK(cudaFuncSetAttribute(sp::wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    // enough voxel chunks to fill the machine, at least 4 sub-tiles each
    const int tiles = (c_in / ci_t) * (c_out / co_t);
    int chunks = div_up(3 * UP3D_NUM_SMS, kernel_volume * tiles);
    chunks = max(1, min(chunks, div_up(n_out, 4 * sp::ROWS)));
    const int rows_per_cta = div_up(div_up(n_out, chunks), sp::ROWS) * sp::ROWS;
    const dim3 grid(div_up(n_out, rows_per_cta), kernel_volume, tiles);
    sp::wgrad_kernel<<<grid, sp::WG_THREADS, smem, (cudaStream_t)stream>>>(n_out, c_in, c_out, rows_per_cta, nbr, in, dout, dweight);
    UP3D_LAUNCH_OK("sp::wgrad_kernel");
    return 0;
}
