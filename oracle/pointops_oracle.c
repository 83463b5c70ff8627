/*
 * pointops_oracle.c -- CPU restatement of the reference's point-op CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (unipre3d_b200/) may import,
 * link or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker / reported CPU baseline.
 *
 * Parity pin: the `-m gpu` tests run this oracle against the reference's OWN kernels
 * (oracle/_ref/pointnet2_batch_cuda.so, compiled unmodified from
 * /root/reference/openpoints/cpp/pointnet2_batch/src by oracle/build_ref.py) on the same
 * inputs, index-for-index.
 *
 * Each function literally replays the reference kernel's per-thread loop and its
 * shared-memory tree reduction, so tie-breaking is identical:
 *   - furthest_point_sampling_kernel   /root/reference/openpoints/cpp/pointnet2_batch/src/sampling_gpu.cu:100-216
 *   - opt_n_threads                    .../src/cuda_utils.h:10-14
 *   - ball_query_kernel_fast           .../src/ball_query_gpu.cu:15-51
 *   - group_points_kernel_fast         .../src/group_points_gpu.cu:53-72
 *   - group_points_grad_kernel_fast    .../src/group_points_gpu.cu:14-31
 *   - gather_points_kernel_fast        .../src/sampling_gpu.cu:15-31
 *
 * Floating point: the squared distance follows the SASS nvcc 12.9 emits for the reference
 * source (checked with cuobjdump on oracle/_ref objects):  d = fma(dz,dz, fma(dx,dx, dy*dy)).
 * Compile with -ffp-contract=off so gcc adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static int opt_n_threads(int work_size) {
    /* cuda_utils.h:10-14 : pow_2 = (int)(log(n)/log(2)); clamp(1 << pow_2, 1, 1024) */
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

int up3d_oracle_fps_block_size(int n) { return opt_n_threads(n); }

static inline float sqdist_ref(float x1, float y1, float z1, float x2, float y2, float z2) {
    const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
    return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* xyz (B,N,3) f32, temp (B,N) f32 in/out (caller fills 1e10 as subsample.py:93 does), idxs (B,M) i32 */
void up3d_oracle_fps(int b, int n, int m, const float *xyz, float *temp, int32_t *idxs) {
    if (m <= 0) return;
    const int bs = opt_n_threads(n);
    float *dists = (float *)malloc(sizeof(float) * bs);
    int *dists_i = (int *)malloc(sizeof(int) * bs);
    for (int bi = 0; bi < b; ++bi) {
        const float *ds = xyz + (size_t)bi * n * 3;
        float *tp = temp + (size_t)bi * n;
        int32_t *out = idxs + (size_t)bi * m;
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = ds[old * 3 + 0], y1 = ds[old * 3 + 1], z1 = ds[old * 3 + 2];
            for (int tid = 0; tid < bs; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += bs) {
                    const float d = sqdist_ref(x1, y1, z1, ds[k * 3 + 0], ds[k * 3 + 1], ds[k * 3 + 2]);
                    const float d2 = fminf(d, tp[k]);
                    tp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            /* __update tree: sampling_gpu.cu:93-98,150-209 */
            for (int s = bs / 2; s >= 1; s >>= 1) {
                for (int tid = 0; tid < s; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + s];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + s];
                    dists[tid] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            }
            old = dists_i[0];
            out[j] = old;
        }
    }
    free(dists);
    free(dists_i);
}

/* new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) i32; caller zero-fills idx (group.py:194) */
void up3d_oracle_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                            const float *xyz, int32_t *idx) {
    const float radius2 = radius * radius;
    for (int bi = 0; bi < b; ++bi)
        for (int pi = 0; pi < m; ++pi) {
            const float *c = new_xyz + ((size_t)bi * m + pi) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int32_t *o = idx + ((size_t)bi * m + pi) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                /* (new_x - x)^2 ... : same contraction as above with d = new - x */
                const float dx = c[0] - pts[k * 3 + 0], dy = c[1] - pts[k * 3 + 1], dz = c[2] - pts[k * 3 + 2];
                const float d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy));
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
}

/* points (B,C,N), idx (B,M,K) -> out (B,C,M,K) */
void up3d_oracle_group(int b, int c, int n, int m, int k, const float *points, const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int pi = 0; pi < m; ++pi)
                for (int s = 0; s < k; ++s)
                    out[(((size_t)bi * c + ci) * m + pi) * k + s] =
                        points[((size_t)bi * c + ci) * n + idx[((size_t)bi * m + pi) * k + s]];
}

/* grad_out (B,C,M,K), idx (B,M,K) -> grad_points (B,C,N) (zero-filled here).
 * The reference uses float atomicAdd in an unspecified order; the oracle sums in double and
 * rounds once, so tests compare with a float tolerance. */
void up3d_oracle_group_grad(int b, int c, int n, int m, int k, const float *grad_out, const int32_t *idx,
                            float *grad_points) {
    double *acc = (double *)calloc((size_t)n, sizeof(double));
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            memset(acc, 0, sizeof(double) * n);
            for (int pi = 0; pi < m; ++pi)
                for (int s = 0; s < k; ++s)
                    acc[idx[((size_t)bi * m + pi) * k + s]] += grad_out[(((size_t)bi * c + ci) * m + pi) * k + s];
            for (int i = 0; i < n; ++i) grad_points[((size_t)bi * c + ci) * n + i] = (float)acc[i];
        }
    free(acc);
}

/* points (B,C,N), idx (B,M) -> out (B,C,M) */
void up3d_oracle_gather(int b, int c, int n, int m, const float *points, const int32_t *idx, float *out) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci)
            for (int pi = 0; pi < m; ++pi)
                out[((size_t)bi * c + ci) * m + pi] = points[((size_t)bi * c + ci) * n + idx[(size_t)bi * m + pi]];
}
