// pointops.cu -- FPS / ball query / grouping / gather for sm_100a.
//
// Replaces /root/reference/openpoints/cpp/pointnet2_batch/src/{sampling_gpu.cu,ball_query_gpu.cu,
// group_points_gpu.cu} (pybind surface pointnet2_api.cpp:10-24).  Index outputs are bit-identical to the
// reference kernels, including its tie-breaking (see fps_priority below); the squared distance uses the
// exact operation order nvcc emits for the reference source: d = fma(dz,dz, fma(dx,dx, dy*dy)).
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace up3d {

// ------------------------------------------------------------------------------------------------
// Furthest point sampling.
//   reference: one CTA of bs = opt_n_threads(N) threads per cloud; every round re-reads all points and
//   the (B,N) temp array from global memory and runs a 10-level __syncthreads tree (sampling_gpu.cu:125-215).
//   here: the cloud and its running min-distances live in REGISTERS (PPT points per thread), the query
//   point is read from a shared-memory copy, the arg-max is two REDUX instructions per level and ONE
//   __syncthreads per round (double-buffered partials).
//
// Tie-breaking of the reference, reproduced exactly: thread t scans k = t, t+bs, ... keeping the first
// maximum (strict >); the tree merges slot i with slot i+s keeping the lower slot on ties.  The winner among
// equal distances is therefore the point with the smallest (bitreverse_{log2 bs}(k mod bs), k div bs).
// We fold that into the low word of a 64-bit key so that a plain max-reduction yields the same index.
// ------------------------------------------------------------------------------------------------
constexpr int FPS_THREADS = 1024;
constexpr int FPS_MAX_PPT = 16;

__device__ __forceinline__ uint32_t fps_priority(int k, int bs, int log2bs) {
    const uint32_t low = (uint32_t)(k & (bs - 1));
    const uint32_t rev = log2bs ? (__brev(low) >> (32 - log2bs)) : 0u;
    return 0xFFFFFFFFu - ((rev << 20) | (uint32_t)(k >> log2bs));
}
__device__ __forceinline__ int fps_index_from_priority(uint32_t inv, int bs, int log2bs) {
    const uint32_t p = 0xFFFFFFFFu - inv;
    const uint32_t rev = p >> 20, q = p & 0xFFFFFu;
    const uint32_t low = log2bs ? (__brev(rev) >> (32 - log2bs)) : 0u;
    return (int)(q * (uint32_t)bs + low);
}

// block-wide arg-max of (hi, lo) with hi compared first; result broadcast to every thread
__device__ __forceinline__ void block_argmax(uint32_t &hi, uint32_t &lo, uint32_t (*red)[2], int lane, int warp,
                                             int nwarps) {
    uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
    uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    if (lane == 0) { red[warp][0] = mh; red[warp][1] = ml; }
    __syncthreads();
    uint32_t h2 = lane < nwarps ? red[lane][0] : 0u, l2 = lane < nwarps ? red[lane][1] : 0u;
    mh = __reduce_max_sync(0xffffffffu, h2);
    ml = __reduce_max_sync(0xffffffffu, h2 == mh ? l2 : 0u);
    hi = mh; lo = ml;
}

template <int PPT, bool REG_COORDS>
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_resident_kernel(int N, int M, int bs, int log2bs, const float *__restrict__ xyz, int32_t *__restrict__ idxs) {
    extern __shared__ float s_pts[];  // [3][N]
    __shared__ uint32_t red[2][32][2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *ds = xyz + (size_t)b * N * 3;
    int32_t *out = idxs + (size_t)b * M;
    float *sx = s_pts, *sy = s_pts + N, *sz = s_pts + 2 * N;
    // PPT <= 8: coordinates + priorities in registers; PPT == 16: only the running min-distances are
    // (64-register budget at 1024 threads), coordinates come from the conflict-free shared copy.
    constexpr int RP = REG_COORDS ? PPT : 1;
    float px[RP], py[RP], pz[RP], tmp[PPT];
    uint32_t pr[RP];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = tid + j * FPS_THREADS;
        float x = 0.f, y = 0.f, z = 0.f;
        if (k < N) {
            x = ds[3 * k]; y = ds[3 * k + 1]; z = ds[3 * k + 2];
            sx[k] = x; sy[k] = y; sz[k] = z;
        }
        if (REG_COORDS) { px[j] = x; py[j] = y; pz[j] = z; pr[j] = k < N ? fps_priority(k, bs, log2bs) : 0u; }
        tmp[j] = 1e10f;
    }
    if (tid == 0) out[0] = 0;
    __syncthreads();
    int old = 0;
    for (int r = 1; r < M; ++r) {
        const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
        uint32_t bh = 0u, bl = 0u;  // every real candidate has lo >= 1 > 0 ... and hi >= 0
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = tid + j * FPS_THREADS;
            const bool valid = k < N;
            float x2, y2, z2;
            uint32_t p;
            if (REG_COORDS) { x2 = px[j]; y2 = py[j]; z2 = pz[j]; p = pr[j]; }
            else { const int kk = valid ? k : 0; x2 = sx[kk]; y2 = sy[kk]; z2 = sz[kk]; p = fps_priority(kk, bs, log2bs); }
            const float dx = xsub(x2, x1), dy = xsub(y2, y1), dz = xsub(z2, z1);
            const float d = xfma(dz, dz, xfma(dx, dx, xmul(dy, dy)));
            const float d2 = fminf(d, tmp[j]);
            tmp[j] = d2;
            const uint32_t h = __float_as_uint(d2);
            const bool better = valid && (h > bh || (h == bh && p > bl));
            bh = better ? h : bh;
            bl = better ? p : bl;
        }
        block_argmax(bh, bl, red[r & 1], lane, warp, FPS_THREADS / 32);
        old = fps_index_from_priority(bl, bs, log2bs);
        if (tid == 0) out[r] = old;
    }
}

// ------------------------------------------------------------------------------------------------
// Cluster-parallel FPS: one thread-block CLUSTER (8 CTAs = 8 SMs) per cloud instead of one CTA.
// Each CTA owns a contiguous 1/8 of the points (registers) and a full shared-memory copy of the cloud (query
// point look-up).  Per round: CTA-local arg-max (REDUX + one __syncthreads) -> 8 lanes push the CTA's candidate
// into every peer's shared memory through DSMEM -> ONE cluster barrier -> every thread picks the winner of the 8
// candidates.  Same 64-bit (distance, tie-priority) key as above, so the index sequence is unchanged.
// ------------------------------------------------------------------------------------------------
constexpr int FPSC_CLUSTER = 8;
constexpr int FPSC_THREADS = 512;
constexpr int FPSC_MAX_PPT = 4;     // N <= 8 * 512 * 4 = 16384

template <int PPT>
__global__ void __cluster_dims__(FPSC_CLUSTER, 1, 1) __launch_bounds__(FPSC_THREADS, 1)
fps_cluster_kernel(int N, int M, int bs, int log2bs, const float *__restrict__ xyz, int32_t *__restrict__ idxs) {
    extern __shared__ float s_pts[];                       // [3][N] full cloud
    __shared__ uint32_t wred[2][FPSC_THREADS / 32][2];     // warp partials (double-buffered)
    __shared__ uint32_t cred[2][FPSC_CLUSTER][2];          // one candidate per CTA of the cluster (double-buffered)
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / FPSC_CLUSTER, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *ds = xyz + (size_t)b * N * 3;
    int32_t *out = idxs + (size_t)b * M;
    float *sx = s_pts, *sy = s_pts + N, *sz = s_pts + 2 * N;
    for (int k = tid; k < N; k += FPSC_THREADS) { sx[k] = ds[3 * k]; sy[k] = ds[3 * k + 1]; sz[k] = ds[3 * k + 2]; }
    const int chunk = (N + FPSC_CLUSTER - 1) / FPSC_CLUSTER;
    const int k_begin = rank * chunk, k_end = min(N, k_begin + chunk);
    __syncthreads();
    float px[PPT], py[PPT], pz[PPT], tmp[PPT];
    uint32_t pr[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = k_begin + tid + j * FPSC_THREADS;
        const bool valid = k < k_end;
        px[j] = valid ? sx[k] : 0.f; py[j] = valid ? sy[k] : 0.f; pz[j] = valid ? sz[k] : 0.f;
        pr[j] = valid ? fps_priority(k, bs, log2bs) : 0u;
        tmp[j] = 1e10f;
    }
    if (rank == 0 && tid == 0) out[0] = 0;
    // peers' candidate slots for this CTA (lanes 0..7 of warp 0 each own one peer)
    uint32_t *peer_slot = nullptr;
    if (warp == 0 && lane < FPSC_CLUSTER) peer_slot = cluster.map_shared_rank(&cred[0][0][0], lane);
    cluster.sync();
    int old = 0;
    for (int r = 1; r < M; ++r) {
        const int buf = r & 1;
        const float x1 = sx[old], y1 = sy[old], z1 = sz[old];
        uint32_t bh = 0u, bl = 0u;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float dx = xsub(px[j], x1), dy = xsub(py[j], y1), dz = xsub(pz[j], z1);
            const float d = xfma(dz, dz, xfma(dx, dx, xmul(dy, dy)));
            const float d2 = fminf(d, tmp[j]);
            tmp[j] = d2;
            const uint32_t h = __float_as_uint(d2);
            const bool better = pr[j] != 0u && (h > bh || (h == bh && pr[j] > bl));
            bh = better ? h : bh;
            bl = better ? pr[j] : bl;
        }
        uint32_t mh = __reduce_max_sync(0xffffffffu, bh);
        uint32_t ml = __reduce_max_sync(0xffffffffu, bh == mh ? bl : 0u);
        if (lane == 0) { wred[buf][warp][0] = mh; wred[buf][warp][1] = ml; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t h2 = lane < FPSC_THREADS / 32 ? wred[buf][lane][0] : 0u;
            const uint32_t l2 = lane < FPSC_THREADS / 32 ? wred[buf][lane][1] : 0u;
            mh = __reduce_max_sync(0xffffffffu, h2);
            ml = __reduce_max_sync(0xffffffffu, h2 == mh ? l2 : 0u);
            if (lane < FPSC_CLUSTER) {      // DSMEM: write this CTA's candidate into peer `lane`
                uint32_t *dst = peer_slot + (buf * FPSC_CLUSTER + rank) * 2;
                dst[0] = mh; dst[1] = ml;
            }
        }
        cluster.sync();                     // release/acquire: all 8 candidates visible everywhere
        const uint32_t h3 = lane < FPSC_CLUSTER ? cred[buf][lane][0] : 0u;
        const uint32_t l3 = lane < FPSC_CLUSTER ? cred[buf][lane][1] : 0u;
        mh = __reduce_max_sync(0xffffffffu, h3);
        ml = __reduce_max_sync(0xffffffffu, h3 == mh ? l3 : 0u);
        old = fps_index_from_priority(ml, bs, log2bs);
        if (rank == 0 && tid == 0) out[r] = old;
    }
    cluster.sync();                         // no CTA may exit while peers can still write into its shared memory
}

// streaming fallback for clouds that do not fit the register file: temp lives in global memory
__global__ void __launch_bounds__(FPS_THREADS, 1)
fps_streaming_kernel(int N, int M, int bs, int log2bs, const float *__restrict__ xyz, float *__restrict__ temp,
                     int32_t *__restrict__ idxs) {
    __shared__ uint32_t red[2][32][2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *ds = xyz + (size_t)b * N * 3;
    float *tp = temp + (size_t)b * N;
    int32_t *out = idxs + (size_t)b * M;
    for (int k = tid; k < N; k += FPS_THREADS) tp[k] = 1e10f;
    if (tid == 0) out[0] = 0;
    __syncthreads();
    int old = 0;
    for (int r = 1; r < M; ++r) {
        const float x1 = ds[3 * old], y1 = ds[3 * old + 1], z1 = ds[3 * old + 2];
        uint32_t bh = 0u, bl = 0u;
        for (int k = tid; k < N; k += FPS_THREADS) {
            const float dx = xsub(ds[3 * k], x1), dy = xsub(ds[3 * k + 1], y1), dz = xsub(ds[3 * k + 2], z1);
            const float d = xfma(dz, dz, xfma(dx, dx, xmul(dy, dy)));
            const float d2 = fminf(d, tp[k]);
            tp[k] = d2;
            const uint32_t h = __float_as_uint(d2), p = fps_priority(k, bs, log2bs);
            const bool better = (h > bh || (h == bh && p > bl));
            bh = better ? h : bh;
            bl = better ? p : bl;
        }
        block_argmax(bh, bl, red[r & 1], lane, warp, FPS_THREADS / 32);
        old = fps_index_from_priority(bl, bs, log2bs);
        if (tid == 0) out[r] = old;
    }
}

// ------------------------------------------------------------------------------------------------
// Ball query (+ optional fused centre gather / grouping / centring): one WARP per query centre.
//   reference: one THREAD per centre scanning all N points (ball_query_gpu.cu:15-51).
//   here: 32 lanes test 32 consecutive points; ballot + popc keep the ascending-index order.
// ------------------------------------------------------------------------------------------------
constexpr int BQ_WARPS = 8;
constexpr int BQ_TILE = 3072;   // points staged per shared-memory tile (36 KB static + K ints per warp dynamic < 48 KB)

template <bool FUSED>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int N, int M, int K, float radius2, const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                  const int32_t *__restrict__ fps_idx, int32_t *__restrict__ idx_out, float *__restrict__ center_out,
                  float *__restrict__ neigh_out) {
    __shared__ __align__(16) float s_pts[BQ_TILE * 3];   // AoS tile of the cloud; lane stride 3 words: conflict-free
    extern __shared__ int32_t s_idx[];                   // [BQ_WARPS][K]
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = blockIdx.x * BQ_WARPS + warp;
    const bool live = m < M;
    const float *pts = xyz + (size_t)b * N * 3;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (live) {
        if (FUSED) {
            const int ci = fps_idx[(size_t)b * M + m];
            cx = pts[3 * ci]; cy = pts[3 * ci + 1]; cz = pts[3 * ci + 2];
            if (lane < 3) center_out[((size_t)b * M + m) * 3 + lane] = lane == 0 ? cx : (lane == 1 ? cy : cz);
        } else {
            const float *c = new_xyz + ((size_t)b * M + m) * 3;
            cx = c[0]; cy = c[1]; cz = c[2];
        }
    }
    int32_t *my = s_idx + warp * K;
    int cnt = live ? 0 : K, first = 0;
    for (int t0 = 0; t0 < N; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, N - t0);
        // every warp of the CTA already has its K neighbours -> stop streaming the cloud
        if (__syncthreads_and(cnt >= K)) break;
        for (int i = tid; i < tn * 3; i += BQ_WARPS * 32) s_pts[i] = pts[(size_t)t0 * 3 + i];
        __syncthreads();
        for (int k0 = 0; k0 < tn && cnt < K; k0 += 32) {
            const int k = k0 + lane;
            bool hit = false;
            if (k < tn) {
                const float dx = xsub(cx, s_pts[3 * k]), dy = xsub(cy, s_pts[3 * k + 1]), dz = xsub(cz, s_pts[3 * k + 2]);
                const float d2 = xfma(dz, dz, xfma(dx, dx, xmul(dy, dy)));
                hit = d2 < radius2;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (bal) {
                if (cnt == 0) first = t0 + k0 + __ffs(bal) - 1;
                const int slot = cnt + __popc(bal & lanemask_lt());
                if (hit && slot < K) my[slot] = t0 + k;
                cnt += __popc(bal);
            }
        }
    }
    if (!live) return;
    __syncwarp();
    cnt = min(cnt, K);
    for (int s = lane; s < K; s += 32) {
        // pad with the first hit; no hit at all -> zeros (the reference's zero-initialised idx, group.py:194)
        const int id = s < cnt ? my[s] : first;
        if (idx_out) idx_out[((size_t)b * M + m) * K + s] = id;
        if (FUSED) {
            const size_t o = (((size_t)b * 3) * M + m) * K + s, cs = (size_t)M * K;
            neigh_out[o] = pts[3 * id] - cx;
            neigh_out[o + cs] = pts[3 * id + 1] - cy;
            neigh_out[o + 2 * cs] = pts[3 * id + 2] - cz;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// k nearest neighbours (pointMLP LocalGrouper, /root/reference/openpoints/models/backbone/pointmlp.py:102-113,
// 160-166) and 3-NN inverse-distance weights (PointNetFeaturePropagation, pointmlp.py:397-409).
//   reference: materialises the (S x N) squared-distance matrix with a batched GEMM (128 MB per object at
//   S=4096, N=8192) and runs torch.topk / a FULL torch.sort over it.
//   here: one WARP per query streams the cloud from a shared-memory tile; each lane keeps the K/32 (rounded up)
//   best candidates of its own strided subset... no: a warp-wide sorted list of the K best is kept distributed over
//   the lanes (slot s lives in lane s%32, register s/32); a candidate better than the current worst is inserted with
//   a ballot + shuffle shift.  The distance matrix is never written.
// Distances are |q|^2 + |p|^2 - 2 q.p evaluated in fp32 like the reference's formula (square_distance); ties at the
// K-th place may resolve differently from cuBLAS+topk (the reference's own order is unspecified: sorted=False).
// ------------------------------------------------------------------------------------------------
constexpr int KNN_WARPS = 8;
constexpr int KNN_TILE = 2048;      // points per shared-memory tile (SoA x,y,z,|p|^2 : 32 KB)
constexpr int KNN_MAXR = 2;         // K <= 64

template <int R, int D>   // R = ceil(K / 32) registers per lane; D = point dimension (3, or 4: pointMLP's xyz+height)
__global__ void __launch_bounds__(KNN_WARPS * 32)
knn_kernel(int N, int S, int K, const float *__restrict__ xyz, const float *__restrict__ query,
           int32_t *__restrict__ idx_out, float *__restrict__ dist_out) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE], sw[D == 4 ? KNN_TILE : 1], sn[KNN_TILE];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x * KNN_WARPS + warp;
    const bool live = q < S;
    const float *pts = xyz + (size_t)b * N * D;
    float qx = 0.f, qy = 0.f, qz = 0.f, qw = 0.f;
    if (live) {
        const float *qq = query + ((size_t)b * S + q) * D;
        qx = qq[0]; qy = qq[1]; qz = qq[2];
        if (D == 4) qw = qq[3];
    }
    const float qn = qx * qx + qy * qy + qz * qz + qw * qw;
    // sorted ascending list of the K best: element s in lane s % 32, register s / 32
    float bd[R];
    int bi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { bd[r] = INFINITY; bi[r] = 0; }
    float worst = INFINITY;   // distance of list element K-1 (warp-uniform)
    for (int t0 = 0; t0 < N; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, N - t0);
        __syncthreads();
        for (int i = tid; i < tn; i += KNN_WARPS * 32) {
            const float *pp = pts + (size_t)(t0 + i) * D;
            const float x = pp[0], y = pp[1], z = pp[2], w = D == 4 ? pp[3] : 0.f;
            sx[i] = x; sy[i] = y; sz[i] = z; sn[i] = x * x + y * y + z * z + w * w;
            if (D == 4) sw[i] = w;
        }
        __syncthreads();
        if (!live) continue;
        for (int k0 = 0; k0 < tn; k0 += 32) {
            const int k = k0 + lane;
            float d = INFINITY;
            if (k < tn) {
                float dot = qx * sx[k] + qy * sy[k] + qz * sz[k];
                if (D == 4) dot += qw * sw[k];
                d = (qn + sn[k]) - 2.f * dot;
            }
            unsigned cand = __ballot_sync(0xffffffffu, d < worst);
            while (cand) {   // insert the candidates one by one (rare after the first few tiles)
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const float dn = __shfl_sync(0xffffffffu, d, src);
                const int in = t0 + k0 + src;
                if (!(dn < worst)) continue;      // the list may have tightened since the ballot
                // position = number of list elements <= dn (stable: equal distances keep the earlier index first)
                int pos = 0;
#pragma unroll
                for (int r = 0; r < R; ++r) pos += __popc(__ballot_sync(0xffffffffu, (r * 32 + lane) < K && bd[r] <= dn));
                // shift elements [pos, K-2] up by one slot, highest register first
#pragma unroll
                for (int r = R - 1; r >= 0; --r) {
                    const int s = r * 32 + lane;
                    // slot s takes slot s-1: lane-1 of the same register, or lane 31 of the register below
                    // (r is a compile-time constant after unrolling, so every shuffle is executed by the full warp)
                    const float up_d = __shfl_up_sync(0xffffffffu, bd[r], 1);
                    const int up_i = __shfl_up_sync(0xffffffffu, bi[r], 1);
                    float lo_d = 0.f;
                    int lo_i = 0;
                    if (r > 0) {
                        lo_d = __shfl_sync(0xffffffffu, bd[r > 0 ? r - 1 : 0], 31);
                        lo_i = __shfl_sync(0xffffffffu, bi[r > 0 ? r - 1 : 0], 31);
                    }
                    const float pd = lane == 0 ? lo_d : up_d;
                    const int pi = lane == 0 ? lo_i : up_i;
                    if (s > pos && s < K) { bd[r] = pd; bi[r] = pi; }
                    else if (s == pos) { bd[r] = dn; bi[r] = in; }
                }
                const int wl = (K - 1) & 31;
                float wsel = bd[0];
#pragma unroll
                for (int r = 1; r < R; ++r) wsel = ((K - 1) >> 5) == r ? bd[r] : wsel;
                worst = __shfl_sync(0xffffffffu, wsel, wl);
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int s = r * 32 + lane;
        if (s < K) {
            idx_out[((size_t)b * S + q) * K + s] = bi[r];
            if (dist_out) dist_out[((size_t)b * S + q) * K + s] = bd[r];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// grouping / gather (+ gradients)
// ------------------------------------------------------------------------------------------------
__global__ void group_points_kernel(int C, int N, int M, int K, const float *__restrict__ points,
                                    const int32_t *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * K) return;
    const int id = idx[(size_t)b * M * K + e];
    out[((size_t)b * C + c) * M * K + e] = points[((size_t)b * C + c) * N + id];
}
__global__ void group_points_grad_kernel(int C, int N, int M, int K, const float *__restrict__ grad_out,
                                         const int32_t *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z, c = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * K) return;
    const int id = idx[(size_t)b * M * K + e];
    atomicAdd(grad_points + ((size_t)b * C + c) * N + id, grad_out[((size_t)b * C + c) * M * K + e]);
}

}  // namespace up3d

using namespace up3d;

static int ref_block_size(int n) {
    // cuda_utils.h:10-14 of the reference, evaluated the same way (double log ratio, truncation)
    const int pow_2 = (int)(log((double)n) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

extern "C" {

int up3d_fps_max_resident_points(void) { return FPS_THREADS * FPS_MAX_PPT; }

int up3d_fps(int B, int N, int M, const float *xyz, float *temp, int32_t *idx, up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && N > 0 && M >= 0, "up3d_fps: bad sizes B=%d N=%d M=%d", B, N, M);
    UP3D_CHECK_ARG(M <= N, "up3d_fps: npoint (%d) must not exceed N (%d)", M, N);
    if (B == 0 || M == 0) return 0;
    UP3D_CHECK_ARG(xyz && idx, "up3d_fps: null pointer");
    UP3D_CHECK_ARG(N < (1 << 30), "up3d_fps: N too large");
    const int bs = ref_block_size(N);
    int log2bs = 0;
    while ((1 << log2bs) < bs) ++log2bs;
    if (N > 4096 && N <= FPSC_CLUSTER * FPSC_THREADS * FPSC_MAX_PPT && M >= 32) {   // measured: the cluster barrier only pays off above 4096 points
        // one 8-CTA cluster per cloud (DSMEM candidate exchange)
        const size_t smem = sizeof(float) * 3 * (size_t)N;
        const int ppt = div_up(div_up(N, FPSC_CLUSTER), FPSC_THREADS);
#define UP3D_FPSC_CASE(P)                                                                                          \
    {                                                                                                              \
        UP3D_CUDA_OK(cudaFuncSetAttribute(fps_cluster_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fps_cluster_kernel<P><<<B * FPSC_CLUSTER, FPSC_THREADS, smem, stream>>>(N, M, bs, log2bs, xyz, idx);        \
    }
        if (ppt <= 1) UP3D_FPSC_CASE(1)
        else if (ppt <= 2) UP3D_FPSC_CASE(2)
        else UP3D_FPSC_CASE(4)
#undef UP3D_FPSC_CASE
        UP3D_LAUNCH_OK("fps_cluster_kernel");
    } else if (N <= FPS_THREADS * FPS_MAX_PPT) {
        const size_t smem = sizeof(float) * 3 * (size_t)N;
#define UP3D_FPS_CASE(P)                                                                                           \
    {                                                                                                              \
        UP3D_CUDA_OK(cudaFuncSetAttribute(fps_resident_kernel<P, (P <= 8)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fps_resident_kernel<P, (P <= 8)><<<B, FPS_THREADS, smem, stream>>>(N, M, bs, log2bs, xyz, idx);             \
    }
        const int ppt = div_up(N, FPS_THREADS);
        if (ppt <= 1) UP3D_FPS_CASE(1)
        else if (ppt <= 2) UP3D_FPS_CASE(2)
        else if (ppt <= 4) UP3D_FPS_CASE(4)
        else if (ppt <= 8) UP3D_FPS_CASE(8)
        else UP3D_FPS_CASE(16)
#undef UP3D_FPS_CASE
        UP3D_LAUNCH_OK("fps_resident_kernel");
    } else {
        UP3D_CHECK_ARG(temp != nullptr, "up3d_fps: temp scratch (B,N) required when N > %d", FPS_THREADS * FPS_MAX_PPT);
        fps_streaming_kernel<<<B, FPS_THREADS, 0, stream>>>(N, M, bs, log2bs, xyz, temp, idx);
        UP3D_LAUNCH_OK("fps_streaming_kernel");
    }
    return 0;
}

int up3d_ball_query(int B, int N, int M, float radius, int nsample, const float *new_xyz, const float *xyz,
                    int32_t *idx, up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && N > 0 && M >= 0 && nsample > 0, "up3d_ball_query: bad sizes B=%d N=%d M=%d nsample=%d", B, N, M, nsample);
    if (B == 0 || M == 0) return 0;
    UP3D_CHECK_ARG(new_xyz && xyz && idx, "up3d_ball_query: null pointer");
    UP3D_CHECK_ARG(nsample <= 1024, "up3d_ball_query: nsample > 1024 not supported");
    volatile float r2 = radius * radius;
    ball_query_kernel<false><<<dim3(div_up(M, BQ_WARPS), B), BQ_WARPS * 32, sizeof(int32_t) * BQ_WARPS * nsample, stream>>>(
        N, M, nsample, r2, new_xyz, xyz, nullptr, idx, nullptr, nullptr);
    UP3D_LAUNCH_OK("ball_query_kernel");
    return 0;
}

int up3d_subsample_group(int B, int N, int G, int K, float radius, const float *xyz, const int32_t *fps_idx,
                         float *center, float *neighborhood, int32_t *idx, up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && N > 0 && G >= 0 && K > 0, "up3d_subsample_group: bad sizes B=%d N=%d G=%d K=%d", B, N, G, K);
    if (B == 0 || G == 0) return 0;
    UP3D_CHECK_ARG(xyz && fps_idx && center && neighborhood, "up3d_subsample_group: null pointer");
    UP3D_CHECK_ARG(K <= 1024, "up3d_subsample_group: K > 1024 not supported");
    volatile float r2 = radius * radius;
    ball_query_kernel<true><<<dim3(div_up(G, BQ_WARPS), B), BQ_WARPS * 32, sizeof(int32_t) * BQ_WARPS * K, stream>>>(
        N, G, K, r2, nullptr, xyz, fps_idx, idx, center, neighborhood);
    UP3D_LAUNCH_OK("ball_query_kernel<fused>");
    return 0;
}


int up3d_knn(int B, int N, int S, int K, int D, const float *xyz, const float *query, int32_t *idx, float *dist,
             up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && N > 0 && S >= 0 && K > 0, "up3d_knn: bad sizes B=%d N=%d S=%d K=%d", B, N, S, K);
    UP3D_CHECK_ARG(K <= N, "up3d_knn: K (%d) must not exceed N (%d)", K, N);
    UP3D_CHECK_ARG(K <= 32 * KNN_MAXR, "up3d_knn: K > %d not supported", 32 * KNN_MAXR);
    UP3D_CHECK_ARG(D == 3 || D == 4, "up3d_knn: point dimension must be 3 or 4 (got %d)", D);
    if (B == 0 || S == 0) return 0;
    UP3D_CHECK_ARG(xyz && query && idx, "up3d_knn: null pointer");
    const dim3 grid(div_up(S, KNN_WARPS), B);
    if (K <= 32 && D == 3) knn_kernel<1, 3><<<grid, KNN_WARPS * 32, 0, stream>>>(N, S, K, xyz, query, idx, dist);
    else if (K <= 32) knn_kernel<1, 4><<<grid, KNN_WARPS * 32, 0, stream>>>(N, S, K, xyz, query, idx, dist);
    else if (D == 3) knn_kernel<2, 3><<<grid, KNN_WARPS * 32, 0, stream>>>(N, S, K, xyz, query, idx, dist);
    else knn_kernel<2, 4><<<grid, KNN_WARPS * 32, 0, stream>>>(N, S, K, xyz, query, idx, dist);
    UP3D_LAUNCH_OK("knn_kernel");
    return 0;
}

int up3d_group_points(int B, int C, int N, int M, int K, const float *points, const int32_t *idx, float *out,
                      up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && C >= 0 && N > 0 && M >= 0 && K >= 0, "up3d_group_points: bad sizes");
    if (B == 0 || C == 0 || M * K == 0) return 0;
    UP3D_CHECK_ARG(points && idx && out, "up3d_group_points: null pointer");
    UP3D_CHECK_ARG(C <= 65535 && B <= 65535, "up3d_group_points: B and C must be <= 65535");
    group_points_kernel<<<dim3(div_up(M * K, 256), C, B), 256, 0, stream>>>(C, N, M, K, points, idx, out);
    UP3D_LAUNCH_OK("group_points_kernel");
    return 0;
}

int up3d_group_points_grad(int B, int C, int N, int M, int K, const float *grad_out, const int32_t *idx,
                           float *grad_points, up3d_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(B >= 0 && C >= 0 && N > 0 && M >= 0 && K >= 0, "up3d_group_points_grad: bad sizes");
    if (B == 0 || C == 0) return 0;
    UP3D_CHECK_ARG(grad_points != nullptr, "up3d_group_points_grad: null pointer");
    UP3D_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)B * C * N, stream));
    if (M * K == 0) return 0;
    UP3D_CHECK_ARG(grad_out && idx, "up3d_group_points_grad: null pointer");
    UP3D_CHECK_ARG(C <= 65535 && B <= 65535, "up3d_group_points_grad: B and C must be <= 65535");
    group_points_grad_kernel<<<dim3(div_up(M * K, 256), C, B), 256, 0, stream>>>(C, N, M, K, grad_out, idx, grad_points);
    UP3D_LAUNCH_OK("group_points_grad_kernel");
    return 0;
}

int up3d_gather_points(int B, int C, int N, int M, const float *points, const int32_t *idx, float *out,
                       up3d_stream_t stream) {
    return up3d_group_points(B, C, N, M, 1, points, idx, out, stream);
}
int up3d_gather_points_grad(int B, int C, int N, int M, const float *grad_out, const int32_t *idx, float *grad_points,
                            up3d_stream_t stream) {
    return up3d_group_points_grad(B, C, N, M, 1, grad_out, idx, grad_points, stream);
}

}  // extern "C"
