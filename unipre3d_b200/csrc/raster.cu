// raster.cu -- batched differentiable 3D-Gaussian-Splatting rasterizer for sm_100a.
//
// Replaces the external diff_gaussian_rasterization CUDA extension the reference binds at
// /root/reference/gaussian_renderer/__init__.py:8,45-61,89-97 (called once per (object, view) from
// /root/reference/train_network.py:418-442).  Arithmetic contract: SURVEY.md Appendix A.
//
// B200-first algorithm (NOT the upstream one):
//   upstream: per view  preprocess -> prefix sum -> D2H sync -> duplicate (tile|depth) keys for every
//             covered tile -> global 64-bit radix sort of I = sum(tiles_touched) pairs -> tile ranges -> blend.
//   here    : ONE launch set for all views of all objects of a step.
//             1. project_kernel : per (view, Gaussian) screen-space record + 32-bit depth key.
//             2. depth_sort_kernel : per view ONE stable warp-radix sort of P depth keys (in shared memory
//                when the view fits), then the records are re-laid-out in depth order.
//                A tile's list in the reference is exactly the depth-ordered sequence filtered by the
//                tile-rectangle test (ties keep ascending Gaussian id), so no per-tile instance list is ever
//                materialised: the I*(12+12+4+4) bytes of key/sort traffic disappear.
//             3. blend_forward_kernel : one CTA per (tile, view) streams the view's depth-ordered records
//                (coalesced 128-bit loads), compacts the ones whose rectangle covers the tile with a warp
//                ballot/scan into shared memory and composites front to back; stops as soon as all 256
//                pixels are saturated (T < 1e-4), which in the reference's regime (sigma >= e^-1) happens
//                after the first chunk.
//             4. blend_backward_kernel : same walk, chunks visited back-to-front, only up to the tile's
//                max n_contrib; per-Gaussian partials are reduced over the 256 pixels with warp shuffles +
//                shared-memory atomics and leave the CTA as ONE global atomic per (tile, Gaussian, term).
//             5. geometry_backward_kernel : per Gaussian, loops over the views of its set in a fixed
//                order and writes dL/d{xyz, scale, rot, opacity, SH}.
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace up3d {

static thread_local char g_err[512] = "";
char *err_buf() { return g_err; }
int set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

__constant__ float c_SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                                 -1.0925484305920792f, 0.5462742152960396f};
__constant__ float c_SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                                 -0.4570457994644658f, 1.445305721320277f,  -0.5900435899266435f};
#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f

constexpr uint32_t KEY_CULLED = 0xFFFFFFFFu;
constexpr int GACC_STRIDE = 12;  // floats per record in the blend-backward accumulator

// ------------------------------------------------------------------------------------------------
// state / scratch layout
// ------------------------------------------------------------------------------------------------
struct State {
    int32_t *n_vis;      // V
    int32_t *radii;      // R
    uint8_t *clamped;    // R (bit ch set when colour channel was clamped at 0)
    uint32_t *key;       // R depth bits or KEY_CULLED (unsorted)
    float2 *u_xy;        // R
    float4 *u_co;        // R conic.xyz, opacity*hs
    float4 *u_rgb;       // R rgb, 1/depth
    uint32_t *u_rect;    // R minx | miny<<8 | maxx<<16 | maxy<<24
    int32_t *s_id;       // R depth-sorted local ids
    float2 *s_xy;
    float4 *s_co;
    float4 *s_rgb;
    uint32_t *s_rect;
    float *final_T;      // V*H*W
    int32_t *n_contrib;  // V*H*W
    // coarse bins (BIN_TILES x BIN_TILES tiles each) for views whose Gaussians have small footprints: bin b of view v lists,
    // in depth order, the positions k (into the view's sorted records) whose rectangle touches the bin.
    int32_t *bin_mode;   // V: 1 = the blend kernels stream bin lists, 0 = they stream the whole view
    int32_t *bin_total;  // V: sum over the view's records of the bins they touch
    int32_t *bin_count;  // V * nbins
    int32_t *bin_start;  // V * nbins: offset of the bin's list inside the view's region of bin_list
    int32_t *bin_list;   // BIN_CAP * R: the view's region starts at BIN_CAP * view_rec_start[v]
    int nbins_x, nbins_y;
    size_t total;
};

constexpr int BIN_TILES = 4;          // a bin is 4 x 4 tiles = 64 x 64 pixels
constexpr int BIN_CAP = 8;            // a view is binned when its records touch <= BIN_CAP bins on average
constexpr int BIN_MIN_SET = 1024;     // smaller sets stream all their records (the object-level transformer regime)
constexpr int BIN_MIN_TILES = 64;

static inline bool bins_enabled(const up3d_raster_desc *d) {
    const int gx = (d->width + UP3D_TILE - 1) / UP3D_TILE, gy = (d->height + UP3D_TILE - 1) / UP3D_TILE;
    return d->max_set_size >= BIN_MIN_SET && gx * gy >= BIN_MIN_TILES;
}

static State carve_state(const up3d_raster_desc *d, void *base) {
    State s;
    size_t off = 0;
    char *b = (char *)base;
    const size_t R = (size_t)(d->n_records > 0 ? d->n_records : 1), V = (size_t)(d->n_views > 0 ? d->n_views : 1);
    const size_t HW = (size_t)d->width * d->height;
#define CARVE(field, type, count)              \
    s.field = (type *)(b + off);               \
    off += align_up(sizeof(type) * (count));
    CARVE(n_vis, int32_t, V)
    CARVE(radii, int32_t, R)
    CARVE(clamped, uint8_t, R)
    CARVE(key, uint32_t, R)
    CARVE(u_xy, float2, R)
    CARVE(u_co, float4, R)
    CARVE(u_rgb, float4, R)
    CARVE(u_rect, uint32_t, R)
    CARVE(s_id, int32_t, R)
    CARVE(s_xy, float2, R)
    CARVE(s_co, float4, R)
    CARVE(s_rgb, float4, R)
    CARVE(s_rect, uint32_t, R)
    CARVE(final_T, float, V * HW)
    CARVE(n_contrib, int32_t, V * HW)
    {
        const int gx = (d->width + UP3D_TILE - 1) / UP3D_TILE, gy = (d->height + UP3D_TILE - 1) / UP3D_TILE;
        s.nbins_x = (gx + BIN_TILES - 1) / BIN_TILES;
        s.nbins_y = (gy + BIN_TILES - 1) / BIN_TILES;
        const size_t nb = (size_t)s.nbins_x * s.nbins_y;
        const bool on = bins_enabled(d);
        CARVE(bin_mode, int32_t, V)
        CARVE(bin_total, int32_t, V)
        CARVE(bin_count, int32_t, on ? V * nb : 1)
        CARVE(bin_start, int32_t, on ? V * nb : 1)
        CARVE(bin_list, int32_t, on ? (size_t)BIN_CAP * R : 1)
    }
#undef CARVE
    s.total = off;
    return s;
}

struct Scratch {
    uint32_t *kA, *kB;  // R each (global-memory sort path)
    int32_t *iA, *iB;   // R each
    float *gacc;        // R * GACC_STRIDE (backward)
    size_t total;
};
static Scratch carve_scratch(const up3d_raster_desc *d, void *base) {
    Scratch s;
    size_t off = 0;
    char *b = (char *)base;
    const size_t R = (size_t)(d->n_records > 0 ? d->n_records : 1);
    s.kA = (uint32_t *)(b + off); off += align_up(4 * R);
    s.kB = (uint32_t *)(b + off); off += align_up(4 * R);
    s.iA = (int32_t *)(b + off); off += align_up(4 * R);
    s.iB = (int32_t *)(b + off); off += align_up(4 * R);
    // gacc aliases the sort buffers' space when larger; keep it simple: separate region
    s.gacc = (float *)(b + off); off += align_up(sizeof(float) * GACC_STRIDE * R);
    s.total = off;
    return s;
}

struct ViewConst {  // host-computed scalars (same float ops as the oracle does on the CPU)
    int W, H, gx, gy;
    float fx, fy, tanfovx, tanfovy, limx, limy, scale_modifier;
    int D, M, antialiasing;
};

// ------------------------------------------------------------------------------------------------
// geometry shared by forward and backward (explicit op order == oracle/raster_oracle.c)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float tp_row(const float *__restrict__ m, int r, float x, float y, float z) {
    return xadd(xfma(m[8 + r], z, xfma(m[r], x, xmul(m[4 + r], y))), m[12 + r]);
}
__device__ __forceinline__ float dot3_c(float a0, float b0, float a1, float b1, float a2, float b2) {
    return xfma(a2, b2, xfma(a0, b0, xmul(a1, b1)));
}

struct Cov3 { float c[6]; float R[9]; float s[3]; };

__device__ __forceinline__ void cov3d_from_scale_rot(const float s3[3], float mod, const float q[4], Cov3 &o) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    float *R = o.R;
    R[0] = xfma(-2.f, xfma(y, y, xmul(z, z)), 1.f);
    R[1] = xmul(2.f, xfma(x, y, -xmul(r, z)));
    R[2] = xmul(2.f, xfma(x, z, xmul(r, y)));
    R[3] = xmul(2.f, xfma(x, y, xmul(r, z)));
    R[4] = xfma(-2.f, xfma(x, x, xmul(z, z)), 1.f);
    R[5] = xmul(2.f, xfma(y, z, -xmul(r, x)));
    R[6] = xmul(2.f, xfma(x, z, -xmul(r, y)));
    R[7] = xmul(2.f, xfma(y, z, xmul(r, x)));
    R[8] = xfma(-2.f, xfma(x, x, xmul(y, y)), 1.f);
    o.s[0] = xmul(mod, s3[0]); o.s[1] = xmul(mod, s3[1]); o.s[2] = xmul(mod, s3[2]);
    float A[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int i = 0; i < 3; ++i) A[a * 3 + i] = xmul(o.s[i], R[a * 3 + i]);
#define SIG(a, b) dot3_c(A[(a)*3 + 0], A[(b)*3 + 0], A[(a)*3 + 1], A[(b)*3 + 1], A[(a)*3 + 2], A[(b)*3 + 2])
    o.c[0] = SIG(0, 0); o.c[1] = SIG(0, 1); o.c[2] = SIG(0, 2);
    o.c[3] = SIG(1, 1); o.c[4] = SIG(1, 2); o.c[5] = SIG(2, 2);
#undef SIG
}

struct Cov2 { float cov[3]; float T[6]; float tc[3]; bool cx, cy; };

__device__ __forceinline__ void cov2d_ewa(const float t_in[3], const ViewConst &vc, const float cov6[6],
                                          const float *__restrict__ view, Cov2 &o) {
    float tx = t_in[0], ty = t_in[1];
    const float tz = t_in[2];
    const float txtz = xdiv(tx, tz), tytz = xdiv(ty, tz);
    tx = xmul(fminf(vc.limx, fmaxf(-vc.limx, txtz)), tz);
    ty = xmul(fminf(vc.limy, fmaxf(-vc.limy, tytz)), tz);
    o.cx = (txtz < -vc.limx || txtz > vc.limx);
    o.cy = (tytz < -vc.limy || tytz > vc.limy);
    const float tz2 = xmul(tz, tz);
    const float J00 = xdiv(vc.fx, tz), J02 = -xdiv(xmul(vc.fx, tx), tz2);
    const float J11 = xdiv(vc.fy, tz), J12 = -xdiv(xmul(vc.fy, ty), tz2);
    float *T = o.T;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        T[i] = xfma(view[4 * i + 2], J02, xmul(view[4 * i + 0], J00));
        T[3 + i] = xfma(view[4 * i + 2], J12, xmul(view[4 * i + 1], J11));
    }
    const float V[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
    float X[6];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            X[r * 3 + c] = dot3_c(T[r * 3 + 0], V[0 * 3 + c], T[r * 3 + 1], V[1 * 3 + c], T[r * 3 + 2], V[2 * 3 + c]);
    o.cov[0] = dot3_c(X[0], T[0], X[1], T[1], X[2], T[2]);
    o.cov[1] = dot3_c(X[3], T[0], X[4], T[1], X[5], T[2]);
    o.cov[2] = dot3_c(X[3], T[3], X[4], T[4], X[5], T[5]);
    o.tc[0] = tx; o.tc[1] = ty; o.tc[2] = tz;
}

__device__ __forceinline__ float ndc2pix(float v, int S) {
    return (float)__dmul_rn(__fma_rn(__dadd_rn((double)v, 1.0), (double)S, -1.0), 0.5);
}

// SH basis evaluation (A.5).  Writes the (deg+1)^2 basis weights; colour = sum_l w_l * sh_l + 0.5.
template <int MAXM>
__device__ __forceinline__ void sh_basis(int deg, float x, float y, float z, float *w) {
    w[0] = SH_C0;
    if (MAXM >= 4 && deg > 0) {
        w[1] = -SH_C1 * y; w[2] = SH_C1 * z; w[3] = -SH_C1 * x;
    }
    if (MAXM >= 9 && deg > 1) {
        const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
        w[4] = c_SH_C2[0] * xy; w[5] = c_SH_C2[1] * yz; w[6] = c_SH_C2[2] * (2.0f * zz - xx - yy);
        w[7] = c_SH_C2[3] * xz; w[8] = c_SH_C2[4] * (xx - yy);
        if (MAXM >= 16 && deg > 2) {
            w[9] = c_SH_C3[0] * y * (3.0f * xx - yy); w[10] = c_SH_C3[1] * xy * z;
            w[11] = c_SH_C3[2] * y * (4.0f * zz - xx - yy); w[12] = c_SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy);
            w[13] = c_SH_C3[4] * x * (4.0f * zz - xx - yy); w[14] = c_SH_C3[5] * z * (xx - yy);
            w[15] = c_SH_C3[6] * x * (xx - 3.0f * yy);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 1. projection (A.1): one thread per (view, Gaussian) record
// ------------------------------------------------------------------------------------------------
struct ProjectArgs {
    ViewConst vc;
    const float *means3D, *shs, *colors, *opacities, *scales, *rotations, *viewmats, *projmats, *campos;
    const int32_t *set_offsets, *view_set, *view_rec_start;
    int32_t *radii_out;
    State st;
};

__global__ void __launch_bounds__(256) project_kernel(const ProjectArgs a) {
    const int v = blockIdx.y;
    const int set = a.view_set[v];
    const int g0 = a.set_offsets[set], P = a.set_offsets[set + 1] - g0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int rec = a.view_rec_start[v] + i;
    const int gi = g0 + i;
    const ViewConst &vc = a.vc;
    const float *view = a.viewmats + 16 * v, *proj = a.projmats + 16 * v;

    int radius_i = 0;
    uint32_t key = KEY_CULLED, rect = 0;
    uint8_t clamp_bits = 0;
    float2 xy = make_float2(0.f, 0.f);
    float4 co = make_float4(0.f, 0.f, 0.f, 0.f), rgb = make_float4(0.f, 0.f, 0.f, 0.f);

    const float px_ = a.means3D[3 * gi], py_ = a.means3D[3 * gi + 1], pz_ = a.means3D[3 * gi + 2];
    float t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = tp_row(view, r, px_, py_, pz_);
    if (t[2] > 0.2f) {
        float ph[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ph[r] = tp_row(proj, r, px_, py_, pz_);
        const float p_w = xdiv(1.0f, xadd(ph[3], 0.0000001f));
        const float ndx = xmul(ph[0], p_w), ndy = xmul(ph[1], p_w);
        const float s3[3] = {a.scales[3 * gi], a.scales[3 * gi + 1], a.scales[3 * gi + 2]};
        const float q[4] = {a.rotations[4 * gi], a.rotations[4 * gi + 1], a.rotations[4 * gi + 2], a.rotations[4 * gi + 3]};
        Cov3 c3;
        cov3d_from_scale_rot(s3, vc.scale_modifier, q, c3);
        Cov2 c2;
        cov2d_ewa(t, vc, c3.c, view, c2);
        const float h_var = 0.3f;
        const float det0 = xfma(c2.cov[0], c2.cov[2], -xmul(c2.cov[1], c2.cov[1]));
        const float ca = xadd(c2.cov[0], h_var), cb = c2.cov[1], cc = xadd(c2.cov[2], h_var);
        const float det = xfma(ca, cc, -xmul(cb, cb));
        float hs = 1.0f;
        if (vc.antialiasing) hs = xsqrt(fmaxf(0.000025f, xdiv(det0, det)));
        if (det != 0.0f) {
            const float det_inv = xdiv(1.f, det);
            const float mid = xmul(0.5f, xadd(ca, cc));
            const float disc = xsqrt(fmaxf(0.1f, xfma(mid, mid, -det)));
            const float l1 = xadd(mid, disc), l2 = xsub(mid, disc);
            const float my_radius = ceilf(xmul(3.f, xsqrt(fmaxf(l1, l2))));
            const float ix = ndc2pix(ndx, vc.W), iy = ndc2pix(ndy, vc.H);
            const int rminx = min(vc.gx, max(0, __float2int_rz(xdiv(xsub(ix, my_radius), 16.f))));
            const int rminy = min(vc.gy, max(0, __float2int_rz(xdiv(xsub(iy, my_radius), 16.f))));
            const int rmaxx = min(vc.gx, max(0, __float2int_rz(xdiv(xadd(xadd(ix, my_radius), 15.f), 16.f))));
            const int rmaxy = min(vc.gy, max(0, __float2int_rz(xdiv(xadd(xadd(iy, my_radius), 15.f), 16.f))));
            if ((rmaxx - rminx) * (rmaxy - rminy) != 0) {
                float col[3];
                if (a.colors) {
                    col[0] = a.colors[3 * gi]; col[1] = a.colors[3 * gi + 1]; col[2] = a.colors[3 * gi + 2];
                } else {
                    const float *cp = a.campos + 3 * v;
                    float dx = px_ - cp[0], dy = py_ - cp[1], dz = pz_ - cp[2];
                    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
                    dx /= len; dy /= len; dz /= len;
                    float w[16];
                    sh_basis<16>(vc.D, dx, dy, dz, w);
                    const int nb = (vc.D + 1) * (vc.D + 1);
                    const float *sh = a.shs + (size_t)gi * vc.M * 3;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        float r = 0.f;
                        for (int l = 0; l < nb; ++l) r += w[l] * sh[l * 3 + ch];
                        r += 0.5f;
                        if (r < 0.f) clamp_bits |= (1 << ch);
                        col[ch] = fmaxf(r, 0.f);
                    }
                }
                radius_i = __float2int_rz(my_radius);
                key = __float_as_uint(t[2]);
                xy = make_float2(ix, iy);
                co = make_float4(xmul(cc, det_inv), xmul(-cb, det_inv), xmul(ca, det_inv), xmul(a.opacities[gi], hs));
                rgb = make_float4(col[0], col[1], col[2], 1.0f / t[2]);
                rect = (uint32_t)rminx | ((uint32_t)rminy << 8) | ((uint32_t)rmaxx << 16) | ((uint32_t)rmaxy << 24);
            }
        }
    }
    a.radii_out[rec] = radius_i;
    a.st.radii[rec] = radius_i;
    a.st.clamped[rec] = clamp_bits;
    a.st.key[rec] = key;
    a.st.u_xy[rec] = xy;
    a.st.u_co[rec] = co;
    a.st.u_rgb[rec] = rgb;
    a.st.u_rect[rec] = rect;
}

// ------------------------------------------------------------------------------------------------
// 2. per-view stable depth sort (warp-radix, LSD 4 x 8 bit) + re-layout in depth order
// ------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 1024;
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_RPW = 8;                                  // rows (of 32 keys) per warp per tile
constexpr int SORT_TILE = SORT_WARPS * SORT_RPW * 32;        // 8192 keys per tile
constexpr int SORT_SMEM_MAX_KEYS = 12288;                    // 4 arrays * 4 B * 12288 = 192 KB
constexpr int SORT_MISC = 16;                                 // [0] survivors, [1..8] scan warp totals
constexpr int SORT_FIXED_SMEM = (SORT_WARPS * 256 + 256 + 256 + SORT_MISC) * 4;

struct SortArgs {
    const int32_t *view_rec_start;
    State st;
    Scratch sc;
    int use_smem;  // 1: all four key/id arrays live in dynamic shared memory
};

__global__ void __launch_bounds__(SORT_THREADS, 1) depth_sort_kernel(const SortArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t *wcnt = (uint32_t *)smem_raw;            // [SORT_WARPS][256]
    uint32_t *hist = wcnt + SORT_WARPS * 256;         // [256]
    uint32_t *dbase = hist + 256;                     // [256]
    uint32_t *misc = dbase + 256;                     // [SORT_MISC]
    const int v = blockIdx.x;
    const int rec0 = a.view_rec_start[v];
    const int n = a.view_rec_start[v + 1] - rec0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    uint32_t *kA, *kB;
    int32_t *iA, *iB;
    if (a.use_smem) {
        // capacity rounded to a multiple of 4 keys so every array stays 16-byte aligned
        const int cap = (n + 3) & ~3;
        kA = misc + SORT_MISC; kB = kA + cap; iA = (int32_t *)(kB + cap); iB = iA + cap;
    } else {
        kA = a.sc.kA + rec0; kB = a.sc.kB + rec0; iA = a.sc.iA + rec0; iB = a.sc.iB + rec0;
    }
    if (tid == 0) misc[0] = 0;
    __syncthreads();
    // load keys, count survivors
    int my_valid = 0;
    for (int i = tid; i < n; i += SORT_THREADS) {
        const uint32_t k = a.st.key[rec0 + i];
        kA[i] = k;
        iA[i] = i;
        my_valid += (k != KEY_CULLED);
    }
    my_valid = __reduce_add_sync(0xffffffffu, my_valid);
    if (lane == 0 && my_valid) atomicAdd(&misc[0], (uint32_t)my_valid);
    __syncthreads();
    const int n_vis = (int)misc[0];
    if (tid == 0) a.st.n_vis[v] = n_vis;

    for (int shift = 0; shift < 32; shift += 8) {
        // ---- digit histogram (warp-aggregated shared atomics)
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int base = 0; base < n; base += SORT_THREADS) {
            const int i = base + tid;
            const uint32_t dgt = (i < n) ? ((kA[i] >> shift) & 0xFFu) : 0xFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, dgt);
            if (dgt != 0xFFFFu && lane == (__ffs(peers) - 1)) atomicAdd(&hist[dgt], (uint32_t)__popc(peers));
        }
        __syncthreads();
        // ---- identity pass? (all keys share this digit)
        const int same = __syncthreads_or(tid < 256 && hist[tid] == (uint32_t)n);
        if (same) continue;
        // ---- exclusive scan of the 256 bins (warp 0..7 each scan 32 bins, then add warp offsets)
        if (tid < 256) {
            uint32_t x = hist[tid], incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            dbase[tid] = incl - x;
            if (lane == 31) misc[1 + warp] = incl;  // warp totals (warp < 8 here)
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t add = 0;
            for (int w = 0; w < warp; ++w) add += misc[1 + w];
            dbase[tid] += add;
        }
        __syncthreads();
        // ---- stable scatter, tile by tile
        for (int tile0 = 0; tile0 < n; tile0 += SORT_TILE) {
            for (int j = tid; j < SORT_WARPS * 256; j += SORT_THREADS) wcnt[j] = 0;
            __syncthreads();
            uint32_t keys[SORT_RPW];
            int32_t ids[SORT_RPW];
            uint32_t rank[SORT_RPW];
            uint32_t *mycnt = wcnt + warp * 256;
#pragma unroll
            for (int r = 0; r < SORT_RPW; ++r) {
                const int i = tile0 + (warp * SORT_RPW + r) * 32 + lane;
                const bool valid = i < n;
                keys[r] = valid ? kA[i] : 0u;
                ids[r] = valid ? iA[i] : 0;
                const uint32_t dgt = valid ? ((keys[r] >> shift) & 0xFFu) : 0xFFFFu;
                const unsigned peers = __match_any_sync(0xffffffffu, dgt);
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (valid && lane == leader) {
                    old = mycnt[dgt];
                    mycnt[dgt] = old + (uint32_t)__popc(peers);
                }
                old = __shfl_sync(0xffffffffu, old, leader);
                rank[r] = old + (uint32_t)__popc(peers & lanemask_lt());
                __syncwarp();
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t run = dbase[tid];
                for (int w = 0; w < SORT_WARPS; ++w) {
                    const uint32_t c = wcnt[w * 256 + tid];
                    wcnt[w * 256 + tid] = run;
                    run += c;
                }
                dbase[tid] = run;
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < SORT_RPW; ++r) {
                const int i = tile0 + (warp * SORT_RPW + r) * 32 + lane;
                if (i < n) {
                    const uint32_t pos = mycnt[(keys[r] >> shift) & 0xFFu] + rank[r];
                    kB[pos] = keys[r];
                    iB[pos] = ids[r];
                }
            }
            __syncthreads();
        }
        uint32_t *tk = kA; kA = kB; kB = tk;
        int32_t *ti = iA; iA = iB; iB = ti;
    }
    // ---- re-layout the surviving records in depth order (coalesced writes, gathered reads)
    for (int k = tid; k < n_vis; k += SORT_THREADS) {
        const int id = iA[k];
        const int src = rec0 + id, dst = rec0 + k;
        a.st.s_id[dst] = id;
        a.st.s_xy[dst] = a.st.u_xy[src];
        a.st.s_co[dst] = a.st.u_co[src];
        a.st.s_rgb[dst] = a.st.u_rgb[src];
        a.st.s_rect[dst] = a.st.u_rect[src];
    }
}

// ------------------------------------------------------------------------------------------------
// 2a'. the same sort for LARGE sets (> SORT_SMEM_MAX_KEYS Gaussians per view: scene level, 10^5 keys): one thread-block
// CLUSTER of SORT_CLUSTER CTAs per view instead of one CTA.  CTA c owns the c-th contiguous slice of the view's keys; per
// 8-bit pass every CTA histograms its slice, the 256-bin histograms are exchanged through distributed shared memory
// (digit d of CTA c starts at  sum_{d' < d} total[d'] + sum_{c' < c} hist[c'][d]  -- a stable LSD radix sort across CTAs),
// and each CTA scatters its slice.  Keys / ids ping-pong in the (L2-resident) global scratch.  Two cluster barriers per
// pass.  Same order as the single-CTA kernel, bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int SORT_CLUSTER = 8;

__device__ __forceinline__ uint32_t sort_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void sort_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t sort_dsmem_ld(const uint32_t *local, uint32_t rank) {
    uint32_t remote, v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"((uint32_t)__cvta_generic_to_shared(local)), "r"(rank));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(remote) : "memory");
    return v;
}

__global__ void __cluster_dims__(SORT_CLUSTER, 1, 1) __launch_bounds__(SORT_THREADS, 1)
depth_sort_cluster_kernel(const SortArgs a) {
    __shared__ uint32_t wcnt[SORT_WARPS * 256];
    __shared__ uint32_t hist[256];          // this CTA's slice histogram (read remotely by the peers)
    __shared__ uint32_t dbase[256];
    __shared__ uint32_t tot[256];
    __shared__ uint32_t misc[SORT_MISC];
    __shared__ uint32_t s_valid;            // un-culled keys of this CTA's slice (read remotely)
    const int v = blockIdx.x / SORT_CLUSTER;
    const int c = (int)sort_cluster_rank();
    const int rec0 = a.view_rec_start[v];
    const int n = a.view_rec_start[v + 1] - rec0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = ((n + SORT_CLUSTER - 1) / SORT_CLUSTER + 31) & ~31;
    const int s0 = min(n, c * per), s1 = min(n, s0 + per);

    uint32_t *kA = a.sc.kA + rec0, *kB = a.sc.kB + rec0;
    int32_t *iA = a.sc.iA + rec0, *iB = a.sc.iB + rec0;
    if (tid == 0) s_valid = 0;
    __syncthreads();
    int my_valid = 0;
    for (int i = s0 + tid; i < s1; i += SORT_THREADS) {
        const uint32_t k = a.st.key[rec0 + i];
        kA[i] = k;
        iA[i] = i;
        my_valid += (k != KEY_CULLED);
    }
    my_valid = __reduce_add_sync(0xffffffffu, my_valid);
    if (lane == 0 && my_valid) atomicAdd(&s_valid, (uint32_t)my_valid);
    __syncthreads();
    sort_cluster_sync();
    int n_vis = 0;
    for (int cc = 0; cc < SORT_CLUSTER; ++cc) n_vis += (int)sort_dsmem_ld(&s_valid, (uint32_t)cc);
    if (c == 0 && tid == 0) a.st.n_vis[v] = n_vis;

    for (int shift = 0; shift < 32; shift += 8) {
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int base = s0; base < s1; base += SORT_THREADS) {
            const int i = base + tid;
            const uint32_t dgt = (i < s1) ? ((kA[i] >> shift) & 0xFFu) : 0xFFFFu;
            const unsigned peers = __match_any_sync(0xffffffffu, dgt);
            if (dgt != 0xFFFFu && lane == (__ffs(peers) - 1)) atomicAdd(&hist[dgt], (uint32_t)__popc(peers));
        }
        __syncthreads();
        sort_cluster_sync();                              // every slice histogram is complete and visible
        uint32_t before = 0, total = 0;
        if (tid < 256) {
            for (int cc = 0; cc < SORT_CLUSTER; ++cc) {
                const uint32_t h = sort_dsmem_ld(&hist[tid], (uint32_t)cc);
                total += h;
                before += (cc < c) ? h : 0u;
            }
            tot[tid] = total;
        }
        const int same = __syncthreads_or(tid < 256 && total == (uint32_t)n);     // identical in every CTA of the cluster
        if (!same) {
            if (tid < 256) {                              // exclusive scan of the 256 totals (as the single-CTA kernel)
                uint32_t x = tot[tid], incl = x;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += y;
                }
                dbase[tid] = incl - x + before;
                if (lane == 31) misc[1 + warp] = incl;
            }
            __syncthreads();
            if (tid < 256) {
                uint32_t add = 0;
                for (int w = 0; w < warp; ++w) add += misc[1 + w];
                dbase[tid] += add;
            }
            __syncthreads();
            for (int tile0 = s0; tile0 < s1; tile0 += SORT_TILE) {
                for (int j = tid; j < SORT_WARPS * 256; j += SORT_THREADS) wcnt[j] = 0;
                __syncthreads();
                uint32_t keys[SORT_RPW];
                int32_t ids[SORT_RPW];
                uint32_t rank[SORT_RPW];
                uint32_t *mycnt = wcnt + warp * 256;
#pragma unroll
                for (int r = 0; r < SORT_RPW; ++r) {
                    const int i = tile0 + (warp * SORT_RPW + r) * 32 + lane;
                    const bool valid = i < s1;
                    keys[r] = valid ? kA[i] : 0u;
                    ids[r] = valid ? iA[i] : 0;
                    const uint32_t dgt = valid ? ((keys[r] >> shift) & 0xFFu) : 0xFFFFu;
                    const unsigned peers = __match_any_sync(0xffffffffu, dgt);
                    const int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if (valid && lane == leader) {
                        old = mycnt[dgt];
                        mycnt[dgt] = old + (uint32_t)__popc(peers);
                    }
                    old = __shfl_sync(0xffffffffu, old, leader);
                    rank[r] = old + (uint32_t)__popc(peers & lanemask_lt());
                    __syncwarp();
                }
                __syncthreads();
                if (tid < 256) {
                    uint32_t run = dbase[tid];
                    for (int w = 0; w < SORT_WARPS; ++w) {
                        const uint32_t cnt = wcnt[w * 256 + tid];
                        wcnt[w * 256 + tid] = run;
                        run += cnt;
                    }
                    dbase[tid] = run;
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < SORT_RPW; ++r) {
                    const int i = tile0 + (warp * SORT_RPW + r) * 32 + lane;
                    if (i < s1) {
                        const uint32_t pos = mycnt[(keys[r] >> shift) & 0xFFu] + rank[r];
                        kB[pos] = keys[r];
                        iB[pos] = ids[r];
                    }
                }
                __syncthreads();
            }
            uint32_t *tk = kA; kA = kB; kB = tk;
            int32_t *ti = iA; iA = iB; iB = ti;
            __threadfence();
        }
        sort_cluster_sync();                              // scatter visible cluster-wide; remote histogram reads are done
    }
    // re-layout: each CTA gathers its slice of the sorted order
    for (int k = s0 + tid; k < min(s1, n_vis); k += SORT_THREADS) {
        const int id = iA[k];
        const int src = rec0 + id, dst = rec0 + k;
        a.st.s_id[dst] = id;
        a.st.s_xy[dst] = a.st.u_xy[src];
        a.st.s_co[dst] = a.st.u_co[src];
        a.st.s_rgb[dst] = a.st.u_rgb[src];
        a.st.s_rect[dst] = a.st.u_rect[src];
    }
}

// ------------------------------------------------------------------------------------------------
// 2b. coarse bins (small-footprint views only)
//
// In the reference's object regime every Gaussian covers every tile and a tile's list IS the view's depth order, so the
// blend kernels stream the view's records.  With 10^5 small splats on 1024 tiles that would stage tiles * P records per
// view; instead each 4x4-tile bin gets the depth-ordered list of record positions touching it (order-preserving ballot
// compaction of the already sorted records, so per-tile lists stay bit-exact), and a tile streams only its bin.
// Three small kernels: per-view total of touched bins -> decision; per-bin counts; per-bin ordered fill.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool rect_covers(uint32_t rect, int tx, int ty) {
    const int minx = rect & 0xFF, miny = (rect >> 8) & 0xFF, maxx = (rect >> 16) & 0xFF, maxy = rect >> 24;
    return tx >= minx && tx < maxx && ty >= miny && ty < maxy;
}
// bin range [bx0, bx1] x [by0, by1] touched by a (non-empty) tile rectangle
__device__ __forceinline__ void rect_bins(uint32_t rect, int &bx0, int &by0, int &bx1, int &by1) {
    const int minx = rect & 0xFF, miny = (rect >> 8) & 0xFF, maxx = (rect >> 16) & 0xFF, maxy = rect >> 24;
    bx0 = minx / BIN_TILES; by0 = miny / BIN_TILES; bx1 = (maxx - 1) / BIN_TILES; by1 = (maxy - 1) / BIN_TILES;
}
__device__ __forceinline__ bool rect_touches_bin(uint32_t rect, int bx, int by) {
    int bx0, by0, bx1, by1;
    rect_bins(rect, bx0, by0, bx1, by1);
    return bx >= bx0 && bx <= bx1 && by >= by0 && by <= by1;
}

struct BinArgs {
    const int32_t *view_rec_start;
    State st;
};

// grid (ceil(max_set/256), V): bin_total[v] = sum over sorted records of the number of bins they touch
__global__ void __launch_bounds__(256) bin_total_kernel(const BinArgs a) {
    const int v = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    const int rec0 = a.view_rec_start[v], n = a.st.n_vis[v];
    int nb = 0;
    if (k < n) {
        int bx0, by0, bx1, by1;
        rect_bins(a.st.s_rect[rec0 + k], bx0, by0, bx1, by1);
        nb = (bx1 - bx0 + 1) * (by1 - by0 + 1);
    }
    nb = __reduce_add_sync(0xffffffffu, nb);
    if ((threadIdx.x & 31) == 0 && nb) atomicAdd(&a.st.bin_total[v], nb);
}

// grid (ceil(max_set/256), V): decides the mode, then (binned views only) per-bin counts
__global__ void __launch_bounds__(256) bin_count_kernel(const BinArgs a) {
    const int v = blockIdx.y, k = blockIdx.x * 256 + threadIdx.x;
    const int rec0 = a.view_rec_start[v], n = a.st.n_vis[v];
    const int cap = BIN_CAP * (a.view_rec_start[v + 1] - rec0);
    const bool binned = a.st.bin_total[v] <= cap;
    if (blockIdx.x == 0 && threadIdx.x == 0) a.st.bin_mode[v] = binned ? 1 : 0;
    if (!binned || k >= n) return;
    int bx0, by0, bx1, by1;
    rect_bins(a.st.s_rect[rec0 + k], bx0, by0, bx1, by1);
    int32_t *cnt = a.st.bin_count + (size_t)v * a.st.nbins_x * a.st.nbins_y;
    for (int by = by0; by <= by1; ++by)
        for (int bx = bx0; bx <= bx1; ++bx) atomicAdd(&cnt[by * a.st.nbins_x + bx], 1);
}

// grid (nbins, V): ordered fill of bin b's list (positions k, ascending = depth order)
__global__ void __launch_bounds__(256) bin_fill_kernel(const BinArgs a) {
    __shared__ int warp_cnt[8];
    __shared__ int s_start;
    const int v = blockIdx.y, b = blockIdx.x;
    if (!a.st.bin_mode[v]) return;
    const int nb = a.st.nbins_x * a.st.nbins_y;
    const int32_t *cnt = a.st.bin_count + (size_t)v * nb;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // exclusive prefix of the counts of bins < b
    int part = 0;
    for (int i = tid; i < b; i += 256) part += cnt[i];
    part = __reduce_add_sync(0xffffffffu, part);
    if (lane == 0) warp_cnt[warp] = part;
    __syncthreads();
    if (tid == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += warp_cnt[w];
        s_start = t;
        a.st.bin_start[(size_t)v * nb + b] = t;
    }
    __syncthreads();
    if (cnt[b] == 0) return;
    const int rec0 = a.view_rec_start[v], n = a.st.n_vis[v];
    int32_t *dst = a.st.bin_list + (size_t)BIN_CAP * rec0 + s_start;
    const int bx = b % a.st.nbins_x, by = b / a.st.nbins_x;
    int run = 0;
    for (int base = 0; base < n; base += 256) {
        const int k = base + tid;
        const bool hit = (k < n) && rect_touches_bin(a.st.s_rect[rec0 + k], bx, by);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        __syncthreads();
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { off += (w < warp) ? warp_cnt[w] : 0; total += warp_cnt[w]; }
        if (hit) dst[run + off + __popc(bal & lanemask_lt())] = k;
        run += total;
    }
}

// ------------------------------------------------------------------------------------------------
// 3. blend forward (A.6)
// ------------------------------------------------------------------------------------------------

// power = -0.5f * (co.x*dx*dx + co.z*dy*dy) - co.y*dx*dy with the operation order nvcc gives the upstream
// expression (so the power > 0 / alpha < 1/255 skip decisions do not depend on this compiler's contraction
// choices and agree with oracle/raster_oracle.c): A = fma(co.x*dx, dx, (co.z*dy)*dy); fma(-0.5, A, -((co.y*dx)*dy)).
__device__ __forceinline__ float gauss_power(const float4 co, float dx, float dy) {
    const float A = xfma(xmul(co.x, dx), dx, xmul(xmul(co.z, dy), dy));
    return xfma(-0.5f, A, -xmul(xmul(co.y, dx), dy));
}

// ---- TMA (bulk async copy) + mbarrier helpers: cp.async.bulk global -> shared, completion on an mbarrier -------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}

// One 256-record chunk of a view's depth-ordered records, staged in shared memory exactly as it lies in HBM
// (five contiguous slices -> five bulk copies on one mbarrier), then compacted IN PLACE to the entries whose
// rectangle covers this tile (order preserved).  In the reference's regime every record covers every tile, the
// compaction is the identity and the staged chunk is consumed as is.
struct __align__(128) StageBuf {
    float2 xy[UP3D_TILE_PIX];
    float4 co[UP3D_TILE_PIX];
    float4 rgb[UP3D_TILE_PIX];
    int32_t id[UP3D_TILE_PIX];
    uint32_t rect[UP3D_TILE_PIX];
};

// thread 0 only: arm the barrier and launch the bulk copies of records [base, base + rows) (rows rounded up to 4 so
// every slice is a multiple of 16 bytes; the over-read stays inside the state blob and is never consumed)
template <bool WITH_ID>
__device__ __forceinline__ void stage_issue(StageBuf &sb, uint64_t *bar, const State &st, int rec0, int n, int base) {
    const int rows = min(UP3D_TILE_PIX, n - base);
    const uint32_t r4 = (uint32_t)((rows + 3) & ~3);
    fence_proxy_async();   // earlier generic-proxy accesses to this buffer are ordered before the async-proxy writes
    mbar_expect_tx(bar, r4 * (8u + 16u + 16u + 4u + (WITH_ID ? 4u : 0u)));
    const size_t o = (size_t)rec0 + base;
    bulk_g2s(sb.xy, st.s_xy + o, r4 * 8u, bar);
    bulk_g2s(sb.co, st.s_co + o, r4 * 16u, bar);
    bulk_g2s(sb.rgb, st.s_rgb + o, r4 * 16u, bar);
    bulk_g2s(sb.rect, st.s_rect + o, r4 * 4u, bar);
    if (WITH_ID) bulk_g2s(sb.id, st.s_id + o, r4 * 4u, bar);
}

// fallback when the view's record range is not 16-byte aligned for every slice: plain coalesced loads
template <bool WITH_ID>
__device__ __forceinline__ void stage_fill_plain(StageBuf &sb, const State &st, int rec0, int n, int base) {
    const int tid = threadIdx.x, k = base + tid;
    if (k < n) {
        sb.xy[tid] = st.s_xy[rec0 + k];
        sb.co[tid] = st.s_co[rec0 + k];
        sb.rgb[tid] = st.s_rgb[rec0 + k];
        sb.rect[tid] = st.s_rect[rec0 + k];
        if (WITH_ID) sb.id[tid] = st.s_id[rec0 + k];
    }
}

// binned views: chunk [base, base + 256) of a bin's list of record positions -> gathered into the staging buffer
template <bool WITH_ID>
__device__ __forceinline__ void stage_fill_indexed(StageBuf &sb, const State &st, int rec0, const int32_t *list, int n, int base) {
    const int tid = threadIdx.x, i = base + tid;
    if (i < n) {
        const int k = rec0 + list[i];
        sb.xy[tid] = st.s_xy[k];
        sb.co[tid] = st.s_co[k];
        sb.rgb[tid] = st.s_rgb[k];
        sb.rect[tid] = st.s_rect[k];
        if (WITH_ID) sb.id[tid] = st.s_id[k];
    }
}

// The candidate records of tile (tx, ty) of view v: the whole view (list == nullptr, n = n_vis) or its bin's list.
struct Candidates {
    const int32_t *list;
    int n;
};
__device__ __forceinline__ Candidates tile_candidates(const State &st, const int32_t *view_rec_start, int v, int tx, int ty) {
    Candidates c;
    if (st.bin_mode[v]) {
        const int nb = st.nbins_x * st.nbins_y, b = (ty / BIN_TILES) * st.nbins_x + tx / BIN_TILES;
        c.list = st.bin_list + (size_t)BIN_CAP * view_rec_start[v] + st.bin_start[(size_t)v * nb + b];
        c.n = st.bin_count[(size_t)v * nb + b];
    } else {
        c.list = nullptr;
        c.n = st.n_vis[v];
    }
    return c;
}

// In-place ordered compaction of a staged chunk to the records covering tile (tx, ty).  Returns the number kept.
// All 256 threads call; the staged data must be visible (mbarrier wait / __syncthreads) to all of them.
template <bool WITH_ID>
__device__ __forceinline__ int stage_compact(StageBuf &sb, int *warp_cnt, int n, int base, int tx, int ty) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = min(UP3D_TILE_PIX, n - base);
    const bool hit = tid < rows && rect_covers(sb.rect[tid], tx, ty);
    const int total = __syncthreads_count(hit);
    if (total == rows) return total;           // dense: identity
    float2 xy; float4 co, rgb; int32_t id = 0;
    if (hit) { xy = sb.xy[tid]; co = sb.co[tid]; rgb = sb.rgb[tid]; if (WITH_ID) id = sb.id[tid]; }
    const unsigned bal = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();                            // all reads done, warp counts visible
    int off = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) off += (w < warp) ? warp_cnt[w] : 0;
    if (hit) {
        const int slot = off + __popc(bal & lanemask_lt());
        sb.xy[slot] = xy; sb.co[slot] = co; sb.rgb[slot] = rgb;
        if (WITH_ID) sb.id[slot] = id;
    }
    __syncthreads();
    return total;
}

struct BlendArgs {
    int W, H;
    const int32_t *view_rec_start;
    const float *bg;
    float *out_color;  // V,3,H,W
    float *invdepth;   // V,1,H,W or null
    State st;
};

// BINS: compiled-in support for binned views (host enables it for large sets on many tiles only; the object-level
// configurations run the leaner BINS = false instantiation: 32 vs 48 registers -> 8 vs 5 resident CTAs per SM).
template <bool BINS>
__global__ void __launch_bounds__(UP3D_TILE_PIX) blend_forward_kernel(const BlendArgs a) {
    __shared__ StageBuf stage[2];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ int warp_cnt[8];
    const int v = blockIdx.z, tx = blockIdx.x, ty = blockIdx.y;
    const int tid = threadIdx.x;
    const int px = tx * UP3D_TILE + (tid & 15), py = ty * UP3D_TILE + (tid >> 4);
    const bool inside = px < a.W && py < a.H;
    const int rec0 = a.view_rec_start[v];
    Candidates cand{nullptr, 0};
    if (BINS) cand = tile_candidates(a.st, a.view_rec_start, v, tx, ty);
    else cand.n = a.st.n_vis[v];
    const int n = cand.n;
    const float pfx = (float)px, pfy = (float)py;
    const bool tma = cand.list == nullptr && (rec0 & 3) == 0;   // every slice of the view's records starts 16-byte aligned
    const int nchunks = (n + UP3D_TILE_PIX - 1) / UP3D_TILE_PIX;
    if (tma && tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    __syncthreads();
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f, ID = 0.f;
    int contributor = 0, last_contributor = 0;
    int issued = 0;                              // chunks whose bulk copy has been launched (block-uniform)
    int c = 0;
    for (; c < nchunks; ++c) {
        // everyone finished reading the buffers of chunk c-1 (and c-2) once this barrier is passed
        if (c > 0 && __syncthreads_count(done) == UP3D_TILE_PIX) break;
        StageBuf &sb = stage[c & 1];
        if (tma) {
            if (issued == c) { if (tid == 0) stage_issue<false>(sb, &bar[c & 1], a.st, rec0, n, c * UP3D_TILE_PIX); issued = c + 1; }
            // prefetch the next chunk only for tiles that already proved to need more than one
            if (c >= 1 && c + 1 < nchunks && issued == c + 1) {
                if (tid == 0) stage_issue<false>(stage[(c + 1) & 1], &bar[(c + 1) & 1], a.st, rec0, n, (c + 1) * UP3D_TILE_PIX);
                issued = c + 2;
            }
            mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
        } else {
            if (BINS && cand.list) stage_fill_indexed<false>(sb, a.st, rec0, cand.list, n, c * UP3D_TILE_PIX);
            else stage_fill_plain<false>(sb, a.st, rec0, n, c * UP3D_TILE_PIX);
            __syncthreads();
        }
        const int cnt = stage_compact<false>(sb, warp_cnt, n, c * UP3D_TILE_PIX, tx, ty);
        for (int j = 0; !done && j < cnt; ++j) {
            contributor++;
            const float2 xy = sb.xy[j];
            const float4 co = sb.co[j];
            const float dx = xy.x - pfx, dy = xy.y - pfy;
            const float power = gauss_power(co, dx, dy);
            if (power > 0.0f) continue;
            const float alpha = fminf(0.99f, co.w * expf(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            const float4 col = sb.rgb[j];
            const float wgt = alpha * T;
            C0 += col.x * wgt; C1 += col.y * wgt; C2 += col.z * wgt; ID += col.w * wgt;
            T = test_T;
            last_contributor = contributor;
        }
    }
    // a prefetched chunk may still be in flight: it must land before this CTA's shared memory is released
    if (tma && issued > c && c < nchunks) mbar_wait(&bar[c & 1], (uint32_t)((c >> 1) & 1));
    if (inside) {
        const size_t HW = (size_t)a.W * a.H, pix = (size_t)py * a.W + px;
        a.st.final_T[v * HW + pix] = T;
        a.st.n_contrib[v * HW + pix] = last_contributor;
        float *o = a.out_color + (size_t)v * 3 * HW + pix;
        o[0] = C0 + T * a.bg[0];
        o[HW] = C1 + T * a.bg[1];
        o[2 * HW] = C2 + T * a.bg[2];
        if (a.invdepth) a.invdepth[v * HW + pix] = ID;
    }
}

// ------------------------------------------------------------------------------------------------
// 4. blend backward (A.7)
// ------------------------------------------------------------------------------------------------
struct BlendBwdArgs {
    int W, H;
    const int32_t *view_rec_start;
    const float *bg;
    const float *dL_dcolor;  // V,3,H,W
    float *gacc;             // R * GACC_STRIDE : mean2D.xy, conic.xyz, opac, rgb.xyz
    State st;
};

// Reduction strategy.  The gradient of Gaussian j sums, over the 256 pixels of the tile, nine terms that all factor
// through two per-(pixel, entry) scalars  u = G * dL/dalpha  and  w = alpha * T :
//     dL/dopac' = S0,   dL/drgb_c = sum w * dLdC_c,
//     dL/dmean2D = -o (cx Sx + cy Sy, cz Sy + cy Sx),  dL/dconic = -o (Sxx/2, Sxy, Syy/2),
// with S0 = sum u, Sx = sum u dx, Sy = sum u dy, Sxx = sum u dx^2, Sxy = sum u dx dy, Syy = sum u dy^2 and
// (cx, cy, cz, o) the conic / opacity of the entry.  Phase 1 (one thread per pixel, sequential over a sub-batch of
// BWD_EB entries, back to front) only produces u and w into shared memory; phase 2 (16 threads per entry) forms the
// nine moment sums over the pixels with 4 shuffle levels, and one lane per entry issues the global atomics.  This
// replaces 45 shuffles + 9 shared atomics per (warp, entry) of the straightforward scheme.
// Sub-batch size BWD_EB: 16 entries (16 threads per entry in phase 2, 35 KB of u / w staging -> 4 CTAs per SM) or 8
// entries (32 threads per entry, 17 KB -> 6 CTAs per SM); selected at run time (UP3D_BWD_EB, default below).
constexpr int BWD_ROW = 256 + 16;     // padded row of the u / w staging arrays (conflict-free phase-2 reads)

template <int BWD_EB>
struct BwdSmem {
    StageBuf ch;                       // staged + in-place compacted chunk (single buffer: see kernel comment)
    uint64_t bar;
    int warp_cnt[8];
    float u[BWD_EB][BWD_ROW];
    float w[BWD_EB][BWD_ROW];
    float dL[3][UP3D_TILE_PIX];
    int red[8];
};

template <bool BINS, int BWD_EB>
__global__ void __launch_bounds__(UP3D_TILE_PIX, BWD_EB == 8 ? 5 : 4) blend_backward_kernel(const BlendBwdArgs a) {
    extern __shared__ __align__(16) unsigned char bwd_smem_raw[];
    BwdSmem<BWD_EB> &sm = *reinterpret_cast<BwdSmem<BWD_EB> *>(bwd_smem_raw);
    constexpr int TPE = UP3D_TILE_PIX / BWD_EB;       // phase-2 threads per entry (16 or 32)
    constexpr int PPT = UP3D_TILE_PIX / TPE;          // pixels per phase-2 thread
    StageBuf &ch = sm.ch;
    const int v = blockIdx.z, tx = blockIdx.x, ty = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int px = tx * UP3D_TILE + (tid & 15), py = ty * UP3D_TILE + (tid >> 4);
    const bool inside = px < a.W && py < a.H;
    const int rec0 = a.view_rec_start[v];
    Candidates cand{nullptr, 0};
    if (BINS) cand = tile_candidates(a.st, a.view_rec_start, v, tx, ty);
    else cand.n = a.st.n_vis[v];
    const int n = cand.n;
    const size_t HW = (size_t)a.W * a.H, pix = (size_t)py * a.W + px;
    const float pfx = (float)px, pfy = (float)py;

    const float T_final = inside ? a.st.final_T[v * HW + pix] : 0.f;
    const int last_contributor = inside ? a.st.n_contrib[v * HW + pix] : 0;
    const bool tma = cand.list == nullptr && (rec0 & 3) == 0;
    if (tma && tid == 0) { mbar_init(&sm.bar, 1); fence_mbar_init(); }
    int m = __reduce_max_sync(0xffffffffu, last_contributor);
    if (lane == 0) sm.red[warp] = m;
    __syncthreads();
    int Lmax = 0;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) Lmax = max(Lmax, sm.red[w8]);
    if (Lmax == 0) return;

    float dLp0 = 0.f, dLp1 = 0.f, dLp2 = 0.f;
    if (inside) {
        const float *g = a.dL_dcolor + (size_t)v * 3 * HW + pix;
        dLp0 = g[0]; dLp1 = g[HW]; dLp2 = g[2 * HW];
    }
    sm.dL[0][tid] = dLp0; sm.dL[1][tid] = dLp1; sm.dL[2][tid] = dLp2;
    const float bg_dot = a.bg[0] * dLp0 + a.bg[1] * dLp1 + a.bg[2] * dLp2;

    // pass 1: how many 256-record chunks hold the first Lmax list entries of this tile
    int c_last = 0, cum = 0;
    for (int base = 0; base < n; base += UP3D_TILE_PIX) {
        const int k = base + tid;
        const bool hit = (k < n) && rect_covers(a.st.s_rect[rec0 + ((BINS && cand.list) ? cand.list[k] : k)], tx, ty);
        cum += __syncthreads_count(hit);
        c_last = base / UP3D_TILE_PIX;
        if (cum >= Lmax) break;
    }
    int running_end = cum;  // list entries in chunks [0, c_last]

    float T = T_final;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;     // accum_rec
    float lc0 = 0.f, lc1 = 0.f, lc2 = 0.f;        // last_color
    float last_alpha = 0.f;
    // phase-2 role of this thread: entry slot e2 of the sub-batch, PPT pixels {part + TPE i}
    const int e2 = tid / TPE, part = tid % TPE;

    // Chunks are visited back to front.  In the reference's regime the first Lmax entries live in chunk 0, so a second
    // staging buffer would only cost occupancy (4 -> 3 CTAs/SM); one buffer, refilled by TMA after the barrier that
    // ends the previous chunk.
    uint32_t loads = 0;
    for (int c = c_last; c >= 0; --c) {
        if (tma) {
            if (tid == 0) stage_issue<true>(ch, &sm.bar, a.st, rec0, n, c * UP3D_TILE_PIX);
            mbar_wait(&sm.bar, loads & 1u);
            ++loads;
        } else {
            if (BINS && cand.list) stage_fill_indexed<true>(ch, a.st, rec0, cand.list, n, c * UP3D_TILE_PIX);
            else stage_fill_plain<true>(ch, a.st, rec0, n, c * UP3D_TILE_PIX);
            __syncthreads();
        }
        const int cnt = stage_compact<true>(ch, sm.warp_cnt, n, c * UP3D_TILE_PIX, tx, ty);
        const int pos_base = running_end - cnt;  // 0-based list position of entry 0 of this chunk
        running_end = pos_base;
        int jtop = min(cnt - 1, Lmax - 1 - pos_base);   // entries beyond the tile's max n_contrib never contribute
        for (; jtop >= 0; jtop -= BWD_EB) {
            // ---- phase 1: per pixel, entries jtop, jtop-1, ... (back to front)
#pragma unroll 4
            for (int e = 0; e < BWD_EB; ++e) {
                const int j = jtop - e;
                float u = 0.f, wgt = 0.f;
                if (j >= 0 && pos_base + j < last_contributor) {
                    const float2 xy = ch.xy[j];
                    const float4 co = ch.co[j];
                    const float dx = xy.x - pfx, dy = xy.y - pfy;
                    const float power = gauss_power(co, dx, dy);
                    if (!(power > 0.0f)) {
                        const float G = expf(power);
                        const float alpha = fminf(0.99f, co.w * G);
                        if (!(alpha < 1.0f / 255.0f)) {
                            T = T / (1.f - alpha);
                            wgt = alpha * T;
                            const float4 col = ch.rgb[j];
                            acc0 = last_alpha * lc0 + (1.f - last_alpha) * acc0;
                            acc1 = last_alpha * lc1 + (1.f - last_alpha) * acc1;
                            acc2 = last_alpha * lc2 + (1.f - last_alpha) * acc2;
                            lc0 = col.x; lc1 = col.y; lc2 = col.z;
                            float dL_dalpha = (col.x - acc0) * dLp0 + (col.y - acc1) * dLp1 + (col.z - acc2) * dLp2;
                            dL_dalpha *= T;
                            last_alpha = alpha;
                            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                            u = G * dL_dalpha;
                        }
                    }
                }
                sm.u[e][tid] = u;
                sm.w[e][tid] = wgt;
            }
            __syncthreads();
            // ---- phase 2: 16 threads per entry reduce the nine moment sums over the 256 pixels
            {
                const int j = jtop - e2;
                float S0 = 0.f, Sx = 0.f, Sy = 0.f, Sxx = 0.f, Sxy = 0.f, Syy = 0.f, R0 = 0.f, R1 = 0.f, R2 = 0.f;
                if (j >= 0) {
                    const float2 xy = ch.xy[j];
                    const float bx = xy.x - (float)(tx * UP3D_TILE + (part & 15));
                    const float by0 = xy.y - (float)(ty * UP3D_TILE + (part >> 4));
#pragma unroll
                    for (int i = 0; i < PPT; ++i) {
                        const int p = part + TPE * i;        // pixel (column = part & 15, row = (part >> 4) + (TPE / 16) i)
                        const float u = sm.u[e2][p], wv = sm.w[e2][p];
                        const float dx = bx, dy = by0 - (float)((TPE / 16) * i);
                        const float ux = u * dx, uy = u * dy;
                        S0 += u; Sx += ux; Sy += uy;
                        Sxx = fmaf(ux, dx, Sxx); Sxy = fmaf(ux, dy, Sxy); Syy = fmaf(uy, dy, Syy);
                        R0 = fmaf(wv, sm.dL[0][p], R0); R1 = fmaf(wv, sm.dL[1][p], R1); R2 = fmaf(wv, sm.dL[2][p], R2);
                    }
                }
#pragma unroll
                for (int o = TPE / 2; o > 0; o >>= 1) {
                    S0 += __shfl_xor_sync(0xffffffffu, S0, o); Sx += __shfl_xor_sync(0xffffffffu, Sx, o);
                    Sy += __shfl_xor_sync(0xffffffffu, Sy, o); Sxx += __shfl_xor_sync(0xffffffffu, Sxx, o);
                    Sxy += __shfl_xor_sync(0xffffffffu, Sxy, o); Syy += __shfl_xor_sync(0xffffffffu, Syy, o);
                    R0 += __shfl_xor_sync(0xffffffffu, R0, o); R1 += __shfl_xor_sync(0xffffffffu, R1, o);
                    R2 += __shfl_xor_sync(0xffffffffu, R2, o);
                }
                if (part == 0 && j >= 0) {
                    const float4 co = ch.co[j];
                    float *g = a.gacc + (size_t)(rec0 + ch.id[j]) * GACC_STRIDE;
                    const float o_ = -co.w;
                    const float v0 = o_ * (co.x * Sx + co.y * Sy), v1 = o_ * (co.z * Sy + co.y * Sx);
                    const float v2 = 0.5f * o_ * Sxx, v3 = o_ * Sxy, v4 = 0.5f * o_ * Syy;
                    if (v0 != 0.f) atomicAdd(g + 0, v0);
                    if (v1 != 0.f) atomicAdd(g + 1, v1);
                    if (v2 != 0.f) atomicAdd(g + 2, v2);
                    if (v3 != 0.f) atomicAdd(g + 3, v3);
                    if (v4 != 0.f) atomicAdd(g + 4, v4);
                    if (S0 != 0.f) atomicAdd(g + 5, S0);
                    if (R0 != 0.f) atomicAdd(g + 6, R0);
                    if (R1 != 0.f) atomicAdd(g + 7, R1);
                    if (R2 != 0.f) atomicAdd(g + 8, R2);
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 5. geometry backward (A.8 + A.9): one thread per Gaussian, fixed-order loop over the set's views
// ------------------------------------------------------------------------------------------------
struct GeomBwdArgs {
    ViewConst vc;
    const float *means3D, *shs, *colors, *opacities, *scales, *rotations, *viewmats, *projmats, *campos;
    const int32_t *set_offsets, *set_view_start, *view_rec_start;
    const float *gacc;
    State st;
    float *dmeans3D, *dmeans2D, *dshs, *dcolors, *dopac, *dscales, *drot;
};

template <int MAXM>
__global__ void __launch_bounds__(128) geometry_backward_kernel(const GeomBwdArgs a) {
    const int set = blockIdx.y;
    const int g0 = a.set_offsets[set], P = a.set_offsets[set + 1] - g0;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int gi = g0 + i;
    const ViewConst &vc = a.vc;
    const float px_ = a.means3D[3 * gi], py_ = a.means3D[3 * gi + 1], pz_ = a.means3D[3 * gi + 2];
    const float s3[3] = {a.scales[3 * gi], a.scales[3 * gi + 1], a.scales[3 * gi + 2]};
    const float q[4] = {a.rotations[4 * gi], a.rotations[4 * gi + 1], a.rotations[4 * gi + 2], a.rotations[4 * gi + 3]};
    const float opac = a.opacities[gi];
    Cov3 c3;
    cov3d_from_scale_rot(s3, vc.scale_modifier, q, c3);
    const float V9[9] = {c3.c[0], c3.c[1], c3.c[2], c3.c[1], c3.c[3], c3.c[4], c3.c[2], c3.c[4], c3.c[5]};

    float dmean[3] = {0.f, 0.f, 0.f}, dS[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dop = 0.f, dcol[3] = {0.f, 0.f, 0.f};
    float dsh[MAXM * 3];
#pragma unroll
    for (int k = 0; k < MAXM * 3; ++k) dsh[k] = 0.f;

    for (int v = a.set_view_start[set]; v < a.set_view_start[set + 1]; ++v) {
        const int rec = a.view_rec_start[v] + i;
        if (a.dmeans2D) { a.dmeans2D[3 * rec] = 0.f; a.dmeans2D[3 * rec + 1] = 0.f; a.dmeans2D[3 * rec + 2] = 0.f; }
        if (a.st.radii[rec] <= 0) continue;
        const float *view = a.viewmats + 16 * v, *proj = a.projmats + 16 * v;
        const float4 ga = *(const float4 *)(a.gacc + (size_t)rec * GACC_STRIDE);
        const float4 gb = *(const float4 *)(a.gacc + (size_t)rec * GACC_STRIDE + 4);
        const float gr2 = a.gacc[(size_t)rec * GACC_STRIDE + 8];
        const float G_mx = ga.x, G_my = ga.y, Gx = ga.z, Gy = ga.w, Gz = gb.x, G_op = gb.y;
        const float G_rgb[3] = {gb.z, gb.w, gr2};

        float t[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) t[r] = tp_row(view, r, px_, py_, pz_);
        Cov2 c2;
        cov2d_ewa(t, vc, c3.c, view, c2);
        const float h = 0.3f;
        const float x = c2.cov[0], z = c2.cov[1], y = c2.cov[2];
        const float ca = x + h, cb = z, cc = y + h;
        const float det0 = x * y - z * z, det = ca * cc - cb * cb;
        float hs = 1.f, dLdr = 0.f;
        if (vc.antialiasing) {
            const float r = det0 / det;
            hs = sqrtf(fmaxf(0.000025f, r));
            const float dL_dhs = G_op * opac;
            dLdr = (r <= 0.000025f) ? 0.f : dL_dhs / (2.f * hs);
        }
        dop += G_op * hs;
        const float id2 = 1.f / (det * det);
        float dL_da = (-cc * cc * Gx + cb * cc * Gy - cb * cb * Gz) * id2;
        float dL_db = (2.f * cb * cc * Gx - (det + 2.f * cb * cb) * Gy + 2.f * ca * cb * Gz) * id2;
        float dL_dc = (-cb * cb * Gx + ca * cb * Gy - ca * ca * Gz) * id2;
        if (dLdr != 0.f) {
            dL_da += dLdr * (y * det - det0 * cc) * id2;
            dL_dc += dLdr * (x * det - det0 * ca) * id2;
            dL_db += dLdr * (-2.f * z * (det - det0)) * id2;
        }
        const float *T = c2.T;
        const float T00 = T[0], T01 = T[1], T02 = T[2], T10 = T[3], T11 = T[4], T12 = T[5];
        dS[0] += T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
        dS[3] += T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
        dS[5] += T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
        dS[1] += 2.f * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2.f * T10 * T11 * dL_dc;
        dS[2] += 2.f * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2.f * T10 * T12 * dL_dc;
        dS[4] += 2.f * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2.f * T11 * T12 * dL_dc;
        float dT[6];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float VT0 = V9[k * 3 + 0] * T00 + V9[k * 3 + 1] * T01 + V9[k * 3 + 2] * T02;
            const float VT1 = V9[k * 3 + 0] * T10 + V9[k * 3 + 1] * T11 + V9[k * 3 + 2] * T12;
            dT[k] = 2.f * dL_da * VT0 + dL_db * VT1;
            dT[3 + k] = 2.f * dL_dc * VT1 + dL_db * VT0;
        }
        const float dJ00 = dT[0] * view[0] + dT[1] * view[4] + dT[2] * view[8];
        const float dJ02 = dT[0] * view[2] + dT[1] * view[6] + dT[2] * view[10];
        const float dJ11 = dT[3] * view[1] + dT[4] * view[5] + dT[5] * view[9];
        const float dJ12 = dT[3] * view[2] + dT[4] * view[6] + dT[5] * view[10];
        const float tzi = 1.f / c2.tc[2], tz2 = tzi * tzi, tz3 = tz2 * tzi;
        const float dtx = (c2.cx ? 0.f : 1.f) * -vc.fx * tz2 * dJ02;
        const float dty = (c2.cy ? 0.f : 1.f) * -vc.fy * tz2 * dJ12;
        const float dtz = -vc.fx * tz2 * dJ00 - vc.fy * tz2 * dJ11 + (2.f * vc.fx * c2.tc[0]) * tz3 * dJ02 +
                          (2.f * vc.fy * c2.tc[1]) * tz3 * dJ12;
#pragma unroll
        for (int c = 0; c < 3; ++c) dmean[c] += view[4 * c + 0] * dtx + view[4 * c + 1] * dty + view[4 * c + 2] * dtz;

        // mean2D -> mean3D through the projection
        {
            float ph[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) ph[r] = tp_row(proj, r, px_, py_, pz_);
            const float m_w = 1.0f / (ph[3] + 0.0000001f);
            const float gxn = G_mx * 0.5f * vc.W, gyn = G_my * 0.5f * vc.H;
            if (a.dmeans2D) { a.dmeans2D[3 * rec] = gxn; a.dmeans2D[3 * rec + 1] = gyn; }
            const float mul1 = ph[0] * m_w * m_w, mul2 = ph[1] * m_w * m_w;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                dmean[k] += (proj[4 * k + 0] * m_w - proj[4 * k + 3] * mul1) * gxn +
                            (proj[4 * k + 1] * m_w - proj[4 * k + 3] * mul2) * gyn;
        }
        // colour
        if (a.colors) {
            dcol[0] += G_rgb[0]; dcol[1] += G_rgb[1]; dcol[2] += G_rgb[2];
        } else {
            const int deg = vc.D;
            const float *sh = a.shs + (size_t)gi * vc.M * 3;
            const float *cp = a.campos + 3 * v;
            const float d0x = px_ - cp[0], d0y = py_ - cp[1], d0z = pz_ - cp[2];
            const float len2 = d0x * d0x + d0y * d0y + d0z * d0z;
            const float len = sqrtf(len2);
            const float X = d0x / len, Y = d0y / len, Z = d0z / len;
            const uint8_t cl = a.st.clamped[rec];
            const float dRGB[3] = {(cl & 1) ? 0.f : G_rgb[0], (cl & 2) ? 0.f : G_rgb[1], (cl & 4) ? 0.f : G_rgb[2]};
            float w[MAXM];
            sh_basis<MAXM>(deg, X, Y, Z, w);
            const int nb = (deg + 1) * (deg + 1);
#pragma unroll
            for (int l = 0; l < MAXM; ++l)
                if (l < nb) {
                    dsh[l * 3 + 0] += w[l] * dRGB[0]; dsh[l * 3 + 1] += w[l] * dRGB[1]; dsh[l * 3 + 2] += w[l] * dRGB[2];
                }
            float ddx = 0.f, ddy = 0.f, ddz = 0.f;
#define DOTRGB(l) (dRGB[0] * sh[(l)*3] + dRGB[1] * sh[(l)*3 + 1] + dRGB[2] * sh[(l)*3 + 2])
            if (MAXM >= 4 && deg > 0) {
                ddx += -SH_C1 * DOTRGB(3); ddy += -SH_C1 * DOTRGB(1); ddz += SH_C1 * DOTRGB(2);
                if (MAXM >= 9 && deg > 1) {
                    const float xx = X * X, yy = Y * Y, zz = Z * Z, xy = X * Y, yz = Y * Z, xz = X * Z;
                    const float s4 = DOTRGB(4), s5 = DOTRGB(5), s6 = DOTRGB(6), s7 = DOTRGB(7), s8 = DOTRGB(8);
                    ddx += c_SH_C2[0] * Y * s4 + c_SH_C2[2] * 2.f * -X * s6 + c_SH_C2[3] * Z * s7 + c_SH_C2[4] * 2.f * X * s8;
                    ddy += c_SH_C2[0] * X * s4 + c_SH_C2[1] * Z * s5 + c_SH_C2[2] * 2.f * -Y * s6 + c_SH_C2[4] * 2.f * -Y * s8;
                    ddz += c_SH_C2[1] * Y * s5 + c_SH_C2[2] * 4.f * Z * s6 + c_SH_C2[3] * X * s7;
                    if (MAXM >= 16 && deg > 2) {
                        const float s9 = DOTRGB(9), s10 = DOTRGB(10), s11 = DOTRGB(11), s12 = DOTRGB(12), s13 = DOTRGB(13),
                                    s14 = DOTRGB(14), s15 = DOTRGB(15);
                        ddx += c_SH_C3[0] * s9 * 6.f * xy + c_SH_C3[1] * s10 * yz + c_SH_C3[2] * s11 * -2.f * xy +
                               c_SH_C3[3] * s12 * -6.f * xz + c_SH_C3[4] * s13 * (-3.f * xx + 4.f * zz - yy) +
                               c_SH_C3[5] * s14 * 2.f * xz + c_SH_C3[6] * s15 * 3.f * (xx - yy);
                        ddy += c_SH_C3[0] * s9 * 3.f * (xx - yy) + c_SH_C3[1] * s10 * xz +
                               c_SH_C3[2] * s11 * (-3.f * yy + 4.f * zz - xx) + c_SH_C3[3] * s12 * -6.f * yz +
                               c_SH_C3[4] * s13 * -2.f * xy + c_SH_C3[5] * s14 * -2.f * yz + c_SH_C3[6] * s15 * -6.f * xy;
                        ddz += c_SH_C3[1] * s10 * xy + c_SH_C3[2] * s11 * 8.f * yz + c_SH_C3[3] * s12 * 3.f * (2.f * zz - xx - yy) +
                               c_SH_C3[4] * s13 * 8.f * xz + c_SH_C3[5] * s14 * (xx - yy);
                    }
                }
            }
#undef DOTRGB
            const float dotv = d0x * ddx + d0y * ddy + d0z * ddz;
            const float inv3 = 1.f / (len2 * len);
            dmean[0] += (ddx * len2 - d0x * dotv) * inv3;
            dmean[1] += (ddy * len2 - d0y * dotv) * inv3;
            dmean[2] += (ddz * len2 - d0z * dotv) * inv3;
        }
    }
    // cov3D -> scale / rotation (once, from the summed dL/dSigma)
    {
        const float Gs[9] = {dS[0], 0.5f * dS[1], 0.5f * dS[2], 0.5f * dS[1], dS[3], 0.5f * dS[4], 0.5f * dS[2], 0.5f * dS[4], dS[5]};
        float dR[9];
        float dsc[3];
#pragma unroll
        for (int ii = 0; ii < 3; ++ii) {
            float ds = 0.f;
#pragma unroll
            for (int aa = 0; aa < 3; ++aa) {
                float dA = 0.f;
#pragma unroll
                for (int bb = 0; bb < 3; ++bb) dA += 2.f * Gs[aa * 3 + bb] * (c3.s[ii] * c3.R[bb * 3 + ii]);
                ds += dA * c3.R[aa * 3 + ii];
                dR[aa * 3 + ii] = dA * c3.s[ii];
            }
            dsc[ii] = ds * vc.scale_modifier;
        }
        const float qr = q[0], qx = q[1], qy = q[2], qz = q[3];
        a.drot[4 * gi + 0] = 2.f * (-qz * dR[1] + qy * dR[2] + qz * dR[3] - qx * dR[5] - qy * dR[6] + qx * dR[7]);
        a.drot[4 * gi + 1] = 2.f * (qy * dR[1] + qz * dR[2] + qy * dR[3] - 2.f * qx * dR[4] - qr * dR[5] + qz * dR[6] + qr * dR[7] - 2.f * qx * dR[8]);
        a.drot[4 * gi + 2] = 2.f * (-2.f * qy * dR[0] + qx * dR[1] + qr * dR[2] + qx * dR[3] + qz * dR[5] - qr * dR[6] + qz * dR[7] - 2.f * qy * dR[8]);
        a.drot[4 * gi + 3] = 2.f * (-2.f * qz * dR[0] - qr * dR[1] + qx * dR[2] + qr * dR[3] - 2.f * qz * dR[4] + qy * dR[5] + qx * dR[6] + qy * dR[7]);
        a.dscales[3 * gi] = dsc[0]; a.dscales[3 * gi + 1] = dsc[1]; a.dscales[3 * gi + 2] = dsc[2];
    }
    a.dmeans3D[3 * gi] = dmean[0]; a.dmeans3D[3 * gi + 1] = dmean[1]; a.dmeans3D[3 * gi + 2] = dmean[2];
    a.dopac[gi] = dop;
    if (a.colors) {
        if (a.dcolors) { a.dcolors[3 * gi] = dcol[0]; a.dcolors[3 * gi + 1] = dcol[1]; a.dcolors[3 * gi + 2] = dcol[2]; }
    } else if (a.dshs) {
        float *o = a.dshs + (size_t)gi * vc.M * 3;
        for (int k = 0; k < vc.M * 3; ++k) o[k] = (k < MAXM * 3) ? dsh[k] : 0.f;
    }
}

// ------------------------------------------------------------------------------------------------
// debug kernels
// ------------------------------------------------------------------------------------------------
struct DebugListArgs {
    int gx, gy;
    const int32_t *view_rec_start;
    State st;
    int32_t *tile_counts;
    const int32_t *tile_offsets;
    int32_t *tile_lists;
};
__global__ void __launch_bounds__(UP3D_TILE_PIX) debug_tile_lists_kernel(const DebugListArgs a) {
    __shared__ int warp_cnt[8];
    const int v = blockIdx.z, tx = blockIdx.x, ty = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rec0 = a.view_rec_start[v], n = a.st.n_vis[v];
    const int tile = v * a.gx * a.gy + ty * a.gx + tx;
    int run = 0;
    for (int base = 0; base < n; base += UP3D_TILE_PIX) {
        const int k = base + tid;
        const bool hit = (k < n) && rect_covers(a.st.s_rect[rec0 + k], tx, ty);
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) warp_cnt[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
        for (int w = 0; w < 8; ++w) { off += (w < warp) ? warp_cnt[w] : 0; total += warp_cnt[w]; }
        if (hit && a.tile_lists) a.tile_lists[a.tile_offsets[tile] + run + off + __popc(bal & lanemask_lt())] = a.st.s_id[rec0 + k];
        run += total;
        __syncthreads();
    }
    if (tid == 0) a.tile_counts[tile] = run;
}

__global__ void debug_unpack_kernel(int R, State st, float *depths, float *xy, float *co, float *rgb, int32_t *rects) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R) return;
    const bool vis = st.key[r] != KEY_CULLED;
    if (depths) depths[r] = vis ? __uint_as_float(st.key[r]) : 0.f;
    if (xy) { xy[2 * r] = st.u_xy[r].x; xy[2 * r + 1] = st.u_xy[r].y; }
    if (co) { const float4 c = st.u_co[r]; co[4 * r] = c.x; co[4 * r + 1] = c.y; co[4 * r + 2] = c.z; co[4 * r + 3] = c.w; }
    if (rgb) { const float4 c = st.u_rgb[r]; rgb[3 * r] = c.x; rgb[3 * r + 1] = c.y; rgb[3 * r + 2] = c.z; }
    if (rects) {
        const uint32_t q = st.u_rect[r];
        rects[4 * r] = q & 0xFF; rects[4 * r + 1] = (q >> 8) & 0xFF; rects[4 * r + 2] = (q >> 16) & 0xFF; rects[4 * r + 3] = q >> 24;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int validate_desc(const up3d_raster_desc *d) {
    UP3D_CHECK_ARG(d != nullptr, "up3d_raster: null descriptor");
    UP3D_CHECK_ARG(d->n_sets >= 0 && d->n_views >= 0 && d->n_gaussians >= 0 && d->n_records >= 0,
                   "up3d_raster: negative sizes in descriptor");
    UP3D_CHECK_ARG(d->width > 0 && d->height > 0, "up3d_raster: image size must be positive (got %dx%d)", d->width, d->height);
    UP3D_CHECK_ARG(div_up(d->width, UP3D_TILE) <= 255 && div_up(d->height, UP3D_TILE) <= 255,
                   "up3d_raster: image larger than 4080x4080 is not supported (tile rect is packed in 8 bits)");
    UP3D_CHECK_ARG(d->sh_degree >= 0 && d->sh_degree <= 3, "up3d_raster: sh_degree must be 0..3 (got %d)", d->sh_degree);
    UP3D_CHECK_ARG(d->sh_coeffs == 0 || d->sh_coeffs >= (d->sh_degree + 1) * (d->sh_degree + 1),
                   "up3d_raster: shs has %d coefficients, degree %d needs %d", d->sh_coeffs, d->sh_degree,
                   (d->sh_degree + 1) * (d->sh_degree + 1));
    UP3D_CHECK_ARG(d->sh_coeffs <= 16, "up3d_raster: at most 16 SH coefficients supported (got %d)", d->sh_coeffs);
    UP3D_CHECK_ARG(d->tanfovx > 0.f && d->tanfovy > 0.f, "up3d_raster: tanfov must be positive");
    if (d->n_views > 0)
        UP3D_CHECK_ARG(d->set_offsets && d->set_view_start && d->view_set && d->view_rec_start,
                       "up3d_raster: descriptor index arrays must not be null");
    return 0;
}

// Optional per-kernel timing with CUDA events recorded on the launching stream (bench.py's roofline leg).
// Slots: 0 project, 1 depth_sort, 2 blend_forward, 3 gacc clear, 4 blend_backward, 5 geometry_backward.
// A DIAGNOSTIC facility, process-wide by design (autograd runs the backward on its own host thread, so the state cannot
// be thread-local): while it is enabled, only ONE stream may issue raster calls; flags are atomics so that concurrent
// untimed callers on other streams are merely not measured.  Every kernel sits between its own two event records with
// nothing but its launch in between (function attributes are configured once, ahead of time).
struct Timing {
    std::atomic<bool> enabled{false}, fwd_valid{false}, bwd_valid{false};
    bool created = false;
    cudaEvent_t ev[8];
};
static Timing g_timing;
static inline void tick(int i, cudaStream_t s) {
    if (g_timing.enabled.load(std::memory_order_relaxed)) cudaEventRecord(g_timing.ev[i], s);
}

// cudaFuncSetAttribute once per growth of the requirement (not on every call; idempotent under races)
static int ensure_dyn_smem(const void *fn, std::atomic<size_t> &configured, size_t bytes) {
    if (bytes <= configured.load(std::memory_order_acquire)) return 0;
    UP3D_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    size_t cur = configured.load(std::memory_order_relaxed);
    while (cur < bytes && !configured.compare_exchange_weak(cur, bytes, std::memory_order_release)) {}
    return 0;
}
static std::atomic<size_t> g_sort_smem{0}, g_bwd_smem[4];
constexpr int BWD_EB_DEFAULT = 16;

static ViewConst make_view_const(const up3d_raster_desc *d) {
    ViewConst vc;
    vc.W = d->width; vc.H = d->height;
    vc.gx = div_up(d->width, UP3D_TILE); vc.gy = div_up(d->height, UP3D_TILE);
    vc.tanfovx = d->tanfovx; vc.tanfovy = d->tanfovy;
    volatile float fx = (float)d->width / (2.0f * d->tanfovx), fy = (float)d->height / (2.0f * d->tanfovy);
    volatile float lx = 1.3f * d->tanfovx, ly = 1.3f * d->tanfovy;
    vc.fx = fx; vc.fy = fy; vc.limx = lx; vc.limy = ly;
    vc.scale_modifier = d->scale_modifier;
    vc.D = d->sh_degree; vc.M = d->sh_coeffs; vc.antialiasing = d->antialiasing;
    return vc;
}

}  // namespace up3d

using namespace up3d;

extern "C" {

const char *up3d_last_error(void) { return up3d::err_buf(); }
int up3d_version(void) { return 100; }

size_t up3d_raster_state_bytes(const up3d_raster_desc *d) {
    if (!d) return 0;
    return carve_state(d, nullptr).total;
}
size_t up3d_raster_scratch_bytes(const up3d_raster_desc *d) {
    if (!d) return 0;
    return carve_scratch(d, nullptr).total;
}

int up3d_raster_forward(const up3d_raster_desc *d, const float *means3D, const float *shs, const float *colors_precomp,
                        const float *opacities, const float *scales, const float *rotations, const float *viewmats,
                        const float *projmats, const float *campos, const float *bg, float *out_color, int32_t *radii,
                        float *invdepth, void *state, void *scratch, up3d_stream_t stream_) {
    if (validate_desc(d)) return 1;
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG((shs != nullptr) != (colors_precomp != nullptr),
                   "Please provide excatly one of either SHs or precomputed colors!");
    UP3D_CHECK_ARG(shs == nullptr || d->sh_coeffs > 0, "up3d_raster_forward: shs given but sh_coeffs == 0");
    UP3D_CHECK_ARG(out_color && radii && state && scratch && bg, "up3d_raster_forward: null output/state/scratch/bg pointer");
    if (d->n_views == 0) return 0;
    UP3D_CHECK_ARG(viewmats && projmats && campos, "up3d_raster_forward: null camera pointer");
    UP3D_CHECK_ARG(d->n_records == 0 || (means3D && opacities && scales && rotations), "up3d_raster_forward: null Gaussian pointer");
    State st = carve_state(d, state);
    Scratch sc = carve_scratch(d, scratch);
    const ViewConst vc = make_view_const(d);
    const int V = d->n_views;
    tick(0, stream);
    if (d->max_set_size > 0) {
        ProjectArgs pa{vc, means3D, shs, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos,
                       d->set_offsets, d->view_set, d->view_rec_start, radii, st};
        project_kernel<<<dim3(div_up(d->max_set_size, 256), V), 256, 0, stream>>>(pa);
        UP3D_LAUNCH_OK("project_kernel");
    }
    SortArgs sa{d->view_rec_start, st, sc, d->max_set_size <= SORT_SMEM_MAX_KEYS ? 1 : 0};
    size_t sort_smem = SORT_FIXED_SMEM;
    if (sa.use_smem) sort_smem += (size_t)((d->max_set_size + 3) & ~3) * 16;
    if (ensure_dyn_smem((const void *)depth_sort_kernel, g_sort_smem, sort_smem)) return 1;
    tick(1, stream);
    if (!sa.use_smem) {
        // large sets: a cluster of SORT_CLUSTER CTAs per view (keys ping-pong in the global scratch)
        depth_sort_cluster_kernel<<<V * SORT_CLUSTER, SORT_THREADS, 0, stream>>>(sa);
        UP3D_LAUNCH_OK("depth_sort_cluster_kernel");
    } else {
        const size_t smem = sort_smem;
        depth_sort_kernel<<<V, SORT_THREADS, smem, stream>>>(sa);
        UP3D_LAUNCH_OK("depth_sort_kernel");
    }
    // bin decision + lists (large, many-tile problems only; otherwise every view streams its records)
    UP3D_CUDA_OK(cudaMemsetAsync(st.bin_mode, 0, (size_t)((char *)st.bin_count - (char *)st.bin_mode), stream));   // mode + total
    if (bins_enabled(d)) {
        const size_t nb = (size_t)st.nbins_x * st.nbins_y;
        UP3D_CUDA_OK(cudaMemsetAsync(st.bin_count, 0, sizeof(int32_t) * V * nb, stream));
        BinArgs ba{d->view_rec_start, st};
        const dim3 grid_r(div_up(d->max_set_size, 256), V);
        bin_total_kernel<<<grid_r, 256, 0, stream>>>(ba);
        bin_count_kernel<<<grid_r, 256, 0, stream>>>(ba);
        bin_fill_kernel<<<dim3((unsigned)nb, V), 256, 0, stream>>>(ba);
        UP3D_LAUNCH_OK("bin kernels");
    }
    tick(2, stream);
    {
        BlendArgs ba{d->width, d->height, d->view_rec_start, bg, out_color, invdepth, st};
        if (bins_enabled(d)) blend_forward_kernel<true><<<dim3(vc.gx, vc.gy, V), UP3D_TILE_PIX, 0, stream>>>(ba);
        else blend_forward_kernel<false><<<dim3(vc.gx, vc.gy, V), UP3D_TILE_PIX, 0, stream>>>(ba);
        UP3D_LAUNCH_OK("blend_forward_kernel");
    }
    tick(3, stream);
    g_timing.fwd_valid = g_timing.enabled.load();
    return 0;
}

int up3d_raster_backward(const up3d_raster_desc *d, const float *means3D, const float *shs, const float *colors_precomp,
                         const float *opacities, const float *scales, const float *rotations, const float *viewmats,
                         const float *projmats, const float *campos, const float *bg, const float *dL_dcolor,
                         const void *state, void *scratch, float *dL_dmeans3D, float *dL_dmeans2D, float *dL_dshs,
                         float *dL_dcolors, float *dL_dopacities, float *dL_dscales, float *dL_drotations,
                         up3d_stream_t stream_) {
    if (validate_desc(d)) return 1;
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG((shs != nullptr) != (colors_precomp != nullptr),
                   "Please provide excatly one of either SHs or precomputed colors!");
    UP3D_CHECK_ARG(state && scratch && bg && dL_dcolor, "up3d_raster_backward: null state/scratch/bg/dL_dcolor pointer");
    UP3D_CHECK_ARG(dL_dmeans3D && dL_dopacities && dL_dscales && dL_drotations, "up3d_raster_backward: null gradient output");
    UP3D_CHECK_ARG(shs == nullptr || dL_dshs != nullptr, "up3d_raster_backward: dL_dshs required when shs are used");
    if (d->n_gaussians == 0) return 0;
    State st = carve_state(d, const_cast<void *>(state));
    Scratch sc = carve_scratch(d, scratch);
    const ViewConst vc = make_view_const(d);
    const int V = d->n_views;
    const bool bins = bins_enabled(d);
    static const int eb = [] { const char *e = getenv("UP3D_BWD_EB"); return (e && atoi(e) == 16) ? 16 : (e && atoi(e) == 8) ? 8 : BWD_EB_DEFAULT; }();
    const int variant = (bins ? 2 : 0) + (eb == 8 ? 1 : 0);
    const void *bwd_fn[4] = {(const void *)blend_backward_kernel<false, 16>, (const void *)blend_backward_kernel<false, 8>,
                             (const void *)blend_backward_kernel<true, 16>, (const void *)blend_backward_kernel<true, 8>};
    const size_t bwd_smem = eb == 8 ? sizeof(BwdSmem<8>) : sizeof(BwdSmem<16>);
    if (ensure_dyn_smem(bwd_fn[variant], g_bwd_smem[variant], bwd_smem)) return 1;
    tick(4, stream);
    UP3D_CUDA_OK(cudaMemsetAsync(sc.gacc, 0, sizeof(float) * GACC_STRIDE * (size_t)d->n_records, stream));
    tick(5, stream);
    if (V > 0) {
        BlendBwdArgs ba{d->width, d->height, d->view_rec_start, bg, dL_dcolor, sc.gacc, st};
        const dim3 grid(vc.gx, vc.gy, V);
        switch (variant) {
            case 0: blend_backward_kernel<false, 16><<<grid, UP3D_TILE_PIX, bwd_smem, stream>>>(ba); break;
            case 1: blend_backward_kernel<false, 8><<<grid, UP3D_TILE_PIX, bwd_smem, stream>>>(ba); break;
            case 2: blend_backward_kernel<true, 16><<<grid, UP3D_TILE_PIX, bwd_smem, stream>>>(ba); break;
            default: blend_backward_kernel<true, 8><<<grid, UP3D_TILE_PIX, bwd_smem, stream>>>(ba); break;
        }
        UP3D_LAUNCH_OK("blend_backward_kernel");
    }
    tick(6, stream);
    GeomBwdArgs ga{vc, means3D, shs, colors_precomp, opacities, scales, rotations, viewmats, projmats, campos,
                   d->set_offsets, d->set_view_start, d->view_rec_start, sc.gacc, st,
                   dL_dmeans3D, dL_dmeans2D, dL_dshs, dL_dcolors, dL_dopacities, dL_dscales, dL_drotations};
    const dim3 grid(div_up(d->max_set_size, 128), d->n_sets);
    if (d->max_set_size > 0 && d->n_sets > 0) {
        if (d->sh_coeffs <= 4 || d->sh_degree <= 1) geometry_backward_kernel<4><<<grid, 128, 0, stream>>>(ga);
        else geometry_backward_kernel<16><<<grid, 128, 0, stream>>>(ga);
        UP3D_LAUNCH_OK("geometry_backward_kernel");
    }
    tick(7, stream);
    g_timing.bwd_valid = g_timing.enabled.load();
    return 0;
}

int up3d_raster_timing_enable(int enable) {
    if (enable && !g_timing.created) {
        for (int i = 0; i < 8; ++i) UP3D_CUDA_OK(cudaEventCreate(&g_timing.ev[i]));
        g_timing.created = true;
    }
    g_timing.enabled = enable != 0;
    g_timing.fwd_valid = false;
    g_timing.bwd_valid = false;
    return 0;
}

int up3d_raster_timing_read(float *ms6) {
    UP3D_CHECK_ARG(ms6 != nullptr, "up3d_raster_timing_read: null output");
    for (int i = 0; i < 6; ++i) ms6[i] = -1.f;
    if (!g_timing.created) return 0;
    if (g_timing.fwd_valid) {
        UP3D_CUDA_OK(cudaEventSynchronize(g_timing.ev[3]));
        for (int i = 0; i < 3; ++i) UP3D_CUDA_OK(cudaEventElapsedTime(&ms6[i], g_timing.ev[i], g_timing.ev[i + 1]));
    }
    if (g_timing.bwd_valid) {
        UP3D_CUDA_OK(cudaEventSynchronize(g_timing.ev[7]));
        for (int i = 0; i < 3; ++i) UP3D_CUDA_OK(cudaEventElapsedTime(&ms6[3 + i], g_timing.ev[4 + i], g_timing.ev[5 + i]));
    }
    return 0;
}

int up3d_raster_debug_state(const up3d_raster_desc *d, const void *state, int32_t *n_visible, int32_t *sorted_ids,
                            float *depths, float *xy, float *conic_opacity, float *rgb, int32_t *rects, float *final_T,
                            int32_t *n_contrib, up3d_stream_t stream_) {
    if (validate_desc(d)) return 1;
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(state != nullptr, "up3d_raster_debug_state: null state");
    State st = carve_state(d, const_cast<void *>(state));
    const size_t HW = (size_t)d->width * d->height;
    if (n_visible) UP3D_CUDA_OK(cudaMemcpyAsync(n_visible, st.n_vis, 4 * (size_t)d->n_views, cudaMemcpyDeviceToDevice, stream));
    if (sorted_ids) UP3D_CUDA_OK(cudaMemcpyAsync(sorted_ids, st.s_id, 4 * (size_t)d->n_records, cudaMemcpyDeviceToDevice, stream));
    if (final_T) UP3D_CUDA_OK(cudaMemcpyAsync(final_T, st.final_T, 4 * HW * d->n_views, cudaMemcpyDeviceToDevice, stream));
    if (n_contrib) UP3D_CUDA_OK(cudaMemcpyAsync(n_contrib, st.n_contrib, 4 * HW * d->n_views, cudaMemcpyDeviceToDevice, stream));
    if ((depths || xy || conic_opacity || rgb || rects) && d->n_records > 0) {
        debug_unpack_kernel<<<div_up(d->n_records, 256), 256, 0, stream>>>(d->n_records, st, depths, xy, conic_opacity, rgb, rects);
        UP3D_LAUNCH_OK("debug_unpack_kernel");
    }
    return 0;
}

int up3d_raster_debug_bins(const up3d_raster_desc *d, const void *state, int32_t *bin_mode, int32_t *bin_total,
                           up3d_stream_t stream_) {
    if (validate_desc(d)) return 1;
    UP3D_CHECK_ARG(state != nullptr, "up3d_raster_debug_bins: null state");
    State st = carve_state(d, const_cast<void *>(state));
    cudaStream_t stream = (cudaStream_t)stream_;
    if (bin_mode) UP3D_CUDA_OK(cudaMemcpyAsync(bin_mode, st.bin_mode, 4 * (size_t)d->n_views, cudaMemcpyDeviceToDevice, stream));
    if (bin_total) UP3D_CUDA_OK(cudaMemcpyAsync(bin_total, st.bin_total, 4 * (size_t)d->n_views, cudaMemcpyDeviceToDevice, stream));
    return 0;
}

int up3d_raster_debug_tile_lists(const up3d_raster_desc *d, const void *state, int32_t *tile_counts,
                                 const int32_t *tile_offsets, int32_t *tile_lists, up3d_stream_t stream_) {
    if (validate_desc(d)) return 1;
    cudaStream_t stream = (cudaStream_t)stream_;
    UP3D_CHECK_ARG(state && tile_counts, "up3d_raster_debug_tile_lists: null state/tile_counts");
    UP3D_CHECK_ARG(tile_lists == nullptr || tile_offsets != nullptr, "up3d_raster_debug_tile_lists: tile_offsets required with tile_lists");
    if (d->n_views == 0) return 0;
    State st = carve_state(d, const_cast<void *>(state));
    const int gx = div_up(d->width, UP3D_TILE), gy = div_up(d->height, UP3D_TILE);
    DebugListArgs a{gx, gy, d->view_rec_start, st, tile_counts, tile_offsets, tile_lists};
    debug_tile_lists_kernel<<<dim3(gx, gy, d->n_views), UP3D_TILE_PIX, 0, stream>>>(a);
    UP3D_LAUNCH_OK("debug_tile_lists_kernel");
    return 0;
}

}  // extern "C"
