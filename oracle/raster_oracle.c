/*
 * raster_oracle.c -- CPU restatement of the differentiable 3D-Gaussian-Splatting rasterizer the
 * reference calls through gaussian_renderer/__init__.py:8,45-61,89-97.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (unipre3d_b200/) may import, link
 * or call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs do, and only as the checker / reported CPU baseline.
 *
 * *** PARITY UNPINNED ***  The arithmetic lives in a third-party dependency that is NOT in
 * /root/reference: graphdeco-inria/diff-gaussian-rasterization, installed from an unpinned clone of
 * gaussian-splatting HEAD (/root/reference/docs/INSTALLATION.md:51-62); the call site's
 * `antialiasing=` kwarg and 3-tuple return (gaussian_renderer/__init__.py:58,89) fix the API
 * generation.  The reference ships no test, golden vector or fixture for this boundary
 * (SURVEY.md §4, §8c).  This file restates the published algorithm as specified in SURVEY.md
 * Appendix A (A.1 preprocess, A.2 cov3D, A.3 cov2D, A.4 binning/sort, A.5 SH, A.6 blend fwd,
 * A.7 blend bwd, A.8/A.9 geometry backward).  Section tags below refer to that appendix.
 * The backward here is hand-derived; tests/ validate it against fp64 torch autograd of the
 * forward restatement (oracle/raster_ref.py), the route Appendix A.9 recommends.
 *
 * Determinism: every quantity that decides tile assignment or sort order (view-space depth,
 * pixel centre, covariance, radius, tile rectangle) is computed with an explicit operation
 * sequence (fmaf / single-rounded ops), chosen to mirror how nvcc contracts the upstream
 * source expressions (rule observed on the reference's own kernels: p0 + p1 + p2 ->
 * fma(p2, fma(p0, round(p1)))).  The CUDA product restates the same sequence with __fmaf_rn
 * etc., so lists can be compared bit-for-bit.  Compile with -ffp-contract=off.
 *
 * Multi-threading: OpenMP over Gaussians / tiles (for the CPU baseline timing only).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLK 16

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

typedef struct {
    int P;              /* Gaussians */
    int M;              /* SH coefficients stored per Gaussian (shs is (P,M,3)); 0 if colors_precomp */
    int D;              /* active SH degree */
    int W, H;
    float tanfovx, tanfovy;
    float scale_modifier;
    int antialiasing;
    const float *bg;             /* 3 */
    const float *means3D;        /* P*3 */
    const float *shs;            /* P*M*3 or NULL */
    const float *colors_precomp; /* P*3 or NULL */
    const float *opacities;      /* P */
    const float *scales;         /* P*3 */
    const float *rotations;      /* P*4 (r,x,y,z), NOT normalised (A.2) */
    const float *viewmatrix;     /* 16, row-vector convention flat (A. conventions) */
    const float *projmatrix;     /* 16 */
    const float *campos;         /* 3 */
} up3d_oracle_scene;

/* per-Gaussian forward state (what upstream keeps in its geometry buffer) */
typedef struct {
    float *depth;        /* P */
    float *xy;           /* P*2 */
    float *conic_opacity;/* P*4 */
    float *rgb;          /* P*3 */
    int32_t *radii;      /* P */
    int32_t *rect;       /* P*4 : minx, miny, maxx, maxy (tile units) */
    int32_t *tiles_touched; /* P */
    uint8_t *clamped;    /* P*3 */
    float *cov3D;        /* P*6 */
} up3d_oracle_geom;

int up3d_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- A. conventions: transformPoint4x3 / 4x4 with nvcc's contraction ---- */
static inline float tp_row(const float *m, int r, float x, float y, float z) {
    /* m[r]*x + m[4+r]*y + m[8+r]*z + m[12+r] */
    return fmaf(m[8 + r], z, fmaf(m[r], x, m[4 + r] * y)) + m[12 + r];
}

static inline float dot3_c(float a0, float b0, float a1, float b1, float a2, float b2) {
    /* a0*b0 + a1*b1 + a2*b2 as nvcc contracts it */
    return fmaf(a2, b2, fmaf(a0, b0, a1 * b1));
}

static inline float clampf(float v, float lo, float hi) { return fminf(hi, fmaxf(lo, v)); }

/* A.2 */
static void cov3d_from_scale_rot(const float *s3, float mod, const float *q, float *cov6, float *Rout, float *sout) {
    const float r = q[0], x = q[1], y = q[2], z = q[3];
    float R[9];
    R[0] = fmaf(-2.f, fmaf(y, y, z * z), 1.f);
    R[1] = 2.f * fmaf(x, y, -(r * z));
    R[2] = 2.f * fmaf(x, z, r * y);
    R[3] = 2.f * fmaf(x, y, r * z);
    R[4] = fmaf(-2.f, fmaf(x, x, z * z), 1.f);
    R[5] = 2.f * fmaf(y, z, -(r * x));
    R[6] = 2.f * fmaf(x, z, -(r * y));
    R[7] = 2.f * fmaf(y, z, r * x);
    R[8] = fmaf(-2.f, fmaf(x, x, y * y), 1.f);
    float s[3] = {mod * s3[0], mod * s3[1], mod * s3[2]};
    float A[9]; /* A[a][i] = s_i * R[a][i] */
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < 3; ++i) A[a * 3 + i] = s[i] * R[a * 3 + i];
#define SIG(a, b) dot3_c(A[(a)*3 + 0], A[(b)*3 + 0], A[(a)*3 + 1], A[(b)*3 + 1], A[(a)*3 + 2], A[(b)*3 + 2])
    cov6[0] = SIG(0, 0);
    cov6[1] = SIG(0, 1);
    cov6[2] = SIG(0, 2);
    cov6[3] = SIG(1, 1);
    cov6[4] = SIG(1, 2);
    cov6[5] = SIG(2, 2);
#undef SIG
    if (Rout) memcpy(Rout, R, sizeof(R));
    if (sout) memcpy(sout, s, sizeof(s));
}

/* A.3 : returns cov2D (xx, xy, yy); optionally the 2x3 T and the clamp masks */
static void cov2d_ewa(const float *t_in, float fx, float fy, float tanfovx, float tanfovy, const float *cov6,
                      const float *view, float *cov3, float *Tout, float *tc_out, int *clampx, int *clampy) {
    float tx = t_in[0], ty = t_in[1];
    const float tz = t_in[2];
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = tx / tz, tytz = ty / tz;
    tx = clampf(txtz, -limx, limx) * tz;
    ty = clampf(tytz, -limy, limy) * tz;
    if (clampx) *clampx = (txtz < -limx || txtz > limx);
    if (clampy) *clampy = (tytz < -limy || tytz > limy);
    const float J00 = fx / tz, J02 = -(fx * tx) / (tz * tz);
    const float J11 = fy / tz, J12 = -(fy * ty) / (tz * tz);
    /* T_ci = sum_k J_ck * view[4i + k]  (Tm = J_std * Rwc, 2x3) */
    float T[6];
    for (int i = 0; i < 3; ++i) {
        T[0 * 3 + i] = fmaf(view[4 * i + 2], J02, view[4 * i + 0] * J00);
        T[1 * 3 + i] = fmaf(view[4 * i + 2], J12, view[4 * i + 1] * J11);
    }
    const float V[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
    float X[6]; /* X = T * V (2x3) */
    for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c)
            X[r * 3 + c] = dot3_c(T[r * 3 + 0], V[0 * 3 + c], T[r * 3 + 1], V[1 * 3 + c], T[r * 3 + 2], V[2 * 3 + c]);
    cov3[0] = dot3_c(X[0], T[0], X[1], T[1], X[2], T[2]);       /* (0,0) */
    cov3[1] = dot3_c(X[3], T[0], X[4], T[1], X[5], T[2]);       /* (1,0) == glm cov[0][1] */
    cov3[2] = dot3_c(X[3], T[3], X[4], T[4], X[5], T[5]);       /* (1,1) */
    if (Tout) memcpy(Tout, T, sizeof(T));
    if (tc_out) { tc_out[0] = tx; tc_out[1] = ty; tc_out[2] = tz; }
}

/* A.5 : SH -> RGB for one Gaussian. dir is normalised inside. */
static void sh_to_rgb(int deg, int M, const float *sh /* M*3 */, const float *p, const float *campos, float *rgb,
                      uint8_t *clamped) {
    float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    dx /= len; dy /= len; dz /= len;
    (void)M;
    for (int ch = 0; ch < 3; ++ch) {
        float r = SH_C0 * sh[0 * 3 + ch];
        if (deg > 0) {
            const float x = dx, y = dy, z = dz;
            r = r - SH_C1 * y * sh[1 * 3 + ch] + SH_C1 * z * sh[2 * 3 + ch] - SH_C1 * x * sh[3 * 3 + ch];
            if (deg > 1) {
                const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + ch] + SH_C2[1] * yz * sh[5 * 3 + ch] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + ch] + SH_C2[3] * xz * sh[7 * 3 + ch] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + ch];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + ch] + SH_C3[1] * xy * z * sh[10 * 3 + ch] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + ch] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + ch] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + ch] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + ch] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + ch];
                }
            }
        }
        r += 0.5f;
        clamped[ch] = (r < 0.0f);
        rgb[ch] = fmaxf(r, 0.0f);
    }
}

/* exported for the golden-vector test against the reference's in-tree SH restatement (utils/sh_utils.py:57-112) */
void up3d_oracle_eval_sh(int deg, int M, int n, const float *sh /* n*M*3 */, const float *pos /* n*3 */,
                         const float *campos /* 3 */, float *rgb /* n*3 */, uint8_t *clamped /* n*3 */) {
    for (int i = 0; i < n; ++i) sh_to_rgb(deg, M, sh + (size_t)i * M * 3, pos + 3 * i, campos, rgb + 3 * i, clamped + 3 * i);
}

static inline float ndc2pix(float v, int S) {
    /* ((v + 1.0) * S - 1.0) * 0.5 in double (A.1 step 7); nvcc contracts (..)*S - 1.0 into one fma */
    return (float)(fma((double)v + 1.0, (double)S, -1.0) * 0.5);
}

/* float -> int with CUDA's saturating semantics (cvt.rzi.s32.f32): NaN -> 0, +-overflow -> INT_MAX/INT_MIN.
 * A plain C cast is undefined there (x86 yields INT_MIN for both), and scales up to e^20 can reach it. */
static inline int f2i_sat(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ A.1 preprocess */
void up3d_oracle_preprocess(const up3d_oracle_scene *sc, up3d_oracle_geom *g) {
    const int P = sc->P;
    const int gx = (sc->W + BLK - 1) / BLK, gy = (sc->H + BLK - 1) / BLK;
    const float fx = sc->W / (2.0f * sc->tanfovx), fy = sc->H / (2.0f * sc->tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        g->radii[i] = 0;
        g->tiles_touched[i] = 0;
        g->depth[i] = 0.f;
        g->xy[2 * i] = g->xy[2 * i + 1] = 0.f;
        for (int k = 0; k < 4; ++k) { g->conic_opacity[4 * i + k] = 0.f; g->rect[4 * i + k] = 0; }
        for (int k = 0; k < 3; ++k) { g->rgb[3 * i + k] = 0.f; g->clamped[3 * i + k] = 0; }
        for (int k = 0; k < 6; ++k) g->cov3D[6 * i + k] = 0.f;
        const float *p = sc->means3D + 3 * i;
        float t[3];
        for (int r = 0; r < 3; ++r) t[r] = tp_row(sc->viewmatrix, r, p[0], p[1], p[2]);
        if (t[2] <= 0.2f) continue;
        float ph[4];
        for (int r = 0; r < 4; ++r) ph[r] = tp_row(sc->projmatrix, r, p[0], p[1], p[2]);
        const float p_w = 1.0f / (ph[3] + 0.0000001f);
        const float px = ph[0] * p_w, py = ph[1] * p_w;
        float cov6[6];
        cov3d_from_scale_rot(sc->scales + 3 * i, sc->scale_modifier, sc->rotations + 4 * i, cov6, NULL, NULL);
        memcpy(g->cov3D + 6 * i, cov6, sizeof(cov6));
        float cov[3];
        cov2d_ewa(t, fx, fy, sc->tanfovx, sc->tanfovy, cov6, sc->viewmatrix, cov, NULL, NULL, NULL, NULL);
        const float h_var = 0.3f;
        const float det0 = fmaf(cov[0], cov[2], -(cov[1] * cov[1]));
        const float a = cov[0] + h_var, b = cov[1], c = cov[2] + h_var;
        const float det = fmaf(a, c, -(b * b));
        float hs = 1.0f;
        if (sc->antialiasing) hs = sqrtf(fmaxf(0.000025f, det0 / det));
        if (det == 0.0f) continue;
        const float det_inv = 1.f / det;
        const float conic[3] = {c * det_inv, -b * det_inv, a * det_inv};
        const float mid = 0.5f * (a + c);
        const float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        const float l1 = mid + disc, l2 = mid - disc;
        const float my_radius = ceilf(3.f * sqrtf(fmaxf(l1, l2)));
        const float ix = ndc2pix(px, sc->W), iy = ndc2pix(py, sc->H);
        const int rminx = imin(gx, imax(0, f2i_sat((ix - my_radius) / (float)BLK)));
        const int rminy = imin(gy, imax(0, f2i_sat((iy - my_radius) / (float)BLK)));
        const int rmaxx = imin(gx, imax(0, f2i_sat((ix + my_radius + (float)(BLK - 1)) / (float)BLK)));
        const int rmaxy = imin(gy, imax(0, f2i_sat((iy + my_radius + (float)(BLK - 1)) / (float)BLK)));
        if ((rmaxx - rminx) * (rmaxy - rminy) == 0) continue;
        if (sc->colors_precomp) {
            for (int k = 0; k < 3; ++k) g->rgb[3 * i + k] = sc->colors_precomp[3 * i + k];
        } else {
            sh_to_rgb(sc->D, sc->M, sc->shs + (size_t)i * sc->M * 3, p, sc->campos, g->rgb + 3 * i, g->clamped + 3 * i);
        }
        g->depth[i] = t[2];
        g->radii[i] = f2i_sat(my_radius);
        g->xy[2 * i] = ix;
        g->xy[2 * i + 1] = iy;
        g->conic_opacity[4 * i + 0] = conic[0];
        g->conic_opacity[4 * i + 1] = conic[1];
        g->conic_opacity[4 * i + 2] = conic[2];
        g->conic_opacity[4 * i + 3] = sc->opacities[i] * hs;
        g->rect[4 * i + 0] = rminx; g->rect[4 * i + 1] = rminy; g->rect[4 * i + 2] = rmaxx; g->rect[4 * i + 3] = rmaxy;
        g->tiles_touched[i] = (rmaxy - rminy) * (rmaxx - rminx);
    }
}

/* ------------------------------------------------------------------ A.4 binning */
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

int64_t up3d_oracle_num_rendered(int P, const int32_t *tiles_touched) {
    int64_t s = 0;
    for (int i = 0; i < P; ++i) s += tiles_touched[i];
    return s;
}

/* keys/vals: caller-allocated num_rendered entries. ranges: (tiles,2) u32, zero-filled here.
 * Stable LSD radix sort (8-bit digits) over bits [0, 32+msb) like cub::DeviceRadixSort::SortPairs. */
void up3d_oracle_bin(int P, int W, int H, const up3d_oracle_geom *g, uint64_t *keys, uint32_t *vals, uint32_t *ranges) {
    const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
    int64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (g->radii[i] <= 0) continue;
        uint32_t dbits;
        memcpy(&dbits, &g->depth[i], 4);
        for (int y = g->rect[4 * i + 1]; y < g->rect[4 * i + 3]; ++y)
            for (int x = g->rect[4 * i + 0]; x < g->rect[4 * i + 2]; ++x) {
                keys[off] = ((uint64_t)(uint32_t)(y * gx + x) << 32) | dbits;
                vals[off] = (uint32_t)i;
                ++off;
            }
    }
    const int64_t L = off;
    const int bits = 32 + (int)higher_msb((uint32_t)(gx * gy));
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(L > 0 ? L : 1));
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(L > 0 ? L : 1));
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < bits; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        int nb = bits - shift < 8 ? bits - shift : 8;
        uint64_t mask = (1ull << nb) - 1;
        for (int64_t j = 0; j < L; ++j) cnt[((ka[j] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t j = 0; j < L; ++j) {
            int64_t dst = cnt[(ka[j] >> shift) & mask]++;
            kb[dst] = ka[j];
            vb[dst] = va[j];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, sizeof(uint64_t) * (size_t)L); memcpy(vals, va, sizeof(uint32_t) * (size_t)L); }
    free(k2);
    free(v2);
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)(gx * gy));
    for (int64_t j = 0; j < L; ++j) {
        uint32_t tile = (uint32_t)(keys[j] >> 32);
        if (j == 0) ranges[2 * tile] = 0;
        else {
            uint32_t prev = (uint32_t)(keys[j - 1] >> 32);
            if (tile != prev) { ranges[2 * prev + 1] = (uint32_t)j; ranges[2 * tile] = (uint32_t)j; }
        }
        if (j == L - 1) ranges[2 * tile + 1] = (uint32_t)L;
    }
}

/* power = -0.5f*(co.x*dx*dx + co.z*dy*dy) - co.y*dx*dy, in the operation order nvcc gives that expression */
static inline float gauss_power(const float *co, float dx, float dy) {
    const float A = fmaf(co[0] * dx, dx, (co[2] * dy) * dy);
    return fmaf(-0.5f, A, -((co[1] * dx) * dy));
}

/* ------------------------------------------------------------------ A.6 blend forward */
void up3d_oracle_blend_forward(int W, int H, const float *bg, const up3d_oracle_geom *g, const uint32_t *point_list,
                               const uint32_t *ranges, float *out_color /*3*H*W*/, float *final_T /*H*W*/,
                               uint32_t *n_contrib /*H*W*/) {
    const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
        const int tx0 = (tile % gx) * BLK, ty0 = (tile / gx) * BLK;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int py = ty0; py < ty0 + BLK && py < H; ++py)
            for (int px = tx0; px < tx0 + BLK && px < W; ++px) {
                const float pfx = (float)px, pfy = (float)py;
                float T = 1.0f, C[3] = {0.f, 0.f, 0.f};
                uint32_t contributor = 0, last_contributor = 0;
                for (uint32_t j = r0; j < r1; ++j) {
                    contributor++;
                    const uint32_t id = point_list[j];
                    const float dx = g->xy[2 * id] - pfx, dy = g->xy[2 * id + 1] - pfy;
                    const float *co = g->conic_opacity + 4 * id;
                    const float power = gauss_power(co, dx, dy);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1.f - alpha);
                    if (test_T < 0.0001f) break;
                    for (int ch = 0; ch < 3; ++ch) C[ch] += g->rgb[3 * id + ch] * alpha * T;
                    T = test_T;
                    last_contributor = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                final_T[pix] = T;
                n_contrib[pix] = last_contributor;
                for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pix] = C[ch] + T * bg[ch];
            }
    }
}

/* ------------------------------------------------------------------ A.7 blend backward
 * Outputs (double accumulators rounded once; upstream uses float atomics in arbitrary order):
 *   dL_dmean2D (P,2) in PIXEL units (the 0.5W / 0.5H factors of upstream are applied in the
 *              geometry backward below), dL_dconic (P,3) = d/d(conic.x, conic.y, conic.z) with the
 *              FULL off-diagonal derivative in .y, dL_dopac (P) w.r.t. opacity*hs, dL_drgb (P,3). */
void up3d_oracle_blend_backward(int P, int W, int H, const float *bg, const up3d_oracle_geom *g,
                                const uint32_t *point_list, const uint32_t *ranges, const float *final_T,
                                const uint32_t *n_contrib, const float *dL_dpix /*3*H*W*/, float *dL_dmean2D,
                                float *dL_dconic, float *dL_dopac, float *dL_drgb) {
    const int gx = (W + BLK - 1) / BLK, gy = (H + BLK - 1) / BLK;
    int nth = up3d_oracle_num_threads();
    double *acc = (double *)calloc((size_t)nth * P * 9, sizeof(double));
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; ++tile) {
#ifdef _OPENMP
        double *A = acc + (size_t)omp_get_thread_num() * P * 9;
#else
        double *A = acc;
#endif
        const int tx0 = (tile % gx) * BLK, ty0 = (tile / gx) * BLK;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int py = ty0; py < ty0 + BLK && py < H; ++py)
            for (int px = tx0; px < tx0 + BLK && px < W; ++px) {
                const size_t pix = (size_t)py * W + px;
                const float pfx = (float)px, pfy = (float)py;
                const float T_final = final_T[pix];
                float T = T_final;
                const uint32_t last_contributor = n_contrib[pix];
                float accum_rec[3] = {0, 0, 0}, last_color[3] = {0, 0, 0}, last_alpha = 0.f;
                float dLp[3];
                for (int ch = 0; ch < 3; ++ch) dLp[ch] = dL_dpix[(size_t)ch * H * W + pix];
                uint32_t contributor = r1 - r0;
                for (uint32_t jj = r1; jj > r0; --jj) {
                    const uint32_t j = jj - 1;
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t id = point_list[j];
                    const float dx = g->xy[2 * id] - pfx, dy = g->xy[2 * id + 1] - pfy;
                    const float *co = g->conic_opacity + 4 * id;
                    const float power = gauss_power(co, dx, dy);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.f;
                    for (int ch = 0; ch < 3; ++ch) {
                        const float c = g->rgb[3 * id + ch];
                        accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
                        last_color[ch] = c;
                        dL_dalpha += (c - accum_rec[ch]) * dLp[ch];
                        A[(size_t)id * 9 + 6 + ch] += (double)(dchannel_dcolor * dLp[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    float bg_dot = 0.f;
                    for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dLp[ch];
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    A[(size_t)id * 9 + 0] += (double)(dL_dG * dG_ddelx);
                    A[(size_t)id * 9 + 1] += (double)(dL_dG * dG_ddely);
                    A[(size_t)id * 9 + 2] += (double)(-0.5f * gdx * dx * dL_dG);
                    A[(size_t)id * 9 + 3] += (double)(-gdx * dy * dL_dG);
                    A[(size_t)id * 9 + 4] += (double)(-0.5f * gdy * dy * dL_dG);
                    A[(size_t)id * 9 + 5] += (double)(G * dL_dalpha);
                }
            }
    }
    for (int i = 0; i < P; ++i) {
        double s[9] = {0};
        for (int t = 0; t < nth; ++t)
            for (int k = 0; k < 9; ++k) s[k] += acc[((size_t)t * P + i) * 9 + k];
        dL_dmean2D[2 * i] = (float)s[0];
        dL_dmean2D[2 * i + 1] = (float)s[1];
        dL_dconic[3 * i] = (float)s[2];
        dL_dconic[3 * i + 1] = (float)s[3];
        dL_dconic[3 * i + 2] = (float)s[4];
        dL_dopac[i] = (float)s[5];
        for (int k = 0; k < 3; ++k) dL_drgb[3 * i + k] = (float)s[6 + k];
    }
    free(acc);
}

/* ------------------------------------------------------------------ A.8 + A.9 geometry backward
 * Inputs: screen-space grads from the blend backward.  Outputs (all zero for radii==0):
 *   dL_dmeans3D (P,3), dL_dmeans2D (P,3) [upstream's viewspace grad: (gx*0.5W, gy*0.5H, 0)],
 *   dL_dsh (P,M,3), dL_dcolors (P,3) [when colors_precomp], dL_dopacity (P), dL_dscales (P,3), dL_drot (P,4). */
void up3d_oracle_preprocess_backward(const up3d_oracle_scene *sc, const up3d_oracle_geom *g, const float *G_mean2D,
                                     const float *G_conic, const float *G_opac, const float *G_rgb, float *dL_dmeans3D,
                                     float *dL_dmeans2D, float *dL_dsh, float *dL_dcolors, float *dL_dopacity,
                                     float *dL_dscales, float *dL_drot) {
    const int P = sc->P, M = sc->M;
    const float fx = sc->W / (2.0f * sc->tanfovx), fy = sc->H / (2.0f * sc->tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; ++i) {
        for (int k = 0; k < 3; ++k) { dL_dmeans3D[3 * i + k] = 0.f; dL_dmeans2D[3 * i + k] = 0.f; dL_dscales[3 * i + k] = 0.f; }
        for (int k = 0; k < 4; ++k) dL_drot[4 * i + k] = 0.f;
        dL_dopacity[i] = 0.f;
        if (dL_dsh) for (int k = 0; k < M * 3; ++k) dL_dsh[(size_t)i * M * 3 + k] = 0.f;
        if (dL_dcolors) for (int k = 0; k < 3; ++k) dL_dcolors[3 * i + k] = 0.f;
        if (g->radii[i] <= 0) continue;
        const float *p = sc->means3D + 3 * i;
        double dmean[3] = {0, 0, 0};

        /* ---- recompute forward intermediates */
        float t[3];
        for (int r = 0; r < 3; ++r) t[r] = tp_row(sc->viewmatrix, r, p[0], p[1], p[2]);
        float cov6[6], R[9], s[3];
        cov3d_from_scale_rot(sc->scales + 3 * i, sc->scale_modifier, sc->rotations + 4 * i, cov6, R, s);
        float cov[3], T[6], tc[3];
        int cx, cy;
        cov2d_ewa(t, fx, fy, sc->tanfovx, sc->tanfovy, cov6, sc->viewmatrix, cov, T, tc, &cx, &cy);
        const double h = 0.3;
        const double x = cov[0], z = cov[1], y = cov[2];
        const double a = x + h, b = z, c = y + h;
        const double det0 = x * y - z * z, det = a * c - b * b;
        double hs = 1.0, dLds_over = 0.0; /* dL/dr where r = det0/det */
        const double opac = sc->opacities[i];
        if (sc->antialiasing) {
            const double r = det0 / det;
            hs = sqrt(fmax(0.000025, r));
            const double dL_dhs = (double)G_opac[i] * opac;
            dLds_over = (r <= 0.000025) ? 0.0 : dL_dhs / (2.0 * hs);
        }
        dL_dopacity[i] = (float)((double)G_opac[i] * hs);

        /* ---- conic -> (a,b,c) */
        const double Gx = G_conic[3 * i], Gy = G_conic[3 * i + 1], Gz = G_conic[3 * i + 2];
        const double id2 = 1.0 / (det * det);
        double dL_da = (-c * c * Gx + b * c * Gy - b * b * Gz) * id2;
        double dL_db = (2 * b * c * Gx - (det + 2 * b * b) * Gy + 2 * a * b * Gz) * id2;
        double dL_dc = (-b * b * Gx + a * b * Gy - a * a * Gz) * id2;
        /* ---- antialias scaling hs -> cov2D */
        if (sc->antialiasing && dLds_over != 0.0) {
            dL_da += dLds_over * (y * det - det0 * c) * id2;
            dL_dc += dLds_over * (x * det - det0 * a) * id2;
            dL_db += dLds_over * (-2.0 * z * (det - det0)) * id2;
        }
        /* ---- cov2D -> cov3D (6 params) and T */
        const double T00 = T[0], T01 = T[1], T02 = T[2], T10 = T[3], T11 = T[4], T12 = T[5];
        double dS[6];
        dS[0] = T00 * T00 * dL_da + T00 * T10 * dL_db + T10 * T10 * dL_dc;
        dS[3] = T01 * T01 * dL_da + T01 * T11 * dL_db + T11 * T11 * dL_dc;
        dS[5] = T02 * T02 * dL_da + T02 * T12 * dL_db + T12 * T12 * dL_dc;
        dS[1] = 2 * T00 * T01 * dL_da + (T00 * T11 + T01 * T10) * dL_db + 2 * T10 * T11 * dL_dc;
        dS[2] = 2 * T00 * T02 * dL_da + (T00 * T12 + T02 * T10) * dL_db + 2 * T10 * T12 * dL_dc;
        dS[4] = 2 * T02 * T01 * dL_da + (T01 * T12 + T02 * T11) * dL_db + 2 * T11 * T12 * dL_dc;
        const double V[9] = {cov6[0], cov6[1], cov6[2], cov6[1], cov6[3], cov6[4], cov6[2], cov6[4], cov6[5]};
        double dT[6];
        for (int k = 0; k < 3; ++k) {
            const double VT0 = V[k * 3 + 0] * T00 + V[k * 3 + 1] * T01 + V[k * 3 + 2] * T02;
            const double VT1 = V[k * 3 + 0] * T10 + V[k * 3 + 1] * T11 + V[k * 3 + 2] * T12;
            dT[0 * 3 + k] = 2 * dL_da * VT0 + dL_db * VT1;
            dT[1 * 3 + k] = 2 * dL_dc * VT1 + dL_db * VT0;
        }
        /* ---- T = J * Rwc : dL/dJ_ck = sum_i dT_ci * view[4i + k] */
        const float *vm = sc->viewmatrix;
        const double dJ00 = dT[0] * vm[0] + dT[1] * vm[4] + dT[2] * vm[8];
        const double dJ02 = dT[0] * vm[2] + dT[1] * vm[6] + dT[2] * vm[10];
        const double dJ11 = dT[3] * vm[1] + dT[4] * vm[5] + dT[5] * vm[9];
        const double dJ12 = dT[3] * vm[2] + dT[4] * vm[6] + dT[5] * vm[10];
        const double tz = 1.0 / tc[2], tz2 = tz * tz, tz3 = tz2 * tz;
        const double dtx = (cx ? 0.0 : 1.0) * -fx * tz2 * dJ02;
        const double dty = (cy ? 0.0 : 1.0) * -fy * tz2 * dJ12;
        const double dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2 * fx * tc[0]) * tz3 * dJ02 + (2 * fy * tc[1]) * tz3 * dJ12;
        /* t = Rwc p + trans : dL/dp_c = sum_r view[4c + r] dL/dt_r */
        for (int cc = 0; cc < 3; ++cc) dmean[cc] += vm[4 * cc + 0] * dtx + vm[4 * cc + 1] * dty + vm[4 * cc + 2] * dtz;

        /* ---- cov3D -> scale, rotation (A.2): Sigma = A A^T, A_ai = s_i R_ai */
        {
            const double Gs[9] = {dS[0], 0.5 * dS[1], 0.5 * dS[2], 0.5 * dS[1], dS[3], 0.5 * dS[4], 0.5 * dS[2], 0.5 * dS[4], dS[5]};
            double dA[9];
            for (int aa = 0; aa < 3; ++aa)
                for (int ii = 0; ii < 3; ++ii) {
                    double v = 0;
                    for (int bb = 0; bb < 3; ++bb) v += 2.0 * Gs[aa * 3 + bb] * ((double)s[ii] * R[bb * 3 + ii]);
                    dA[aa * 3 + ii] = v;
                }
            double dR[9];
            for (int ii = 0; ii < 3; ++ii) {
                double ds = 0;
                for (int aa = 0; aa < 3; ++aa) { ds += dA[aa * 3 + ii] * R[aa * 3 + ii]; dR[aa * 3 + ii] = dA[aa * 3 + ii] * s[ii]; }
                dL_dscales[3 * i + ii] = (float)(ds * sc->scale_modifier);
            }
            const double qr = sc->rotations[4 * i], qx = sc->rotations[4 * i + 1], qy = sc->rotations[4 * i + 2], qz = sc->rotations[4 * i + 3];
            /* R00=1-2(yy+zz) R01=2(xy-rz) R02=2(xz+ry) R10=2(xy+rz) R11=1-2(xx+zz) R12=2(yz-rx) R20=2(xz-ry) R21=2(yz+rx) R22=1-2(xx+yy) */
            const double dqr = 2 * (-qz * dR[1] + qy * dR[2] + qz * dR[3] - qx * dR[5] - qy * dR[6] + qx * dR[7]);
            const double dqx = 2 * (qy * dR[1] + qz * dR[2] + qy * dR[3] - 2 * qx * dR[4] - qr * dR[5] + qz * dR[6] + qr * dR[7] - 2 * qx * dR[8]);
            const double dqy = 2 * (-2 * qy * dR[0] + qx * dR[1] + qr * dR[2] + qx * dR[3] + qz * dR[5] - qr * dR[6] + qz * dR[7] - 2 * qy * dR[8]);
            const double dqz = 2 * (-2 * qz * dR[0] - qr * dR[1] + qx * dR[2] + qr * dR[3] - 2 * qz * dR[4] + qy * dR[5] + qx * dR[6] + qy * dR[7]);
            dL_drot[4 * i + 0] = (float)dqr; dL_drot[4 * i + 1] = (float)dqx; dL_drot[4 * i + 2] = (float)dqy; dL_drot[4 * i + 3] = (float)dqz;
        }

        /* ---- mean2D -> mean3D through the projection (A.9) */
        {
            const float *pm = sc->projmatrix;
            float ph[4];
            for (int r = 0; r < 4; ++r) ph[r] = tp_row(pm, r, p[0], p[1], p[2]);
            const double m_w = 1.0 / ((double)ph[3] + 0.0000001);
            const double gxn = (double)G_mean2D[2 * i] * 0.5 * sc->W, gyn = (double)G_mean2D[2 * i + 1] * 0.5 * sc->H;
            dL_dmeans2D[3 * i + 0] = (float)gxn;
            dL_dmeans2D[3 * i + 1] = (float)gyn;
            const double mul1 = ph[0] * m_w * m_w, mul2 = ph[1] * m_w * m_w;
            for (int k = 0; k < 3; ++k)
                dmean[k] += (pm[4 * k + 0] * m_w - pm[4 * k + 3] * mul1) * gxn + (pm[4 * k + 1] * m_w - pm[4 * k + 3] * mul2) * gyn;
        }

        /* ---- colour (A.5 backward) */
        if (sc->colors_precomp) {
            if (dL_dcolors) for (int k = 0; k < 3; ++k) dL_dcolors[3 * i + k] = G_rgb[3 * i + k];
        } else {
            const int deg = sc->D;
            const float *sh = sc->shs + (size_t)i * M * 3;
            float *dsh = dL_dsh + (size_t)i * M * 3;
            double d0[3] = {(double)p[0] - sc->campos[0], (double)p[1] - sc->campos[1], (double)p[2] - sc->campos[2]};
            const double len = sqrt(d0[0] * d0[0] + d0[1] * d0[1] + d0[2] * d0[2]);
            const double xx_ = d0[0] / len, yy_ = d0[1] / len, zz_ = d0[2] / len;
            const double X = xx_, Y = yy_, Z = zz_;
            double dRGB[3];
            for (int ch = 0; ch < 3; ++ch) dRGB[ch] = g->clamped[3 * i + ch] ? 0.0 : (double)G_rgb[3 * i + ch];
            double ddir[3] = {0, 0, 0};
#define SHV(l, ch) ((double)sh[(l)*3 + (ch)])
#define DOTRGB(l) (dRGB[0] * SHV(l, 0) + dRGB[1] * SHV(l, 1) + dRGB[2] * SHV(l, 2))
#define SETSH(l, w) for (int ch = 0; ch < 3; ++ch) dsh[(l)*3 + ch] = (float)((w) * dRGB[ch])
            SETSH(0, (double)SH_C0);
            if (deg > 0) {
                SETSH(1, -(double)SH_C1 * Y);
                SETSH(2, (double)SH_C1 * Z);
                SETSH(3, -(double)SH_C1 * X);
                ddir[0] += -(double)SH_C1 * DOTRGB(3);
                ddir[1] += -(double)SH_C1 * DOTRGB(1);
                ddir[2] += (double)SH_C1 * DOTRGB(2);
                if (deg > 1) {
                    const double xx = X * X, yy = Y * Y, zz = Z * Z, xy = X * Y, yz = Y * Z, xz = X * Z;
                    SETSH(4, (double)SH_C2[0] * xy);
                    SETSH(5, (double)SH_C2[1] * yz);
                    SETSH(6, (double)SH_C2[2] * (2.0 * zz - xx - yy));
                    SETSH(7, (double)SH_C2[3] * xz);
                    SETSH(8, (double)SH_C2[4] * (xx - yy));
                    ddir[0] += (double)SH_C2[0] * Y * DOTRGB(4) + (double)SH_C2[2] * 2.0 * -X * DOTRGB(6) + (double)SH_C2[3] * Z * DOTRGB(7) + (double)SH_C2[4] * 2.0 * X * DOTRGB(8);
                    ddir[1] += (double)SH_C2[0] * X * DOTRGB(4) + (double)SH_C2[1] * Z * DOTRGB(5) + (double)SH_C2[2] * 2.0 * -Y * DOTRGB(6) + (double)SH_C2[4] * 2.0 * -Y * DOTRGB(8);
                    ddir[2] += (double)SH_C2[1] * Y * DOTRGB(5) + (double)SH_C2[2] * 2.0 * 2.0 * Z * DOTRGB(6) + (double)SH_C2[3] * X * DOTRGB(7);
                    if (deg > 2) {
                        SETSH(9, (double)SH_C3[0] * Y * (3.0 * xx - yy));
                        SETSH(10, (double)SH_C3[1] * xy * Z);
                        SETSH(11, (double)SH_C3[2] * Y * (4.0 * zz - xx - yy));
                        SETSH(12, (double)SH_C3[3] * Z * (2.0 * zz - 3.0 * xx - 3.0 * yy));
                        SETSH(13, (double)SH_C3[4] * X * (4.0 * zz - xx - yy));
                        SETSH(14, (double)SH_C3[5] * Z * (xx - yy));
                        SETSH(15, (double)SH_C3[6] * X * (xx - 3.0 * yy));
                        ddir[0] += (double)SH_C3[0] * DOTRGB(9) * 3.0 * 2.0 * xy + (double)SH_C3[1] * DOTRGB(10) * yz +
                                   (double)SH_C3[2] * DOTRGB(11) * -2.0 * xy + (double)SH_C3[3] * DOTRGB(12) * -3.0 * 2.0 * xz +
                                   (double)SH_C3[4] * DOTRGB(13) * (-3.0 * xx + 4.0 * zz - yy) + (double)SH_C3[5] * DOTRGB(14) * 2.0 * xz +
                                   (double)SH_C3[6] * DOTRGB(15) * 3.0 * (xx - yy);
                        ddir[1] += (double)SH_C3[0] * DOTRGB(9) * 3.0 * (xx - yy) + (double)SH_C3[1] * DOTRGB(10) * xz +
                                   (double)SH_C3[2] * DOTRGB(11) * (-3.0 * yy + 4.0 * zz - xx) + (double)SH_C3[3] * DOTRGB(12) * -3.0 * 2.0 * yz +
                                   (double)SH_C3[4] * DOTRGB(13) * -2.0 * xy + (double)SH_C3[5] * DOTRGB(14) * -2.0 * yz +
                                   (double)SH_C3[6] * DOTRGB(15) * -3.0 * 2.0 * xy;
                        ddir[2] += (double)SH_C3[1] * DOTRGB(10) * xy + (double)SH_C3[2] * DOTRGB(11) * 4.0 * 2.0 * yz +
                                   (double)SH_C3[3] * DOTRGB(12) * 3.0 * (2.0 * zz - xx - yy) + (double)SH_C3[4] * DOTRGB(13) * 4.0 * 2.0 * xz +
                                   (double)SH_C3[5] * DOTRGB(14) * (xx - yy);
                    }
                }
            }
#undef SHV
#undef DOTRGB
#undef SETSH
            /* through dir = d0/|d0| */
            const double dotv = (d0[0] * ddir[0] + d0[1] * ddir[1] + d0[2] * ddir[2]);
            const double inv3 = 1.0 / (len * len * len);
            for (int k = 0; k < 3; ++k) dmean[k] += (ddir[k] * len * len - d0[k] * dotv) * inv3;
        }
        for (int k = 0; k < 3; ++k) dL_dmeans3D[3 * i + k] = (float)dmean[k];
    }
}

/* ------------------------------------------------------------------ whole-pipeline helpers
 * (used for the CPU baseline timing and by tests that only need images / final grads) */
static void geom_alloc(up3d_oracle_geom *g, int P) {
    g->depth = (float *)calloc(P, sizeof(float));
    g->xy = (float *)calloc((size_t)P * 2, sizeof(float));
    g->conic_opacity = (float *)calloc((size_t)P * 4, sizeof(float));
    g->rgb = (float *)calloc((size_t)P * 3, sizeof(float));
    g->radii = (int32_t *)calloc(P, sizeof(int32_t));
    g->rect = (int32_t *)calloc((size_t)P * 4, sizeof(int32_t));
    g->tiles_touched = (int32_t *)calloc(P, sizeof(int32_t));
    g->clamped = (uint8_t *)calloc((size_t)P * 3, 1);
    g->cov3D = (float *)calloc((size_t)P * 6, sizeof(float));
}
static void geom_free(up3d_oracle_geom *g) {
    free(g->depth); free(g->xy); free(g->conic_opacity); free(g->rgb); free(g->radii);
    free(g->rect); free(g->tiles_touched); free(g->clamped); free(g->cov3D);
}

/* forward (+ optional backward when dL_dpix != NULL) of one view; returns num_rendered */
int64_t up3d_oracle_render(const up3d_oracle_scene *sc, float *out_color, int32_t *radii_out, const float *dL_dpix,
                           float *dL_dmeans3D, float *dL_dmeans2D, float *dL_dsh, float *dL_dcolors, float *dL_dopacity,
                           float *dL_dscales, float *dL_drot) {
    const int P = sc->P, W = sc->W, H = sc->H;
    const int tiles = ((W + BLK - 1) / BLK) * ((H + BLK - 1) / BLK);
    up3d_oracle_geom g;
    geom_alloc(&g, P > 0 ? P : 1);
    up3d_oracle_preprocess(sc, &g);
    const int64_t L = up3d_oracle_num_rendered(P, g.tiles_touched);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(L > 0 ? L : 1));
    uint32_t *vals = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(L > 0 ? L : 1));
    uint32_t *ranges = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (size_t)tiles);
    up3d_oracle_bin(P, W, H, &g, keys, vals, ranges);
    float *final_T = (float *)malloc(sizeof(float) * (size_t)W * H);
    uint32_t *n_contrib = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)W * H);
    up3d_oracle_blend_forward(W, H, sc->bg, &g, vals, ranges, out_color, final_T, n_contrib);
    if (radii_out) memcpy(radii_out, g.radii, sizeof(int32_t) * (size_t)P);
    if (dL_dpix) {
        float *G2 = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1) * 2);
        float *Gc = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1) * 3);
        float *Go = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1));
        float *Gr = (float *)malloc(sizeof(float) * (size_t)(P > 0 ? P : 1) * 3);
        up3d_oracle_blend_backward(P, W, H, sc->bg, &g, vals, ranges, final_T, n_contrib, dL_dpix, G2, Gc, Go, Gr);
        up3d_oracle_preprocess_backward(sc, &g, G2, Gc, Go, Gr, dL_dmeans3D, dL_dmeans2D, dL_dsh, dL_dcolors,
                                        dL_dopacity, dL_dscales, dL_drot);
        free(G2); free(Gc); free(Go); free(Gr);
    }
    free(keys); free(vals); free(ranges); free(final_T); free(n_contrib);
    geom_free(&g);
    return L;
}

size_t up3d_oracle_sizeof_scene(void) { return sizeof(up3d_oracle_scene); }
size_t up3d_oracle_sizeof_geom(void) { return sizeof(up3d_oracle_geom); }
