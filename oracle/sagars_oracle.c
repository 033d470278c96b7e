/*
 * sagars_oracle.c -- CPU restatement of the reference rasterizer's algorithm.   TEST INFRASTRUCTURE.
 *
 * This file is the parity oracle and the `cpu_baseline` leg of bench.py.  Nothing under
 * seganygaussians_b200/ may include, link or call it: only tests/, __graft_entry__.smoke() and
 * bench.py do, and only as the checker / the timed CPU baseline.
 *
 * The reference has NO CPU implementation of this path (its only implementation is the CUDA extension),
 * so this is a plain-C restatement of that CUDA code, function by function.  Citations are paths
 * relative to /root/reference/submodules/diff-gaussian-rasterization_contrastive_f/ (CF; BASE is the
 * same source at 3 channels) and .../diff-gaussian-rasterization-depth/ (DEPTH).
 *
 * PINNING: the reference ships no golden vectors or tests for this path.  The oracle is pinned against
 * outputs of the reference itself: tests/golden/*.npz were produced by running the unmodified reference
 * CUDA extension (rebuilt for sm_100a by oracle/build_ref.py) on a B200 with tests/golden/make_golden.py;
 * tests/test_oracle_golden.py checks this file against them on every CPU run.
 *
 * Arithmetic: fp32 throughout, compiled with -ffp-contract=off; every place where nvcc fuses a multiply
 * and an add in the reference's expression trees is written as an explicit fmaf() (nvcc/LLVM contraction
 * rule: in `a*b + c*d` the LEFT product is fused, fma(a, b, c*d)).  expf() is the host libm's, which may
 * differ from CUDA's by an ulp, so alpha thresholds can in principle flip on a measure-zero set of pairs.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define BLOCK_X 16
#define BLOCK_Y 16

typedef struct {
    int P, D, M, C, W, H;
    float tan_fovx, tan_fovy, scale_modifier;
    int has_mask_depth;
    const float *bg, *means3D, *shs, *colors_precomp, *opacities, *mask, *scales, *rotations, *cov3D_precomp;
    const float *view, *proj, *campos;
} orc_in;

/* ---- column-major 3x3 with glm's product order (third_party/glm/glm/detail/type_mat3x3.inl:486-518) ---- */
typedef struct { float c[3][3]; } m3;

static m3 m3_cols(float a0, float a1, float a2, float b0, float b1, float b2, float c0, float c1, float c2)
{
    m3 m;
    m.c[0][0] = a0; m.c[0][1] = a1; m.c[0][2] = a2;
    m.c[1][0] = b0; m.c[1][1] = b1; m.c[1][2] = b2;
    m.c[2][0] = c0; m.c[2][1] = c1; m.c[2][2] = c2;
    return m;
}
/* R[c][r] = A[0][r]*B[c][0] + A[1][r]*B[c][1] + A[2][r]*B[c][2]  ->  fma(A2,B2, fma(A0,B0, A1*B1)) */
static m3 m3_mul(const m3* A, const m3* B)
{
    m3 R;
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++)
            R.c[c][r] = fmaf(A->c[2][r], B->c[c][2], fmaf(A->c[0][r], B->c[c][0], A->c[1][r] * B->c[c][1]));
    return R;
}
static m3 m3_t(const m3* A)
{
    m3 R;
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) R.c[c][r] = A->c[r][c];
    return R;
}

/* auxiliary.h:58-77 : m0*x + m4*y + m8*z + m12  ->  fma(m8,z, fma(m0,x, m4*y)) + m12 */
static float row_affine(const float* m, int r, float x, float y, float z)
{
    return fmaf(m[8 + r], z, fmaf(m[r], x, m[4 + r] * y)) + m[12 + r];
}
/* auxiliary.h:41-44, evaluated in double */
static float ndc2pix(float v, int S) { return (float)((((double)v + 1.0) * (double)S - 1.0) * 0.5); }

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* auxiliary.h:46-56 */
static void get_rect(float px, float py, int max_radius, int gx, int gy, int* x0, int* y0, int* x1, int* y1)
{
    *x0 = imin(gx, imax(0, (int)((px - max_radius) / BLOCK_X)));
    *y0 = imin(gy, imax(0, (int)((py - max_radius) / BLOCK_Y)));
    *x1 = imin(gx, imax(0, (int)((px + max_radius + BLOCK_X - 1) / BLOCK_X)));
    *y1 = imin(gy, imax(0, (int)((py + max_radius + BLOCK_Y - 1) / BLOCK_Y)));
}

static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f, -1.0925484305920792f,
                               0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f, 0.3731763325901154f,
                               -0.4570457994644658f, 1.445305721320277f, -0.5900435899266435f};

/* forward.cu:23-74 (per channel; plain fp32, colours only need 1e-4) */
static void sh_to_rgb(const orc_in* in, int idx, float* rgb, uint8_t* clamped)
{
    const float* p = in->means3D + 3 * idx;
    float dx = p[0] - in->campos[0], dy = p[1] - in->campos[1], dz = p[2] - in->campos[2];
    float len = sqrtf(fmaf(dz, dz, fmaf(dx, dx, dy * dy)));
    float x = dx / len, y = dy / len, z = dz / len;
    const float* sh = in->shs + (size_t)idx * in->M * 3;
    for (int c = 0; c < 3; c++) {
#define S(k) sh[(k) * 3 + c]
        float r = SH_C0 * S(0);
        if (in->D > 0) {
            r = r - SH_C1 * y * S(1) + SH_C1 * z * S(2) - SH_C1 * x * S(3);
            if (in->D > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * S(4) + SH_C2[1] * yz * S(5) + SH_C2[2] * (2.0f * zz - xx - yy) * S(6) +
                    SH_C2[3] * xz * S(7) + SH_C2[4] * (xx - yy) * S(8);
                if (in->D > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * S(9) + SH_C3[1] * xy * z * S(10) +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * S(11) + SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * S(12) +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * S(13) + SH_C3[5] * z * (xx - yy) * S(14) +
                        SH_C3[6] * x * (xx - 3.0f * yy) * S(15);
                }
            }
        }
#undef S
        r += 0.5f;
        clamped[3 * idx + c] = (r < 0);
        rgb[3 * idx + c] = r > 0.0f ? r : 0.0f;
    }
}

/* forward.cu:121-155 */
static void cov3d(const float* scale, float mod, const float* rot, float* out6)
{
    m3 S = m3_cols(1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f);
    S.c[0][0] = mod * scale[0];
    S.c[1][1] = mod * scale[1];
    S.c[2][2] = mod * scale[2];
    float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
    /* 1 - 2*(y*y + z*z) -> fma(-2, fma(y,y, z*z), 1);  2*(x*y - r*z) -> 2 * fma(x,y, -(r*z)) */
    m3 R = m3_cols(fmaf(-2.f, fmaf(y, y, z * z), 1.f), 2.f * fmaf(x, y, -(r * z)), 2.f * fmaf(x, z, r * y),
                   2.f * fmaf(x, y, r * z), fmaf(-2.f, fmaf(x, x, z * z), 1.f), 2.f * fmaf(y, z, -(r * x)),
                   2.f * fmaf(x, z, -(r * y)), 2.f * fmaf(y, z, r * x), fmaf(-2.f, fmaf(x, x, y * y), 1.f));
    m3 Mm = m3_mul(&S, &R);
    m3 Mt = m3_t(&Mm);
    m3 Sg = m3_mul(&Mt, &Mm);
    out6[0] = Sg.c[0][0]; out6[1] = Sg.c[0][1]; out6[2] = Sg.c[0][2];
    out6[3] = Sg.c[1][1]; out6[4] = Sg.c[1][2]; out6[5] = Sg.c[2][2];
}

/* forward.cu:77-116 */
static void cov2d(const orc_in* in, const float* mean, float fx, float fy, const float* c3, float* a, float* b, float* c,
                  m3* T_out, float* t_out, float* txtz_o, float* tytz_o)
{
    const float* v = in->view;
    float t[3] = {row_affine(v, 0, mean[0], mean[1], mean[2]), row_affine(v, 1, mean[0], mean[1], mean[2]),
                  row_affine(v, 2, mean[0], mean[1], mean[2])};
    float limx = 1.3f * in->tan_fovx, limy = 1.3f * in->tan_fovy;
    float txtz = t[0] / t[2], tytz = t[1] / t[2];
    t[0] = fminf(limx, fmaxf(-limx, txtz)) * t[2];
    t[1] = fminf(limy, fmaxf(-limy, tytz)) * t[2];
    m3 J = m3_cols(fx / t[2], 0.0f, -(fx * t[0]) / (t[2] * t[2]), 0.0f, fy / t[2], -(fy * t[1]) / (t[2] * t[2]), 0.f, 0.f, 0.f);
    m3 Wm = m3_cols(v[0], v[4], v[8], v[1], v[5], v[9], v[2], v[6], v[10]);
    m3 T = m3_mul(&Wm, &J);
    m3 Vrk = m3_cols(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
    m3 Tt = m3_t(&T), Vt = m3_t(&Vrk);
    m3 TV = m3_mul(&Tt, &Vt);
    m3 cov = m3_mul(&TV, &T);
    *a = cov.c[0][0] + 0.3f;
    *b = cov.c[0][1];
    *c = cov.c[1][1] + 0.3f;
    if (T_out) *T_out = T;
    if (t_out) { t_out[0] = t[0]; t_out[1] = t[1]; t_out[2] = t[2]; }
    if (txtz_o) *txtz_o = txtz;
    if (tytz_o) *tytz_o = tytz;
}

/* forward.cu:158-259 + rasterizer_impl.cu:277 (inclusive scan).  Returns num_rendered. */
int orc_preprocess(const orc_in* in, int32_t* radii, float* means2D, float* depths, float* cov3D, float* conic_opacity,
                   float* rgb, uint8_t* clamped, uint32_t* tiles_touched, uint32_t* point_offsets)
{
    const int P = in->P;
    const float focal_y = in->H / (2.0f * in->tan_fovy);
    const float focal_x = in->W / (2.0f * in->tan_fovx);
    const int gx = (in->W + BLOCK_X - 1) / BLOCK_X, gy = (in->H + BLOCK_Y - 1) / BLOCK_Y;
    for (int idx = 0; idx < P; idx++) {
        radii[idx] = 0;
        tiles_touched[idx] = 0;
        const float* p = in->means3D + 3 * idx;
        float hx = row_affine(in->proj, 0, p[0], p[1], p[2]);
        float hy = row_affine(in->proj, 1, p[0], p[1], p[2]);
        float hw = row_affine(in->proj, 3, p[0], p[1], p[2]);
        float p_w = 1.0f / (hw + 0.0000001f);
        float projx = hx * p_w, projy = hy * p_w;
        float view_z = row_affine(in->view, 2, p[0], p[1], p[2]);
        if (view_z <= 0.2f) continue;
        const float* c3;
        if (in->cov3D_precomp) {
            c3 = in->cov3D_precomp + 6 * idx;
        } else {
            cov3d(in->scales + 3 * idx, in->scale_modifier, in->rotations + 4 * idx, cov3D + 6 * idx);
            c3 = cov3D + 6 * idx;
        }
        float a, b, c;
        cov2d(in, p, focal_x, focal_y, c3, &a, &b, &c, NULL, NULL, NULL, NULL);
        float det = fmaf(a, c, -(b * b));
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float conx = c * det_inv, cony = -b * det_inv, conz = a * det_inv;
        float mid = 0.5f * (a + c);
        float disc = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float lambda1 = mid + disc, lambda2 = mid - disc;
        float my_radius = ceilf(3.f * sqrtf(fmaxf(lambda1, lambda2)));
        float pix_x = ndc2pix(projx, in->W), pix_y = ndc2pix(projy, in->H);
        int x0, y0, x1, y1;
        get_rect(pix_x, pix_y, (int)my_radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (in->colors_precomp == NULL && in->shs != NULL) sh_to_rgb(in, idx, rgb, clamped);
        depths[idx] = view_z;
        radii[idx] = (int32_t)my_radius;
        means2D[2 * idx] = pix_x;
        means2D[2 * idx + 1] = pix_y;
        conic_opacity[4 * idx] = conx;
        conic_opacity[4 * idx + 1] = cony;
        conic_opacity[4 * idx + 2] = conz;
        conic_opacity[4 * idx + 3] = in->opacities[idx];
        tiles_touched[idx] = (uint32_t)((y1 - y0) * (x1 - x0));
    }
    uint32_t run = 0;
    for (int i = 0; i < P; i++) {
        run += tiles_touched[i];
        point_offsets[i] = run;
    }
    return (int)run;
}

/* rasterizer_impl.cu:35-50 */
static uint32_t higher_msb(uint32_t n)
{
    uint32_t msb = sizeof(n) * 4, step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

/* rasterizer_impl.cu:70-138,289-317: keys, stable sort on bits [0, 32+msb(tiles)), tile ranges */
void orc_bin(const orc_in* in, const int32_t* radii, const float* means2D, const float* depths, const uint32_t* point_offsets,
             int R, uint64_t* keys_sorted, uint32_t* point_list, uint32_t* ranges /* [tiles][2] */)
{
    const int gx = (in->W + BLOCK_X - 1) / BLOCK_X, gy = (in->H + BLOCK_Y - 1) / BLOCK_Y;
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)gx * gy);
    if (R <= 0) return;
    uint64_t* k0 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
    uint32_t* v0 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
    for (int idx = 0; idx < in->P; idx++) {
        if (radii[idx] > 0) {
            uint32_t off = (idx == 0) ? 0 : point_offsets[idx - 1];
            int x0, y0, x1, y1;
            get_rect(means2D[2 * idx], means2D[2 * idx + 1], radii[idx], gx, gy, &x0, &y0, &x1, &y1);
            uint32_t dbits;
            memcpy(&dbits, &depths[idx], 4);
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) {
                    uint64_t key = (uint64_t)(y * gx + x);
                    key <<= 32;
                    key |= dbits;
                    k0[off] = key;
                    v0[off] = (uint32_t)idx;
                    off++;
                }
        }
    }
    /* stable LSD radix sort, 8-bit digits, on the low end_bit bits */
    const int end_bit = 32 + (int)higher_msb((uint32_t)(gx * gy));
    uint64_t* k1 = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)R);
    uint32_t* v1 = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)R);
    uint64_t *ka = k0, *kb = k1;
    uint32_t *va = v0, *vb = v1;
    for (int shift = 0; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint32_t msk = (1u << bits) - 1u;
        size_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int i = 0; i < R; i++) cnt[((ka[i] >> shift) & msk) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int i = 0; i < R; i++) {
            size_t pos = cnt[(ka[i] >> shift) & msk]++;
            kb[pos] = ka[i];
            vb[pos] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    memcpy(keys_sorted, ka, sizeof(uint64_t) * (size_t)R);
    memcpy(point_list, va, sizeof(uint32_t) * (size_t)R);
    free(k0); free(v0); free(k1); free(v1);
    for (int i = 0; i < R; i++) {
        uint32_t cur = (uint32_t)(keys_sorted[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(keys_sorted[i - 1] >> 32);
            if (cur != prev) {
                ranges[2 * prev + 1] = (uint32_t)i;
                ranges[2 * cur] = (uint32_t)i;
            }
        }
        if (i == R - 1) ranges[2 * cur + 1] = (uint32_t)R;
    }
}

/* power of forward.cu:339 / backward.cu:495 with nvcc's contraction */
static inline float pair_power(float cx, float cy, float cz, float dx, float dy)
{
    float s = fmaf(cx * dx, dx, (cz * dy) * dy);
    return fmaf(-0.5f, s, -((cy * dx) * dy));
}

/* forward.cu:264-385 (+ DEPTH forward.cu:364-365,384-385) */
void orc_render_forward(const orc_in* in, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                        const float* conic_opacity, const float* features, const float* depths, float* final_T,
                        uint32_t* n_contrib, float* out_color, float* out_mask, float* out_depth, int nthreads)
{
    const int W = in->W, H = in->H, C = in->C;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    (void)nthreads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        float Cacc[64];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                const float pfx = (float)px, pfy = (float)py;
                float T = 1.0f, Macc = 0.f, Dacc = 0.f;
                uint32_t contributor = 0, last_contributor = 0;
                for (int ch = 0; ch < C; ch++) Cacc[ch] = 0.f;
                for (uint32_t i = r0; i < r1; i++) {
                    contributor++;
                    const uint32_t id = point_list[i];
                    const float dx = means2D[2 * id] - pfx, dy = means2D[2 * id + 1] - pfy;
                    const float* co = conic_opacity + 4 * (size_t)id;
                    const float power = pair_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    const float alpha = fminf(0.99f, co[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    const float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;
                    for (int ch = 0; ch < C; ch++) Cacc[ch] = fmaf(features[(size_t)id * C + ch] * alpha, T, Cacc[ch]);
                    if (in->has_mask_depth) {
                        Macc = fmaf(in->mask[id] * alpha, T, Macc);
                        Dacc = fmaf(depths[id] * alpha, T, Dacc);
                    }
                    T = test_T;
                    last_contributor = contributor;
                }
                const size_t pix = (size_t)W * py + px;
                final_T[pix] = T;
                n_contrib[pix] = last_contributor;
                for (int ch = 0; ch < C; ch++) out_color[(size_t)ch * H * W + pix] = fmaf(T, in->bg[ch], Cacc[ch]);
                if (in->has_mask_depth) {
                    out_mask[pix] = Macc;
                    out_depth[pix] = Dacc;
                }
            }
    }
}

static inline void acc_add(float* p, float v, int par)
{
    if (par) {
#ifdef _OPENMP
#pragma omp atomic
#endif
        *p += v;
    } else {
        *p += v;
    }
}

/* backward.cu:399-559 (+ DEPTH backward.cu:457,516).  dL_dmean2D is [P][3], dL_dconic [P][4] (x, y, -, w). */
void orc_render_backward(const orc_in* in, const uint32_t* ranges, const uint32_t* point_list, const float* means2D,
                         const float* conic_opacity, const float* colors, const float* final_Ts, const uint32_t* n_contrib,
                         const float* dL_dpixels, const float* dL_dout_mask, float* dL_dmean2D, float* dL_dconic2D,
                         float* dL_dopacity, float* dL_dcolors, float* dL_dmask, int nthreads)
{
    const int W = in->W, H = in->H, C = in->C;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const int par = nthreads > 1;
    const float ddelx_dx = (float)(0.5 * W), ddely_dy = (float)(0.5 * H);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads > 0 ? nthreads : 1)
#endif
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        float accum_rec[64], last_color[64], dL_dpixel[64];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int px = tx * BLOCK_X + lx, py = ty * BLOCK_Y + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)W * py + px;
                const float pfx = (float)px, pfy = (float)py;
                const float T_final = final_Ts[pix];
                float T = T_final;
                const uint32_t last_contributor = n_contrib[pix];
                float last_alpha = 0.f;
                float bg_dot_dpixel = 0.f;
                for (int ch = 0; ch < C; ch++) {
                    accum_rec[ch] = 0.f;
                    last_color[ch] = 0.f;
                    dL_dpixel[ch] = dL_dpixels[(size_t)ch * H * W + pix];
                    bg_dot_dpixel = fmaf(in->bg[ch], dL_dpixel[ch], bg_dot_dpixel);
                }
                const float dmask_i = in->has_mask_depth ? dL_dout_mask[pix] : 0.f;
                /* back to front: position `contributor` (0-based) is skipped while >= last_contributor */
                for (uint32_t pos = last_contributor; pos-- > 0;) {
                    const uint32_t id = point_list[r0 + pos];
                    const float dx = means2D[2 * id] - pfx, dy = means2D[2 * id + 1] - pfy;
                    const float* co = conic_opacity + 4 * (size_t)id;
                    const float power = pair_power(co[0], co[1], co[2], dx, dy);
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < C; ch++) {
                        const float c = colors[(size_t)id * C + ch];
                        accum_rec[ch] = fmaf(last_alpha, last_color[ch], (1.f - last_alpha) * accum_rec[ch]);
                        last_color[ch] = c;
                        const float dL_dchannel = dL_dpixel[ch];
                        dL_dalpha = fmaf(c - accum_rec[ch], dL_dchannel, dL_dalpha);
                        acc_add(&dL_dcolors[(size_t)id * C + ch], dchannel_dcolor * dL_dchannel, par);
                    }
                    if (in->has_mask_depth) acc_add(&dL_dmask[id], dchannel_dcolor * dmask_i, par);
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha = fmaf(-T_final / (1.f - alpha), bg_dot_dpixel, dL_dalpha);
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = fmaf(-gdx, co[0], -(gdy * co[1]));
                    const float dG_ddely = fmaf(-gdy, co[2], -(gdx * co[1]));
                    acc_add(&dL_dmean2D[3 * (size_t)id + 0], dL_dG * dG_ddelx * ddelx_dx, par);
                    acc_add(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy, par);
                    acc_add(&dL_dconic2D[4 * (size_t)id + 0], -0.5f * gdx * dx * dL_dG, par);
                    acc_add(&dL_dconic2D[4 * (size_t)id + 1], -0.5f * gdx * dy * dL_dG, par);
                    acc_add(&dL_dconic2D[4 * (size_t)id + 3], -0.5f * gdy * dy * dL_dG, par);
                    acc_add(&dL_dopacity[id], G * dL_dalpha, par);
                }
                (void)r1;
            }
    }
}

static void dnormvdv3(const float* v, const float* dv, float* out)
{
    float sum2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    float inv = 1.0f / sqrtf(sum2 * sum2 * sum2);
    out[0] = ((+sum2 - v[0] * v[0]) * dv[0] - v[1] * v[0] * dv[1] - v[2] * v[0] * dv[2]) * inv;
    out[1] = (-v[0] * v[1] * dv[0] + (sum2 - v[1] * v[1]) * dv[1] - v[2] * v[1] * dv[2]) * inv;
    out[2] = (-v[0] * v[2] * dv[0] - v[1] * v[2] * dv[1] + (sum2 - v[2] * v[2]) * dv[2]) * inv;
}

/* backward.cu:20-139 */
static void sh_backward(const orc_in* in, int idx, const uint8_t* clamped, const float* dL_dcolor, float* dL_dmeans, float* dL_dshs)
{
    const float* p = in->means3D + 3 * idx;
    float dorig[3] = {p[0] - in->campos[0], p[1] - in->campos[1], p[2] - in->campos[2]};
    float len = sqrtf(dorig[0] * dorig[0] + dorig[1] * dorig[1] + dorig[2] * dorig[2]);
    float x = dorig[0] / len, y = dorig[1] / len, z = dorig[2] / len;
    const float* sh = in->shs + (size_t)idx * in->M * 3;
    float* out = dL_dshs + (size_t)idx * in->M * 3;
    float ddir[3] = {0, 0, 0};
    for (int c = 0; c < 3; c++) {
        float dL = dL_dcolor[3 * idx + c] * (clamped[3 * idx + c] ? 0.f : 1.f);
        float dRx = 0, dRy = 0, dRz = 0;
#define S(k) sh[(k) * 3 + c]
#define O(k) out[(k) * 3 + c]
        O(0) = SH_C0 * dL;
        if (in->D > 0) {
            O(1) = (-SH_C1 * y) * dL; O(2) = (SH_C1 * z) * dL; O(3) = (-SH_C1 * x) * dL;
            dRx = -SH_C1 * S(3); dRy = -SH_C1 * S(1); dRz = SH_C1 * S(2);
            if (in->D > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                O(4) = (SH_C2[0] * xy) * dL; O(5) = (SH_C2[1] * yz) * dL; O(6) = (SH_C2[2] * (2.f * zz - xx - yy)) * dL;
                O(7) = (SH_C2[3] * xz) * dL; O(8) = (SH_C2[4] * (xx - yy)) * dL;
                dRx += SH_C2[0] * y * S(4) + SH_C2[2] * 2.f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.f * x * S(8);
                dRy += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.f * -y * S(6) + SH_C2[4] * 2.f * -y * S(8);
                dRz += SH_C2[1] * y * S(5) + SH_C2[2] * 2.f * 2.f * z * S(6) + SH_C2[3] * x * S(7);
                if (in->D > 2) {
                    O(9) = (SH_C3[0] * y * (3.f * xx - yy)) * dL; O(10) = (SH_C3[1] * xy * z) * dL;
                    O(11) = (SH_C3[2] * y * (4.f * zz - xx - yy)) * dL; O(12) = (SH_C3[3] * z * (2.f * zz - 3.f * xx - 3.f * yy)) * dL;
                    O(13) = (SH_C3[4] * x * (4.f * zz - xx - yy)) * dL; O(14) = (SH_C3[5] * z * (xx - yy)) * dL;
                    O(15) = (SH_C3[6] * x * (xx - 3.f * yy)) * dL;
                    dRx += (SH_C3[0] * S(9) * 3.f * 2.f * xy + SH_C3[1] * S(10) * yz + SH_C3[2] * S(11) * -2.f * xy +
                            SH_C3[3] * S(12) * -3.f * 2.f * xz + SH_C3[4] * S(13) * (-3.f * xx + 4.f * zz - yy) +
                            SH_C3[5] * S(14) * 2.f * xz + SH_C3[6] * S(15) * 3.f * (xx - yy));
                    dRy += (SH_C3[0] * S(9) * 3.f * (xx - yy) + SH_C3[1] * S(10) * xz + SH_C3[2] * S(11) * (-3.f * yy + 4.f * zz - xx) +
                            SH_C3[3] * S(12) * -3.f * 2.f * yz + SH_C3[4] * S(13) * -2.f * xy + SH_C3[5] * S(14) * -2.f * yz +
                            SH_C3[6] * S(15) * -3.f * 2.f * xy);
                    dRz += (SH_C3[1] * S(10) * xy + SH_C3[2] * S(11) * 4.f * 2.f * yz + SH_C3[3] * S(12) * 3.f * (2.f * zz - xx - yy) +
                            SH_C3[4] * S(13) * 4.f * 2.f * xz + SH_C3[5] * S(14) * (xx - yy));
                }
            }
        }
#undef S
#undef O
        ddir[0] += dRx * dL; ddir[1] += dRy * dL; ddir[2] += dRz * dL;
    }
    float dm[3];
    dnormvdv3(dorig, ddir, dm);
    dL_dmeans[3 * idx] += dm[0]; dL_dmeans[3 * idx + 1] += dm[1]; dL_dmeans[3 * idx + 2] += dm[2];
}

/* backward.cu:144-274 (computeCov2DCUDA), :346-396 (preprocessCUDA), :278-341 (computeCov3D).
 * All outputs must be zero-initialised by the caller (the reference uses torch::zeros). */
void orc_geom_backward(const orc_in* in, const int32_t* radii, const float* cov3Ds, const uint8_t* clamped,
                       const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor, float* dL_dmeans3D,
                       float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots)
{
    const float h_y = in->H / (2.0f * in->tan_fovy);
    const float h_x = in->W / (2.0f * in->tan_fovx);
    const float* view = in->view;
    const float* proj = in->proj;
    for (int idx = 0; idx < in->P; idx++) {
        if (!(radii[idx] > 0)) continue;
        const float* c3 = cov3Ds + 6 * idx;
        const float* mean = in->means3D + 3 * idx;
        float dcx = dL_dconic[4 * idx], dcy = dL_dconic[4 * idx + 1], dcz = dL_dconic[4 * idx + 3];
        float a, b, c, t[3], txtz, tytz;
        m3 T;
        cov2d(in, mean, h_x, h_y, c3, &a, &b, &c, &T, t, &txtz, &tytz);
        const float limx = 1.3f * in->tan_fovx, limy = 1.3f * in->tan_fovy;
        const float x_grad_mul = (txtz < -limx || txtz > limx) ? 0.f : 1.f;
        const float y_grad_mul = (tytz < -limy || tytz > limy) ? 0.f : 1.f;
        m3 Wm = m3_cols(view[0], view[4], view[8], view[1], view[5], view[9], view[2], view[6], view[10]);
        m3 Vrk = m3_cols(c3[0], c3[1], c3[2], c3[1], c3[3], c3[4], c3[2], c3[4], c3[5]);
        float denom = a * c - b * b;
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
        float* dcov = dL_dcov3D + 6 * idx;
        if (denom2inv != 0) {
            dL_da = denom2inv * (-c * c * dcx + 2 * b * c * dcy + (denom - a * c) * dcz);
            dL_dc = denom2inv * (-a * a * dcz + 2 * a * b * dcy + (denom - a * c) * dcx);
            dL_db = denom2inv * 2 * (b * c * dcx - (denom + 2 * b * b) * dcy + a * b * dcz);
            dcov[0] = (T.c[0][0] * T.c[0][0] * dL_da + T.c[0][0] * T.c[1][0] * dL_db + T.c[1][0] * T.c[1][0] * dL_dc);
            dcov[3] = (T.c[0][1] * T.c[0][1] * dL_da + T.c[0][1] * T.c[1][1] * dL_db + T.c[1][1] * T.c[1][1] * dL_dc);
            dcov[5] = (T.c[0][2] * T.c[0][2] * dL_da + T.c[0][2] * T.c[1][2] * dL_db + T.c[1][2] * T.c[1][2] * dL_dc);
            dcov[1] = 2 * T.c[0][0] * T.c[0][1] * dL_da + (T.c[0][0] * T.c[1][1] + T.c[0][1] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][1] * dL_dc;
            dcov[2] = 2 * T.c[0][0] * T.c[0][2] * dL_da + (T.c[0][0] * T.c[1][2] + T.c[0][2] * T.c[1][0]) * dL_db + 2 * T.c[1][0] * T.c[1][2] * dL_dc;
            dcov[4] = 2 * T.c[0][2] * T.c[0][1] * dL_da + (T.c[0][1] * T.c[1][2] + T.c[0][2] * T.c[1][1]) * dL_db + 2 * T.c[1][1] * T.c[1][2] * dL_dc;
        } else {
            for (int i = 0; i < 6; i++) dcov[i] = 0;
        }
        float dL_dT00 = 2 * (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_db;
        float dL_dT01 = 2 * (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_db;
        float dL_dT02 = 2 * (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_da +
                        (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_db;
        float dL_dT10 = 2 * (T.c[1][0] * Vrk.c[0][0] + T.c[1][1] * Vrk.c[0][1] + T.c[1][2] * Vrk.c[0][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[0][0] + T.c[0][1] * Vrk.c[0][1] + T.c[0][2] * Vrk.c[0][2]) * dL_db;
        float dL_dT11 = 2 * (T.c[1][0] * Vrk.c[1][0] + T.c[1][1] * Vrk.c[1][1] + T.c[1][2] * Vrk.c[1][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[1][0] + T.c[0][1] * Vrk.c[1][1] + T.c[0][2] * Vrk.c[1][2]) * dL_db;
        float dL_dT12 = 2 * (T.c[1][0] * Vrk.c[2][0] + T.c[1][1] * Vrk.c[2][1] + T.c[1][2] * Vrk.c[2][2]) * dL_dc +
                        (T.c[0][0] * Vrk.c[2][0] + T.c[0][1] * Vrk.c[2][1] + T.c[0][2] * Vrk.c[2][2]) * dL_db;
        float dL_dJ00 = Wm.c[0][0] * dL_dT00 + Wm.c[0][1] * dL_dT01 + Wm.c[0][2] * dL_dT02;
        float dL_dJ02 = Wm.c[2][0] * dL_dT00 + Wm.c[2][1] * dL_dT01 + Wm.c[2][2] * dL_dT02;
        float dL_dJ11 = Wm.c[1][0] * dL_dT10 + Wm.c[1][1] * dL_dT11 + Wm.c[1][2] * dL_dT12;
        float dL_dJ12 = Wm.c[2][0] * dL_dT10 + Wm.c[2][1] * dL_dT11 + Wm.c[2][2] * dL_dT12;
        float tz = 1.f / t[2], tz2 = tz * tz, tz3 = tz2 * tz;
        float dL_dtx = x_grad_mul * -h_x * tz2 * dL_dJ02;
        float dL_dty = y_grad_mul * -h_y * tz2 * dL_dJ12;
        float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t[0]) * tz3 * dL_dJ02 + (2 * h_y * t[1]) * tz3 * dL_dJ12;
        float* dm = dL_dmeans3D + 3 * idx;
        dm[0] = view[0] * dL_dtx + view[1] * dL_dty + view[2] * dL_dtz;
        dm[1] = view[4] * dL_dtx + view[5] * dL_dty + view[6] * dL_dtz;
        dm[2] = view[8] * dL_dtx + view[9] * dL_dty + view[10] * dL_dtz;

        /* preprocessCUDA (backward.cu:346-396) */
        float hw = proj[3] * mean[0] + proj[7] * mean[1] + proj[11] * mean[2] + proj[15];
        float m_w = 1.0f / (hw + 0.0000001f);
        float mul1 = (proj[0] * mean[0] + proj[4] * mean[1] + proj[8] * mean[2] + proj[12]) * m_w * m_w;
        float mul2 = (proj[1] * mean[0] + proj[5] * mean[1] + proj[9] * mean[2] + proj[13]) * m_w * m_w;
        float gx2 = dL_dmean2D[3 * idx], gy2 = dL_dmean2D[3 * idx + 1];
        dm[0] += (proj[0] * m_w - proj[3] * mul1) * gx2 + (proj[1] * m_w - proj[3] * mul2) * gy2;
        dm[1] += (proj[4] * m_w - proj[7] * mul1) * gx2 + (proj[5] * m_w - proj[7] * mul2) * gy2;
        dm[2] += (proj[8] * m_w - proj[11] * mul1) * gx2 + (proj[9] * m_w - proj[11] * mul2) * gy2;
        if (in->shs) sh_backward(in, idx, clamped, dL_dcolor, dL_dmeans3D, dL_dsh);
        if (in->scales) {
            const float* rot = in->rotations + 4 * idx;
            float r = rot[0], x = rot[1], y = rot[2], z = rot[3];
            m3 R = m3_cols(1.f - 2.f * (y * y + z * z), 2.f * (x * y - r * z), 2.f * (x * z + r * y),
                           2.f * (x * y + r * z), 1.f - 2.f * (x * x + z * z), 2.f * (y * z - r * x),
                           2.f * (x * z - r * y), 2.f * (y * z + r * x), 1.f - 2.f * (x * x + y * y));
            float s[3] = {in->scale_modifier * in->scales[3 * idx], in->scale_modifier * in->scales[3 * idx + 1],
                          in->scale_modifier * in->scales[3 * idx + 2]};
            m3 S = m3_cols(s[0], 0, 0, 0, s[1], 0, 0, 0, s[2]);
            m3 Mm = m3_mul(&S, &R);
            m3 dSig = m3_cols(dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3], 0.5f * dcov[4],
                              0.5f * dcov[2], 0.5f * dcov[4], dcov[5]);
            m3 M2;
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M2.c[i][j] = 2.0f * Mm.c[i][j];
            m3 dM = m3_mul(&M2, &dSig);
            m3 Rt = m3_t(&R);
            m3 dMt = m3_t(&dM);
            for (int k = 0; k < 3; k++)
                dL_dscales[3 * idx + k] = Rt.c[k][0] * dMt.c[k][0] + Rt.c[k][1] * dMt.c[k][1] + Rt.c[k][2] * dMt.c[k][2];
            for (int k = 0; k < 3; k++) for (int j = 0; j < 3; j++) dMt.c[k][j] *= s[k];
            float* dq = dL_drots + 4 * idx;
            dq[0] = 2 * z * (dMt.c[0][1] - dMt.c[1][0]) + 2 * y * (dMt.c[2][0] - dMt.c[0][2]) + 2 * x * (dMt.c[1][2] - dMt.c[2][1]);
            dq[1] = 2 * y * (dMt.c[1][0] + dMt.c[0][1]) + 2 * z * (dMt.c[2][0] + dMt.c[0][2]) + 2 * r * (dMt.c[1][2] - dMt.c[2][1]) - 4 * x * (dMt.c[2][2] + dMt.c[1][1]);
            dq[2] = 2 * x * (dMt.c[1][0] + dMt.c[0][1]) + 2 * r * (dMt.c[2][0] - dMt.c[0][2]) + 2 * z * (dMt.c[1][2] + dMt.c[2][1]) - 4 * y * (dMt.c[2][2] + dMt.c[0][0]);
            dq[3] = 2 * r * (dMt.c[0][1] - dMt.c[1][0]) + 2 * x * (dMt.c[2][0] + dMt.c[0][2]) + 2 * y * (dMt.c[1][2] + dMt.c[2][1]) - 4 * z * (dMt.c[1][1] + dMt.c[0][0]);
        }
    }
}

/* rasterizer_impl.cu:54-66 */
void orc_mark_visible(int P, const float* means3D, const float* view, uint8_t* present)
{
    for (int i = 0; i < P; i++) {
        const float* p = means3D + 3 * i;
        present[i] = !(row_affine(view, 2, p[0], p[1], p[2]) <= 0.2f);
    }
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
