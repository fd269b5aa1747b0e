// Crop / augment of rendered views on the device: what RenderedDataset.__getitem__ does per sample on the CPU with PIL
// in the reference (anakin/artiboost/rendered_dataset.py:127-133,155-274; utils/transform.py:425-470;
// utils/img_augment.py:6-80; datasets/hodata.py:161-186), batched and host-free.
//
//   augment_prelude_kernel   one thread per view: projection, crop box, centre / scale jitter, forward affine (fp64 in a
//                            fixed operation order, rounded to fp32), closed-form inverse, Pillow's 16.16 fixed-point
//                            warp coefficients (or its running-sum tables when there is no rotation), blur weights,
//                            and every transformed annotation (intrinsics, joints, corners, visibility, object pose)
//   augment_blur_kernel      one CTA per 32 x 32 tile of SOURCE pixels: GaussianBlur(radius <= 0.1) = 3 horizontal + 3 vertical
//                            box passes with Pillow's 8.24 fixed-point weights and byte rounding after every pass, the
//                            horizontal results staged in shared memory; also the luma sum the Contrast enhancer needs (after the colour
//                            operations that precede it in this view's shuffled order)
//   augment_warp_kernel      one thread per OUTPUT pixel: nearest-neighbour AFFINE warp (Pillow's rules), the colour
//                            operations in order (Brightness / Color / Contrast = Blend.c arithmetic, hue through
//                            Pillow's rgb2hsv / hsv2rgb), `to_tensor` and the -0.5 normalisation, fp32 NCHW out
//
// The byte arithmetic is specified by oracle/augment.py (pinned bit-for-bit against Pillow); this file must be compiled
// with -fmad=false so that every fp32 / fp64 operation is rounded on its own, as in the oracle.
#include <math.h>

#include "common.cuh"

namespace ab {

struct AugView {
    long long a[6];       // 16.16 fixed-point inverse affine (a2, a5 include the half-pixel terms)
    int scale_mode;       // 1: no rotation -> xs / ys tables (ImagingScaleAffine)
    int blur_on;
    unsigned ww, fw;      // box-blur weights
    float alpha[3];       // brightness, contrast, saturation
    int hue_shift;
    int order[4];         // execution order, entries index {brightness, saturation, hue, contrast}
    int n_before_contrast;
};

// ------------------------------------------------------------------------------------------ Pillow byte arithmetic
__device__ __forceinline__ int rgb_to_l(int r, int g, int b) { return (r * 19595 + g * 38470 + b * 7471 + 0x8000) >> 16; }

__device__ __forceinline__ int blend_byte(int deg, int v, float alpha) {
    const float d = (float)deg;
    const float t = d + alpha * ((float)v - d);
    if (alpha >= 0.0f && alpha <= 1.0f) return (int)t;
    if (t <= 0.0f) return 0;
    if (t >= 255.0f) return 255;
    return (int)t;
}

__device__ __forceinline__ int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

__device__ void rgb_to_hsv(int r, int g, int b, int& uh, int& us, int& uv) {
    const int maxc = max(r, max(g, b)), minc = min(r, min(g, b));
    uv = maxc;
    if (minc == maxc) { uh = 0; us = 0; return; }
    const float cr = (float)(maxc - minc);
    const float s = cr / (float)maxc;
    const float rc = (float)(maxc - r) / cr, gc = (float)(maxc - g) / cr, bc = (float)(maxc - b) / cr;
    float h;
    if (r == maxc) h = bc - gc;
    else if (g == maxc) h = (float)(2.0 + (double)rc - (double)bc);
    else h = (float)(4.0 + (double)gc - (double)rc);
    h = (float)fmod((double)h / 6.0 + 1.0, 1.0);
    uh = clip8((int)((double)h * 255.0));
    us = clip8((int)((double)s * 255.0));
}

__device__ void hsv_to_rgb(int h, int s, int v, int& r, int& g, int& b) {
    if (s == 0) { r = g = b = v; return; }
    const double hf = (double)(float)h * 6.0 / 255.0;
    const int i = (int)floor(hf);
    const float f = (float)(hf - (double)(float)i);
    const float fs = (float)((double)(float)s / 255.0);
    const double vf = (double)(float)v, f64 = (double)f, fs64 = (double)fs;
    const int p = clip8((int)floor(vf * (1.0 - fs64) + 0.5));
    const int q = clip8((int)floor(vf * (1.0 - fs64 * f64) + 0.5));
    const int t = clip8((int)floor(vf * (1.0 - fs64 * (1.0 - f64)) + 0.5));
    switch (i % 6) {
        case 0: r = v; g = t; b = p; break;
        case 1: r = q; g = v; b = p; break;
        case 2: r = p; g = v; b = t; break;
        case 3: r = p; g = q; b = v; break;
        case 4: r = t; g = p; b = v; break;
        default: r = v; g = p; b = q; break;
    }
}

// ops [first, last) of the view's execution order; `mean` is only read by the contrast op
__device__ void apply_ops(const AugView& vw, int first, int last, int mean, int& r, int& g, int& b) {
    for (int k = first; k < last; ++k) {
        const int op = vw.order[k];
        if (op == 0) {
            r = blend_byte(0, r, vw.alpha[0]); g = blend_byte(0, g, vw.alpha[0]); b = blend_byte(0, b, vw.alpha[0]);
        } else if (op == 1) {
            const int l = rgb_to_l(r, g, b);
            r = blend_byte(l, r, vw.alpha[2]); g = blend_byte(l, g, vw.alpha[2]); b = blend_byte(l, b, vw.alpha[2]);
        } else if (op == 2) {
            int h, s, v;
            rgb_to_hsv(r, g, b, h, s, v);
            hsv_to_rgb((h + vw.hue_shift) & 0xFF, s, v, r, g, b);
        } else {
            r = blend_byte(mean, r, vw.alpha[1]); g = blend_byte(mean, g, vw.alpha[1]); b = blend_byte(mean, b, vw.alpha[1]);
        }
    }
}

// --------------------------------------------------------------------------------------------------- prelude
struct AugOut {
    float *image, *cam_intr, *root_joint, *joints_3d, *joints_2d, *joints_vis, *corners_3d, *corners_2d, *corners_vis, *obj_transf,
        *affine, *inv_affine;
};

__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return (a0 * b0 + a1 * b1) + a2 * b2;
}

__device__ void affine_no_rot(double cx, double cy, double scale, int res0, int res1, double& m00, double& m11, double& m02,
                              double& m12) {
    const double ratio = (double)res0 / (double)res1;
    m00 = (double)res0 / scale;
    m11 = (double)res1 / scale * ratio;
    m02 = (double)res0 * (-cx / scale + 0.5);
    m12 = (double)res1 * (-cy / scale * ratio + 0.5);
}

__global__ void augment_prelude_kernel(const ab_augment_cfg cfg, int B, const float* __restrict__ joints,
                                       const float* __restrict__ obj_pose, const float* __restrict__ corners_can,
                                       const float* __restrict__ draws, const int* __restrict__ order, AugView* views,
                                       int* xs_tab, int* ys_tab, unsigned* lsum, int* status, AugOut out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= B) return;
    const float* J = joints + (size_t)v * 63;
    const float* Pm = obj_pose + (size_t)v * 16;
    const float* CC = corners_can + (size_t)v * 24;
    const float* dr = draws + (size_t)v * AB_AUG_DRAWS;
    double K[9];
    for (int i = 0; i < 9; ++i) K[i] = (double)cfg.K[i];
    lsum[v] = 0u;

    // ---- projections (rendered_dataset.py:127-133)
    double c3[8][3], j2[21][2], c2[8][2];
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 3; ++i)
            c3[k][i] = dot3((double)Pm[4 * i], (double)Pm[4 * i + 1], (double)Pm[4 * i + 2], (double)CC[3 * k], (double)CC[3 * k + 1],
                            (double)CC[3 * k + 2]) + (double)Pm[4 * i + 3];
    for (int k = 0; k < 21; ++k) {
        const double x = J[3 * k], y = J[3 * k + 1], z = J[3 * k + 2];
        const double hz = dot3(K[6], K[7], K[8], x, y, z) + 1e-8;
        j2[k][0] = dot3(K[0], K[1], K[2], x, y, z) / hz;
        j2[k][1] = dot3(K[3], K[4], K[5], x, y, z) / hz;
    }
    for (int k = 0; k < 8; ++k) {
        const double hz = dot3(K[6], K[7], K[8], c3[k][0], c3[k][1], c3[k][2]) + 1e-8;
        c2[k][0] = dot3(K[0], K[1], K[2], c3[k][0], c3[k][1], c3[k][2]) / hz;
        c2[k][1] = dot3(K[3], K[4], K[5], c3[k][0], c3[k][1], c3[k][2]) / hz;
    }

    // ---- crop box (hodata.py:161-186) + jitter (rendered_dataset.py:175-190)
    double cx, cy, scale;
    if (cfg.full_image) {
        cx = cfg.raw_w / 2.0; cy = cfg.raw_h / 2.0; scale = (double)cfg.raw_w;
    } else {
        double mnx = 1e300, mny = 1e300, mxx = -1e300, mxy = -1e300;
        const int nj = cfg.crop_model == 0 ? 1 : 21, nc = cfg.crop_model == 2 ? 0 : 8;
        for (int k = 0; k < nj; ++k) { mnx = fmin(mnx, j2[k][0]); mxx = fmax(mxx, j2[k][0]); mny = fmin(mny, j2[k][1]); mxy = fmax(mxy, j2[k][1]); }
        for (int k = 0; k < nc; ++k) { mnx = fmin(mnx, c2[k][0]); mxx = fmax(mxx, c2[k][0]); mny = fmin(mny, c2[k][1]); mxy = fmax(mxy, c2[k][1]); }
        cx = (double)(int)((mxx + mnx) / 2.0);
        cy = (double)(int)((mxy + mny) / 2.0);
        scale = fmax(mxx - mnx, mxy - mny) * (double)cfg.bbox_expand_ratio;
    }
    float cs = 1.0f, sn = 0.0f;
    if (cfg.aug) {
        const double f = (double)cfg.center_jit * scale;
        cx += (double)(int)(f * (double)dr[0]);
        cy += (double)(int)(f * (double)dr[1]);
        double jit = (double)dr[2] + 1.0;
        jit = fmin(fmax(jit, 1.0 - (double)cfg.scale_jit), 1.0 + (double)cfg.scale_jit);
        scale = scale * jit;
        cs = dr[3]; sn = dr[4];
    }

    // ---- forward affine (transform.py:434-460), rounded to fp32; closed-form inverse rounded to fp32
    const double c64 = (double)cs, s64 = (double)sn, ox = K[2], oy = K[5];
    const double orcx = c64 * cx + (-s64) * cy, orcy = s64 * cx + c64 * cy;
    const double dx = cx - ox, dy = cy - oy;
    const double tcx = (c64 * dx + (-s64) * dy) + ox, tcy = (s64 * dx + c64 * dy) + oy;
    double m00, m11, m02, m12, p00, p11, p02, p12;
    affine_no_rot(orcx, orcy, scale, cfg.out_w, cfg.out_h, m00, m11, m02, m12);
    affine_no_rot(tcx, tcy, scale, cfg.out_w, cfg.out_h, p00, p11, p02, p12);
    const float T[6] = {(float)(m00 * c64), (float)(m00 * (-s64)), (float)m02, (float)(m11 * s64), (float)(m11 * c64), (float)m12};
    const float Pp[9] = {(float)p00, 0.f, (float)p02, 0.f, (float)p11, (float)p12, 0.f, 0.f, 1.f};
    float inv[6];
    {
        const double a = T[0], b = T[1], c = T[2], d = T[3], e = T[4], f = T[5];
        const double det = a * e - b * d;
        const double ia = e / det, ib = -b / det, id = -d / det, ie = a / det;
        inv[0] = (float)ia; inv[1] = (float)ib; inv[2] = (float)(-(ia * c + ib * f));
        inv[3] = (float)id; inv[4] = (float)ie; inv[5] = (float)(-(id * c + ie * f));
    }
    if (out.affine) for (int i = 0; i < 6; ++i) out.affine[(size_t)v * 6 + i] = T[i];
    if (out.inv_affine) for (int i = 0; i < 6; ++i) out.inv_affine[(size_t)v * 6 + i] = inv[i];

    // ---- Pillow's warp set-up (Geometry.c): pure scaling -> running double sums; otherwise 16.16 fixed point
    AugView vw;
    double A[6];
    for (int i = 0; i < 6; ++i) A[i] = (double)inv[i];
    vw.scale_mode = (A[1] == 0.0 && A[3] == 0.0) ? 1 : 0;
    if (vw.scale_mode) {
        double xo = A[2] + A[0] * 0.5;
        for (int x = 0; x < cfg.out_w; ++x) { xs_tab[(size_t)v * cfg.out_w + x] = xo < 0.0 ? -1 : (int)xo; xo += A[0]; }
        double yo = A[5] + A[4] * 0.5;
        for (int y = 0; y < cfg.out_h; ++y) { ys_tab[(size_t)v * cfg.out_h + y] = yo < 0.0 ? -1 : (int)yo; yo += A[4]; }
        for (int i = 0; i < 6; ++i) vw.a[i] = 0;
    } else {
        vw.a[0] = (long long)floor(A[0] * 65536.0 + 0.5);
        vw.a[1] = (long long)floor(A[1] * 65536.0 + 0.5);
        vw.a[3] = (long long)floor(A[3] * 65536.0 + 0.5);
        vw.a[4] = (long long)floor(A[4] * 65536.0 + 0.5);
        vw.a[2] = (long long)floor((A[2] + A[0] * 0.5 + A[1] * 0.5) * 65536.0 + 0.5);
        vw.a[5] = (long long)floor((A[5] + A[3] * 0.5 + A[4] * 0.5) * 65536.0 + 0.5);
        // Pillow's check_fixed: the four output corners must map inside +-32768 source pixels
        const double w = cfg.out_w, h = cfg.out_h;
        const double ex = fmax(fmax(fabs(A[2]), fabs(w * A[0] + h * A[1] + A[2])), fmax(fabs(h * A[1] + A[2]), fabs(w * A[0] + A[2])));
        const double ey = fmax(fmax(fabs(A[5]), fabs(w * A[3] + h * A[4] + A[5])), fmax(fabs(h * A[4] + A[5]), fabs(w * A[3] + A[5])));
        if (!(ex < 32768.0 && ey < 32768.0)) atomicOr(status, 2);
    }

    // ---- blur weights (BoxBlur.c) and colour factors
    vw.blur_on = 0; vw.ww = 1u << 24; vw.fw = 0u;
    vw.alpha[0] = vw.alpha[1] = vw.alpha[2] = 1.0f;
    vw.hue_shift = 0; vw.n_before_contrast = 0;
    for (int k = 0; k < 4; ++k) vw.order[k] = -1;
    if (cfg.aug) {
        const float radius = dr[5];
        if (radius != 0.0f) {
            const float sigma2 = radius * radius / 3.0f;
            const float L = (float)sqrt(12.0 * (double)sigma2 + 1.0);
            const float l = (float)floor(((double)L - 1.0) / 2.0);
            float a = (2.0f * l + 1.0f) * (l * (l + 1.0f) - 3.0f * sigma2);
            a = a / (6.0f * (sigma2 - (l + 1.0f) * (l + 1.0f)));
            const float fr = l + a;
            if ((int)fr != 0) atomicOr(status, 1);  // box radius >= 1: outside the reference's range (radius <= 0.1)
            vw.ww = (unsigned)(16777216.0f / (fr * 2.0f + 1.0f));
            vw.fw = ((1u << 24) - vw.ww) / 2u;
            vw.blur_on = 1;
        }
        vw.alpha[0] = dr[6]; vw.alpha[1] = dr[7]; vw.alpha[2] = dr[8];
        vw.hue_shift = ((int)trunc((double)dr[9] * 255.0)) & 0xFF;
        vw.n_before_contrast = 4;
        for (int k = 0; k < 4; ++k) {
            vw.order[k] = order[(size_t)v * 4 + k];
            if (vw.order[k] == 3 && vw.n_before_contrast == 4) vw.n_before_contrast = k;
        }
    }
    views[v] = vw;

    // ---- annotations (rendered_dataset.py:207-253)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            out.cam_intr[(size_t)v * 9 + 3 * i + j] = (float)dot3((double)Pp[3 * i], (double)Pp[3 * i + 1], (double)Pp[3 * i + 2], K[j], K[3 + j], K[6 + j]);
    double j3[21][3];
    for (int k = 0; k < 21; ++k) {
        const double x = J[3 * k], y = J[3 * k + 1];
        j3[k][0] = c64 * x + (-s64) * y; j3[k][1] = s64 * x + c64 * y; j3[k][2] = (double)J[3 * k + 2];
    }
    const double rx = j3[cfg.center_idx][0], ry = j3[cfg.center_idx][1], rz = j3[cfg.center_idx][2];
    out.root_joint[(size_t)v * 3] = (float)rx; out.root_joint[(size_t)v * 3 + 1] = (float)ry; out.root_joint[(size_t)v * 3 + 2] = (float)rz;
    const double t00 = T[0], t01 = T[1], t02 = T[2], t10 = T[3], t11 = T[4], t12 = T[5];
    int raw_vis = 0, aug_vis = 0;
    float visbuf[21];
    for (int k = 0; k < 21; ++k) {
        out.joints_3d[(size_t)v * 63 + 3 * k] = (float)(j3[k][0] - rx);
        out.joints_3d[(size_t)v * 63 + 3 * k + 1] = (float)(j3[k][1] - ry);
        out.joints_3d[(size_t)v * 63 + 3 * k + 2] = (float)(j3[k][2] - rz);
        const float u = (float)((t00 * j2[k][0] + t01 * j2[k][1]) + t02), w = (float)((t10 * j2[k][0] + t11 * j2[k][1]) + t12);
        out.joints_2d[(size_t)v * 42 + 2 * k] = u; out.joints_2d[(size_t)v * 42 + 2 * k + 1] = w;
        raw_vis += (j2[k][0] >= 0 && j2[k][0] < cfg.raw_w && j2[k][1] >= 0 && j2[k][1] < cfg.raw_h);
        visbuf[k] = (u >= 0 && u < cfg.out_w && w >= 0 && w < cfg.out_h) ? 1.0f : 0.0f;
        aug_vis += (int)visbuf[k];
    }
    const bool jv = !((double)raw_vis < 21 * 0.4 || (double)aug_vis < 21 * 0.4);
    for (int k = 0; k < 21; ++k) out.joints_vis[(size_t)v * 21 + k] = jv ? visbuf[k] : 0.0f;
    raw_vis = aug_vis = 0;
    for (int k = 0; k < 8; ++k) {
        const double x = c3[k][0], y = c3[k][1];
        out.corners_3d[(size_t)v * 24 + 3 * k] = (float)((c64 * x + (-s64) * y) - rx);
        out.corners_3d[(size_t)v * 24 + 3 * k + 1] = (float)((s64 * x + c64 * y) - ry);
        out.corners_3d[(size_t)v * 24 + 3 * k + 2] = (float)(c3[k][2] - rz);
        const float u = (float)((t00 * c2[k][0] + t01 * c2[k][1]) + t02), w = (float)((t10 * c2[k][0] + t11 * c2[k][1]) + t12);
        out.corners_2d[(size_t)v * 16 + 2 * k] = u; out.corners_2d[(size_t)v * 16 + 2 * k + 1] = w;
        raw_vis += (c2[k][0] >= 0 && c2[k][0] < cfg.raw_w && c2[k][1] >= 0 && c2[k][1] < cfg.raw_h);
        visbuf[k] = (u >= 0 && u < cfg.out_w && w >= 0 && w < cfg.out_h) ? 1.0f : 0.0f;
        aug_vis += (int)visbuf[k];
    }
    const bool cv = !((double)raw_vis < 8 * 0.4 || (double)aug_vis < 8 * 0.4);
    for (int k = 0; k < 8; ++k) out.corners_vis[(size_t)v * 8 + k] = cv ? visbuf[k] : 0.0f;
    for (int j = 0; j < 4; ++j) {
        out.obj_transf[(size_t)v * 16 + j] = (float)(c64 * (double)Pm[j] + (-s64) * (double)Pm[4 + j]);
        out.obj_transf[(size_t)v * 16 + 4 + j] = (float)(s64 * (double)Pm[j] + c64 * (double)Pm[4 + j]);
        out.obj_transf[(size_t)v * 16 + 8 + j] = Pm[8 + j];
        out.obj_transf[(size_t)v * 16 + 12 + j] = j == 3 ? 1.0f : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------------------ blur
__device__ __forceinline__ unsigned blur3(unsigned l, unsigned c, unsigned r, unsigned ww, unsigned fw) {
    return (c * ww + (l + r) * fw + (1u << 23)) >> 24;  // UINT32 arithmetic, as in ImagingLineBoxBlur
}

// v[k] holds a level's value at position p + k - 2 (k - 1 for the 3-wide one).  Positions outside [0, n) are the edge pixel
// of THAT level: copy inwards-to-outwards with static indices (the arrays stay in registers).
__device__ __forceinline__ void replicate5(unsigned (&v)[5], int p, int n) {
#pragma unroll
    for (int k = 3; k >= 0; --k) if (p + k - 2 < 0) v[k] = v[k + 1];
#pragma unroll
    for (int k = 1; k <= 4; ++k) if (p + k - 2 > n - 1) v[k] = v[k - 1];
}
__device__ __forceinline__ void replicate3(unsigned (&v)[3], int p, int n) {
#pragma unroll
    for (int k = 1; k >= 0; --k) if (p + k - 1 < 0) v[k] = v[k + 1];
#pragma unroll
    for (int k = 1; k <= 2; ++k) if (p + k - 1 > n - 1) v[k] = v[k - 1];
}

// One CTA per 32 x 32 tile of source pixels.  The three horizontal passes of a pixel depend on its row alone, so they are
// evaluated once per (row, column) of the tile and its 3-row aprons into shared memory (38 x 32 results), and the three
// vertical passes read their 7-row column from there: 59 box evaluations per pixel and channel instead of the 216 of a
// thread that redoes the horizontal passes of its whole 7 x 7 window (210 -> ~60 us per 48 views of 256 x 256).
// Window positions that fall outside the image are replicas of the edge pixel AT EVERY PASS (each pass of Pillow clamps
// its own input), hence the re-clamping between the levels; rows outside the image are the clamped row's results.
constexpr int kBlurTile = 32;

__global__ void __launch_bounds__(256)
augment_blur_kernel(const uchar4* __restrict__ src, int B, int H, int W, const AugView* __restrict__ views,
                    uchar4* __restrict__ blurred, unsigned* __restrict__ lsum) {
    __shared__ unsigned red[8];
    __shared__ uchar4 hres[kBlurTile + 6][kBlurTile];
    const int v = blockIdx.y;
    const int tiles_x = (W + kBlurTile - 1) / kBlurTile;
    const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
    const int x0 = tx * kBlurTile, y0 = ty * kBlurTile;
    const AugView& vw = views[v];
    const uchar4* img = src + (size_t)v * H * W;
    const bool blur = vw.blur_on != 0;
    const unsigned ww = vw.ww, fw = vw.fw;
    if (blur) {
        for (int it = threadIdx.x; it < (kBlurTile + 6) * kBlurTile; it += 256) {
            const int ry = it / kBlurTile, cx = it - ry * kBlurTile;
            const int x = x0 + cx;
            if (x >= W) continue;
            const int yy = min(max(y0 + ry - 3, 0), H - 1);
            unsigned l0[3][7];
#pragma unroll
            for (int dx = 0; dx < 7; ++dx) {
                const uchar4 p = img[(size_t)yy * W + min(max(x + dx - 3, 0), W - 1)];
                l0[0][dx] = p.x; l0[1][dx] = p.y; l0[2][dx] = p.z;
            }
            unsigned o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                unsigned l1[5], l2[3];
#pragma unroll
                for (int k = 0; k < 5; ++k) l1[k] = blur3(l0[c][k], l0[c][k + 1], l0[c][k + 2], ww, fw);      // x-2 .. x+2
                replicate5(l1, x, W);  // a position outside the image IS the edge pixel of this level
#pragma unroll
                for (int k = 0; k < 3; ++k) l2[k] = blur3(l1[k], l1[k + 1], l1[k + 2], ww, fw);              // x-1 .. x+1
                replicate3(l2, x, W);
                o[c] = blur3(l2[0], l2[1], l2[2], ww, fw);
            }
            hres[ry][cx] = make_uchar4((unsigned char)o[0], (unsigned char)o[1], (unsigned char)o[2], 0);
        }
    }
    __syncthreads();
    unsigned lum = 0;
    const int cx = threadIdx.x & (kBlurTile - 1), x = x0 + cx;
#pragma unroll 1
    for (int cy = threadIdx.x / kBlurTile; cy < kBlurTile; cy += 256 / kBlurTile) {
        const int y = y0 + cy;
        if (x >= W || y >= H) continue;
        const int i = y * W + x;
        int r, g, b;
        if (!blur) {
            const uchar4 p = img[i];
            r = p.x; g = p.y; b = p.z;
        } else {
            unsigned col[3][7];  // after the horizontal passes: column x, rows y-3 .. y+3
#pragma unroll
            for (int dy = 0; dy < 7; ++dy) {
                const uchar4 q = hres[cy + dy][cx];
                col[0][dy] = q.x; col[1][dy] = q.y; col[2][dy] = q.z;
            }
            int out3[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                unsigned l1[5], l2[3];
#pragma unroll
                for (int k = 0; k < 5; ++k) l1[k] = blur3(col[c][k], col[c][k + 1], col[c][k + 2], ww, fw);
                replicate5(l1, y, H);
#pragma unroll
                for (int k = 0; k < 3; ++k) l2[k] = blur3(l1[k], l1[k + 1], l1[k + 2], ww, fw);
                replicate3(l2, y, H);
                out3[c] = (int)blur3(l2[0], l2[1], l2[2], ww, fw);
            }
            r = out3[0]; g = out3[1]; b = out3[2];
        }
        blurred[(size_t)v * H * W + i] = make_uchar4((unsigned char)r, (unsigned char)g, (unsigned char)b, 255);
        if (vw.n_before_contrast < 4) {  // the Contrast enhancer's mean: luma of the image as it is when contrast runs
            apply_ops(vw, 0, vw.n_before_contrast, 0, r, g, b);
            lum += (unsigned)rgb_to_l(r, g, b);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lum += __shfl_xor_sync(0xffffffffu, lum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned s = 0;
        for (int w = 0; w < 8; ++w) s += red[w];
        if (s) atomicAdd(lsum + v, s);  // integer sum: order-independent
    }
}

// ------------------------------------------------------------------------------------------------------ warp
__global__ void __launch_bounds__(256)
augment_warp_kernel(const uchar4* __restrict__ blurred, int B, int H, int W, int Ho, int Wo, const AugView* __restrict__ views,
                    const int* __restrict__ xs_tab, const int* __restrict__ ys_tab, const unsigned* __restrict__ lsum,
                    float* __restrict__ image) {
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ho * Wo) return;
    const int y = i / Wo, x = i - y * Wo;
    const AugView& vw = views[v];
    int xin, yin;
    if (vw.scale_mode) {
        xin = xs_tab[(size_t)v * Wo + x];
        yin = ys_tab[(size_t)v * Ho + y];
    } else {
        xin = (int)((vw.a[2] + (long long)x * vw.a[0] + (long long)y * vw.a[1]) >> 16);
        yin = (int)((vw.a[5] + (long long)x * vw.a[3] + (long long)y * vw.a[4]) >> 16);
    }
    int r = 0, g = 0, b = 0;
    if (xin >= 0 && xin < W && yin >= 0 && yin < H) {
        const uchar4 p = blurred[((size_t)v * H + yin) * W + xin];
        r = p.x; g = p.y; b = p.z;
        // ImageStat mean: float(sum) / float(count) + 0.5, truncated
        const int mean = (int)((double)lsum[v] / (double)(H * W) + 0.5);
        int n_ops = 0;
        while (n_ops < 4 && vw.order[n_ops] >= 0) ++n_ops;
        apply_ops(vw, 0, n_ops, mean, r, g, b);
    }
    // pixels the warp leaves unfilled are 0 BEFORE normalisation (Pillow fills, then to_tensor): 0 / 255 - 0.5
    float* o = image + (size_t)v * 3 * Ho * Wo + i;
    o[0] = (float)r / 255.0f - 0.5f;
    o[(size_t)Ho * Wo] = (float)g / 255.0f - 0.5f;
    o[(size_t)2 * Ho * Wo] = (float)b / 255.0f - 0.5f;
}

static size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }

}  // namespace ab

using namespace ab;

extern "C" uint64_t ab_augment_workspace_bytes(const ab_augment_cfg* cfg, int B) {
    if (!cfg || B <= 0) return 0;
    return align256(sizeof(AugView) * B) + align256(sizeof(int) * (size_t)B * cfg->out_w) + align256(sizeof(int) * (size_t)B * cfg->out_h) +
           align256(sizeof(unsigned) * B) + align256(sizeof(int)) + align256((size_t)B * cfg->raw_w * cfg->raw_h * 4);
}

extern "C" int ab_crop_augment(const ab_augment_cfg* cfg, int B, const uint8_t* rgba, const float* joints, const float* obj_pose,
                               const float* corners_can, const float* draws, const int32_t* order, float* image, float* cam_intr,
                               float* root_joint, float* joints_3d, float* joints_2d, float* joints_vis, float* corners_3d,
                               float* corners_2d, float* corners_vis, float* obj_transf, float* affine, float* inv_affine,
                               int32_t* status, void* ws, void* stream) {
    AB_REQUIRE(cfg, "null config");
    AB_REQUIRE(B >= 0 && cfg->raw_w > 0 && cfg->raw_h > 0 && cfg->out_w > 0 && cfg->out_h > 0, "bad shape");
    AB_REQUIRE(cfg->center_idx >= 0 && cfg->center_idx < 21 && cfg->crop_model >= 0 && cfg->crop_model <= 2, "bad config");
    if (B == 0) return AB_OK;
    AB_REQUIRE(rgba && joints && obj_pose && corners_can && image && cam_intr && root_joint && joints_3d && joints_2d && joints_vis &&
                   corners_3d && corners_2d && corners_vis && obj_transf && ws, "null pointer");
    AB_REQUIRE(!cfg->aug || (draws && order), "augmentation needs the random draws");
    AB_REQUIRE(((uintptr_t)rgba & 3) == 0 && ((uintptr_t)ws & 255) == 0, "rgba must be 4-byte aligned, ws 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* p = (uint8_t*)ws;
    AugView* views = (AugView*)p;            p += align256(sizeof(AugView) * B);
    int* xs = (int*)p;                       p += align256(sizeof(int) * (size_t)B * cfg->out_w);
    int* ys = (int*)p;                       p += align256(sizeof(int) * (size_t)B * cfg->out_h);
    unsigned* lsum = (unsigned*)p;           p += align256(sizeof(unsigned) * B);
    int* st_word = (int*)p;                  p += align256(sizeof(int));
    uchar4* blurred = (uchar4*)p;
    AB_REQUIRE((long long)cfg->raw_w * cfg->raw_h * 255 < 0xFFFFFFFFll, "image too large for the 32-bit luma sum");
    AB_CUDA(cudaMemsetAsync(st_word, 0, sizeof(int), st));
    AugOut out{image, cam_intr, root_joint, joints_3d, joints_2d, joints_vis, corners_3d, corners_2d, corners_vis, obj_transf, affine,
               inv_affine};
    StageTimer tm(AB_STAGE_AUGMENT, st);
    augment_prelude_kernel<<<cdiv(B, 64), 64, 0, st>>>(*cfg, B, joints, obj_pose, corners_can, draws, order, views, xs, ys, lsum, st_word, out);
    augment_blur_kernel<<<dim3(cdiv(cfg->raw_w, kBlurTile) * cdiv(cfg->raw_h, kBlurTile), B), 256, 0, st>>>((const uchar4*)rgba, B, cfg->raw_h, cfg->raw_w, views,
                                                                                      blurred, lsum);
    augment_warp_kernel<<<dim3(cdiv(cfg->out_w * cfg->out_h, 256), B), 256, 0, st>>>(blurred, B, cfg->raw_h, cfg->raw_w, cfg->out_h, cfg->out_w,
                                                                                      views, xs, ys, lsum, image);
    if (status) AB_CUDA(cudaMemcpyAsync(status, st_word, sizeof(int), cudaMemcpyDeviceToDevice, st));
    count_launch(3);
    return check_launch("augment kernels");
}
