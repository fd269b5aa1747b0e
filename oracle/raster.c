/* ORACLE (test infrastructure; never linked or called by the product path).
 *
 * Scalar CPU restatement of the hand+object view rasteriser: one view, z-buffer, one thread.
 *
 * Reference path: anakin/utils/renderer.py:101-123 (Renderer.__call__) over pyrender 0.1.43 / OpenGL
 * (anakin/utils/frender_utils.py:179-205).  pyrender, PyOpenGL and EGL are third-party, absent from
 * /root/reference and from this image, and GL rasterisation + 4x MSAA resolve is not bit-defined across
 * drivers, so there is NO golden image: "parity unpinned" for this file.  What is anchored on the reference:
 *   - pixel convention u = fx*X/Z + cx, v = fy*Y/Z + cy, camera looks down +z, y down
 *     (renderer.py:76-78 + CONST.PYRENDER_EXTRINSIC misc.py:87-95; rendered_dataset.py:127-133)
 *   - object drawn with model matrix obj_pose, hand drawn with posed vertices (renderer.py:106-109)
 *   - draw order objects then hand (scene insertion order, renderer.py:90-93) => depth ties go to the lower
 *     primitive id, object faces numbered before hand faces
 *   - depth 0 = background, bg composited where depth == 0 (renderer.py:111-119)
 *   - ambient 0.8, one point light at the camera origin whose intensity is redrawn per call (renderer.py:77,103-104)
 * The exact rules below ARE the specification both this file and the CUDA path implement; every fp32
 * operation is individually rounded (compile with -ffp-contract=off; fmaf is a true fused op).
 *
 *  vertex   : Xc = fma(R02,z, fma(R01,y, fma(R00,x, t0))) (object only);  iz = 1/Z;
 *             u = fma(fx, X*iz, cx); v = fma(fy, Y*iz, cy);
 *             xi = rint(clamp(u*256, -2^22, 2^22))  (24.8 fixed point, round-half-even); invalid if !(Z >= znear)
 *  triangle : discarded if any vertex invalid; area2 = (x1-x0)(y2-y0) - (x2-x0)(y1-y0) in int64;
 *             area2 == 0 discarded; area2 > 0 is BACK-facing (y-down image), culled when cull_backface
 *  coverage : sample at pixel centre (256*px+128, 256*py+128); E_i integer edge functions times sign(area2);
 *             inside iff E_i > 0, or E_i == 0 on a top/left edge (D3D top-left rule in y-down image space)
 *  depth    : f_i = (float)E_i; num = fma(f2,iz2, fma(f1,iz1, f0*iz0)); depth = (float)|area2| / num
 *  z test   : min over key = (float_bits(depth) << 32) | prim_id
 *  shading  : see shade() below.  seg: 0 bg, 1 hand, 2 object.  rgba.a = 255 covered / 0 background.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int width, height;
    float fx, fy, cx, cy;
    float znear;
    int cull_backface;
    float ambient, diffuse;
    int bg_r, bg_g, bg_b;
} ab_oracle_cfg;

typedef struct { int32_t x, y; float iz; int32_t ok; } pvert;

static inline int32_t snap(float u) {
    float s = u * 256.0f;
    if (!(s >= -4194304.0f)) s = -4194304.0f; /* also catches NaN */
    if (s > 4194304.0f) s = 4194304.0f;
    return (int32_t)rintf(s);
}

static inline pvert project(const ab_oracle_cfg* c, float X, float Y, float Z) {
    pvert p;
    p.ok = (Z >= c->znear) ? 1 : 0;
    float iz = 1.0f / Z;
    p.iz = iz;
    p.x = snap(fmaf(c->fx, X * iz, c->cx));
    p.y = snap(fmaf(c->fy, Y * iz, c->cy));
    return p;
}

static inline void xform(const float* M, const float* v, float* o) {
    o[0] = fmaf(M[2], v[2], fmaf(M[1], v[1], fmaf(M[0], v[0], M[3])));
    o[1] = fmaf(M[6], v[2], fmaf(M[5], v[1], fmaf(M[4], v[0], M[7])));
    o[2] = fmaf(M[10], v[2], fmaf(M[9], v[1], fmaf(M[8], v[0], M[11])));
}

typedef struct {
    int64_t sarea;
    int32_t x[3], y[3];
    float iz[3];
    int s;
    int tl[3];
} tri_setup;

/* returns 0 if discarded */
static int setup(const ab_oracle_cfg* c, const pvert* a, const pvert* b, const pvert* d, tri_setup* t) {
    if (!(a->ok && b->ok && d->ok)) return 0;
    int64_t area2 = (int64_t)(b->x - a->x) * (int64_t)(d->y - a->y) - (int64_t)(d->x - a->x) * (int64_t)(b->y - a->y);
    if (area2 == 0) return 0;
    if (area2 > 0 && c->cull_backface) return 0;
    t->s = area2 > 0 ? 1 : -1;
    t->sarea = area2 > 0 ? area2 : -area2;
    t->x[0] = a->x; t->x[1] = b->x; t->x[2] = d->x;
    t->y[0] = a->y; t->y[1] = b->y; t->y[2] = d->y;
    t->iz[0] = a->iz; t->iz[1] = b->iz; t->iz[2] = d->iz;
    for (int i = 0; i < 3; ++i) { /* edge i runs v[i+1] -> v[i+2], opposite vertex i */
        int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        int64_t dx = (int64_t)t->s * (t->x[i2] - t->x[i1]);
        int64_t dy = (int64_t)t->s * (t->y[i2] - t->y[i1]);
        t->tl[i] = (dy < 0) || (dy == 0 && dx > 0);
    }
    return 1;
}

static inline int64_t edge(const tri_setup* t, int i, int64_t px, int64_t py) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    int64_t dx = t->x[i2] - t->x[i1], dy = t->y[i2] - t->y[i1];
    return (int64_t)t->s * (dx * (py - t->y[i1]) - dy * (px - t->x[i1]));
}

/* 1 if pixel centre covered; fills e[3] */
static inline int cover(const tri_setup* t, int px, int py, int64_t* e) {
    int64_t cx = 256 * (int64_t)px + 128, cy = 256 * (int64_t)py + 128;
    for (int i = 0; i < 3; ++i) {
        e[i] = edge(t, i, cx, cy);
        if (e[i] < 0 || (e[i] == 0 && !t->tl[i])) return 0;
    }
    return 1;
}

static inline float depth_at(const tri_setup* t, const int64_t* e, float* num_out, float* a) {
    float f0 = (float)e[0], f1 = (float)e[1], f2 = (float)e[2];
    a[0] = f0 * t->iz[0];
    a[1] = f1 * t->iz[1];
    a[2] = f2 * t->iz[2];
    float num = fmaf(f2, t->iz[2], fmaf(f1, t->iz[1], a[0]));
    *num_out = num;
    return (float)t->sarea / num;
}

static inline int floordiv256(int64_t v) { return (int)(v >> 8); } /* arithmetic shift = floor */

/* shading of one covered pixel.  p[3][3]: camera-space corner positions, col[3][3]: vertex colours 0..255 */
static void shade(const ab_oracle_cfg* c, int px, int py, float depth, const float* a, float num, const float p[3][3],
                  const uint8_t col[3][4], float light, uint8_t* rgba) {
    float inv = 1.0f / num;
    float w0 = a[0] * inv, w1 = a[1] * inv, w2 = a[2] * inv;
    float e1x = p[1][0] - p[0][0], e1y = p[1][1] - p[0][1], e1z = p[1][2] - p[0][2];
    float e2x = p[2][0] - p[0][0], e2y = p[2][1] - p[0][1], e2z = p[2][2] - p[0][2];
    float nx = fmaf(e1y, e2z, -(e1z * e2y));
    float ny = fmaf(e1z, e2x, -(e1x * e2z));
    float nz = fmaf(e1x, e2y, -(e1y * e2x));
    float rx = (((float)px + 0.5f) - c->cx) / c->fx;
    float ry = (((float)py + 0.5f) - c->cy) / c->fy;
    float Px = rx * depth, Py = ry * depth, Pz = depth;
    float d2 = fmaf(Px, Px, fmaf(Py, Py, Pz * Pz));
    float n2 = fmaf(nx, nx, fmaf(ny, ny, nz * nz));
    float ndp = fabsf(fmaf(nx, Px, fmaf(ny, Py, nz * Pz)));
    float den = sqrtf(n2 * d2);
    float cosv = den > 0.0f ? ndp / den : 0.0f;
    float sh = fmaf(c->diffuse * light, cosv / d2, c->ambient);
    for (int ch = 0; ch < 3; ++ch) {
        float cc = fmaf(w2, (float)col[2][ch], fmaf(w1, (float)col[1][ch], w0 * (float)col[0][ch]));
        float v = cc * sh;
        if (!(v >= 0.0f)) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        rgba[ch] = (uint8_t)(int)rintf(v);
    }
    rgba[3] = 255;
}

/* Render one view.
 *  hand_verts  f32[n_hv,3] camera space; hand_faces i32[n_hf,3]; hand_cols u8[n_hv,4]
 *  obj_verts   f32[n_ov,3] canonical;    obj_faces  i32[n_of,3]; obj_cols  u8[n_ov,4]; obj_pose f32[16] row-major
 *              (n_of == 0 => hand only, the CONST.DUMMY case renderer.py:107)
 *  bg          u8[bg_h,bg_w,3] or NULL; bg_sel = {x0, y0, crop_w, crop_h}: nearest-neighbour resize of the crop
 *  out: rgba u8[H,W,4], depth f32[H,W], seg u8[H,W], key u64[H,W] (scratch, also returned for inspection)
 */
void ab_oracle_render(const ab_oracle_cfg* c, int n_hv, const float* hand_verts, int n_hf, const int32_t* hand_faces,
                      const uint8_t* hand_cols, int n_ov, const float* obj_verts, int n_of, const int32_t* obj_faces,
                      const uint8_t* obj_cols, const float* obj_pose, float light, const uint8_t* bg, int bg_h, int bg_w,
                      const int32_t* bg_sel, uint8_t* rgba, float* depth, uint8_t* seg, uint64_t* key) {
    const int W = c->width, H = c->height;
    (void)bg_h;
    /* per-thread scratch that only grows: no mmap/munmap per view (they serialise threads on the process mm lock) */
    static __thread pvert* pv = NULL;
    static __thread float* cam = NULL;
    static __thread size_t cap = 0;
    if ((size_t)(n_ov + n_hv) > cap) {
        cap = (size_t)(n_ov + n_hv);
        pv = (pvert*)realloc(pv, sizeof(pvert) * cap);
        cam = (float*)realloc(cam, sizeof(float) * 3 * cap);
    }
    for (int v = 0; v < n_ov; ++v) {
        xform(obj_pose, obj_verts + 3 * v, cam + 3 * v);
        pv[v] = project(c, cam[3 * v], cam[3 * v + 1], cam[3 * v + 2]);
    }
    for (int v = 0; v < n_hv; ++v) {
        float* o = cam + 3 * (n_ov + v);
        memcpy(o, hand_verts + 3 * v, 12);
        pv[n_ov + v] = project(c, o[0], o[1], o[2]);
    }
    for (int i = 0; i < W * H; ++i) key[i] = ~(uint64_t)0;
    const int n_prim = n_of + n_hf;
    for (int f = 0; f < n_prim; ++f) {
        const int32_t* idx = f < n_of ? obj_faces + 3 * f : hand_faces + 3 * (f - n_of);
        const int off = f < n_of ? 0 : n_ov;
        tri_setup t;
        if (!setup(c, &pv[off + idx[0]], &pv[off + idx[1]], &pv[off + idx[2]], &t)) continue;
        int32_t minx = t.x[0], maxx = t.x[0], miny = t.y[0], maxy = t.y[0];
        for (int i = 1; i < 3; ++i) {
            if (t.x[i] < minx) minx = t.x[i];
            if (t.x[i] > maxx) maxx = t.x[i];
            if (t.y[i] < miny) miny = t.y[i];
            if (t.y[i] > maxy) maxy = t.y[i];
        }
        int x0 = floordiv256((int64_t)minx - 128 + 255), x1 = floordiv256((int64_t)maxx - 128);
        int y0 = floordiv256((int64_t)miny - 128 + 255), y1 = floordiv256((int64_t)maxy - 128);
        if (x0 < 0) x0 = 0;
        if (y0 < 0) y0 = 0;
        if (x1 > W - 1) x1 = W - 1;
        if (y1 > H - 1) y1 = H - 1;
        for (int py = y0; py <= y1; ++py)
            for (int px = x0; px <= x1; ++px) {
                int64_t e[3];
                if (!cover(&t, px, py, e)) continue;
                float num, a[3];
                float z = depth_at(&t, e, &num, a);
                uint32_t zb;
                memcpy(&zb, &z, 4);
                uint64_t k = ((uint64_t)zb << 32) | (uint32_t)f;
                if (k < key[py * W + px]) key[py * W + px] = k;
            }
    }
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            const int i = py * W + px;
            uint64_t k = key[i];
            if (k == ~(uint64_t)0) {
                uint8_t r = (uint8_t)c->bg_r, g = (uint8_t)c->bg_g, b = (uint8_t)c->bg_b;
                if (bg) {
                    int sx = bg_sel[0] + (int)(((int64_t)(2 * px + 1) * bg_sel[2]) / (2 * W));
                    int sy = bg_sel[1] + (int)(((int64_t)(2 * py + 1) * bg_sel[3]) / (2 * H));
                    const uint8_t* s = bg + 3 * ((size_t)sy * bg_w + sx);
                    r = s[0]; g = s[1]; b = s[2];
                }
                rgba[4 * i] = r; rgba[4 * i + 1] = g; rgba[4 * i + 2] = b; rgba[4 * i + 3] = 0;
                depth[i] = 0.0f;
                seg[i] = 0;
                continue;
            }
            int f = (int)(uint32_t)(k & 0xffffffffu);
            const int32_t* idx = f < n_of ? obj_faces + 3 * f : hand_faces + 3 * (f - n_of);
            const int off = f < n_of ? 0 : n_ov;
            const uint8_t* cols = f < n_of ? obj_cols : hand_cols;
            tri_setup t;
            setup(c, &pv[off + idx[0]], &pv[off + idx[1]], &pv[off + idx[2]], &t);
            int64_t e[3];
            cover(&t, px, py, e);
            float num, a[3];
            float z = depth_at(&t, e, &num, a);
            float p[3][3];
            uint8_t col[3][4];
            for (int j = 0; j < 3; ++j) {
                memcpy(p[j], cam + 3 * (off + idx[j]), 12);
                memcpy(col[j], cols + 4 * idx[j], 4);
            }
            shade(c, px, py, z, a, num, p, col, light, rgba + 4 * i);
            depth[i] = z;
            seg[i] = f < n_of ? 2 : 1;
        }
}

/* Batch driver for the CPU baseline: the views of a batch are independent, one OpenMP task per view.
 * Scene layout as in include/artiboost_b200.h (objects concatenated with prefix offsets, faces as 3 ints here).
 * key_scratch: u64[n_threads][H*W].  Timed by bench.py as the host-core baseline; never used by the product. */
void ab_oracle_render_batch(const ab_oracle_cfg* c, int n_views, int n_hv, const float* hand_verts, int n_hf,
                            const int32_t* hand_faces, const uint8_t* hand_cols_all, const int32_t* hand_tex,
                            const float* obj_verts, const int32_t* obj_vert_off, const int32_t* obj_faces,
                            const int32_t* obj_face_off, const uint8_t* obj_cols, const int32_t* obj_id,
                            const float* obj_pose, const float* light, const uint8_t* bgs, int bg_h, int bg_w,
                            const int32_t* bg_sel, uint8_t* rgba, float* depth, uint8_t* seg, uint64_t* key_scratch,
                            int n_threads) {
    const size_t npx = (size_t)c->width * c->height;
#pragma omp parallel for schedule(dynamic, 1) num_threads(n_threads)
    for (int v = 0; v < n_views; ++v) {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        const int oid = obj_id[v];
        int n_ov = 0, n_of = 0;
        const float* ov = NULL;
        const int32_t* of = NULL;
        const uint8_t* oc = NULL;
        if (oid >= 0) {
            n_ov = obj_vert_off[oid + 1] - obj_vert_off[oid];
            n_of = obj_face_off[oid + 1] - obj_face_off[oid];
            ov = obj_verts + 3 * (size_t)obj_vert_off[oid];
            of = obj_faces + 3 * (size_t)obj_face_off[oid];
            oc = obj_cols + 4 * (size_t)obj_vert_off[oid];
        }
        const uint8_t* bg = NULL;
        const int32_t* sel = NULL;
        if (bgs && bg_sel && bg_sel[5 * v] >= 0) {
            bg = bgs + 3 * (size_t)bg_sel[5 * v] * bg_h * bg_w;
            sel = bg_sel + 5 * v + 1;
        }
        ab_oracle_render(c, n_hv, hand_verts + 3 * (size_t)v * n_hv, n_hf, hand_faces,
                         hand_cols_all + 4 * (size_t)hand_tex[v] * n_hv, n_ov, ov, n_of, of, oc, obj_pose + 16 * (size_t)v,
                         light[v], bg, bg_h, bg_w, sel, rgba + 4 * npx * v, depth + npx * v, seg + npx * v,
                         key_scratch + npx * tid);
    }
}
