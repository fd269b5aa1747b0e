// Batched hand+object view rasteriser (sm_100a).
//
// Replaces Renderer.__call__ (anakin/utils/renderer.py:101-123) over pyrender's OffscreenRenderer
// (anakin/utils/frender_utils.py:179-205): B views per call instead of one GL draw + two glReadPixels per call.
// The rule set (fixed-point snapping, top-left fill rule, 1/z depth, z-test key, shading) is the one stated at the
// top of oracle/raster.c; every fp32 operation below is individually rounded in the same order (this file is
// compiled with -fmad=false and uses explicit __fmaf_rn where the rules say "fma").
//
// Views are processed in chunks; per chunk three kernels run back to back on the caller's stream:
//   1. raster_vertex_kernel    one thread per (view, vertex): object model matrix, pinhole projection, snapping
//                              to 24.8 fixed point.  16 B per projected vertex, coalesced, into an L2-resident scratch.
//   2. raster_triangle_kernel  one thread per (view, triangle): integer set-up, cull, pixel bbox, coverage with
//                              exact integer edge functions (32-bit when the bbox is < 128 px, else 64-bit),
//                              z-test by RED.MIN.U64 on key = depth_bits << 32 | prim_id into the chunk's key
//                              buffer (8 B / pixel, L2-resident).  At 256^2 most triangles cover 0-2 pixel centres,
//                              so the pass is set-up bound, not fill bound.
//   3. raster_resolve_kernel   one thread per 4 consecutive pixels: winner triangle re-set-up, exact depth,
//                              shading, background compositing; 16-byte RGBA / depth stores, 4-byte seg stores;
//                              restores the key buffer to "empty" where it was hit.
// HBM traffic per view is the output (9 B / pixel) plus the 9.4 KB of per-view inputs; scratch never needs to
// leave L2 when `chunk` is sized so (chunk * (8*W*H + 16*verts)) stays well under the 126 MB L2.
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "common.cuh"

namespace ab {

constexpr int kMaxObjects = 64;
constexpr unsigned long long kEmptyKey = ~0ull;

// floor(n / d) for n < 2^30 as (n * m) >> s with m = ceil(2^s / d), s = 31 + ceil(log2 d): m < 2^32, and with
// e = m d - 2^s in [0, d) the quotient is exact because n e < 2^30 2^ceil(log2 d) <= 2^s.
static void make_fastdiv(unsigned d, unsigned* m, int* s) {
    int L = 0;
    while ((1ull << L) < d) ++L;
    *s = 31 + L;
    *m = (unsigned)(((1ull << *s) + d - 1) / d);
}
__device__ __forceinline__ unsigned fastdiv(unsigned n, unsigned m, int s) {
    return (unsigned)(((unsigned long long)n * m) >> s);
}

struct RasterParams {
    // scene
    const float* obj_verts;
    const int4* obj_faces;
    const uchar4* obj_colors;
    const int4* hand_faces;
    const uchar4* hand_colors;
    const uint8_t* bgs;
    int n_obj, n_hv, n_hf, n_tex, n_bg, bg_h, bg_w, bg_ch;
    unsigned div_w_m, div_2w_m, div_2h_m;  // exact division by W, 2W, 2H as multiply + shift (FastDiv below)
    int div_w_s, div_2w_s, div_2h_s;
    int max_ov, max_of;  // launch sizing
    // camera
    int W, H;
    float fx, fy, cx, cy, znear, ambient, diffuse;
    int cull;
    int bg_r, bg_g, bg_b;
    // per-view inputs (already offset to the chunk's first view)
    const float* hand_verts;
    const int32_t* hand_tex;
    const int32_t* obj_id;
    const float* obj_pose;
    const float* light;
    const int32_t* bg_sel;
    // outputs (offset to the chunk)
    uint8_t* rgba;
    float* depth;
    uint8_t* seg;
    // scratch
    unsigned long long* keys;  // [chunk][H][W]
    int4* pv;                  // [chunk][pv_stride]  {x, y, iz bits, ok}
    int pv_stride;
    int n_views;
    int vert_off[kMaxObjects + 1];
    int face_off[kMaxObjects + 1];
};

// ---- rule: vertex ------------------------------------------------------------------------------------------
__device__ __forceinline__ int snap(float u) {
    float s = __fmul_rn(u, 256.0f);
    if (!(s >= -4194304.0f)) s = -4194304.0f;  // also catches NaN
    if (s > 4194304.0f) s = 4194304.0f;
    return __float2int_rn(s);
}

__device__ __forceinline__ int4 project(const RasterParams& P, float X, float Y, float Z) {
    int4 r;
    r.w = (Z >= P.znear) ? 1 : 0;
    float iz = __fdiv_rn(1.0f, Z);
    r.z = __float_as_int(iz);
    r.x = snap(__fmaf_rn(P.fx, __fmul_rn(X, iz), P.cx));
    r.y = snap(__fmaf_rn(P.fy, __fmul_rn(Y, iz), P.cy));
    return r;
}

__device__ __forceinline__ void xform(const float* __restrict__ M, float x, float y, float z, float* o) {
    o[0] = __fmaf_rn(M[2], z, __fmaf_rn(M[1], y, __fmaf_rn(M[0], x, M[3])));
    o[1] = __fmaf_rn(M[6], z, __fmaf_rn(M[5], y, __fmaf_rn(M[4], x, M[7])));
    o[2] = __fmaf_rn(M[10], z, __fmaf_rn(M[9], y, __fmaf_rn(M[8], x, M[11])));
}

__global__ void __launch_bounds__(256)
raster_vertex_kernel(const __grid_constant__ RasterParams P) {
    const int view = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int oid = P.obj_id[view];
    const int n_ov = oid >= 0 ? P.vert_off[oid + 1] - P.vert_off[oid] : 0;
    if (t >= n_ov + P.n_hv) return;
    float c[3];
    if (t < n_ov) {
        const float* v = P.obj_verts + 3 * (size_t)(P.vert_off[oid] + t);
        xform(P.obj_pose + 16 * (size_t)view, v[0], v[1], v[2], c);
    } else {
        const float* v = P.hand_verts + 3 * ((size_t)view * P.n_hv + (t - n_ov));
        c[0] = v[0]; c[1] = v[1]; c[2] = v[2];
    }
    P.pv[(size_t)view * P.pv_stride + t] = project(P, c[0], c[1], c[2]);
}

// ---- rule: triangle ----------------------------------------------------------------------------------------
// Triangles whose snapped bbox is under 64 px in both axes (all of them in valid ArtiBoost views) take the "small"
// path: every factor of the edge functions is below 2^14 + 2^8 in magnitude, so products and their differences are
// exact in int32.  Larger ones take the int64 path.  Both produce the same integers, hence the same floats.
constexpr int kSmallExtent = 16384;  // 64 px in 24.8 fixed point
constexpr int kTriThreads = 256;

struct TriRec {            // one per surviving triangle of a CTA (shared memory); 27 words: odd stride, no bank pattern
    int x0, y0, w;         // pixel bbox origin, width
    int f;                 // primitive id
    int e[3];              // small path: edge values at the centre of pixel (x0, y0)
    int ex[3], ey[3];      // small path: edge steps per pixel in x / y
    float iz[3];
    float sarea;           // (float)|area2|
    int nbias;             // bit i set <=> edge i is not a top/left edge (E_i == 0 is outside)
    int big;               // 1 -> int64 path from vx, vy, s
    int vx[3], vy[3], s;
    int pad;
};

__device__ __forceinline__ float depth_from(const float* iz, float sarea, float f0, float f1, float f2, float* num_out,
                                            float* a) {
    a[0] = __fmul_rn(f0, iz[0]);
    a[1] = __fmul_rn(f1, iz[1]);
    a[2] = __fmul_rn(f2, iz[2]);
    const float num = __fmaf_rn(f2, iz[2], __fmaf_rn(f1, iz[1], a[0]));
    *num_out = num;
    return __fdiv_rn(sarea, num);
}

__device__ __forceinline__ int floordiv256(int v) { return v >> 8; }

__device__ __forceinline__ void emit(unsigned long long* __restrict__ keys, int idx, float z, int f) {
    const unsigned long long k = ((unsigned long long)__float_as_uint(z) << 32) | (unsigned)f;
    atomicMin(keys + idx, k);  // result unused -> RED.E.MIN.64
}

// signed 2*area with the sign convention of the rule set; `small` selects the exact-in-int32 evaluation
__device__ __forceinline__ long long area2_of(const int4& a, const int4& b, const int4& d, bool small) {
    if (small) return (long long)((b.x - a.x) * (d.y - a.y) - (d.x - a.x) * (b.y - a.y));
    return (long long)(b.x - a.x) * (long long)(d.y - a.y) - (long long)(d.x - a.x) * (long long)(b.y - a.y);
}

__device__ __forceinline__ long long edge64(const int* x, const int* y, int s, int i, long long px, long long py) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    const long long dx = x[i2] - x[i1], dy = y[i2] - y[i1];
    return (long long)s * (dx * (py - y[i1]) - dy * (px - x[i1]));
}

// Phase A: one thread per triangle -- gather, pixel bbox (most sub-pixel triangles stop here), area / cull, edge
// set-up into shared memory.  Phase B: the CTA's surviving bbox ROWS are dealt out evenly to all 256 threads (block
// scan + binary search), so lanes stay busy although triangles differ in size by two orders of magnitude.
template <int kMinBlocks>
__global__ void __launch_bounds__(kTriThreads, kMinBlocks)
raster_triangle_kernel(const __grid_constant__ RasterParams P) {
    __shared__ TriRec recs[kTriThreads];
    __shared__ int prefix[kTriThreads + 1];
    __shared__ int warp_tot[kTriThreads / 32];
    const int view = blockIdx.y;
    const int t = threadIdx.x;
    const int f = blockIdx.x * kTriThreads + t;
    const int oid = P.obj_id[view];
    int n_of = 0, n_ov = 0;
    if (oid >= 0) {
        n_of = P.face_off[oid + 1] - P.face_off[oid];
        n_ov = P.vert_off[oid + 1] - P.vert_off[oid];
    }
    if (blockIdx.x * kTriThreads >= n_of + P.n_hf) return;  // whole CTA past the end (launch sized for the largest object)
    int rows = 0;
    if (f < n_of + P.n_hf) {
        int4 idx;
        int off;
        if (f < n_of) { idx = __ldg(P.obj_faces + P.face_off[oid] + f); off = 0; }
        else { idx = __ldg(P.hand_faces + (f - n_of)); off = n_ov; }
        const int4* pv = P.pv + (size_t)view * P.pv_stride + off;
        const int4 a = pv[idx.x], b = pv[idx.y], d = pv[idx.z];
        if (a.w & b.w & d.w) {
            const int minx = min(a.x, min(b.x, d.x)), maxx = max(a.x, max(b.x, d.x));
            const int miny = min(a.y, min(b.y, d.y)), maxy = max(a.y, max(b.y, d.y));
            const int x0 = max(floordiv256(minx - 128 + 255), 0), x1 = min(floordiv256(maxx - 128), P.W - 1);
            const int y0 = max(floordiv256(miny - 128 + 255), 0), y1 = min(floordiv256(maxy - 128), P.H - 1);
            if (x0 <= x1 && y0 <= y1) {
                const bool small = (maxx - minx < kSmallExtent) && (maxy - miny < kSmallExtent);
                const long long area2 = area2_of(a, b, d, small);
                if (area2 != 0 && !(area2 > 0 && P.cull)) {
                    TriRec& r = recs[t];
                    const int s = area2 > 0 ? 1 : -1;
                    const int vx[3] = {a.x, b.x, d.x}, vy[3] = {a.y, b.y, d.y};
                    r.x0 = x0; r.y0 = y0; r.w = x1 - x0 + 1; r.f = f;
                    r.iz[0] = __int_as_float(a.z); r.iz[1] = __int_as_float(b.z); r.iz[2] = __int_as_float(d.z);
                    r.sarea = __ll2float_rn(area2 > 0 ? area2 : -area2);
                    r.big = small ? 0 : 1;
                    r.s = s;
                    int nbias = 0;
                    const int cx0 = 256 * x0 + 128, cy0 = 256 * y0 + 128;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // edge i runs v[i+1] -> v[i+2], opposite vertex i
                        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                        const int dx = s * (vx[i2] - vx[i1]), dy = s * (vy[i2] - vy[i1]);  // |.| <= 2^23: no overflow
                        if (!((dy < 0) || (dy == 0 && dx > 0))) nbias |= 1 << i;
                        r.vx[i] = vx[i]; r.vy[i] = vy[i];
                        if (small) {
                            r.e[i] = dx * (cy0 - vy[i1]) - dy * (cx0 - vx[i1]);
                            r.ex[i] = -256 * dy;
                            r.ey[i] = 256 * dx;
                        }
                    }
                    r.nbias = nbias;
                    rows = y1 - y0 + 1;
                }
            }
        }
    }
    // ---- exclusive block scan of the row counts
    const int lane = t & 31, wid = t >> 5;
    int incl = rows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int w = 0; w < kTriThreads / 32; ++w) base += (w < wid) ? warp_tot[w] : 0;
    prefix[t] = base + incl - rows;
    if (t == kTriThreads - 1) prefix[kTriThreads] = base + incl;
    __syncthreads();
    const int total = prefix[kTriThreads];
    unsigned long long* keys = P.keys + (size_t)view * P.W * P.H;
    // ---- phase B: one bbox row per thread per iteration
    for (int c = t; c < total; c += kTriThreads) {
        int lo = 1, hi = kTriThreads;  // first index with prefix[idx] > c; prefix[0] = 0 <= c, 256 candidates -> 8 halvings
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int mid = (lo + hi) >> 1;
            if (prefix[mid] > c) hi = mid; else lo = mid + 1;
        }
        const TriRec& r = recs[lo - 1];
        const int row = c - prefix[lo - 1];
        const int py = r.y0 + row;
        const int b0 = -(r.nbias & 1), b1 = -((r.nbias >> 1) & 1), b2 = -((r.nbias >> 2) & 1);
        const float iz[3] = {r.iz[0], r.iz[1], r.iz[2]};
        const float sarea = r.sarea;
        const int fid = r.f, w = r.w, kbase = py * P.W + r.x0;
        if (!r.big) {
            int e0 = r.e[0] + row * r.ey[0], e1 = r.e[1] + row * r.ey[1], e2 = r.e[2] + row * r.ey[2];
            const int ex0 = r.ex[0], ex1 = r.ex[1], ex2 = r.ex[2];
            for (int i = 0; i < w; ++i) {
                if (((e0 + b0) | (e1 + b1) | (e2 + b2)) >= 0) {
                    float num, aa[3];
                    const float z = depth_from(iz, sarea, __int2float_rn(e0), __int2float_rn(e1), __int2float_rn(e2), &num, aa);
                    emit(keys, kbase + i, z, fid);
                }
                e0 += ex0; e1 += ex1; e2 += ex2;
            }
        } else {
            const int vx[3] = {r.vx[0], r.vx[1], r.vx[2]}, vy[3] = {r.vy[0], r.vy[1], r.vy[2]};
            const int s = r.s;
            const long long cy = 256ll * py + 128;
            for (int i = 0; i < w; ++i) {
                const long long cx = 256ll * (r.x0 + i) + 128;
                const long long e0 = edge64(vx, vy, s, 0, cx, cy), e1 = edge64(vx, vy, s, 1, cx, cy),
                                e2 = edge64(vx, vy, s, 2, cx, cy);
                if (((e0 + b0) | (e1 + b1) | (e2 + b2)) >= 0) {
                    float num, aa[3];
                    const float z = depth_from(iz, sarea, __ll2float_rn(e0), __ll2float_rn(e1), __ll2float_rn(e2), &num, aa);
                    emit(keys, kbase + i, z, fid);
                }
            }
        }
    }
}

// ---- rule: shading + resolve -------------------------------------------------------------------------------
struct PixelOut { uchar4 rgba; float depth; uint8_t seg; };

__device__ __forceinline__ PixelOut shade_pixel(const RasterParams& P, int view, int oid, int n_ov, int n_of, int px,
                                                int py, unsigned f) {
    int4 idx;
    int off;
    const bool is_obj = (int)f < n_of;
    if (is_obj) { idx = __ldg(P.obj_faces + P.face_off[oid] + f); off = 0; }
    else { idx = __ldg(P.hand_faces + (f - n_of)); off = n_ov; }
    const int4* pv = P.pv + (size_t)view * P.pv_stride + off;
    const int vi[3] = {idx.x, idx.y, idx.z};
    const int4 a4 = pv[vi[0]], b4 = pv[vi[1]], d4 = pv[vi[2]];
    // the winner passed set-up in the triangle pass: recompute its edge values at this pixel the same way
    const int vx[3] = {a4.x, b4.x, d4.x}, vy[3] = {a4.y, b4.y, d4.y};
    const float iz[3] = {__int_as_float(a4.z), __int_as_float(b4.z), __int_as_float(d4.z)};
    const int minx = min(vx[0], min(vx[1], vx[2])), maxx = max(vx[0], max(vx[1], vx[2]));
    const int miny = min(vy[0], min(vy[1], vy[2])), maxy = max(vy[0], max(vy[1], vy[2]));
    const bool small = (maxx - minx < kSmallExtent) && (maxy - miny < kSmallExtent);
    const long long area2 = area2_of(a4, b4, d4, small);
    const int s = area2 > 0 ? 1 : -1;
    float fe[3];
    if (small) {
        const int cx = 256 * px + 128, cy = 256 * py + 128;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
            const int dx = s * (vx[i2] - vx[i1]), dy = s * (vy[i2] - vy[i1]);
            fe[i] = __int2float_rn(dx * (cy - vy[i1]) - dy * (cx - vx[i1]));
        }
    } else {
        const long long cx = 256ll * px + 128, cy = 256ll * py + 128;
#pragma unroll
        for (int i = 0; i < 3; ++i) fe[i] = __ll2float_rn(edge64(vx, vy, s, i, cx, cy));
    }
    float num, a[3];
    const float z = depth_from(iz, __ll2float_rn(area2 > 0 ? area2 : -area2), fe[0], fe[1], fe[2], &num, a);
    float p[3][3];
    uchar4 col[3];
    if (is_obj) {
        const float* M = P.obj_pose + 16 * (size_t)view;
        const int vo = P.vert_off[oid];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float* v = P.obj_verts + 3 * (size_t)(vo + vi[j]);
            xform(M, v[0], v[1], v[2], p[j]);
            col[j] = __ldg(P.obj_colors + vo + vi[j]);
        }
    } else {
        const float* hv = P.hand_verts + 3 * (size_t)view * P.n_hv;
        const uchar4* hc = P.hand_colors + (size_t)P.hand_tex[view] * P.n_hv;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            p[j][0] = hv[3 * vi[j]]; p[j][1] = hv[3 * vi[j] + 1]; p[j][2] = hv[3 * vi[j] + 2];
            col[j] = __ldg(hc + vi[j]);
        }
    }
    const float inv = __fdiv_rn(1.0f, num);
    const float w0 = __fmul_rn(a[0], inv), w1 = __fmul_rn(a[1], inv), w2 = __fmul_rn(a[2], inv);
    const float e1x = __fsub_rn(p[1][0], p[0][0]), e1y = __fsub_rn(p[1][1], p[0][1]), e1z = __fsub_rn(p[1][2], p[0][2]);
    const float e2x = __fsub_rn(p[2][0], p[0][0]), e2y = __fsub_rn(p[2][1], p[0][1]), e2z = __fsub_rn(p[2][2], p[0][2]);
    const float nx = __fmaf_rn(e1y, e2z, -__fmul_rn(e1z, e2y));
    const float ny = __fmaf_rn(e1z, e2x, -__fmul_rn(e1x, e2z));
    const float nz = __fmaf_rn(e1x, e2y, -__fmul_rn(e1y, e2x));
    const float rx = __fdiv_rn(__fsub_rn(__fadd_rn((float)px, 0.5f), P.cx), P.fx);
    const float ry = __fdiv_rn(__fsub_rn(__fadd_rn((float)py, 0.5f), P.cy), P.fy);
    const float Px = __fmul_rn(rx, z), Py = __fmul_rn(ry, z), Pz = z;
    const float d2 = __fmaf_rn(Px, Px, __fmaf_rn(Py, Py, __fmul_rn(Pz, Pz)));
    const float n2 = __fmaf_rn(nx, nx, __fmaf_rn(ny, ny, __fmul_rn(nz, nz)));
    const float ndp = fabsf(__fmaf_rn(nx, Px, __fmaf_rn(ny, Py, __fmul_rn(nz, Pz))));
    const float den = __fsqrt_rn(__fmul_rn(n2, d2));
    const float cosv = den > 0.0f ? __fdiv_rn(ndp, den) : 0.0f;
    const float sh = __fmaf_rn(__fmul_rn(P.diffuse, P.light[view]), __fdiv_rn(cosv, d2), P.ambient);
    const float c0[3] = {(float)col[0].x, (float)col[0].y, (float)col[0].z};
    const float c1[3] = {(float)col[1].x, (float)col[1].y, (float)col[1].z};
    const float c2[3] = {(float)col[2].x, (float)col[2].y, (float)col[2].z};
    uint8_t o[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float cc = __fmaf_rn(w2, c2[ch], __fmaf_rn(w1, c1[ch], __fmul_rn(w0, c0[ch])));
        float v = __fmul_rn(cc, sh);
        if (!(v >= 0.0f)) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        o[ch] = (uint8_t)__float2int_rn(v);
    }
    PixelOut r;
    r.rgba = make_uchar4(o[0], o[1], o[2], 255);
    r.depth = z;
    r.seg = is_obj ? 2 : 1;
    return r;
}

// Phase 1, one thread per PX consecutive pixels of a row (PX = 4 when W % 4 == 0, else 1): 32-byte key loads; background
// pixels are finished here -- crop fetch with 32-bit divisions (operands < 2^30), one 16-byte RGBA store, one 16-byte
// depth store and one 4-byte seg store per thread, i.e. 512 + 512 + 128 contiguous bytes per warp; covered pixels are
// appended to a CTA-local list.  Phase 2: the CTA's covered pixels are dealt one per thread, so the long shading path runs
// on fully populated warps instead of on the few lanes of each row segment that touch the hand or the object
// (~9 % of the frame); their outputs overwrite the placeholders of phase 1 after the barrier.
template <int PX, int G>
__global__ void __launch_bounds__(256)
raster_resolve_kernel(const __grid_constant__ RasterParams P) {
    __shared__ unsigned hit_prim[256 * PX * G];
    __shared__ unsigned short hit_px[256 * PX * G];
    __shared__ int n_hit;
    const int view = blockIdx.y;
    const int cta_p0 = blockIdx.x * 256 * PX * G;
    const int npx = P.W * P.H;
    if (threadIdx.x == 0) n_hit = 0;
    __syncthreads();
    unsigned long long* kbase = P.keys + (size_t)view * npx;
    const size_t obase = (size_t)view * npx;
    {
        // G independent groups of PX pixels per thread (group g of a warp = 32 * PX consecutive pixels): all key loads
        // first, then all background fetches, then the stores -- G * (PX / 2 + PX) loads in flight per thread
        unsigned long long k[G][PX];
        uchar4 c[G][PX];
        int p0[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            p0[g] = cta_p0 + (g * 256 + threadIdx.x) * PX;
#pragma unroll
            for (int j = 0; j < PX; ++j) k[g][j] = kEmptyKey;
            if (p0[g] < npx) {
                if constexpr (PX == 4) {
                    const ulonglong2 k01 = *reinterpret_cast<const ulonglong2*>(kbase + p0[g]),
                                     k23 = *reinterpret_cast<const ulonglong2*>(kbase + p0[g] + 2);
                    k[g][0] = k01.x; k[g][1] = k01.y; k[g][2] = k23.x; k[g][3] = k23.y;
                } else {
                    k[g][0] = kbase[p0[g]];
                }
            }
        }
        const int32_t* sel = P.bg_sel ? P.bg_sel + 5 * (size_t)view : nullptr;
        const bool has_bg = sel && P.bgs && sel[0] >= 0;
        // background texels are fetched for every pixel, covered or not (phase 2 overwrites the covered ones): the fetch
        // does not wait for the key loads, so the L2 round trips of a thread overlap instead of chaining
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int j = 0; j < PX; ++j) c[g][j] = make_uchar4((uint8_t)P.bg_r, (uint8_t)P.bg_g, (uint8_t)P.bg_b, 0);
            if (has_bg && p0[g] < npx) {
                const int py = (int)fastdiv((unsigned)p0[g], P.div_w_m, P.div_w_s), px0 = p0[g] - py * P.W;
                const unsigned sy = (unsigned)sel[2] + fastdiv((unsigned)(2 * py + 1) * (unsigned)sel[4], P.div_2h_m, P.div_2h_s);
                const uint8_t* bg_row = P.bgs + (size_t)P.bg_ch * (((size_t)sel[0] * P.bg_h + sy) * P.bg_w);
                const unsigned bg_x0 = (unsigned)sel[1], bg_cw = (unsigned)sel[3];
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    const unsigned sx = bg_x0 + fastdiv((unsigned)(2 * (px0 + j) + 1) * bg_cw, P.div_2w_m, P.div_2w_s);
                    if (P.bg_ch == 4) {
                        c[g][j] = __ldg(reinterpret_cast<const uchar4*>(bg_row) + sx);
                        c[g][j].w = 0;
                    } else {
                        const uint8_t* src = bg_row + 3 * (size_t)sx;
                        c[g][j] = make_uchar4(src[0], src[1], src[2], 0);
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g) {
            if (p0[g] >= npx) continue;
#pragma unroll
            for (int j = 0; j < PX; ++j) {
                if (k[g][j] != kEmptyKey) {
                    const int slot = atomicAdd(&n_hit, 1);
                    hit_px[slot] = (unsigned short)((g * 256 + threadIdx.x) * PX + j);
                    hit_prim[slot] = (unsigned)(k[g][j] & 0xffffffffull);
                }
            }
            const size_t o = obase + p0[g];
            if constexpr (PX == 4) {
                if (P.rgba) {
                    uint4 v;
                    v.x = *reinterpret_cast<unsigned*>(&c[g][0]); v.y = *reinterpret_cast<unsigned*>(&c[g][1]);
                    v.z = *reinterpret_cast<unsigned*>(&c[g][2]); v.w = *reinterpret_cast<unsigned*>(&c[g][3]);
                    *reinterpret_cast<uint4*>(P.rgba + 4 * o) = v;
                }
                if (P.depth) *reinterpret_cast<float4*>(P.depth + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.seg) *reinterpret_cast<unsigned*>(P.seg + o) = 0u;
            } else {
                if (P.rgba) reinterpret_cast<uchar4*>(P.rgba)[o] = c[g][0];
                if (P.depth) P.depth[o] = 0.0f;
                if (P.seg) P.seg[o] = 0;
            }
        }
    }
    __syncthreads();
    // Phase 2: the CTA's covered pixels, one per thread, so the long shading path runs on fully populated warps; it
    // overlaps with the streaming phase 1 of the other CTAs resident on the SM (measured: a separate grid-wide shading
    // kernel over a compacted list is slower in total, 19 + 18 us against 30 us per 64 views).
    const int total = n_hit;
    if (total == 0) return;
    const int oid = P.obj_id[view];
    int n_of = 0, n_ov = 0;
    if (oid >= 0) {
        n_of = P.face_off[oid + 1] - P.face_off[oid];
        n_ov = P.vert_off[oid + 1] - P.vert_off[oid];
    }
    for (int i = threadIdx.x; i < total; i += 256) {
        const int p = cta_p0 + hit_px[i];
        const int py = (int)fastdiv((unsigned)p, P.div_w_m, P.div_w_s), px = p - py * P.W;
        const PixelOut r = shade_pixel(P, view, oid, n_ov, n_of, px, py, hit_prim[i]);
        kbase[p] = kEmptyKey;  // leave the key buffer empty for the next chunk
        const size_t o = obase + p;
        if (P.rgba) reinterpret_cast<uchar4*>(P.rgba)[o] = r.rgba;
        if (P.depth) P.depth[o] = r.depth;
        if (P.seg) P.seg[o] = r.seg;
    }
}

static int max_hand_obj_verts(const ab_scene* s) {
    int m = 0;
    for (int i = 0; i < s->n_obj; ++i) m = max(m, s->obj_vert_off_host[i + 1] - s->obj_vert_off_host[i]);
    return m;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Chunks alternate between the caller's stream and one library-owned auxiliary stream per device (each with its own
// scratch set), so the set-up-bound triangle pass of one chunk shares the SMs with the write-bound resolve pass of the
// other and the tails of the kernels overlap.  Fork / join is by events, so the call stays asynchronous and ordered on
// the caller's stream (and capturable into a CUDA graph).
constexpr int kMaxDevices = 16;
constexpr int kMaxSets = 4;
struct AuxStreams {
    std::mutex mu;
    cudaStream_t aux[kMaxDevices][kMaxSets] = {};
    cudaEvent_t fork[kMaxDevices] = {}, join[kMaxDevices][kMaxSets] = {};
};
static AuxStreams g_aux;

constexpr int kWsSets = 4;  // scratch sets every workspace is sized for
static std::atomic<int> g_sets{0};
static int raster_sets() {  // chunks in flight: 4 by default, AB_RASTER_STREAMS / ab_set_raster_streams override (1..4)
    int n = g_sets.load(std::memory_order_relaxed);
    if (n == 0) {
        const char* e = getenv("AB_RASTER_STREAMS");
        n = e ? atoi(e) : kWsSets;
        n = n < 1 ? 1 : (n > kWsSets ? kWsSets : n);
        g_sets.store(n, std::memory_order_relaxed);
    }
    return n;
}

static size_t scratch_set_bytes(int chunk, int npx, int pv_stride) {
    return align_up((size_t)chunk * npx * 8, 256) + align_up((size_t)chunk * pv_stride * 16, 256);
}

}  // namespace ab

extern "C" int ab_set_raster_streams(int n) {
    AB_REQUIRE(n >= 1 && n <= ab::kWsSets, "n must be in 1..4");
    ab::g_sets.store(n, std::memory_order_relaxed);
    return AB_OK;
}

extern "C" uint64_t ab_render_workspace_bytes(const ab_scene* scene, const ab_camera* cam, int chunk) {
    if (!scene || !cam || chunk <= 0 || cam->width <= 0 || cam->height <= 0) return 0;
    if (scene->n_obj > 0 && !scene->obj_vert_off_host) return 0;
    const int pv_stride = (int)ab::align_up((size_t)ab::max_hand_obj_verts(scene) + scene->n_hand_verts, 8);
    return (size_t)ab::kWsSets * ab::scratch_set_bytes(chunk, cam->width * cam->height, pv_stride);
}

extern "C" int ab_render_batch(const ab_scene* scene, const ab_camera* cam, int batch, int chunk,
                               const float* hand_verts, const int32_t* hand_tex, const int32_t* obj_id,
                               const int32_t* obj_id_host, const float* obj_pose, const float* light,
                               const int32_t* bg_sel, uint8_t* rgba, float* depth, uint8_t* seg, void* ws,
                               void* stream) {
    using namespace ab;
    AB_REQUIRE(scene && cam, "null scene / camera");
    AB_REQUIRE(batch >= 0 && chunk > 0, "bad batch / chunk");
    AB_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 4096 && cam->height <= 4096, "bad image size");
    AB_REQUIRE(scene->n_obj >= 0 && scene->n_obj <= kMaxObjects, "n_obj out of range (max 64)");
    AB_REQUIRE(scene->n_obj == 0 || (scene->obj_verts && scene->obj_faces && scene->obj_colors &&
                                     scene->obj_vert_off_host && scene->obj_face_off_host), "null object arrays");
    AB_REQUIRE(!scene->bgs || (scene->bg_w > 0 && scene->bg_h > 0 && scene->bg_w <= 65535 && scene->bg_h <= 65535), "bad background size");
    AB_REQUIRE(scene->n_hand_verts > 0 && scene->n_hand_faces > 0 && scene->n_hand_tex > 0 && scene->hand_faces &&
                   scene->hand_colors, "bad hand mesh");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(hand_verts && hand_tex && obj_id && obj_pose && light && ws, "null per-view input / workspace");
    AB_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    AB_REQUIRE(!rgba || ((uintptr_t)rgba & 3) == 0, "rgba must be 4-byte aligned");
    AB_REQUIRE(!depth || ((uintptr_t)depth & 3) == 0, "depth must be 4-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;

    RasterParams P;
    P.obj_verts = scene->obj_verts;
    P.obj_faces = reinterpret_cast<const int4*>(scene->obj_faces);
    P.obj_colors = reinterpret_cast<const uchar4*>(scene->obj_colors);
    P.hand_faces = reinterpret_cast<const int4*>(scene->hand_faces);
    P.hand_colors = reinterpret_cast<const uchar4*>(scene->hand_colors);
    P.bgs = scene->bgs;
    P.n_obj = scene->n_obj; P.n_hv = scene->n_hand_verts; P.n_hf = scene->n_hand_faces; P.n_tex = scene->n_hand_tex;
    P.n_bg = scene->n_bg; P.bg_h = scene->bg_h; P.bg_w = scene->bg_w;
    P.bg_ch = scene->bg_channels == 4 ? 4 : 3;
    AB_REQUIRE(scene->bg_channels == 0 || scene->bg_channels == 3 || scene->bg_channels == 4, "bg_channels must be 3 or 4");
    AB_REQUIRE(P.bg_ch == 3 || ((uintptr_t)scene->bgs & 3) == 0, "RGBX backgrounds must be 4-byte aligned");
    make_fastdiv((unsigned)cam->width, &P.div_w_m, &P.div_w_s);
    make_fastdiv(2u * (unsigned)cam->width, &P.div_2w_m, &P.div_2w_s);
    make_fastdiv(2u * (unsigned)cam->height, &P.div_2h_m, &P.div_2h_s);
    P.W = cam->width; P.H = cam->height;
    P.fx = cam->fx; P.fy = cam->fy; P.cx = cam->cx; P.cy = cam->cy; P.znear = cam->znear;
    P.ambient = cam->ambient; P.diffuse = cam->diffuse; P.cull = cam->cull_backface;
    P.bg_r = cam->bg_r; P.bg_g = cam->bg_g; P.bg_b = cam->bg_b;
    int scene_max_ov = 0, scene_max_of = 0;
    for (int i = 0; i <= scene->n_obj; ++i) {
        P.vert_off[i] = scene->n_obj ? scene->obj_vert_off_host[i] : 0;
        P.face_off[i] = scene->n_obj ? scene->obj_face_off_host[i] : 0;
        if (i > 0) {
            AB_REQUIRE(P.vert_off[i] >= P.vert_off[i - 1] && P.face_off[i] >= P.face_off[i - 1], "offsets not sorted");
            scene_max_ov = max(scene_max_ov, P.vert_off[i] - P.vert_off[i - 1]);
            scene_max_of = max(scene_max_of, P.face_off[i] - P.face_off[i - 1]);
        }
    }
    const int npx = P.W * P.H;
    P.pv_stride = (int)align_up((size_t)scene_max_ov + P.n_hv, 8);
    const size_t set_bytes = scratch_set_bytes(chunk, npx, P.pv_stride);
    const size_t keys_bytes = align_up((size_t)chunk * npx * 8, 256);
    const int n_chunks = cdiv(batch, chunk);
    const int sets = min(raster_sets(), n_chunks);
    const bool dual = sets > 1;
    int dev = 0;
    std::unique_lock<std::mutex> lock(g_aux.mu, std::defer_lock);
    if (dual) {
        AB_CUDA(cudaGetDevice(&dev));
        AB_REQUIRE(dev < kMaxDevices, "device index out of range");
        lock.lock();  // the per-device events are reused by every call
        if (!g_aux.fork[dev]) AB_CUDA(cudaEventCreateWithFlags(&g_aux.fork[dev], cudaEventDisableTiming));
        for (int i = 1; i < sets; ++i) {
            if (!g_aux.aux[dev][i]) {
                AB_CUDA(cudaStreamCreateWithFlags(&g_aux.aux[dev][i], cudaStreamNonBlocking));
                AB_CUDA(cudaEventCreateWithFlags(&g_aux.join[dev][i], cudaEventDisableTiming));
            }
        }
    }
    for (int i = 0; i < sets; ++i) AB_CUDA(cudaMemsetAsync((char*)ws + i * set_bytes, 0xFF, keys_bytes, st));
    if (dual) {
        AB_CUDA(cudaEventRecord(g_aux.fork[dev], st));
        for (int i = 1; i < sets; ++i) AB_CUDA(cudaStreamWaitEvent(g_aux.aux[dev][i], g_aux.fork[dev], 0));
    }
    const cudaStream_t caller = st;
    int chunk_idx = 0;
    for (int v0 = 0; v0 < batch; v0 += chunk, ++chunk_idx) {
        const int n = min(chunk, batch - v0);
        const int set = chunk_idx % sets;
        st = set ? g_aux.aux[dev][set] : caller;
        P.keys = (unsigned long long*)((char*)ws + set * set_bytes);
        P.pv = (int4*)((char*)ws + set * set_bytes + keys_bytes);
        int max_ov = scene_max_ov, max_of = scene_max_of;
        if (obj_id_host) {
            max_ov = max_of = 0;
            for (int i = 0; i < n; ++i) {
                const int o = obj_id_host[v0 + i];
                AB_REQUIRE(o < scene->n_obj, "obj_id out of range");
                if (o >= 0) {
                    max_ov = max(max_ov, P.vert_off[o + 1] - P.vert_off[o]);
                    max_of = max(max_of, P.face_off[o + 1] - P.face_off[o]);
                }
            }
        }
        P.max_ov = max_ov; P.max_of = max_of;
        P.n_views = n;
        P.hand_verts = hand_verts + (size_t)v0 * P.n_hv * 3;
        P.hand_tex = hand_tex + v0;
        P.obj_id = obj_id + v0;
        P.obj_pose = obj_pose + (size_t)v0 * 16;
        P.light = light + v0;
        P.bg_sel = bg_sel ? bg_sel + (size_t)v0 * 5 : nullptr;
        P.rgba = rgba ? rgba + (size_t)v0 * npx * 4 : nullptr;
        P.depth = depth ? depth + (size_t)v0 * npx : nullptr;
        P.seg = seg ? seg + (size_t)v0 * npx : nullptr;
        {
            StageTimer tm(AB_STAGE_RASTER_VERTEX, st);
            raster_vertex_kernel<<<dim3(cdiv(max_ov + P.n_hv, 256), n), 256, 0, st>>>(P);
        }
        {
            StageTimer tm(AB_STAGE_RASTER_TRIANGLE, st);
            // 6 resident CTAs per SM (40 registers, a few spilled words): with several chunks in flight the pass is
            // latency-bound, and occupancy buys more than the spills cost (measured 1.15 -> 1.24 M views/s against 4 CTAs)
            static const int minb = getenv("AB_TRI_MINB") ? atoi(getenv("AB_TRI_MINB")) : 6;
            const dim3 grid(cdiv(max_of + P.n_hf, kTriThreads), n);
            if (minb == 4) raster_triangle_kernel<4><<<grid, kTriThreads, 0, st>>>(P);
            else if (minb == 8) raster_triangle_kernel<8><<<grid, kTriThreads, 0, st>>>(P);
            else raster_triangle_kernel<6><<<grid, kTriThreads, 0, st>>>(P);
        }
        {
            StageTimer tm(AB_STAGE_RASTER_RESOLVE, st);
            // 4 pixels per thread needs 16-byte aligned output rows: W % 4 == 0 and 16 / 16 / 4-byte aligned bases
            const bool vec = (P.W % 4 == 0) && (((uintptr_t)P.rgba & 15) == 0) && (((uintptr_t)P.depth & 15) == 0) &&
                             (((uintptr_t)P.seg & 3) == 0);
            if (vec) raster_resolve_kernel<4, 1><<<dim3(cdiv(npx, 1024), n), 256, 0, st>>>(P);
            else raster_resolve_kernel<1, 1><<<dim3(cdiv(npx, 256), n), 256, 0, st>>>(P);

        }
        count_launch(3);
        int rc = check_launch("ab_render_batch");
        if (rc) return rc;
    }
    for (int i = 1; i < sets; ++i) {
        AB_CUDA(cudaEventRecord(g_aux.join[dev][i], g_aux.aux[dev][i]));
        AB_CUDA(cudaStreamWaitEvent(caller, g_aux.join[dev][i], 0));
    }
    return AB_OK;
}
