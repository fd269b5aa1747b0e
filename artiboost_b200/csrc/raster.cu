// Batched hand+object view rasteriser (sm_100a): tile-binned, z-buffer in shared memory.
//
// Replaces Renderer.__call__ (anakin/utils/renderer.py:101-123) over pyrender's OffscreenRenderer
// (anakin/utils/frender_utils.py:179-205): B views per call instead of one GL draw + two glReadPixels per call.
// The rule set (fixed-point snapping, top-left fill rule, 1/z depth, z-test key, shading) is the one stated at the
// top of oracle/raster.c; every fp32 operation of the rules is individually rounded in the same order (this file is
// compiled with -fmad=false and uses explicit __fmaf_rn where the rules say "fma").
//
// Meshes are walked as patches (patches.cu: <= 32 faces over <= 32 vertices, bounding sphere + normal cone).
// Two kernels per group of views:
//   1. raster_bin_kernel   object patches, one thread each: sphere centre + cone axis into camera space, conservative
//                          back-face test of the whole patch (margin derived from the 24.8 snapping error and the
//                          patch's worst perimeter / area ratio), conservative screen box of the sphere -> appended to
//                          the list of every 64x64 tile it may touch.  Hand patches (vertices change per view), one
//                          warp each: exact box of the projected vertices by warp reductions.
//   2. raster_tile_kernel  one CTA per (view, tile).  The tile's z-buffer (64-bit key = depth bits << 32 | primitive
//                          id per pixel, 32 KB) lives in SHARED memory.  Warps pull patches from the tile's list; per
//                          patch a lane per vertex (model matrix, projection, snapping -> 16 B in a per-warp slab),
//                          then a lane per face (integer set-up, cull, bbox clipped to the tile; exact integer edge
//                          functions; small boxes walked by the lane itself, large ones by the whole warp), z-test by
//                          a shared-memory 64-bit atomic min.  Then the CTA streams the tile out: background crop /
//                          flat colour, zero depth and seg with 16-byte stores, covered pixels compacted into a list
//                          and shaded on full warps (winner re-set-up, flat normal, point light).
// Nothing but the per-view inputs, the tile lists (a few KB per view) and the output image (9 B / pixel) touches HBM or
// L2: no projected-vertex scratch, no key buffer, no memset of either.  Tiles without geometry (most of the frame) skip
// the z-buffer altogether and are pure write streams that overlap with the set-up work of the busy tiles on the same SM.
#include <limits.h>
#include <stdlib.h>

#include <algorithm>
#include <utility>
#include <vector>

#include "common.cuh"

namespace ab {

constexpr int kMaxObjects = 64;
constexpr unsigned long long kEmptyKey = ~0ull;
constexpr int kTile = 64;                 // tile edge in pixels
constexpr int kTilePx = kTile * kTile;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr unsigned kHandBit = 0x8000u;    // list entries: patch index (relative to its mesh) | kHandBit for hand patches
constexpr unsigned kMultiBit = 0x4000u;   //               | kMultiBit when the binning box spans several tiles (the tile kernel then
constexpr unsigned kIndexMask = 0x3fffu;  //               tests the exact box of the projected vertices before the face phase)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kBigCap = 512;              // queue of triangles of >= 64 px per tile (more than that: the tile is redone slowly, see below)
constexpr float kSnapEps = 0.0032f;       // bound on the displacement (pixels) of a projected vertex by fp32 evaluation + 24.8 snapping

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

constexpr int kClasses = 32;    // work classes of the tile kernel's launch order: 0 = most work ... 30, 31 = empty tiles
constexpr int kMaxOrder = 256;  // frames of up to 1024 x 1024 pixels get the ranked tile order (below)

struct RasterParams {
    // scene
    const float* obj_verts;
    const int4* obj_faces;
    const uchar4* obj_colors;
    const int4* hand_faces;
    const uchar4* hand_colors;
    const uint8_t* bgs;
    // patches
    const float4* op_pos;     // [P][32]
    const uint32_t* op_face;  // [P][32]
    const int32_t* op_prim;   // [P][32]
    const float4* op_bound;   // [P][3]
    const int32_t* hp_vid;
    const uint32_t* hp_face;
    const int32_t* hp_prim;
    int n_hp;                 // hand patches
    int n_obj, n_hv, n_hf, n_tex, n_bg, bg_h, bg_w, bg_ch;
    unsigned div_w_m, div_2w_m, div_2h_m;  // exact division by W, 2W, 2H as multiply + shift
    int div_w_s, div_2w_s, div_2h_s;
    // camera
    int W, H;
    float fx, fy, cx, cy, znear, ambient, diffuse;
    int cull;
    int bg_r, bg_g, bg_b;
    // tiles
    int tiles_x, tiles_y, n_tiles, list_cap;
    int near_tiles;             // tiles of a view rendered in the first pass over the views (see raster_tile_kernel)
    unsigned div_n1_m, div_n2_m, div_tx_m;  // exact division by near_tiles, n_tiles - near_tiles, tiles_x
    int div_n1_s, div_n2_s, div_tx_s;
    int use_order;              // 0: more than kMaxOrder tiles, natural order
    unsigned short tile_order[kMaxOrder];  // tiles by distance from the principal point
    int* bin_count;             // [views][n_tiles]
    unsigned short* bin_list;   // [views][n_tiles][list_cap]
    // work-ordered launch (see raster_tile_kernel): the tiles of the group filed into classes by estimated work
    int sorted;                 // 0: launch order from the ranked tile order above
    int class_width;            // work units (list entries) per class
    int area_shift, area_mul;   // covered-pixel estimate = summed patch area >> area_shift, worth area_mul / 64 list entries per pixel
    int class_cap;              // entries per class list = views * n_tiles of a full group
    int* class_total;           // [kClasses] tiles per class, cleared before the binning kernel
    unsigned* class_list;       // [kClasses][class_cap] bin index (view * n_tiles + tile)
    unsigned div_nt_m;          // exact division by n_tiles
    int div_nt_s;
    // per-view inputs (already offset to the group's first view)
    const float* hand_verts;
    const int32_t* hand_tex;
    const int32_t* obj_id;
    const float* obj_pose;
    const float* light;
    const int32_t* bg_sel;
    // outputs (offset to the group)
    uint8_t* rgba;
    float* depth;
    uint8_t* seg;
    int n_views;
    int vert_off[kMaxObjects + 1];
    int face_off[kMaxObjects + 1];
    int patch_off[kMaxObjects + 1];
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
    float iz = __frcp_rn(Z);  // correctly rounded, i.e. the rule's 1.0f / Z
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

__device__ __forceinline__ int floordiv256(int v) { return v >> 8; }

// ---- binning -------------------------------------------------------------------------------------------------
constexpr int kMaxTiles = 4096;  // 4096 x 4096 pixels

__device__ __forceinline__ void bin_append(const RasterParams& P, int* cnt, int* area, int view, int px_lo, int px_hi, int py_lo,
                                           int py_hi, unsigned entry) {
    const int tx_lo = px_lo / kTile, tx_hi = px_hi / kTile, ty_lo = py_lo / kTile, ty_hi = py_hi / kTile;
    if (tx_lo != tx_hi || ty_lo != ty_hi) entry |= kMultiBit;
    for (int ty = ty_lo; ty <= ty_hi; ++ty)
        for (int tx = tx_lo; tx <= tx_hi; ++tx) {
            const int tile = ty * P.tiles_x + tx;
            const int slot = atomicAdd(cnt + tile, 1);  // shared memory: this CTA owns every list of the view
            if (P.sorted)  // screen area of the patch inside the tile: with the list length, the tile's work estimate
                atomicAdd(area + tile, (min(px_hi, tx * kTile + kTile - 1) - max(px_lo, tx * kTile) + 1) *
                                           (min(py_hi, ty * kTile + kTile - 1) - max(py_lo, ty * kTile) + 1));
            P.bin_list[((size_t)view * P.n_tiles + tile) * P.list_cap + slot] = (unsigned short)entry;
        }
}

// Object patch: everything happens on the patch's bounding sphere (centre s, radius r, camera space) and its normal
// cone (unit axis a, every face normal n within acos(c) of it).
//  * Back-face test.  With p0, p1, p2 the camera-space corners of a face and n = (p1-p0) x (p2-p0) = 2 A n^ its winding
//    normal, the projected signed area is exactly  area2 = fx fy (p0 . n) / (z0 z1 z2)  pixels^2, and the rule culls
//    area2 > 0.  For p in the sphere and n^ in the cone,  n^ . p >= |s| cos(phi + theta) - r =: Bnd  (phi = angle(a, s),
//    theta = acos c; needs phi + theta <= pi, hence c > 0 and cos phi > 0).  Snapping (and fp32 evaluation) moves each
//    projected corner by at most eps pixels, which changes area2 by at most 2 eps (projected perimeter) + 4 eps^2, and a
//    world-space segment of length L projects to at most L g pixels, g = max(fx, fy) (1 + T) / zmin, T = the largest
//    |(x, y)| / z over the sphere.  So every face of the patch is certainly culled when
//        Bnd > 1.5 zmax^3 / (fx fy) (eps q g + 4 eps^2 ia) + 2e-5,    q = max perimeter / area, ia = max 1 / (2 area).
//  * Screen box.  x / z over the box [sx - r, sx + r] x [zmin, zmax] is extremal at its corners; 0.05 px of slack covers
//    snapping and fp32.  Pixels whose CENTRE lies in the box are the only ones a face of the patch can cover.
// One CTA per view: its threads take the object patches, its warps the hand patches (four at a time, all loads of the
// four issued before any is used: the vertex id -> position gathers are two dependent round trips); the per-tile counters
// live in shared memory (no global atomics, and no clearing between calls: every count is written).
__global__ void __launch_bounds__(kThreads)
raster_bin_kernel(const __grid_constant__ RasterParams P) {
    __shared__ int cnt[kMaxTiles];
    __shared__ int area[kMaxTiles];
    const int view = blockIdx.x;
    const int oid = P.obj_id[view];
    for (int i = threadIdx.x; i < P.n_tiles; i += kThreads) { cnt[i] = 0; area[i] = 0; }
    float M[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) M[i] = P.obj_pose[16 * (size_t)view + i];
    __syncthreads();
    const int n_op = oid >= 0 ? P.patch_off[oid + 1] - P.patch_off[oid] : 0;
    const float4* bounds = P.op_bound + 3 * (size_t)(oid >= 0 ? P.patch_off[oid] : 0);
#pragma unroll 2
    for (int i = threadIdx.x; i < n_op; i += kThreads) {
        const float4 b0 = __ldg(bounds + 3 * i), b1 = __ldg(bounds + 3 * i + 1), b2 = __ldg(bounds + 3 * i + 2);
        float s[3];
        xform(M, b0.x, b0.y, b0.z, s);
        const float r = b0.w;
        const float zmin = s[2] - r, zmax = s[2] + r;
        if (P.cull && b1.w > 0.0f && zmin > 1e-4f) {
            const float ax = M[0] * b1.x + M[1] * b1.y + M[2] * b1.z;
            const float ay = M[4] * b1.x + M[5] * b1.y + M[6] * b1.z;
            const float az = M[8] * b1.x + M[9] * b1.y + M[10] * b1.z;
            const float ls = sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
            const float la = sqrtf(ax * ax + ay * ay + az * az);  // 1 for a rigid pose; a scaled pose keeps the test valid
            const float cos_phi = (ax * s[0] + ay * s[1] + az * s[2]) / (ls * la);
            if (cos_phi > 0.0f) {
                const float c = b1.w;
                const float sin_phi = sqrtf(fmaxf(0.0f, 1.0f - cos_phi * cos_phi));
                const float sin_th = sqrtf(fmaxf(0.0f, 1.0f - c * c));
                const float bnd = ls * (cos_phi * c - sin_phi * sin_th) - r;
                const float T = (sqrtf(s[0] * s[0] + s[1] * s[1]) + r) / zmin;
                const float g = fmaxf(P.fx, P.fy) * (1.0f + T) / zmin;
                const float need = 1.5f * zmax * zmax * zmax / (P.fx * P.fy) * (kSnapEps * b2.x * g + 4.0f * kSnapEps * kSnapEps * b2.y) + 2e-5f;
                if (bnd > need) continue;  // inf / NaN in `need` (degenerate faces) compares false: not culled
            }
        }
        int px_lo = 0, px_hi = P.W - 1, py_lo = 0, py_hi = P.H - 1;
        if (zmin > 1e-4f) {
            const float ilo = 1.0f / zmin, ihi = 1.0f / zmax;
            const float xl = s[0] - r, xh = s[0] + r, yl = s[1] - r, yh = s[1] + r;
            float umin = P.fx * fminf(xl * ilo, xl * ihi) + P.cx, umax = P.fx * fmaxf(xh * ilo, xh * ihi) + P.cx;
            float vmin = P.fy * fminf(yl * ilo, yl * ihi) + P.cy, vmax = P.fy * fmaxf(yh * ilo, yh * ihi) + P.cy;
            umin = fminf(fmaxf(ceilf(umin - 0.55f), -1.0e6f), 1.0e6f); umax = fminf(fmaxf(floorf(umax - 0.45f), -1.0e6f), 1.0e6f);
            vmin = fminf(fmaxf(ceilf(vmin - 0.55f), -1.0e6f), 1.0e6f); vmax = fminf(fmaxf(floorf(vmax - 0.45f), -1.0e6f), 1.0e6f);
            px_lo = max((int)umin, 0); px_hi = min((int)umax, P.W - 1);
            py_lo = max((int)vmin, 0); py_hi = min((int)vmax, P.H - 1);
            if (px_lo > px_hi || py_lo > py_hi) continue;
        }
        bin_append(P, cnt, area, view, px_lo, px_hi, py_lo, py_hi, (unsigned)i);
    }
    const int lane = threadIdx.x & 31;
    const float* hverts = P.hand_verts + 3 * (size_t)view * P.n_hv;
    constexpr int kAhead = 4;
    for (int w0 = threadIdx.x >> 5; w0 < P.n_hp; w0 += kAhead * kWarps) {
        int v[kAhead];
        float X[kAhead], Y[kAhead], Z[kAhead];
#pragma unroll
        for (int k = 0; k < kAhead; ++k) {
            const int w = w0 + k * kWarps;
            v[k] = w < P.n_hp ? __ldg(P.hp_vid + w * 32 + lane) : -1;
        }
#pragma unroll
        for (int k = 0; k < kAhead; ++k) {
            X[k] = Y[k] = Z[k] = 0.0f;
            if (v[k] >= 0) { X[k] = hverts[3 * v[k]]; Y[k] = hverts[3 * v[k] + 1]; Z[k] = hverts[3 * v[k] + 2]; }
        }
#pragma unroll
        for (int k = 0; k < kAhead; ++k) {
            const int w = w0 + k * kWarps;
            int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
            if (v[k] >= 0) {
                const int4 p = project(P, X[k], Y[k], Z[k]);
                if (p.w) { mnx = mxx = p.x; mny = mxy = p.y; }  // faces with an invalid corner are never drawn
            }
            mnx = __reduce_min_sync(kFull, mnx); mxx = __reduce_max_sync(kFull, mxx);
            mny = __reduce_min_sync(kFull, mny); mxy = __reduce_max_sync(kFull, mxy);
            if (lane != 0 || w >= P.n_hp || mnx > mxx) continue;
            const int px_lo = max(floordiv256(mnx + 127), 0), px_hi = min(floordiv256(mxx - 128), P.W - 1);
            const int py_lo = max(floordiv256(mny + 127), 0), py_hi = min(floordiv256(mxy - 128), P.H - 1);
            if (px_lo > px_hi || py_lo > py_hi) continue;
            bin_append(P, cnt, area, view, px_lo, px_hi, py_lo, py_hi, (unsigned)w | kHandBit);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P.n_tiles; i += kThreads) {
        const int c = cnt[i];
        P.bin_count[(size_t)view * P.n_tiles + i] = c;
        if (P.sorted) {
            // File the tile under its work class (most work in class 0, empty tiles in the last class).  Work in list
            // entries: measured on B200 (tools/trace_raster.py), a tile CTA takes ~0.25 us per entry in the patch loop and
            // ~0.02 us per covered pixel in the shading pass; the covered pixels are estimated from the summed screen
            // boxes of the tile's patches.
            const int est = c + ((min(area[i] >> P.area_shift, kTilePx) * P.area_mul) >> 6);
            const int cls = c == 0 ? kClasses - 1 : kClasses - 2 - min(kClasses - 2, (est - 1) / P.class_width);
            const int pos = atomicAdd(P.class_total + cls, 1);
            P.class_list[(size_t)cls * P.class_cap + pos] = (unsigned)(view * P.n_tiles + i);
        }
    }
}

// ---- rule: triangle ----------------------------------------------------------------------------------------
// Triangles whose snapped bbox is under 64 px in both axes (all of them in valid ArtiBoost views) take the int32
// path: every factor of the edge functions is below 2^14 + 2^8 in magnitude, so products and their differences are
// exact in int32.  Larger ones take the int64 path.  Both produce the same integers, hence the same floats.
constexpr int kSmallExtent = 16384;  // 64 px in 24.8 fixed point

__device__ __forceinline__ long long area2_of(const int4& a, const int4& b, const int4& d, bool small) {
    if (small) return (long long)((b.x - a.x) * (d.y - a.y) - (d.x - a.x) * (b.y - a.y));
    return (long long)(b.x - a.x) * (long long)(d.y - a.y) - (long long)(d.x - a.x) * (long long)(b.y - a.y);
}

__device__ __forceinline__ long long edge64(const int* x, const int* y, int s, int i, long long px, long long py) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    const long long dx = x[i2] - x[i1], dy = y[i2] - y[i1];
    return (long long)s * (dx * (py - y[i1]) - dy * (px - x[i1]));
}

// z-test: key = depth bits << 32 | primitive id, minimum wins (nearest; ties to the lower id)
__device__ __forceinline__ void emit(unsigned long long* zbuf, int idx, float f0, float f1, float f2, float iz0, float iz1,
                                     float iz2, float sarea, unsigned prim) {
    const float num = __fmaf_rn(f2, iz2, __fmaf_rn(f1, iz1, __fmul_rn(f0, iz0)));
    const float z = __fdiv_rn(sarea, num);
    const unsigned long long k = ((unsigned long long)__float_as_uint(z) << 32) | prim;
    // shared memory has no native 64-bit min: compare-and-swap, entered only when the fragment is in front
    unsigned long long old = zbuf[idx];
    while (k < old) {
        const unsigned long long prev = atomicCAS(zbuf + idx, old, k);
        if (prev == old) break;
        old = prev;
    }
}

// ---- rule: shading -------------------------------------------------------------------------------------------
struct PixelOut { uchar4 rgba; float depth; uint8_t seg; };

// The winner passed set-up in the patch loop: recompute its edge values at this pixel the same way.  `z` is the
// depth of the z-test key (the bits the patch loop computed).
__device__ __forceinline__ PixelOut shade_pixel(const RasterParams& P, const float* __restrict__ M, int view, int oid,
                                                int n_of, int px, int py, unsigned f, float z) {
    int4 idx;
    const bool is_obj = (int)f < n_of;
    if (is_obj) idx = __ldg(P.obj_faces + P.face_off[oid] + f);
    else idx = __ldg(P.hand_faces + (f - n_of));
    const int vi[3] = {idx.x, idx.y, idx.z};
    float p[3][3];
    uchar4 col[3];
    if (is_obj) {
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
    const int4 a4 = project(P, p[0][0], p[0][1], p[0][2]), b4 = project(P, p[1][0], p[1][1], p[1][2]),
               d4 = project(P, p[2][0], p[2][1], p[2][2]);
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
    const float a0 = __fmul_rn(fe[0], iz[0]), a1 = __fmul_rn(fe[1], iz[1]), a2 = __fmul_rn(fe[2], iz[2]);
    const float num = __fmaf_rn(fe[2], iz[2], __fmaf_rn(fe[1], iz[1], a0));
    const float inv = __fdiv_rn(1.0f, num);
    const float w0 = __fmul_rn(a0, inv), w1 = __fmul_rn(a1, inv), w2 = __fmul_rn(a2, inv);
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

#ifdef AB_RASTER_TRACE
// Debug build only (tools/trace_raster.py): start / end time, SM, list length and covered pixels of every tile CTA.
__device__ unsigned long long g_trace[6 * 32768];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#endif

// ---- triangles of 64 px and more in either axis (close-ups; the cap fans of a mesh seen face-on) --------------------------
// int64 edge functions E_i(px, py) = A_i px + B_i py + C_i over the pixel centres of the triangle's box clipped to the tile.
// Such a triangle's box is up to the whole tile while the triangle itself is often a sliver, so the box is cut into 8 x 8
// pixel blocks first: a lane per block evaluates each edge at the block corner where it is largest (exact integers) and
// drops the block when one edge is negative there; the warp then walks the surviving blocks, two pixels per lane.  The
// integers that reach the z-test are the ones a per-pixel evaluation produces.  Kept out of line: it is rare, and inlined
// its registers cost the common path of the kernel (a tile crossed by 53 such triangles took 362 us when every pixel of
// every box was tested with three int64 products inside the patch loop: the whole launch waited for it).
__device__ __noinline__ void raster_big_triangle(unsigned long long* zbuf, int4 ca, int4 cb, int4 cd, int cx0, int cy0, int cw, int ch,
                                                 int tx0, int ty0, unsigned cprim, int lane) {
    const int vx[3] = {ca.x, cb.x, cd.x}, vy[3] = {ca.y, cb.y, cd.y};
    const float jz0 = __int_as_float(ca.z), jz1 = __int_as_float(cb.z), jz2 = __int_as_float(cd.z);
    const long long carea2 = area2_of(ca, cb, cd, false);
    const int s = carea2 > 0 ? 1 : -1;
    const float jarea = __ll2float_rn(carea2 > 0 ? carea2 : -carea2);
    long long A[3], B[3], C[3];   // E_i = A_i px + B_i py + C_i with px, py in 24.8 units; the bias of the fill rule folded into C
    int bias[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        const int dx = s * (vx[i2] - vx[i1]), dy = s * (vy[i2] - vy[i1]);  // |.| <= 2^23: no overflow
        bias[i] = ((dy < 0) || (dy == 0 && dx > 0)) ? 0 : 1;
        A[i] = -(long long)dy;
        B[i] = (long long)dx;
        C[i] = (long long)dy * vx[i1] - (long long)dx * vy[i1];
    }
    const int nbx = (cw + 7) >> 3, nby = (ch + 7) >> 3, nblk = nbx * nby;   // <= 64 blocks
    for (int b0 = 0; b0 < nblk; b0 += 32) {
        const int bi = b0 + lane;
        bool keep = false;
        int bx = 0, by = 0;
        if (bi < nblk) {
            by = bi / nbx; bx = bi - by * nbx;
            const int px0 = cx0 + 8 * bx, py0 = cy0 + 8 * by;
            const int px1 = min(px0 + 7, cx0 + cw - 1), py1 = min(py0 + 7, cy0 + ch - 1);
            keep = true;
#pragma unroll
            for (int i = 0; i < 3; ++i) {   // the largest value of edge i over the block's pixel centres
                const long long qx = 256ll * (A[i] > 0 ? px1 : px0) + 128, qy = 256ll * (B[i] > 0 ? py1 : py0) + 128;
                if (A[i] * qx + B[i] * qy + C[i] - bias[i] < 0) keep = false;
            }
        }
        unsigned live = __ballot_sync(kFull, keep);
        while (live) {
            const int l = __ffs(live) - 1;
            live &= live - 1;
            const int kbx = __shfl_sync(kFull, bx, l), kby = __shfl_sync(kFull, by, l);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int px = cx0 + 8 * kbx + (lane & 7), py = cy0 + 8 * kby + (lane >> 3) + 4 * j;
                if (px >= cx0 + cw || py >= cy0 + ch) continue;
                const long long qx = 256ll * px + 128, qy = 256ll * py + 128;
                const long long e0 = A[0] * qx + B[0] * qy + C[0], e1 = A[1] * qx + B[1] * qy + C[1], e2 = A[2] * qx + B[2] * qy + C[2];
                if (((e0 - bias[0]) | (e1 - bias[1]) | (e2 - bias[2])) >= 0)
                    emit(zbuf, (py - ty0) * kTile + (px - tx0), __ll2float_rn(e0), __ll2float_rn(e1), __ll2float_rn(e2), jz0, jz1, jz2,
                         jarea, cprim);
            }
        }
    }
}

// The triangles the patch loop queued (their projected vertices are gone with the warp's slab): a warp per triangle re-projects
// its three vertices the way the vertex phase did -- the same operations on the same inputs, hence the same integers -- and
// hands it to raster_big_triangle.  If the queue overflowed (more than kBigCap such triangles in one tile), every patch of the
// tile's list is searched again for them instead: slow, correct, and never seen on ArtiBoost views.
__device__ __noinline__ void raster_big_queue(const RasterParams& P, unsigned long long* zbuf, const unsigned* big_q, int n_q, bool overflow,
                                              int view, int bin, int count, int tx0, int ty0, int tx1, int ty1, const float* Msh) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int oid = P.obj_id[view];
    const int n_of = oid >= 0 ? P.face_off[oid + 1] - P.face_off[oid] : 0;
    const int poff = oid >= 0 ? P.patch_off[oid] : 0;
    const float* hverts = P.hand_verts + 3 * (size_t)view * P.n_hv;
    const unsigned short* list = P.bin_list + (size_t)bin * P.list_cap;
    const int total = overflow ? count * 32 : n_q;   // overflow: every (list entry, face lane) pair is a candidate
    for (int i = wid; i < total; i += kWarps) {
        const unsigned q = overflow ? ((unsigned)list[i >> 5] | ((unsigned)(i & 31) << 16)) : big_q[i];
        const unsigned entry = q & 0xffffu, fl = q >> 16;
        const bool hand = (entry & kHandBit) != 0;
        const size_t poffs = (size_t)(entry & kIndexMask) * 32u;
        const unsigned fw = hand ? __ldg(P.hp_face + poffs + fl) : __ldg(P.op_face + (size_t)poff * 32 + poffs + fl);
        if (fw == 0xFFFFFFFFu) continue;   // warp-uniform
        const int prim_local = hand ? __ldg(P.hp_prim + poffs + fl) : __ldg(P.op_prim + (size_t)poff * 32 + poffs + fl);
        int4 pv[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const unsigned vs = (fw >> (8 * j)) & 31u;
            float c[3] = {0.0f, 0.0f, 0.0f};
            bool on;
            if (!hand) {
                const float4 v4 = __ldg(P.op_pos + (size_t)poff * 32 + poffs + vs);
                on = v4.w != 0.0f;
                xform(Msh, v4.x, v4.y, v4.z, c);
            } else {
                const int v = __ldg(P.hp_vid + poffs + vs);
                on = v >= 0;
                if (on) { c[0] = hverts[3 * v]; c[1] = hverts[3 * v + 1]; c[2] = hverts[3 * v + 2]; }
            }
            pv[j] = project(P, c[0], c[1], c[2]);
            if (!on) pv[j].w = 0;
        }
        if (!(pv[0].w & pv[1].w & pv[2].w)) continue;
        const int minx = min(pv[0].x, min(pv[1].x, pv[2].x)), maxx = max(pv[0].x, max(pv[1].x, pv[2].x));
        const int miny = min(pv[0].y, min(pv[1].y, pv[2].y)), maxy = max(pv[0].y, max(pv[1].y, pv[2].y));
        if ((maxx - minx < kSmallExtent) && (maxy - miny < kSmallExtent)) continue;   // (overflow search) drawn by the patch loop
        const int x0 = max(floordiv256(minx + 127), tx0), y0 = max(floordiv256(miny + 127), ty0);
        const int x1 = min(floordiv256(maxx - 128), tx1), y1 = min(floordiv256(maxy - 128), ty1);
        if (x0 > x1 || y0 > y1) continue;
        const long long a2 = area2_of(pv[0], pv[1], pv[2], false);
        if (a2 == 0 || (a2 > 0 && P.cull)) continue;
        raster_big_triangle(zbuf, pv[0], pv[1], pv[2], x0, y0, x1 - x0 + 1, y1 - y0 + 1, tx0, ty0,
                            (unsigned)(prim_local + (hand ? n_of : 0)), lane);
    }
}

// ---- the tile kernel -----------------------------------------------------------------------------------------
// PX = pixels per thread in the output stream: 4 when W % 4 == 0 and the output rows are 16-byte aligned, else 1.
// BG4: backgrounds are RGBX (one aligned 32-bit load per pixel) instead of packed RGB.
#ifndef AB_TILE_CTAS
#define AB_TILE_CTAS 4   // CTAs per SM the register budget is cut for (64 registers); 3 and 5 measured slower
#endif
template <int PX, bool BG4>
__global__ void __launch_bounds__(kThreads, AB_TILE_CTAS)
raster_tile_kernel(const __grid_constant__ RasterParams P) {
    __shared__ __align__(16) unsigned long long zbuf[kTilePx];  // 32 KB
    __shared__ __align__(16) int4 slab[kWarps][32];             // projected vertices of the patch a warp is on
    __shared__ unsigned short hit_px[kTilePx];                  // covered pixels of the tile (tile-local index)
    __shared__ __align__(16) float Msh[12];
    __shared__ unsigned char rank_lane[kWarps][32];             // per warp: lane of the k-th box that has rows
    __shared__ __align__(16) int colmap[kTile];                 // background source column of every tile column
    __shared__ int n_hit, next_entry, n_bigq;
    __shared__ unsigned big_q[kBigCap];                        // (list entry | face lane << 16) of the triangles of >= 64 px

    // CTAs are handed out in launch order, so the order of the work decides how the grid ends.  Groups too large for the
    // work-ordered launch fall back to a static order: the tiles of a view ranked by their distance from the principal point
    // (ArtiBoost places the object there, yaml:13-20 CAMERA_Z_RANGE on the optical axis), the first `near_tiles` of every
    // view first, view by view, the outermost tiles of all views -- almost always pure background -- last.
    int view, tile;
    if (P.sorted) {
        // Work-ordered launch.  The binning kernel filed every tile of the group under a class by its estimated work.
        // Busy tiles are handed out most work first, so that the CTAs still running when the grid drains are the short
        // ones (a busy tile runs 10-50x longer than an empty one; in launch order by view the long tiles of the last views
        // were the tail of the kernel: per-CTA timeline in profiles/); the empty tiles -- pure write streams -- follow and
        // fill the SMs while the last busy tiles finish.
        const unsigned b = blockIdx.x, n0 = (unsigned)P.class_total[kClasses - 1], nb = gridDim.x - n0;
        unsigned bin_id;
        if (b >= nb) {
            bin_id = P.class_list[(size_t)(kClasses - 1) * P.class_cap + (b - nb)];
        } else {
            const int l = threadIdx.x & 31;
            const int tot = l < kClasses - 1 ? P.class_total[l] : 0;
            int incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (l >= o) incl += v;
            }
            const int cls = __popc(__ballot_sync(kFull, incl <= (int)b));   // first class whose inclusive count exceeds b
            const int excl = __shfl_sync(kFull, incl - tot, cls);
            bin_id = P.class_list[(size_t)cls * P.class_cap + ((int)b - excl)];
        }
        view = (int)fastdiv(bin_id, P.div_nt_m, P.div_nt_s);
        tile = (int)(bin_id - (unsigned)view * (unsigned)P.n_tiles);
    } else {
        const unsigned b = blockIdx.x, n1 = (unsigned)P.near_tiles, total1 = (unsigned)P.n_views * n1;
        if (b < total1) {
            view = (int)fastdiv(b, P.div_n1_m, P.div_n1_s);
            tile = (int)(b - (unsigned)view * n1);
        } else {
            const unsigned r = b - total1, n2 = (unsigned)P.n_tiles - n1;
            view = (int)fastdiv(r, P.div_n2_m, P.div_n2_s);
            tile = (int)(n1 + (r - (unsigned)view * n2));
        }
        if (P.use_order) tile = P.tile_order[tile];
    }
    const int tile_y = (int)fastdiv((unsigned)tile, P.div_tx_m, P.div_tx_s), tile_x = tile - tile_y * P.tiles_x;
    const int tx0 = tile_x * kTile, ty0 = tile_y * kTile;
    const int tx1 = min(tx0 + kTile, P.W) - 1, ty1 = min(ty0 + kTile, P.H) - 1;  // inclusive
    const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
    const int bin = view * P.n_tiles + tile;
    const int count = P.bin_count[bin];
#ifdef AB_RASTER_TRACE
    unsigned long long t_start = gtime(), t_patch = 0;
#endif
    const int32_t* sel = P.bg_sel ? P.bg_sel + 5 * (size_t)view : nullptr;
    const bool has_bg = sel && P.bgs && sel[0] >= 0;
    if (has_bg && t < kTile)  // nearest-neighbour source column of every tile column, once per CTA (exact multiply-shift division)
        colmap[t] = sel[1] + (int)fastdiv((unsigned)(2 * min(tx0 + t, P.W - 1) + 1) * (unsigned)sel[3], P.div_2w_m, P.div_2w_s);

    if (count > 0) {
        const int oid = P.obj_id[view];
        const int n_of = oid >= 0 ? P.face_off[oid + 1] - P.face_off[oid] : 0;
        if (t == 0) { n_hit = 0; next_entry = kWarps; n_bigq = 0; }
        if (t < 12) Msh[t] = oid >= 0 ? P.obj_pose[16 * (size_t)view + t] : 0.0f;
        ulonglong2* z2 = reinterpret_cast<ulonglong2*>(zbuf);
        for (int i = t; i < kTilePx / 2; i += kThreads) z2[i] = make_ulonglong2(kEmptyKey, kEmptyKey);
        __syncthreads();
        const unsigned short* list = P.bin_list + (size_t)bin * P.list_cap;
        const int poff = oid >= 0 ? P.patch_off[oid] : 0;
        // per-lane base pointers: a patch is a 32-bit element offset from them
        const float4* op_pos = P.op_pos + (size_t)poff * 32 + lane;
        const uint32_t* op_face = P.op_face + (size_t)poff * 32 + lane;
        const int32_t* op_prim = P.op_prim + (size_t)poff * 32 + lane;
        const int32_t* hp_vid = P.hp_vid + lane;
        const uint32_t* hp_face = P.hp_face + lane;
        const int32_t* hp_prim = P.hp_prim + lane;
        const float* hverts = P.hand_verts + 3 * (size_t)view * P.n_hv;
        int4* my_slab = slab[wid];
        unsigned char* my_rank = rank_lane[wid];
        const unsigned lt_mask = (1u << lane) - 1u;
        // warps claim list entries one at a time from a shared counter (the first by warp index; claiming pairs left the
        // warps of a CTA further apart at the barrier below); the next entry is fetched one patch ahead.  A warp that runs out
        // goes on to the background below instead of waiting for the others.
        int li = wid;
        unsigned entry = li < count ? list[li] : 0u;
        while (li < count) {
            const unsigned cur_entry = entry;
            const bool hand = (entry & kHandBit) != 0;
            const unsigned poffs = (entry & kIndexMask) * 32u;
            const bool multi = (entry & kMultiBit) != 0;
            // the patch's face words and primitive ids do not depend on the vertex phase: issue their loads first, so that
            // one round of memory latency per patch is exposed instead of three
            const unsigned fw = __ldg((hand ? hp_face : op_face) + poffs);
            const int prim_local = __ldg((hand ? hp_prim : op_prim) + poffs);
            {
                int nli = 0;
                if (lane == 0) nli = atomicAdd(&next_entry, 1);
                li = __shfl_sync(kFull, nli, 0);
            }
            entry = li < count ? list[li] : 0u;
            // ---- a lane per vertex
            float c[3] = {0.0f, 0.0f, 0.0f};
            bool on;
            if (!hand) {
                const float4 q = __ldg(op_pos + poffs);
                on = q.w != 0.0f;
                xform(Msh, q.x, q.y, q.z, c);
            } else {
                const int v = __ldg(hp_vid + poffs);
                on = v >= 0;
                if (on) {
                    const float* hv = hverts + 3 * v;
                    c[0] = hv[0]; c[1] = hv[1]; c[2] = hv[2];
                }
            }
            int4 pvv = project(P, c[0], c[1], c[2]);
            if (!on) pvv.w = 0;
            my_slab[lane] = pvv;
            bool overlap = true;
            if (multi) {  // warp-uniform
                const int mnx = __reduce_min_sync(kFull, on ? pvv.x : INT_MAX), mxx = __reduce_max_sync(kFull, on ? pvv.x : INT_MIN);
                const int mny = __reduce_min_sync(kFull, on ? pvv.y : INT_MAX), mxy = __reduce_max_sync(kFull, on ? pvv.y : INT_MIN);
                overlap = floordiv256(mnx + 127) <= tx1 && floordiv256(mxx - 128) >= tx0 &&
                          floordiv256(mny + 127) <= ty1 && floordiv256(mxy - 128) >= ty0;
            }
            __syncwarp();
            if (overlap) {
                // ---- a lane per face
                int n = 0, x0 = 0, y0 = 0, w = 0, h = 0, area = 0;  // area: the int32 value when `small`, else only its sign
                bool small = false;
                int4 a, b, d;
                a = b = d = make_int4(0, 0, 0, 0);
                if (fw != 0xFFFFFFFFu) {
                    a = my_slab[fw & 31]; b = my_slab[(fw >> 8) & 31]; d = my_slab[(fw >> 16) & 31];
                    if (a.w & b.w & d.w) {
                        const int minx = min(a.x, min(b.x, d.x)), maxx = max(a.x, max(b.x, d.x));
                        const int miny = min(a.y, min(b.y, d.y)), maxy = max(a.y, max(b.y, d.y));
                        x0 = max(floordiv256(minx + 127), tx0);
                        y0 = max(floordiv256(miny + 127), ty0);
                        const int x1 = min(floordiv256(maxx - 128), tx1), y1 = min(floordiv256(maxy - 128), ty1);
                        if (x0 <= x1 && y0 <= y1) {
                            small = (maxx - minx < kSmallExtent) && (maxy - miny < kSmallExtent);
                            if (small) {
                                area = (b.x - a.x) * (d.y - a.y) - (d.x - a.x) * (b.y - a.y);
                            } else {
                                const long long a2 = area2_of(a, b, d, false);
                                area = a2 > 0 ? 1 : (a2 < 0 ? -1 : 0);
                            }
                            if (area != 0 && !(area > 0 && P.cull)) { w = x1 - x0 + 1; h = y1 - y0 + 1; n = w * h; }
                        }
                    }
                }
                const unsigned prim = (unsigned)(prim_local + (hand ? n_of : 0));
                // int32 set-up of every surviving lane: edge values at the centre of pixel (x0, y0), biased so that "inside"
                // is e >= 0 on all three; the steps per pixel are -256 dy (x) and 256 dx (y), kept as (dx, dy) in 16 + 16 bits
                int e[3] = {0, 0, 0}, dd[3] = {0, 0, 0}, nb = 0;
                float iz0 = 0.f, iz1 = 0.f, iz2 = 0.f, sarea = 0.f;
                if (n > 0 && small) {
                    const int s = area > 0 ? 1 : -1;
                    const int vx[3] = {a.x, b.x, d.x}, vy[3] = {a.y, b.y, d.y};
                    iz0 = __int_as_float(a.z); iz1 = __int_as_float(b.z); iz2 = __int_as_float(d.z);
                    sarea = __int2float_rn(abs(area));
                    const int cx0 = 256 * x0 + 128, cy0 = 256 * y0 + 128;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {  // edge i runs v[i+1] -> v[i+2], opposite vertex i
                        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
                        const int dx = s * (vx[i2] - vx[i1]), dy = s * (vy[i2] - vy[i1]);  // |.| < 2^14
                        const int bias = ((dy < 0) || (dy == 0 && dx > 0)) ? 0 : 1;  // 1: not a top/left edge, E == 0 is outside
                        nb |= bias << i;
                        e[i] = dx * (cy0 - vy[i1]) - dy * (cx0 - vx[i1]) - bias;
                        dd[i] = (dx & 0xffff) | (dy << 16);
                    }
                }
                // ---- coverage: the ROWS of all surviving boxes of the patch are dealt to the lanes.  Warp scan of the row
                // counts; the owner of row j of a chunk of 32 rows is found from a bit mask of the positions where a box starts
                // (one warp reduction) and a rank -> lane table; its set-up comes by indexed shuffles.  A row's candidate span
                // follows from the edge functions: e_i + x ex_i >= 0 bounds x from below (ex_i > 0) or from above (ex_i < 0); the
                // fp32 guess is at most one pixel wide of the truth on either side (rows of up to 4 pixels skip the guess), and
                // every candidate is then tested with the exact integers.
                const int rows = (n > 0 && small) ? h : 0;
                int incl = rows;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(kFull, incl, o);
                    if (lane >= o) incl += v;
                }
                const int total = __shfl_sync(kFull, incl, 31);
                if (total > 0) {  // warp-uniform
                    const int pre = incl - rows;
                    const unsigned owners = __ballot_sync(kFull, rows > 0);
                    if (rows > 0) my_rank[__popc(owners & lt_mask)] = (unsigned char)lane;
                    __syncwarp();
                    const int last_rank = __popc(owners) - 1;
                    const int pack = (x0 - tx0) | ((y0 - ty0) << 6) | (w << 12) | (nb << 19);
                    for (int j0 = 0; j0 < total; j0 += 32) {
                        const int rel = pre - j0;
                        const unsigned starts = __reduce_or_sync(kFull, (rows > 0 && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
                        const int before = __popc(__ballot_sync(kFull, rows > 0 && rel < 0));
                        const int rank = before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1;
                        const int own = my_rank[min(max(rank, 0), last_rank)];
                        const int j = j0 + lane;
                        const int ry = j - __shfl_sync(kFull, pre, own);
                        const int ce0 = __shfl_sync(kFull, e[0], own), ce1 = __shfl_sync(kFull, e[1], own), ce2 = __shfl_sync(kFull, e[2], own);
                        const int cd0 = __shfl_sync(kFull, dd[0], own), cd1 = __shfl_sync(kFull, dd[1], own), cd2 = __shfl_sync(kFull, dd[2], own);
                        const float ciz0 = __shfl_sync(kFull, iz0, own), ciz1 = __shfl_sync(kFull, iz1, own), ciz2 = __shfl_sync(kFull, iz2, own);
                        const float csarea = __shfl_sync(kFull, sarea, own);
                        const unsigned cprim = __shfl_sync(kFull, prim, own);
                        const int cpack = __shfl_sync(kFull, pack, own);
                        const int cw = (cpack >> 12) & 127, cnb = cpack >> 19;
                        const bool act = j < total;
                        const int cex0 = -256 * (cd0 >> 16), cex1 = -256 * (cd1 >> 16), cex2 = -256 * (cd2 >> 16);
                        const int r0 = ce0 + ry * (256 * (int)(short)cd0), r1 = ce1 + ry * (256 * (int)(short)cd1),
                                  r2 = ce2 + ry * (256 * (int)(short)cd2);  // biased edge values at x = 0 of the row
                        int lo = 0, hi = act ? cw - 1 : -1;
                        if (__any_sync(kFull, act && cw > 4)) {
                            if (cex0 != 0) {
                                const int fl = min(__float2int_rd(__fdividef(-(float)r0, (float)cex0)), 1 << 20);
                                if (cex0 > 0) lo = max(lo, fl); else hi = min(hi, fl + 1);
                            }
                            if (cex1 != 0) {
                                const int fl = min(__float2int_rd(__fdividef(-(float)r1, (float)cex1)), 1 << 20);
                                if (cex1 > 0) lo = max(lo, fl); else hi = min(hi, fl + 1);
                            }
                            if (cex2 != 0) {
                                const int fl = min(__float2int_rd(__fdividef(-(float)r2, (float)cex2)), 1 << 20);
                                if (cex2 > 0) lo = max(lo, fl); else hi = min(hi, fl + 1);
                            }
                        }
                        const int base = (((cpack >> 6) & 63) + ry) * kTile + (cpack & 63);
                        for (int x = lo; x <= hi; ++x) {
                            const int f0 = r0 + x * cex0, f1 = r1 + x * cex1, f2 = r2 + x * cex2;
                            if ((f0 | f1 | f2) >= 0)
                                emit(zbuf, base + x, __int2float_rn(f0 + (cnb & 1)), __int2float_rn(f1 + ((cnb >> 1) & 1)),
                                     __int2float_rn(f2 + ((cnb >> 2) & 1)), ciz0, ciz1, ciz2, csarea, cprim);
                        }
                    }
                }
                // ---- extents of 64 px and more (close-ups, cap fans seen face-on): queued for the pass after the patch loop
                if (n > 0 && !small) {
                    const int slot = atomicAdd(&n_bigq, 1);
                    if (slot < kBigCap) big_q[slot] = (cur_entry & 0xffffu) | ((unsigned)lane << 16);
                }
            }
            __syncwarp();  // the slab and the rank table are rewritten by the next patch
        }
    } else {
        __syncthreads();  // colmap
    }

    // ---- stream the background out: flat colour or the nearest-neighbour crop, zero depth and seg.  This does not depend on
    // the z-buffer (covered pixels are overwritten by the shading pass below, after a barrier), so a warp that runs out of
    // patches starts it at once instead of waiting for the others.
    constexpr int TPR = kTile / PX;        // threads per tile row
    constexpr int RPP = kThreads / TPR;    // rows per pass
    const size_t obase = (size_t)view * P.W * P.H;
    const unsigned flat = (unsigned)(P.bg_r & 255) | ((unsigned)(P.bg_g & 255) << 8) | ((unsigned)(P.bg_b & 255) << 16);
    const int col = (t % TPR) * PX, px0 = tx0 + col;
    if (px0 <= tx1) {  // PX == 4 only when W % 4 == 0: a group of 4 pixels is entirely in or out
        int sx[PX];
        unsigned bg_y0 = 0, bg_ch_ = 0;
        const uint8_t* bg_img = nullptr;
        if (has_bg) {
            if constexpr (PX == 4) {
                const int4 q = *reinterpret_cast<const int4*>(colmap + col);
                sx[0] = q.x; sx[1] = q.y; sx[2] = q.z; sx[3] = q.w;
            } else {
                sx[0] = colmap[col];
            }
            bg_y0 = (unsigned)sel[2]; bg_ch_ = (unsigned)sel[4];
            bg_img = P.bgs + (size_t)(BG4 ? 4 : 3) * ((size_t)sel[0] * P.bg_h * P.bg_w);
        }
        size_t o = obase + (size_t)(ty0 + t / TPR) * P.W + px0;
#pragma unroll
        for (int pass = 0; pass < kTile / RPP; ++pass, o += (size_t)RPP * P.W) {
            const int py = ty0 + pass * RPP + t / TPR;
            if (py > ty1) break;
            unsigned c[PX];
#pragma unroll
            for (int j = 0; j < PX; ++j) c[j] = flat;
            if (has_bg) {
                const unsigned sy = bg_y0 + fastdiv((unsigned)(2 * py + 1) * bg_ch_, P.div_2h_m, P.div_2h_s);
#pragma unroll
                for (int j = 0; j < PX; ++j) {
                    if constexpr (BG4) {
                        c[j] = __ldg(reinterpret_cast<const unsigned*>(bg_img) + sy * (unsigned)P.bg_w + (unsigned)sx[j]) & 0x00ffffffu;
                    } else {
                        const uint8_t* src = bg_img + 3 * (size_t)(sy * (unsigned)P.bg_w + (unsigned)sx[j]);
                        c[j] = (unsigned)src[0] | ((unsigned)src[1] << 8) | ((unsigned)src[2] << 16);
                    }
                }
            }
            if constexpr (PX == 4) {
                if (P.rgba) *reinterpret_cast<uint4*>(P.rgba + 4 * o) = make_uint4(c[0], c[1], c[2], c[3]);
                if (P.depth) *reinterpret_cast<float4*>(P.depth + o) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (P.seg) *reinterpret_cast<unsigned*>(P.seg + o) = 0u;
            } else {
                if (P.rgba) reinterpret_cast<unsigned*>(P.rgba)[o] = c[0];
                if (P.depth) P.depth[o] = 0.0f;
                if (P.seg) P.seg[o] = 0;
            }
        }
    }
#ifdef AB_RASTER_TRACE
    if (count == 0 && threadIdx.x == 0 && blockIdx.x < 32768) {
        unsigned long long* q = g_trace + 6 * blockIdx.x;
        q[0] = t_start; q[1] = gtime(); q[2] = smid(); q[3] = 0; q[4] = 0; q[5] = q[1];
    }
#endif
    if (count == 0) return;
    __syncthreads();  // every warp's fragments are in the z-buffer, every placeholder store has been issued
    if (n_bigq > 0) {     // CTA-uniform
        raster_big_queue(P, zbuf, big_q, min(n_bigq, kBigCap), n_bigq > kBigCap, view, bin, count, tx0, ty0, tx1, ty1, Msh);
        __syncthreads();
    }
#ifdef AB_RASTER_TRACE
    t_patch = gtime();
#endif
    // ---- collect the covered pixels of the tile ...
    for (int i = t; i < kTilePx / 2; i += kThreads) {
        const ulonglong2 k = reinterpret_cast<const ulonglong2*>(zbuf)[i];
        const bool h0 = k.x != kEmptyKey, h1 = k.y != kEmptyKey;
        const unsigned m0 = __ballot_sync(kFull, h0), m1 = __ballot_sync(kFull, h1);
        if (m0 | m1) {  // warp-uniform
            int base = 0;
            if (lane == 0) base = atomicAdd(&n_hit, __popc(m0) + __popc(m1));
            base = __shfl_sync(kFull, base, 0);
            const unsigned lt = (1u << lane) - 1u;
            if (h0) hit_px[base + __popc(m0 & lt)] = (unsigned short)(2 * i);
            if (h1) hit_px[base + __popc(m0) + __popc(m1 & lt)] = (unsigned short)(2 * i + 1);
        }
    }
    __syncthreads();
    // ---- ... and shade them, one per thread, on fully populated warps; their outputs overwrite the placeholders
    const int oid = P.obj_id[view];
    const int n_of = oid >= 0 ? P.face_off[oid + 1] - P.face_off[oid] : 0;
    const int total = n_hit;
    for (int i = t; i < total; i += kThreads) {
        const int idx = hit_px[i];
        const unsigned long long key = zbuf[idx];
        const int px = tx0 + (idx & (kTile - 1)), py = ty0 + (idx >> 6);
        const PixelOut r = shade_pixel(P, Msh, view, oid, n_of, px, py, (unsigned)(key & 0xffffffffull),
                                       __uint_as_float((unsigned)(key >> 32)));
        const size_t o = obase + (size_t)py * P.W + px;
        if (P.rgba) reinterpret_cast<uchar4*>(P.rgba)[o] = r.rgba;
        if (P.depth) P.depth[o] = r.depth;
        if (P.seg) P.seg[o] = r.seg;
    }
#ifdef AB_RASTER_TRACE
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x < 32768) {
        unsigned long long* q = g_trace + 6 * blockIdx.x;
        q[0] = t_start; q[1] = gtime(); q[2] = smid() | ((unsigned long long)n_bigq << 32); q[3] = count; q[4] = total; q[5] = t_patch;
    }
#endif
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Geometry {  // what the workspace layout depends on
    int tiles_x, tiles_y, n_tiles, list_cap, max_op;
};

static int geometry_of(const ab_scene* s, const ab_camera* cam, Geometry* g) {
    if (!s || !cam || cam->width <= 0 || cam->height <= 0) return -1;
    if (!s->hand_patches || !s->hand_patches->patch_off_host || s->hand_patches->n_mesh != 1) return -1;
    if (s->n_obj > 0 && (!s->obj_patches || !s->obj_patches->patch_off_host || s->obj_patches->n_mesh != s->n_obj)) return -1;
    g->tiles_x = cdiv(cam->width, kTile);
    g->tiles_y = cdiv(cam->height, kTile);
    g->n_tiles = g->tiles_x * g->tiles_y;
    g->max_op = 0;
    for (int i = 0; i < s->n_obj; ++i)
        g->max_op = max(g->max_op, s->obj_patches->patch_off_host[i + 1] - s->obj_patches->patch_off_host[i]);
    const int n_hp = s->hand_patches->patch_off_host[1] - s->hand_patches->patch_off_host[0];
    if (g->max_op > (int)kIndexMask || n_hp > (int)kIndexMask || n_hp <= 0) return -1;
    g->list_cap = (int)align_up((size_t)g->max_op + n_hp, 8);
    return 0;
}

// The work-ordered launch keeps its class lists in the workspace; groups of more than 32767 tiles use the static order.
static bool sorted_ok(int chunk, int n_tiles) { return (size_t)chunk * n_tiles <= 32767; }
static size_t sorted_bytes(int chunk, int n_tiles) {
    if (!sorted_ok(chunk, n_tiles)) return 0;
    return 256 + align_up((size_t)kClasses * chunk * n_tiles * sizeof(unsigned), 256);
}

}  // namespace ab

#ifdef AB_RASTER_TRACE
extern "C" __attribute__((visibility("default"))) int ab_debug_raster_trace(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, ab::g_trace, sizeof(unsigned long long) * 6 * (size_t)n);
}
#endif

extern "C" uint64_t ab_render_workspace_bytes(const ab_scene* scene, const ab_camera* cam, int chunk) {
    ab::Geometry g;
    if (chunk <= 0 || ab::geometry_of(scene, cam, &g)) return 0;
    return ab::align_up((size_t)chunk * g.n_tiles * sizeof(int), 256) +
           ab::align_up((size_t)chunk * g.n_tiles * g.list_cap * sizeof(unsigned short), 256) +
           ab::sorted_bytes(chunk, g.n_tiles);
}

extern "C" int ab_render_batch(const ab_scene* scene, const ab_camera* cam, int batch, int chunk,
                               const float* hand_verts, const int32_t* hand_tex, const int32_t* obj_id,
                               const int32_t* obj_id_host, const float* obj_pose, const float* light,
                               const int32_t* bg_sel, uint8_t* rgba, float* depth, uint8_t* seg, void* ws,
                               void* stream) {
    using namespace ab;
    AB_REQUIRE(scene && cam, "null scene / camera");
    AB_REQUIRE(batch >= 0 && chunk > 0 && chunk <= 65535, "bad batch / chunk (at most 65535 views per group)");
    AB_REQUIRE(cam->width > 0 && cam->height > 0 && cam->width <= 4096 && cam->height <= 4096, "bad image size");
    AB_REQUIRE(cam->fx > 0.0f && cam->fy > 0.0f, "focal lengths must be positive");
    AB_REQUIRE(scene->n_obj >= 0 && scene->n_obj <= kMaxObjects, "n_obj out of range (max 64)");
    AB_REQUIRE(scene->n_obj == 0 || (scene->obj_verts && scene->obj_faces && scene->obj_colors &&
                                     scene->obj_vert_off_host && scene->obj_face_off_host), "null object arrays");
    AB_REQUIRE(!scene->bgs || (scene->bg_w > 0 && scene->bg_h > 0 && scene->bg_w <= 65535 && scene->bg_h <= 65535), "bad background size");
    AB_REQUIRE(scene->n_hand_verts > 0 && scene->n_hand_faces > 0 && scene->n_hand_tex > 0 && scene->hand_faces &&
                   scene->hand_colors, "bad hand mesh");
    Geometry g;
    AB_REQUIRE(geometry_of(scene, cam, &g) == 0, "missing / inconsistent patch tables (ab_build_patches_host; at most 16383 patches per mesh)");
    const ab_patch_table* hp = scene->hand_patches;
    const ab_patch_table* op = scene->obj_patches;
    AB_REQUIRE(hp->vid && hp->face && hp->prim, "null hand patch arrays");
    AB_REQUIRE(scene->n_obj == 0 || (op->pos && op->face && op->prim && op->bound), "null object patch arrays");
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
    P.op_pos = scene->n_obj ? reinterpret_cast<const float4*>(op->pos) : nullptr;
    P.op_face = scene->n_obj ? op->face : nullptr;
    P.op_prim = scene->n_obj ? op->prim : nullptr;
    P.op_bound = scene->n_obj ? reinterpret_cast<const float4*>(op->bound) : nullptr;
    AB_REQUIRE(((uintptr_t)P.op_pos & 15) == 0 && ((uintptr_t)P.op_bound & 15) == 0, "patch pos / bound must be 16-byte aligned");
    P.hp_vid = hp->vid; P.hp_face = hp->face; P.hp_prim = hp->prim;
    P.n_hp = hp->patch_off_host[1] - hp->patch_off_host[0];
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
    for (int i = 0; i <= scene->n_obj; ++i) {
        P.vert_off[i] = scene->n_obj ? scene->obj_vert_off_host[i] : 0;
        P.face_off[i] = scene->n_obj ? scene->obj_face_off_host[i] : 0;
        P.patch_off[i] = scene->n_obj ? op->patch_off_host[i] : 0;
        if (i > 0)
            AB_REQUIRE(P.vert_off[i] >= P.vert_off[i - 1] && P.face_off[i] >= P.face_off[i - 1] &&
                           P.patch_off[i] >= P.patch_off[i - 1], "offsets not sorted");
    }
    P.tiles_x = g.tiles_x; P.tiles_y = g.tiles_y; P.n_tiles = g.n_tiles; P.list_cap = g.list_cap;
    const size_t count_bytes = align_up((size_t)chunk * g.n_tiles * sizeof(int), 256);
    P.bin_count = (int*)ws;
    P.bin_list = (unsigned short*)((char*)ws + count_bytes);
    {
        static const int want_sorted = getenv("AB_RASTER_SORTED") ? atoi(getenv("AB_RASTER_SORTED")) : 1;
        const size_t list_bytes = align_up((size_t)chunk * g.n_tiles * g.list_cap * sizeof(unsigned short), 256);
        P.sorted = sorted_ok(chunk, g.n_tiles) ? (want_sorted != 0) : 0;
        P.class_total = (int*)((char*)ws + count_bytes + list_bytes);
        P.class_list = (unsigned*)((char*)P.class_total + 256);
        P.class_cap = chunk * g.n_tiles;
        P.area_shift = 1; P.area_mul = 5;  // half the summed boxes ~ covered pixels; 5 / 64 entries per pixel (0.02 us / 0.25 us)
        P.class_width = max(1, (g.list_cap / 2 + ((kTilePx * P.area_mul) >> 6) + kClasses - 2) / (kClasses - 1));
        make_fastdiv((unsigned)g.n_tiles, &P.div_nt_m, &P.div_nt_s);
    }
    // tiles by distance of their centre from the principal point; the outer `defer` of them go last
    {
        P.use_order = g.n_tiles <= kMaxOrder;
        int defer = 0;
        if (P.use_order) {
            std::vector<std::pair<float, unsigned short>> rank((size_t)g.n_tiles);
            for (int i = 0; i < g.n_tiles; ++i) {
                const float dx = ((i % g.tiles_x) + 0.5f) * kTile - cam->cx, dy = ((i / g.tiles_x) + 0.5f) * kTile - cam->cy;
                rank[i] = {dx * dx + dy * dy, (unsigned short)i};
            }
            std::stable_sort(rank.begin(), rank.end());
            for (int i = 0; i < g.n_tiles; ++i) P.tile_order[i] = rank[i].second;
            static const int defer_pct = getenv("AB_RASTER_DEFER_PCT") ? atoi(getenv("AB_RASTER_DEFER_PCT")) : 50;
            defer = max(0, min(g.n_tiles * defer_pct / 100, g.n_tiles - 1));
        }
        P.near_tiles = g.n_tiles - defer;
        make_fastdiv((unsigned)P.near_tiles, &P.div_n1_m, &P.div_n1_s);
        make_fastdiv((unsigned)max(defer, 1), &P.div_n2_m, &P.div_n2_s);
        make_fastdiv((unsigned)g.tiles_x, &P.div_tx_m, &P.div_tx_s);
    }
    const int npx = P.W * P.H;
    for (int v0 = 0; v0 < batch; v0 += chunk) {
        const int n = min(chunk, batch - v0);
        if (obj_id_host)
            for (int i = 0; i < n; ++i) AB_REQUIRE(obj_id_host[v0 + i] < scene->n_obj, "obj_id out of range");
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
        if (P.sorted) AB_CUDA(cudaMemsetAsync(P.class_total, 0, kClasses * sizeof(int), st));
        {
            StageTimer tm(AB_STAGE_RASTER_BIN, st);
            raster_bin_kernel<<<n, kThreads, 0, st>>>(P);
        }
        {
            StageTimer tm(AB_STAGE_RASTER_TILE, st);
            // 4 pixels per thread needs 16-byte aligned output rows: W % 4 == 0 and 16 / 16 / 4-byte aligned bases
            const bool vec = (P.W % 4 == 0) && (((uintptr_t)P.rgba & 15) == 0) && (((uintptr_t)P.depth & 15) == 0) &&
                             (((uintptr_t)P.seg & 3) == 0);
            const dim3 grid(g.n_tiles * n);
            if (vec && P.bg_ch == 4) raster_tile_kernel<4, true><<<grid, kThreads, 0, st>>>(P);
            else if (vec) raster_tile_kernel<4, false><<<grid, kThreads, 0, st>>>(P);
            else if (P.bg_ch == 4) raster_tile_kernel<1, true><<<grid, kThreads, 0, st>>>(P);
            else raster_tile_kernel<1, false><<<grid, kThreads, 0, st>>>(P);
        }
        count_launch(2);
        int rc = check_launch("ab_render_batch");
        if (rc) return rc;
    }
    return AB_OK;
}
