// Host-side preparation of the rasteriser's mesh patches (no device code in this file).
//
// The tile rasteriser (raster.cu) walks a mesh as patches of at most 32 faces over at most 32 vertices: one warp
// stages the patch's vertices in shared memory (a lane per vertex), then takes a lane per face.  A patch also carries a
// bounding sphere and a normal cone, so that the binning pass can drop patches whose faces are all back-facing and place
// the others into screen tiles before a single vertex of them has been fetched.  Patches are built once per mesh.
//
// What the reference does instead: pyrender uploads the whole mesh as one VBO per object and lets the GL pipeline cull
// and clip per triangle (anakin/utils/renderer.py:79-93, frender_utils.py:16-20).  Face ORDER is part of the image
// (z-test ties go to the lower primitive id = scene insertion order, renderer.py:90-93), so every face keeps its
// original index as `prim`; grouping faces into patches changes which faces are visited, never how one is drawn.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace ab {
namespace {

struct V3 {
    double x, y, z;
};
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return sqrt(dot(a, a)); }

inline uint32_t spread10(uint32_t x) {  // 10 bits -> every third bit
    x &= 0x3ff;
    x = (x | (x << 16)) & 0x030000FF;
    x = (x | (x << 8)) & 0x0300F00F;
    x = (x | (x << 4)) & 0x030C30C3;
    x = (x | (x << 2)) & 0x09249249;
    return x;
}

inline float round_up(double v) {  // smallest float >= v (for bounds that must not shrink)
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}
inline float round_down(double v) {
    float f = (float)v;
    if ((double)f > v) f = nextafterf(f, -INFINITY);
    return f;
}

}  // namespace
}  // namespace ab

extern "C" int ab_build_patches_host(const float* verts, int n_verts, const int32_t* faces, int n_faces, int face_stride,
                                     int capacity, float* pos, int32_t* vid, uint32_t* face, int32_t* prim, float* bound,
                                     int32_t* n_patches) {
    using namespace ab;
    AB_REQUIRE(verts && faces && vid && face && prim && bound && n_patches, "null argument");
    AB_REQUIRE(n_verts > 0 && n_faces > 0 && face_stride >= 3, "empty mesh / bad face stride");
    AB_REQUIRE(capacity >= n_faces, "capacity must be at least n_faces (ab_patch_capacity)");
    const int F = n_faces, V = n_verts;
    auto fv = [&](int f, int k) { return faces[(size_t)f * face_stride + k]; };
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) AB_REQUIRE(fv(f, k) >= 0 && fv(f, k) < V, "face index out of range");
    auto P = [&](int v) { return V3{verts[3 * (size_t)v], verts[3 * (size_t)v + 1], verts[3 * (size_t)v + 2]}; };

    // per-face centroid, unit normal (winding order), area, perimeter
    std::vector<V3> cen(F), nrm(F);
    std::vector<double> area(F), perim(F);
    V3 lo = P(fv(0, 0)), hi = lo;
    for (int f = 0; f < F; ++f) {
        const V3 a = P(fv(f, 0)), b = P(fv(f, 1)), c = P(fv(f, 2));
        cen[f] = (a + b + c) * (1.0 / 3.0);
        const V3 n = cross(b - a, c - a);
        const double ln = norm(n);
        area[f] = 0.5 * ln;
        nrm[f] = ln > 0 ? n * (1.0 / ln) : V3{0, 0, 0};
        perim[f] = norm(b - a) + norm(c - b) + norm(a - c);
        lo = {std::min(lo.x, cen[f].x), std::min(lo.y, cen[f].y), std::min(lo.z, cen[f].z)};
        hi = {std::max(hi.x, cen[f].x), std::max(hi.y, cen[f].y), std::max(hi.z, cen[f].z)};
    }
    // fallback seed order: Morton curve over the centroids
    std::vector<std::pair<uint32_t, int>> order(F);
    for (int f = 0; f < F; ++f) {
        const double ex = std::max(hi.x - lo.x, 1e-12), ey = std::max(hi.y - lo.y, 1e-12), ez = std::max(hi.z - lo.z, 1e-12);
        const uint32_t qx = (uint32_t)std::min(1023.0, std::max(0.0, (cen[f].x - lo.x) / ex * 1023.0));
        const uint32_t qy = (uint32_t)std::min(1023.0, std::max(0.0, (cen[f].y - lo.y) / ey * 1023.0));
        const uint32_t qz = (uint32_t)std::min(1023.0, std::max(0.0, (cen[f].z - lo.z) / ez * 1023.0));
        order[f] = {spread10(qx) | (spread10(qy) << 1) | (spread10(qz) << 2), f};
    }
    std::stable_sort(order.begin(), order.end(), [](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) { return a.first < b.first; });
    // vertex -> incident faces (CSR)
    std::vector<int> vf_off(V + 1, 0), vf(3 * (size_t)F);
    for (int f = 0; f < F; ++f)
        for (int k = 0; k < 3; ++k) ++vf_off[fv(f, k) + 1];
    for (int v = 0; v < V; ++v) vf_off[v + 1] += vf_off[v];
    {
        std::vector<int> fill(vf_off.begin(), vf_off.end() - 1);
        for (int f = 0; f < F; ++f)
            for (int k = 0; k < 3; ++k) vf[fill[fv(f, k)]++] = f;
    }

    std::vector<char> assigned(F, 0);
    std::vector<int> cand_stamp(F, -1), local_of(V, -1);
    std::vector<int> frontier;  // faces that touched a finished patch and are still free: preferred seeds
    size_t frontier_head = 0;
    int pos_in_order = 0, n_p = 0;
    std::vector<int> cand, pf, pv;
    while (true) {
        // seed: among the first live frontier faces the one with the most assigned neighbours (fills pockets before
        // they become islands); else the next free face along the Morton curve
        int seed = -1, best_k = -1, seen = 0;
        while (frontier_head < frontier.size() && assigned[frontier[frontier_head]]) ++frontier_head;
        for (size_t i = frontier_head; i < frontier.size() && seen < 64; ++i) {
            const int g = frontier[i];
            if (assigned[g]) continue;
            ++seen;
            int k = 0;
            for (int c = 0; c < 3; ++c)
                for (int j = vf_off[fv(g, c)]; j < vf_off[fv(g, c) + 1]; ++j) k += assigned[vf[j]];
            if (k > best_k) { best_k = k; seed = g; }
        }
        if (seed < 0) {
            while (pos_in_order < F && assigned[order[pos_in_order].second]) ++pos_in_order;
            if (pos_in_order >= F) break;
            seed = order[pos_in_order].second;
        }
        AB_REQUIRE(n_p < capacity, "patch capacity exceeded");
        cand.clear(); pf.clear(); pv.clear();
        cand.push_back(seed);
        cand_stamp[seed] = n_p;
        V3 csum{0, 0, 0}, nsum{0, 0, 0};
        while (!cand.empty() && (int)pf.size() < 32) {
            const V3 c = pf.empty() ? cen[seed] : csum * (1.0 / pf.size());
            const double ln = norm(nsum);
            const V3 na = ln > 0 ? nsum * (1.0 / ln) : nrm[seed];
            int best = -1, best_new = 4;
            double best_d = 0;
            for (size_t i = 0; i < cand.size(); ++i) {
                const int g = cand[i];
                int nw = 0;
                const int a = fv(g, 0), b = fv(g, 1), d = fv(g, 2);
                nw += local_of[a] < 0;
                nw += (local_of[b] < 0 && b != a);
                nw += (local_of[d] < 0 && d != a && d != b);
                if ((int)pv.size() + nw > 32) continue;
                const double dist = norm(cen[g] - c) * (2.0 - dot(nrm[g], na));  // near the patch and aligned with its cone
                if (nw < best_new || (nw == best_new && dist < best_d)) { best = (int)i; best_new = nw; best_d = dist; }
            }
            if (best < 0) break;
            const int g = cand[best];
            cand[best] = cand.back();
            cand.pop_back();
            assigned[g] = 1;
            pf.push_back(g);
            csum = csum + cen[g];
            nsum = nsum + nrm[g];
            for (int k = 0; k < 3; ++k) {
                const int v = fv(g, k);
                if (local_of[v] < 0) { local_of[v] = (int)pv.size(); pv.push_back(v); }
                for (int j = vf_off[v]; j < vf_off[v + 1]; ++j) {
                    const int h = vf[j];
                    if (!assigned[h] && cand_stamp[h] != n_p) { cand_stamp[h] = n_p; cand.push_back(h); }
                }
            }
        }
        for (int g : cand) if (!assigned[g]) frontier.push_back(g);
        // ---- emit the patch
        float* ppos = pos ? pos + (size_t)n_p * 128 : nullptr;
        int32_t* pvid = vid + (size_t)n_p * 32;
        uint32_t* pface = face + (size_t)n_p * 32;
        int32_t* pprim = prim + (size_t)n_p * 32;
        for (int l = 0; l < 32; ++l) {
            pvid[l] = l < (int)pv.size() ? pv[l] : -1;
            if (ppos) {
                const bool on = l < (int)pv.size();
                ppos[4 * l + 0] = on ? verts[3 * (size_t)pv[l]] : 0.f;
                ppos[4 * l + 1] = on ? verts[3 * (size_t)pv[l] + 1] : 0.f;
                ppos[4 * l + 2] = on ? verts[3 * (size_t)pv[l] + 2] : 0.f;
                ppos[4 * l + 3] = on ? 1.f : 0.f;
            }
            if (l < (int)pf.size()) {
                const int g = pf[l];
                pface[l] = (uint32_t)local_of[fv(g, 0)] | ((uint32_t)local_of[fv(g, 1)] << 8) | ((uint32_t)local_of[fv(g, 2)] << 16);
                pprim[l] = g;
            } else {
                pface[l] = 0xFFFFFFFFu;
                pprim[l] = -1;
            }
        }
        // bounding sphere (centre of the vertex box), normal cone, worst perimeter / area and 1 / (2 area) of a face
        V3 blo = P(pv[0]), bhi = blo;
        for (int v : pv) {
            const V3 p = P(v);
            blo = {std::min(blo.x, p.x), std::min(blo.y, p.y), std::min(blo.z, p.z)};
            bhi = {std::max(bhi.x, p.x), std::max(bhi.y, p.y), std::max(bhi.z, p.z)};
        }
        const V3 cf{(double)(float)(0.5 * (blo.x + bhi.x)), (double)(float)(0.5 * (blo.y + bhi.y)), (double)(float)(0.5 * (blo.z + bhi.z))};
        double r = 0;
        for (int v : pv) r = std::max(r, norm(P(v) - cf));
        V3 axis = nsum;
        const double la = norm(axis);
        axis = la > 0 ? axis * (1.0 / la) : V3{0, 0, 1};
        const V3 af{(double)(float)axis.x, (double)(float)axis.y, (double)(float)axis.z};  // the axis the kernel will see
        const double laf = norm(af);
        double cutoff = 1.0, q = 0.0, ia = 0.0;
        for (int g : pf) {
            if (!(area[g] > 0)) { q = INFINITY; ia = INFINITY; cutoff = -1.0; continue; }  // degenerate in 3-D: never cull
            cutoff = std::min(cutoff, dot(nrm[g], af) / laf);
            q = std::max(q, perim[g] / area[g]);
            ia = std::max(ia, 0.5 / area[g]);
        }
        float* pb = bound + (size_t)n_p * 12;
        pb[0] = (float)cf.x; pb[1] = (float)cf.y; pb[2] = (float)cf.z;
        pb[3] = round_up(r * (1.0 + 1e-6) + 1e-7);
        pb[4] = (float)(af.x / laf); pb[5] = (float)(af.y / laf); pb[6] = (float)(af.z / laf);
        pb[7] = round_down(cutoff - 1e-6);
        pb[8] = round_up(q);
        pb[9] = round_up(ia);
        pb[10] = (float)pv.size();
        pb[11] = (float)pf.size();
        for (int v : pv) local_of[v] = -1;
        ++n_p;
    }
    *n_patches = n_p;
    return AB_OK;
}

extern "C" int ab_patch_capacity(int n_faces) { return n_faces > 0 ? n_faces : 0; }
