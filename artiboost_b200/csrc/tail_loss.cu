// Fused clasbased tail + training criterion, forward AND gradient in one launch (sm_100a).
//
// Replaces, for the training step, the ~320 small torch launches of
//   * the HybridBaseline tail (anakin/models/hybridbaseline.py:41-96): uvd -> xyz (utils/transform.py:512-546), 6D -> rotation
//     (:578-598), box corners, corner projection, the seven output tensors, and
//   * the Criterion of the clasbased configs (anakin/criterions/criterion.py:57-67): JointsLoss (jointloss.py:14-67),
//     HandOrdLoss joint- and part-level (ordinal.py:75-227), SceneOrdLoss (ordinal.py:231-306), SymCornerLoss
//     (symcornerloss.py:18-108),
// and their autograd backward.  The loss is a scalar whose only trainable inputs here are kp3d [B,22,3] (heatmap decode)
// and the 6-D box rotation [B,6] (MLP_O), so the kernel returns d loss / d kp3d and d loss / d rot6d directly: every mean's
// normaliser is known before the launch, hence the gradient needs no second pass.
//
// One CTA per sample.  All reductions run in a fixed order (per-thread strided sums, shuffle tree, per-warp slots summed by
// one thread; per-sample partial sums added over the batch by a second, single-CTA kernel): no atomics, bit-reproducible.
// Random draws (virtual view vectors, pair subsets) are inputs: they come from the caller's generator in the same order as
// the unfused criterion, so both paths consume the same stream.
#include "common.cuh"

namespace ab {

constexpr int kTailThreads = 128;
constexpr int kParts = 8;  // joints, corners, joint_ord, part_ord, scene_ord, sym, (unused), total

__constant__ int c_parents[21] = {0, 0, 1, 2, 3, 0, 5, 6, 7, 0, 9, 10, 11, 0, 13, 14, 15, 0, 17, 18, 19};  // anakin/utils/misc.py:71

struct TailArgs {
    ab_tail_cfg cfg;
    const float *kp3d, *rot6d, *root_joint, *cam_intr, *corners_can, *joints_3d, *corners_3d, *joints_vis, *corners_vis;
    const float* vv_hand;
    const int32_t *jp, *pp;
    const float* vv_scene;
    const int32_t* hp;
    const float *sym_R, *sym_t;
    const int32_t* obj_idx;
    const float* obj_transf;
    float *o_joints_abs, *o_corners_abs, *o_joints_rel, *o_corners_rel, *o_uvd, *o_boxroot, *o_rotmat;
    float *d_kp3d, *d_rot6d, *partial;
};

__device__ __forceinline__ float sgn(float x) { return (x > 0.0f) - (x < 0.0f); }

// deterministic block sum: strided per-thread value -> warp tree -> warp slots added by thread 0 in index order
__device__ float block_sum(float v, float* slots) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) slots[w] = v;
    __syncthreads();
    float s = 0.0f;
    for (int i = 0; i < kTailThreads / 32; ++i) s += slots[i];
    return s;
}

struct V3 { float x, y, z; };
__device__ __forceinline__ V3 ld3(const float* p) { return {p[0], p[1], p[2]}; }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Ordinal term over (pair, view): x = -sign(t) * p; joint / scene level: log(1 + relu(x)); returns the term and d term / d p
__device__ __forceinline__ float ord_log(float t, float p, float* dp) {
    const float s = sgn(t), x = -s * p;
    if (x > 0.0f) { *dp = -s / (1.0f + x); return log1pf(x); }
    *dp = 0.0f;
    return 0.0f;
}

__global__ void __launch_bounds__(kTailThreads)
tail_loss_kernel(const TailArgs A) {
    extern __shared__ float dyn[];  // coefficient scratch: max(pairs * views) floats, then 3 floats per pair
    __shared__ float pj[21][3], tj[21][3], pc[8][3], tc[8][3], mj[21], mc[8];
    __shared__ float ppart[20][3], tpart[20][3], gpart[20][3];
    __shared__ float gj[21][3], gc[8][3];      // d loss / d joints_3d_abs, d loss / d corners_3d_abs
    __shared__ float Rm[9], can[8][3], slots[kTailThreads / 32], sums[kParts];
    __shared__ float best_err;
    __shared__ int best_k;
    const ab_tail_cfg& c = A.cfg;
    const int b = blockIdx.x, t = threadIdx.x, B = c.batch;
    const float* kp = A.kp3d + (size_t)b * 66;
    const float* K = A.cam_intr + (size_t)b * 9;
    const float fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const float rootz = A.root_joint[b * 3 + 2];

    // ---- tail forward
    if (t < 22) {  // uvd -> xyz
        const float u = kp[3 * t] * c.inp_w, v = kp[3 * t + 1] * c.inp_h;
        const float z = (kp[3 * t + 2] - 0.5f) * c.depth_range + rootz;
        const float x = (u - cx) / fx * z, y = (v - cy) / fy * z;
        if (t < 21) { pj[t][0] = x; pj[t][1] = y; pj[t][2] = z; }
        else { Rm[0] = x; Rm[1] = y; Rm[2] = z; }  // box root, parked until the rotation is written
    }
    if (t >= 32 && t < 32 + 21) {
        const int j = t - 32;
        mj[j] = A.joints_vis[b * 21 + j];
        for (int k = 0; k < 3; ++k) tj[j][k] = A.joints_3d[(b * 21 + j) * 3 + k] + A.root_joint[b * 3 + k];
    }
    if (t >= 64 && t < 72) {
        const int j = t - 64;
        mc[j] = A.corners_vis[b * 8 + j];
        for (int k = 0; k < 3; ++k) {
            tc[j][k] = A.corners_3d[(b * 8 + j) * 3 + k] + A.root_joint[b * 3 + k];
            can[j][k] = A.corners_can[(b * 8 + j) * 3 + k];
        }
    }
    __syncthreads();
    float boxroot[3] = {Rm[0], Rm[1], Rm[2]};
    __syncthreads();
    // 6D -> rotation (columns x, y, z); every thread keeps the intermediates the gradient needs
    const float* r6 = A.rot6d + (size_t)b * 6;
    const V3 a6 = ld3(r6), b6 = ld3(r6 + 3);
    const float na = fmaxf(sqrtf(dot(a6, a6)), 1e-8f);
    const V3 xv = (1.0f / na) * a6;
    const V3 cv = cross(xv, b6);
    const float nc = fmaxf(sqrtf(dot(cv, cv)), 1e-8f);
    const V3 zv = (1.0f / nc) * cv;
    const V3 yv = cross(zv, xv);
    if (t == 0) {
        Rm[0] = xv.x; Rm[1] = yv.x; Rm[2] = zv.x;
        Rm[3] = xv.y; Rm[4] = yv.y; Rm[5] = zv.y;
        Rm[6] = xv.z; Rm[7] = yv.z; Rm[8] = zv.z;
    }
    __syncthreads();
    if (t < 8)
        for (int k = 0; k < 3; ++k) pc[t][k] = Rm[3 * k] * can[t][0] + Rm[3 * k + 1] * can[t][1] + Rm[3 * k + 2] * can[t][2] + boxroot[k];
    if (t < 63) gj[t / 3][t % 3] = 0.0f;
    if (t >= 64 && t < 88) gc[(t - 64) / 3][(t - 64) % 3] = 0.0f;
    if (t < kParts) sums[t] = 0.0f;
    __syncthreads();
    // ---- the seven outputs
    {
        const float* rj = pj[c.center_idx];
        if (t < 63) {
            const int j = t / 3, k = t % 3;
            A.o_joints_abs[(size_t)b * 63 + t] = pj[j][k];
            A.o_joints_rel[(size_t)b * 63 + t] = pj[j][k] - rj[k];
        }
        if (t >= 64 && t < 88) {
            const int i = t - 64, j = i / 3, k = i % 3;
            A.o_corners_abs[(size_t)b * 24 + i] = pc[j][k];
            A.o_corners_rel[(size_t)b * 24 + i] = pc[j][k] - rj[k];
        }
        if (t >= 96 && t < 105) A.o_rotmat[(size_t)b * 9 + (t - 96)] = Rm[t - 96];
        if (t >= 105 && t < 108) A.o_boxroot[(size_t)b * 3 + (t - 105)] = boxroot[t - 105];
        // 2d_uvd [30,3]: kp3d[0:21], projected corners (u / W, v / H, 0), kp3d[21]
        float* uvd = A.o_uvd + (size_t)b * 90;
        if (t < 63) uvd[t] = kp[t];
        if (t >= 64 && t < 72) {
            const int j = t - 64;
            const float X = pc[j][0], Y = pc[j][1], Z = pc[j][2];
            const float pu = K[0] * X + K[1] * Y + K[2] * Z, pv = K[3] * X + K[4] * Y + K[5] * Z, pw = K[6] * X + K[7] * Y + K[8] * Z;
            uvd[63 + 3 * j] = pu / pw / c.img_w; uvd[63 + 3 * j + 1] = pv / pw / c.img_h; uvd[63 + 3 * j + 2] = 0.0f;
        }
        if (t >= 72 && t < 75) uvd[87 + (t - 72)] = kp[63 + (t - 72)];
    }
    // masked copies used by the ordinal losses (ordinal.py:118-121,255-262): p * vis, t * vis
    __syncthreads();
    float raw_pj = 0.0f, raw_pc = 0.0f;   // unmasked predictions of this thread's coordinate, for the MSE terms
    if (t < 63) raw_pj = pj[t / 3][t % 3];
    if (t >= 64 && t < 88) raw_pc = pc[(t - 64) / 3][(t - 64) % 3];
    __syncthreads();
    if (t < 63) { pj[t / 3][t % 3] *= mj[t / 3]; tj[t / 3][t % 3] *= mj[t / 3]; }
    if (t >= 64 && t < 88) { pc[(t - 64) / 3][(t - 64) % 3] *= mc[(t - 64) / 3]; tc[(t - 64) / 3][(t - 64) % 3] *= mc[(t - 64) / 3]; }
    __syncthreads();

    // ---- JointsLoss: mse(p * m, t * m), mean over B * n * 3
    {
        float e = 0.0f;
        if (c.w_joints != 0.0f && t < 63) {
            const float d = pj[t / 3][t % 3] - tj[t / 3][t % 3];
            e = d * d;
            gj[t / 3][t % 3] += c.w_joints * 2.0f * d * mj[t / 3] / (float)(B * 63);
        }
        const float s = block_sum(e, slots);
        if (t == 0) sums[0] = s;
        e = 0.0f;
        if (c.w_corners != 0.0f && t >= 64 && t < 88) {
            const int i = t - 64;
            const float d = pc[i / 3][i % 3] - tc[i / 3][i % 3];
            e = d * d;
            gc[i / 3][i % 3] += c.w_corners * 2.0f * d * mc[i / 3] / (float)(B * 24);
        }
        const float s2 = block_sum(e, slots);
        if (t == 0) sums[1] = s2;
    }
    (void)raw_pj; (void)raw_pc;

    // ---- HandOrdLoss, joint level
    if (c.w_joint_ord != 0.0f && c.n_pairs_joint > 0) {
        const int n = c.n_pairs_joint, V = c.n_views_hand;
        float acc = 0.0f;
        for (int i = t; i < n * V; i += kTailThreads) {
            const int p = i / V, v = i - p * V;
            const int ja = A.jp[2 * p], jb = A.jp[2 * p + 1];
            const V3 vv = ld3(A.vv_hand + 3 * v);
            const V3 dt = ld3(tj[ja]) - ld3(tj[jb]), dp = ld3(pj[ja]) - ld3(pj[jb]);
            float g;
            acc += ord_log(dot(dt, vv), dot(dp, vv), &g);
            dyn[i] = g;
        }
        const float s = block_sum(acc, slots);
        if (t == 0) sums[2] = s;
        __syncthreads();
        float* W = dyn + n * V;  // per pair: sum_v coef * view
        for (int i = t; i < 3 * n; i += kTailThreads) {
            const int p = i / 3, k = i - 3 * p;
            float w = 0.0f;
            for (int v = 0; v < V; ++v) w += dyn[p * V + v] * A.vv_hand[3 * v + k];
            W[i] = w;
        }
        __syncthreads();
        if (t < 63) {
            const int j = t / 3, k = t % 3;
            float g = 0.0f;
            for (int p = 0; p < n; ++p) {
                if (A.jp[2 * p] == j) g += W[3 * p + k];
                if (A.jp[2 * p + 1] == j) g -= W[3 * p + k];
            }
            gj[j][k] += c.w_joint_ord * g * mj[j] / (float)(B * n * V);
        }
        __syncthreads();
    }
    // ---- HandOrdLoss, part level: parts = (p - p[parent])[1:], cross products of part pairs against the views
    if (c.w_part_ord != 0.0f && c.n_pairs_part > 0) {
        const int n = c.n_pairs_part, V = c.n_views_hand;
        if (t < 60) {
            const int i = t / 3, k = t % 3;
            ppart[i][k] = pj[i + 1][k] - pj[c_parents[i + 1]][k];
            tpart[i][k] = tj[i + 1][k] - tj[c_parents[i + 1]][k];
            gpart[i][k] = 0.0f;
        }
        __syncthreads();
        float acc = 0.0f;
        for (int i = t; i < n * V; i += kTailThreads) {
            const int p = i / V, v = i - p * V;
            const int pa = A.pp[2 * p], pb = A.pp[2 * p + 1];
            const V3 vv = ld3(A.vv_hand + 3 * v);
            const float to = dot(cross(ld3(tpart[pa]), ld3(tpart[pb])), vv), po = dot(cross(ld3(ppart[pa]), ld3(ppart[pb])), vv);
            const float s = sgn(to), x = -s * po;
            acc += fmaxf(x, 0.0f);
            dyn[i] = x > 0.0f ? -s : 0.0f;
        }
        const float s = block_sum(acc, slots);
        if (t == 0) sums[3] = s;
        __syncthreads();
        float* W = dyn + n * V;
        for (int i = t; i < 3 * n; i += kTailThreads) {
            const int p = i / 3, k = i - 3 * p;
            float w = 0.0f;
            for (int v = 0; v < V; ++v) w += dyn[p * V + v] * A.vv_hand[3 * v + k];
            W[i] = w;
        }
        __syncthreads();
        // d((a x b) . w) / da = b x w,  / db = w x a; one thread per part sums its pairs in list order
        if (t < 20) {
            V3 g = {0.f, 0.f, 0.f};
            for (int p = 0; p < n; ++p) {
                const int pa = A.pp[2 * p], pb = A.pp[2 * p + 1];
                const V3 w = ld3(W + 3 * p);
                if (pa == t) g = g + cross(ld3(ppart[pb]), w);
                if (pb == t) g = g + cross(w, ld3(ppart[pa]));
            }
            gpart[t][0] = g.x; gpart[t][1] = g.y; gpart[t][2] = g.z;
        }
        __syncthreads();
        if (t < 63) {  // part i belongs to joint i + 1 (+) and to its parent (-)
            const int j = t / 3, k = t % 3;
            float g = j >= 1 ? gpart[j - 1][k] : 0.0f;
            for (int i = 0; i < 20; ++i)
                if (c_parents[i + 1] == j) g -= gpart[i][k];
            gj[j][k] += c.w_part_ord * g * mj[j] / (float)(B * n * V);
        }
        __syncthreads();
    }
    // ---- SceneOrdLoss: (hand joint, box corner) pairs
    if (c.w_scene_ord != 0.0f && c.n_pairs_scene > 0) {
        const int n = c.n_pairs_scene, V = c.n_views_scene;
        float acc = 0.0f;
        for (int i = t; i < n * V; i += kTailThreads) {
            const int p = i / V, v = i - p * V;
            const int ja = A.hp[2 * p], cb = A.hp[2 * p + 1];
            const V3 vv = ld3(A.vv_scene + 3 * v);
            const V3 dt = ld3(tj[ja]) - ld3(tc[cb]), dp = ld3(pj[ja]) - ld3(pc[cb]);
            float g;
            acc += ord_log(dot(dt, vv), dot(dp, vv), &g);
            dyn[i] = g;
        }
        const float s = block_sum(acc, slots);
        if (t == 0) sums[4] = s;
        __syncthreads();
        float* W = dyn + n * V;
        for (int i = t; i < 3 * n; i += kTailThreads) {
            const int p = i / 3, k = i - 3 * p;
            float w = 0.0f;
            for (int v = 0; v < V; ++v) w += dyn[p * V + v] * A.vv_scene[3 * v + k];
            W[i] = w;
        }
        __syncthreads();
        const float inv = c.w_scene_ord / (float)(B * n * V);
        if (t < 63) {
            const int j = t / 3, k = t % 3;
            float g = 0.0f;
            for (int p = 0; p < n; ++p)
                if (A.hp[2 * p] == j) g += W[3 * p + k];
            gj[j][k] += inv * g * mj[j];
        }
        if (t >= 64 && t < 88) {
            const int i = t - 64, j = i / 3, k = i % 3;
            float g = 0.0f;
            for (int p = 0; p < n; ++p)
                if (A.hp[2 * p + 1] == j) g -= W[3 * p + k];
            gc[j][k] += inv * g * mc[j];
        }
        __syncthreads();
    }
    // ---- SymCornerLoss: mse to the closest of the object's symmetric corner sets
    if (c.w_sym != 0.0f && c.n_sym > 0) {
        const int Ks = c.n_sym, oi = A.obj_idx[b] - 1;
        const float* T = A.obj_transf + (size_t)b * 16;
        float my_err = 3.4e38f;
        int my_k = 0x7fffffff;
        for (int k = t; k < Ks; k += kTailThreads) {
            const float* R = A.sym_R + ((size_t)oi * Ks + k) * 9;
            const float* tr = A.sym_t + ((size_t)oi * Ks + k) * 3;
            float e = 0.0f;
            for (int j = 0; j < 8; ++j) {
                float q[3], s[3];
                if (!c.sym_ho3d) {
                    for (int m = 0; m < 3; ++m) q[m] = R[3 * m] * can[j][0] + R[3 * m + 1] * can[j][1] + R[3 * m + 2] * can[j][2] + tr[m];
                } else {  // ext (R (ext c) + t), ext = diag(1, -1, -1)
                    const float e0 = can[j][0], e1 = -can[j][1], e2 = -can[j][2];
                    for (int m = 0; m < 3; ++m) q[m] = R[3 * m] * e0 + R[3 * m + 1] * e1 + R[3 * m + 2] * e2 + tr[m];
                    q[1] = -q[1]; q[2] = -q[2];
                }
                for (int m = 0; m < 3; ++m) {
                    s[m] = (T[4 * m] * q[0] + T[4 * m + 1] * q[1] + T[4 * m + 2] * q[2] + T[4 * m + 3]) * mc[j];
                    const float d = s[m] - pc[j][m];
                    e += d * d;
                }
            }
            e *= 1.0f / 24.0f;
            if (e < my_err) { my_err = e; my_k = k; }
        }
        // arg-min over the CTA: smallest error, ties to the lower index
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oe = __shfl_down_sync(0xffffffffu, my_err, o);
            const int ok = __shfl_down_sync(0xffffffffu, my_k, o);
            if (oe < my_err || (oe == my_err && ok < my_k)) { my_err = oe; my_k = ok; }
        }
        __shared__ float werr[kTailThreads / 32];
        __shared__ int wk[kTailThreads / 32];
        if ((t & 31) == 0) { werr[t >> 5] = my_err; wk[t >> 5] = my_k; }
        __syncthreads();
        if (t == 0) {
            float be = werr[0];
            int bk = wk[0];
            for (int i = 1; i < kTailThreads / 32; ++i)
                if (werr[i] < be || (werr[i] == be && wk[i] < bk)) { be = werr[i]; bk = wk[i]; }
            best_err = be; best_k = bk;
            sums[5] = be;
        }
        __syncthreads();
        if (t >= 64 && t < 72) {
            const int j = t - 64, k = best_k;
            const float* R = A.sym_R + ((size_t)oi * Ks + k) * 9;
            const float* tr = A.sym_t + ((size_t)oi * Ks + k) * 3;
            float q[3];
            if (!c.sym_ho3d) {
                for (int m = 0; m < 3; ++m) q[m] = R[3 * m] * can[j][0] + R[3 * m + 1] * can[j][1] + R[3 * m + 2] * can[j][2] + tr[m];
            } else {
                const float e0 = can[j][0], e1 = -can[j][1], e2 = -can[j][2];
                for (int m = 0; m < 3; ++m) q[m] = R[3 * m] * e0 + R[3 * m + 1] * e1 + R[3 * m + 2] * e2 + tr[m];
                q[1] = -q[1]; q[2] = -q[2];
            }
            for (int m = 0; m < 3; ++m) {
                const float s = (T[4 * m] * q[0] + T[4 * m + 1] * q[1] + T[4 * m + 2] * q[2] + T[4 * m + 3]) * mc[j];
                gc[j][m] += c.w_sym * 2.0f * (pc[j][m] - s) * mc[j] / (24.0f * (float)B);
            }
        }
        __syncthreads();
    }

    // ---- chain rule back to rot6d and kp3d
    if (t == 0) {
        // corners = R can + boxroot:  dR = sum_c g_c (x) can_c,  d boxroot = sum_c g_c
        V3 gx = {0, 0, 0}, gy = {0, 0, 0}, gz = {0, 0, 0};  // gradients of the columns x, y, z of R
        for (int j = 0; j < 8; ++j) {
            const V3 g = ld3(gc[j]);
            gx = gx + can[j][0] * g; gy = gy + can[j][1] * g; gz = gz + can[j][2] * g;
        }
        // y = z x x
        gz = gz + cross(xv, gy);
        gx = gx + cross(gy, zv);
        // z = cv / |cv| (norm clamped at 1e-8: constant denominator below the clamp)
        V3 gcv = sqrtf(dot(cv, cv)) > 1e-8f ? (1.0f / nc) * (gz - dot(zv, gz) * zv) : (1.0f / nc) * gz;
        // cv = x x b
        gx = gx + cross(b6, gcv);
        const V3 gb = cross(gcv, xv);
        // x = a / |a|
        const V3 ga = sqrtf(dot(a6, a6)) > 1e-8f ? (1.0f / na) * (gx - dot(xv, gx) * xv) : (1.0f / na) * gx;
        float* d6 = A.d_rot6d + (size_t)b * 6;
        d6[0] = ga.x; d6[1] = ga.y; d6[2] = ga.z; d6[3] = gb.x; d6[4] = gb.y; d6[5] = gb.z;
    }
    if (t < 22) {
        V3 g;
        if (t < 21) g = ld3(gj[t]);
        else {
            g = {0, 0, 0};
            for (int j = 0; j < 8; ++j) g = g + ld3(gc[j]);
        }
        const float u = kp[3 * t] * c.inp_w, v = kp[3 * t + 1] * c.inp_h;
        const float z = (kp[3 * t + 2] - 0.5f) * c.depth_range + rootz;
        float* d = A.d_kp3d + (size_t)b * 66 + 3 * t;
        d[0] = g.x * c.inp_w / fx * z;
        d[1] = g.y * c.inp_h / fy * z;
        d[2] = c.depth_range * (g.x * (u - cx) / fx + g.y * (v - cy) / fy + g.z);
    }
    if (t < kParts) A.partial[(size_t)b * kParts + t] = sums[t];
}

// parts[i] = (sum over the batch, in index order) / normaliser; parts[7] = weighted total
__global__ void tail_loss_finalize_kernel(const float* __restrict__ partial, const ab_tail_cfg c, float* __restrict__ parts) {
    __shared__ float s[kParts];
    const int t = threadIdx.x;
    if (t < 6) {
        float a = 0.0f;
        for (int b = 0; b < c.batch; ++b) a += partial[(size_t)b * kParts + t];
        const float B = (float)c.batch;
        const float norm[6] = {B * 63.0f, B * 24.0f, B * (float)(c.n_pairs_joint * c.n_views_hand), B * (float)(c.n_pairs_part * c.n_views_hand),
                               B * (float)(c.n_pairs_scene * c.n_views_scene), B};
        s[t] = norm[t] > 0.0f ? a / norm[t] : 0.0f;
        parts[t] = s[t];
    }
    __syncthreads();
    if (t == 0) {
        parts[6] = 0.0f;
        parts[7] = c.w_joints * s[0] + c.w_corners * s[1] + c.w_joint_ord * s[2] + c.w_part_ord * s[3] + c.w_scene_ord * s[4] + c.w_sym * s[5];
    }
}

}  // namespace ab

extern "C" uint64_t ab_tail_losses_workspace_bytes(int batch) { return batch > 0 ? (uint64_t)batch * ab::kParts * sizeof(float) : 0; }

extern "C" int ab_tail_losses(const ab_tail_cfg* cfg, const float* kp3d, const float* rot6d, const float* root_joint,
                              const float* cam_intr, const float* corners_can, const float* joints_3d, const float* corners_3d,
                              const float* joints_vis, const float* corners_vis, const float* vv_hand, const int32_t* jp,
                              const int32_t* pp, const float* vv_scene, const int32_t* hp, const float* sym_R, const float* sym_t,
                              const int32_t* obj_idx, const float* obj_transf, float* joints_abs, float* corners_abs,
                              float* joints_rel, float* corners_rel, float* uvd2d, float* boxroot, float* rotmat, float* parts,
                              float* d_kp3d, float* d_rot6d, void* ws, void* stream) {
    using namespace ab;
    AB_REQUIRE(cfg != nullptr && cfg->batch >= 0, "bad config");
    if (cfg->batch == 0) return AB_OK;
    AB_REQUIRE(kp3d && rot6d && root_joint && cam_intr && corners_can && joints_3d && corners_3d && joints_vis && corners_vis,
               "null input");
    AB_REQUIRE(joints_abs && corners_abs && joints_rel && corners_rel && uvd2d && boxroot && rotmat && parts && d_kp3d && d_rot6d && ws,
               "null output / workspace");
    AB_REQUIRE(cfg->center_idx >= 0 && cfg->center_idx < 21, "center_idx out of range");
    const bool hand = cfg->w_joint_ord != 0.0f || cfg->w_part_ord != 0.0f;
    AB_REQUIRE(!hand || (vv_hand && cfg->n_views_hand > 0), "HandOrdLoss needs its view vectors");
    AB_REQUIRE(cfg->w_joint_ord == 0.0f || cfg->n_pairs_joint == 0 || jp, "null joint pair list");
    AB_REQUIRE(cfg->w_part_ord == 0.0f || cfg->n_pairs_part == 0 || pp, "null part pair list");
    AB_REQUIRE(cfg->w_scene_ord == 0.0f || cfg->n_pairs_scene == 0 || (vv_scene && hp && cfg->n_views_scene > 0), "SceneOrdLoss needs views and pairs");
    AB_REQUIRE(cfg->w_sym == 0.0f || cfg->n_sym == 0 || (sym_R && sym_t && obj_idx && obj_transf), "SymCornerLoss needs its tables");
    const int n1 = cfg->n_pairs_joint * (cfg->n_views_hand + 3), n2 = cfg->n_pairs_part * (cfg->n_views_hand + 3),
              n3 = cfg->n_pairs_scene * (cfg->n_views_scene + 3);
    const size_t dyn = sizeof(float) * (size_t)max(1, max(n1, max(n2, n3)));
    AB_REQUIRE(dyn <= 160 * 1024, "too many (pair, view) terms for one CTA");
    if (dyn > 40 * 1024) AB_CUDA(cudaFuncSetAttribute(tail_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    TailArgs A;
    A.cfg = *cfg;
    A.kp3d = kp3d; A.rot6d = rot6d; A.root_joint = root_joint; A.cam_intr = cam_intr; A.corners_can = corners_can;
    A.joints_3d = joints_3d; A.corners_3d = corners_3d; A.joints_vis = joints_vis; A.corners_vis = corners_vis;
    A.vv_hand = vv_hand; A.jp = jp; A.pp = pp; A.vv_scene = vv_scene; A.hp = hp;
    A.sym_R = sym_R; A.sym_t = sym_t; A.obj_idx = obj_idx; A.obj_transf = obj_transf;
    A.o_joints_abs = joints_abs; A.o_corners_abs = corners_abs; A.o_joints_rel = joints_rel; A.o_corners_rel = corners_rel;
    A.o_uvd = uvd2d; A.o_boxroot = boxroot; A.o_rotmat = rotmat;
    A.d_kp3d = d_kp3d; A.d_rot6d = d_rot6d; A.partial = (float*)ws;
    cudaStream_t st = (cudaStream_t)stream;
    {
        StageTimer tm(AB_STAGE_TAIL_LOSS, st);
        tail_loss_kernel<<<cfg->batch, kTailThreads, dyn, st>>>(A);
        tail_loss_finalize_kernel<<<1, 32, 0, st>>>((const float*)ws, *cfg, parts);
    }
    count_launch(2);
    return check_launch("ab_tail_losses");
}
