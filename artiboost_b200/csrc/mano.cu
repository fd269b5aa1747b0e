// MANO linear-blend skinning as one fused fp32 kernel (sm_100a).
//
// Replaces manotorch.ManoLayer.forward as the hot path calls it (anakin/artiboost/preprocessor.py:25,62,
// anakin/artiboost/refiner.py:138); algorithm per anakin/postprocess/iknet/manolayer.py:182-276.
//
// Work split: one CTA = kS = 8 samples x 128 of the vertices.  The CTA's slice of the blend-shape matrix (145 x 384
// fp32 = 223 KB of the 1.35 MB, L2 resident) is streamed once per CTA with column-contiguous (coalesced) loads, 27 loads
// in flight per thread, and applied to the 8 samples at a time from registers (round 1 ran 4 samples x a third of the
// vertices: 173 MB of L2 reads per 512 samples; now 100 MB); the per-sample 16-bone chain is redone by every CTA of a sample group (a few hundred
// flops) so no inter-CTA exchange is needed.  Outputs are staged in shared memory and written as contiguous
// float runs.  The optional rigid map `post_rt` fuses the pose generator's camera transform
// (preprocessor.py:84-88) into the store.
#include "mano_math.cuh"

namespace ab {

constexpr int kV = AB_MANO_VERTS;
constexpr int kS = 8;          // samples per CTA: every blend-shape weight fetched from L2 feeds 8 FMAs
constexpr int kSplit = 7;      // vertex ranges per sample group
constexpr int kVPer = 128;     // vertices per range (7*128 >= 778): 384 columns = 3 rounds of the 128 threads
                               // (19 ranges of 42 vertices -- 1216 CTAs, 24 warps per SM -- measured the same 46 us: the launch
                               // is bound by the dependent phases of a CTA, not by occupancy or L2 bandwidth, 2 TB/s)
constexpr int kMaxExtra = 6;   // 5 tips + centre tip
constexpr int kThreads = 128;
constexpr int kNCoef = 10 + AB_MANO_POSE_FEAT;

struct alignas(16) ManoSmem {
    float R[kS][16][9];
    float J[kS][16][3];
    float G[kS][16][12];
    float A[kS][16][12];
    float coef[kNCoef][kS];                 // [0,10) betas, [10,145) pose map
    float vp[kS][(kVPer + kMaxExtra) * 3];  // v_posed, then skinned verts in place
    float post[kS][12];
    float centre[kS][3];
    int extra[kMaxExtra];
    int n_extra;
};

__global__ void __launch_bounds__(kThreads)
mano_lbs_kernel(ab_mano_model m, int batch, const float* __restrict__ pose, const float* __restrict__ betas,
                const float* __restrict__ post_rt, int center_idx, float* __restrict__ verts,
                float* __restrict__ joints, float* __restrict__ transforms_abs) {
    __shared__ ManoSmem sm;
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * kS;
    const int split = blockIdx.y;
    const int v0 = split * kVPer;
    const int nv_own = min(kV, v0 + kVPer) - v0;
    const int centre_chain = center_idx < 0 ? -1 : kJointReorder[center_idx];  // index into [16 chain, 5 tips]

    // ---- extra vertices this CTA must also skin: the 5 tips (split 0 writes the joints) and a tip centre
    if (tid == 0) {
        int n = 0;
        if (split == 0)
            for (int i = 0; i < 5; ++i) sm.extra[n++] = kTipVerts[i];
        if (centre_chain >= 16) sm.extra[n++] = kTipVerts[centre_chain - 16];
        sm.n_extra = n;
    }
    // ---- per-joint rotations, pose map, betas, rest joints
    for (int i = tid; i < kS * 16; i += kThreads) {
        int s = i >> 4, k = i & 15;
        int b = min(b0 + s, batch - 1);
        const float* p = pose + (size_t)b * 48 + 3 * k;
        float R[9];
        rodrigues(p[0], p[1], p[2], R);
#pragma unroll
        for (int j = 0; j < 9; ++j) sm.R[s][k][j] = R[j];
        if (k > 0) {
#pragma unroll
            for (int j = 0; j < 9; ++j) sm.coef[10 + 9 * (k - 1) + j][s] = R[j] - ((j == 0 || j == 4 || j == 8) ? 1.0f : 0.0f);
        }
    }
    for (int i = tid; i < kS * 10; i += kThreads) {
        int s = i / 10, k = i % 10;
        int b = min(b0 + s, batch - 1);
        sm.coef[k][s] = betas ? betas[(size_t)b * 10 + k] : 0.0f;
    }
    for (int i = tid; i < kS * 12; i += kThreads) {
        int s = i / 12, k = i % 12;
        int b = min(b0 + s, batch - 1);
        sm.post[s][k] = post_rt ? post_rt[(size_t)b * 12 + k] : ((k == 0 || k == 4 || k == 8) ? 1.0f : 0.0f);
    }
    __syncthreads();
    for (int i = tid; i < kS * 48; i += kThreads) {
        int s = i / 48, r = i % 48;
        float acc = m.j_template[r];
#pragma unroll
        for (int k = 0; k < 10; ++k) acc += m.j_shapedirs[r * 10 + k] * sm.coef[k][s];
        sm.J[s][r / 3][r % 3] = acc;
    }
    __syncthreads();
    // ---- kinematic chain, one thread per (sample, finger): the root, then the finger's three joints; A_k = [G_R | G_t - G_R J_k]
    if (tid < kS * 5) {
        const int s = tid / 5, f = tid - 5 * s;
        const float* R = &sm.R[s][0][0];
        const float* J = &sm.J[s][0][0];
        float g[12], a[12];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            g[4 * i] = R[3 * i]; g[4 * i + 1] = R[3 * i + 1]; g[4 * i + 2] = R[3 * i + 2];
            g[4 * i + 3] = J[i];
        }
        auto emit_joint = [&](int k, const float* gk) {
            const float* j = J + 3 * k;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                a[4 * i] = gk[4 * i]; a[4 * i + 1] = gk[4 * i + 1]; a[4 * i + 2] = gk[4 * i + 2];
                a[4 * i + 3] = gk[4 * i + 3] - (gk[4 * i] * j[0] + gk[4 * i + 1] * j[1] + gk[4 * i + 2] * j[2]);
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) { sm.G[s][k][i] = gk[i]; sm.A[s][k][i] = a[i]; }
        };
        if (f == 0) emit_joint(0, g);
        int p = 0;
#pragma unroll 1
        for (int q = 0; q < 3; ++q) {   // joints 1 + 3f .. 3 + 3f, each the child of the previous (kManoParents)
            const int k = 1 + 3 * f + q;
            const float* r = R + 9 * k;
            const float rel[3] = {J[3 * k] - J[3 * p], J[3 * k + 1] - J[3 * p + 1], J[3 * k + 2] - J[3 * p + 2]};
            float n[12];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float x = g[4 * i], y = g[4 * i + 1], z = g[4 * i + 2];
                n[4 * i] = x * r[0] + y * r[3] + z * r[6];
                n[4 * i + 1] = x * r[1] + y * r[4] + z * r[7];
                n[4 * i + 2] = x * r[2] + y * r[5] + z * r[8];
                n[4 * i + 3] = x * rel[0] + y * rel[1] + z * rel[2] + g[4 * i + 3];
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) g[i] = n[i];
            emit_joint(k, g);
            p = k;
        }
    }
    // ---- blend shapes: v_posed[col] = template[col] + sum_k dirs[k][col] * coef[k], kS samples per load
    const int n_extra = sm.n_extra;  // written before the first barrier
    const int nv = nv_own + n_extra;
    const int ncol = nv * 3;
    for (int c = tid; c < ncol; c += kThreads) {
        int lv = c / 3, d = c - 3 * lv;
        int gcol = (lv < nv_own ? v0 + lv : sm.extra[lv - nv_own]) * 3 + d;
        float acc[kS];
        float t = m.v_template[gcol];
#pragma unroll
        for (int s = 0; s < kS; ++s) acc[s] = t;
        static_assert(kS == 8, "the blend loop reads the coefficients of 8 samples as two float4");
#pragma unroll 10
        for (int k = 0; k < 10; ++k) {
            const float w = __ldg(m.shapedirs_t + (size_t)k * (kV * 3) + gcol);
            const float4 c0 = *reinterpret_cast<const float4*>(&sm.coef[k][0]), c1 = *reinterpret_cast<const float4*>(&sm.coef[k][4]);
            acc[0] += w * c0.x; acc[1] += w * c0.y; acc[2] += w * c0.z; acc[3] += w * c0.w;
            acc[4] += w * c1.x; acc[5] += w * c1.y; acc[6] += w * c1.z; acc[7] += w * c1.w;
        }
        // 135 = 5 x 27: 27 independent L2 loads in flight per thread (the kernel runs at 12 warps / SM: latency, not issue)
#pragma unroll 27
        for (int k = 0; k < AB_MANO_POSE_FEAT; ++k) {
            const float w = __ldg(m.posedirs_t + (size_t)k * (kV * 3) + gcol);
            const float4 c0 = *reinterpret_cast<const float4*>(&sm.coef[10 + k][0]), c1 = *reinterpret_cast<const float4*>(&sm.coef[10 + k][4]);
            acc[0] += w * c0.x; acc[1] += w * c0.y; acc[2] += w * c0.z; acc[3] += w * c0.w;
            acc[4] += w * c1.x; acc[5] += w * c1.y; acc[6] += w * c1.z; acc[7] += w * c1.w;
        }
#pragma unroll
        for (int s = 0; s < kS; ++s) sm.vp[s][c] = acc[s];
    }
    __syncthreads();
    // ---- skinning: x = (sum_k w_vk A_k) [v_posed; 1], in place; a thread per vertex, its 16 weights loaded once for
    // the kS samples
    for (int lv = tid; lv < nv; lv += kThreads) {
        const int gv = lv < nv_own ? v0 + lv : sm.extra[lv - nv_own];
        const float4* wp = reinterpret_cast<const float4*>(m.weights + (size_t)gv * 16);
        float w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w4 = __ldg(wp + q);
            w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
        }
#pragma unroll 1
        for (int s = 0; s < kS; ++s) {
            float T[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) T[j] = 0.0f;
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                if (w[r] != 0.0f) {
                    const float* a = sm.A[s][r];
#pragma unroll
                    for (int j = 0; j < 12; ++j) T[j] += w[r] * a[j];
                }
            }
            float* p = &sm.vp[s][3 * lv];
            const float x = p[0], y = p[1], z = p[2];
            p[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
            p[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
            p[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
        }
    }
    __syncthreads();
    if (tid < kS * 3) {
        int s = tid / 3, d = tid % 3;
        float c = 0.0f;
        if (centre_chain >= 16) c = sm.vp[s][3 * (nv - 1) + d];
        else if (centre_chain >= 0) c = sm.G[s][centre_chain][4 * d + 3];
        sm.centre[s][d] = c;
    }
    __syncthreads();
    // ---- joints + transforms (split 0), from the un-centred values
    if (split == 0) {
        for (int i = tid; i < kS * 21; i += kThreads) {
            int s = i / 21, jn = i - 21 * s;
            if (b0 + s >= batch) continue;
            int src = kJointReorder[jn];
            float p[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
                p[d] = (src < 16 ? sm.G[s][src][4 * d + 3] : sm.vp[s][3 * (nv_own + src - 16) + d]) - sm.centre[s][d];
            const float* q = sm.post[s];
            float* o = joints + ((size_t)(b0 + s) * 21 + jn) * 3;
            o[0] = q[0] * p[0] + q[1] * p[1] + q[2] * p[2] + q[9];
            o[1] = q[3] * p[0] + q[4] * p[1] + q[5] * p[2] + q[10];
            o[2] = q[6] * p[0] + q[7] * p[1] + q[8] * p[2] + q[11];
        }
        if (transforms_abs) {
            for (int i = tid; i < kS * 256; i += kThreads) {
                int s = i >> 8, r = i & 255;
                if (b0 + s >= batch) continue;
                int k = r >> 4, e = r & 15;
                float v = e < 12 ? sm.G[s][kTransfReorder[k]][e] : (e == 15 ? 1.0f : 0.0f);
                transforms_abs[(size_t)(b0 + s) * 256 + r] = v;
            }
        }
    }
    // ---- centre + rigid map per own vertex (in place), then contiguous stores
    for (int i = tid; i < nv_own * kS; i += kThreads) {
        int s = i / nv_own, lv = i - s * nv_own;
        float* p = &sm.vp[s][3 * lv];
        const float* q = sm.post[s];
        float x = p[0] - sm.centre[s][0], y = p[1] - sm.centre[s][1], z = p[2] - sm.centre[s][2];
        p[0] = q[0] * x + q[1] * y + q[2] * z + q[9];
        p[1] = q[3] * x + q[4] * y + q[5] * z + q[10];
        p[2] = q[6] * x + q[7] * y + q[8] * z + q[11];
    }
    __syncthreads();
    for (int s = 0; s < kS; ++s) {
        if (b0 + s >= batch) break;
        float* o = verts + ((size_t)(b0 + s) * kV + v0) * 3;
        for (int c = tid; c < nv_own * 3; c += kThreads) o[c] = sm.vp[s][c];
    }
}

int launch_mano(const ab_mano_model* model, int batch, const float* pose, const float* betas, const float* post_rt,
                int center_idx, float* verts, float* joints, float* transforms_abs, cudaStream_t st) {
    dim3 grid(cdiv(batch, kS), kSplit);
    StageTimer tm(AB_STAGE_MANO_LBS, st);
    mano_lbs_kernel<<<grid, kThreads, 0, st>>>(*model, batch, pose, betas, post_rt, center_idx, verts, joints,
                                              transforms_abs);
    count_launch();
    return check_launch("mano_lbs_kernel");
}

}  // namespace ab

extern "C" int ab_mano_forward(const ab_mano_model* model, int batch, const float* pose, const float* betas,
                               const float* post_rt, int center_idx, float* verts, float* joints,
                               float* transforms_abs, void* stream) {
    AB_REQUIRE(model && model->v_template && model->shapedirs_t && model->posedirs_t && model->j_template &&
                   model->j_shapedirs && model->weights, "null model array");
    AB_REQUIRE(batch >= 0, "negative batch");
    AB_REQUIRE(center_idx < AB_MANO_KEYPOINTS, "center_idx out of range");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(pose && verts && joints, "null pose/verts/joints");
    return ab::launch_mano(model, batch, pose, betas, post_rt, center_idx, verts, joints, transforms_abs,
                           (cudaStream_t)stream);
}
