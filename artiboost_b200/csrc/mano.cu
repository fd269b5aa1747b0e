// MANO linear-blend skinning as one fused fp32 kernel (sm_100a).
//
// Replaces manotorch.ManoLayer.forward as the hot path calls it (anakin/artiboost/preprocessor.py:25,62,
// anakin/artiboost/refiner.py:138); algorithm per anakin/postprocess/iknet/manolayer.py:182-276.
//
// Work split: one CTA = kS = 16 samples x 42 of the vertices (19 vertex ranges), 128 threads, four CTAs per SM.
//   * blend shapes: a thread owns one output column (vertex coordinate) and the 16 samples' accumulators as eight packed
//     pairs (FFMA2); every weight of the 145 x 2334 blend-shape matrix it fetches from L2 feeds 16 FMAs, and the loads
//     of 15 rows are in flight before the first is used.  L2 reads per 512 samples: 43 MB (round 1: 4 samples per CTA,
//     173 MB; 8 samples x 128 vertices: 100 MB at 12 warps per SM and 46 us per launch -- latency bound).
//   * the per-sample prelude (Rodrigues, pose map, rest joints, 16-bone chain) is redone by every CTA of a sample group
//     (a few hundred flops per sample) so no inter-CTA exchange or scratch buffer is needed; rest joints are computed
//     with one thread per regressed coordinate looping over the samples (its 10 shape coefficients stay in registers).
//   * skinning: one (vertex, sample) pair per thread; the vertex's non-zero skinning weights are compacted once per CTA
//     (MANO's weight rows are sparse), so the pair loop runs over 1-4 bones instead of branching over 16.
// Outputs are staged in shared memory and written as contiguous float runs.  The optional rigid map `post_rt` fuses the
// pose generator's camera transform (preprocessor.py:84-88) into the store.
#include <atomic>

#include "mano_math.cuh"

namespace ab {

constexpr int kV = AB_MANO_VERTS;
#ifndef AB_LBS_VPER
#define AB_LBS_VPER 42
#endif
#ifndef AB_LBS_THREADS
#define AB_LBS_THREADS 128
#endif
constexpr int kS = 16;         // samples per CTA
constexpr int kVPer = AB_LBS_VPER;   // vertices per range: its columns (+ the extra vertices') fit one round of the threads
constexpr int kSplit = (kV + kVPer - 1) / kVPer;
constexpr int kMaxExtra = 6;   // 5 tips + centre tip
constexpr int kMaxV = kVPer + kMaxExtra;
constexpr int kThreads = AB_LBS_THREADS;
static_assert(kMaxV <= kThreads - 64 && kS * 5 <= kThreads && 48 <= 64, "phase thread ranges");
constexpr int kNCoef = 10 + AB_MANO_POSE_FEAT;
constexpr int kAhead = 27;     // blend-shape rows in flight per thread (135 = 5 x 27)

struct alignas(16) ManoSmem {
    float coef[kNCoef][kS];                 // [0,10) betas, [10,145) pose map
    float A[kS][16][12];                    // skinning transforms [G_R | G_t - G_R J]
    union {                                 // R is dead once the chain has run; v_posed is written after the barrier
        float R[kS][16][9];                 // that follows the chain
        float vp[kS][kMaxV * 3];            // v_posed, then skinned verts in place
    };
    float J[kS][16][3];
    float Gt[kS][16][3];                    // joint positions (translation column of G)
    float post[kS][12];
    float centre[kS][3];
    float cw[kMaxV][16];                    // compacted skinning weights of the CTA's vertices ...
    unsigned char cb[kMaxV][16];            // ... their bones ...
    int cn[kMaxV];                          // ... and how many
    int extra[kMaxExtra];
    int n_extra;
};

#ifdef AB_LBS_TRACE
// Debug build only (tools/trace_lbs.py): globaltimer at the phase boundaries of every CTA, as seen by thread 0.
__device__ unsigned long long g_lbs_trace[8 * 4096];
__device__ __forceinline__ unsigned long long lbs_time() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define LBS_MARK(i) do { if (threadIdx.x == 0 && cta_lin < 4096) g_lbs_trace[8 * cta_lin + (i)] = lbs_time(); } while (0)
#else
#define LBS_MARK(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 4)
mano_lbs_kernel(ab_mano_model m, int batch, const float* __restrict__ pose, const float* __restrict__ betas,
                const float* __restrict__ post_rt, int center_idx, float* __restrict__ verts,
                float* __restrict__ joints, float* __restrict__ transforms_abs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ManoSmem& sm = *reinterpret_cast<ManoSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * kS;
    const int split = blockIdx.y;
    const int v0 = split * kVPer;
    const int nv_own = min(kV, v0 + kVPer) - v0;
    const int centre_chain = center_idx < 0 ? -1 : kJointReorder[center_idx];  // index into [16 chain, 5 tips]
#ifdef AB_LBS_TRACE
    const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
#endif
    LBS_MARK(0);

    // ---- extra vertices this CTA must also skin: the 5 tips (split 0 writes the joints) and a tip centre
    if (tid == 0) {
        int n = 0;
        if (split == 0)
            for (int i = 0; i < 5; ++i) sm.extra[n++] = kTipVerts[i];
        if (centre_chain >= 16) sm.extra[n++] = kTipVerts[centre_chain - 16];
        sm.n_extra = n;
    }
    // ---- per-joint rotations, pose map, betas, rigid maps
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
    LBS_MARK(1);
    const int n_extra = sm.n_extra;
    const int nv = nv_own + n_extra;
    // ---- rest joints: thread r owns regressed coordinate r (its shape row in registers) for all samples;
    //      the other threads compact the skinning weights of the CTA's vertices meanwhile
    if (tid < 48) {
        float sd[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) sd[k] = m.j_shapedirs[tid * 10 + k];
        const float jt = m.j_template[tid];
#pragma unroll 4
        for (int s = 0; s < kS; ++s) {
            float acc = jt;
#pragma unroll
            for (int k = 0; k < 10; ++k) acc += sd[k] * sm.coef[k][s];
            sm.J[s][tid / 3][tid % 3] = acc;
        }
    } else if (tid >= 64 && tid - 64 < nv) {
        const int lv = tid - 64;
        const int gv = lv < nv_own ? v0 + lv : sm.extra[lv - nv_own];
        const float4* wp = reinterpret_cast<const float4*>(m.weights + (size_t)gv * 16);
        float w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w4 = __ldg(wp + q);
            w[4 * q] = w4.x; w[4 * q + 1] = w4.y; w[4 * q + 2] = w4.z; w[4 * q + 3] = w4.w;
        }
        int n = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (w[r] != 0.0f) { sm.cw[lv][n] = w[r]; sm.cb[lv][n] = (unsigned char)r; ++n; }
        sm.cn[lv] = n;
    }
    __syncthreads();
    LBS_MARK(2);
    // ---- kinematic chain, one thread per (sample, finger): the root, then the finger's three joints; A_k = [G_R | G_t - G_R J_k]
    if (tid < kS * 5) {
        const int s = tid / 5, f = tid - 5 * s;
        const float* R = &sm.R[s][0][0];
        const float* J = &sm.J[s][0][0];
        float* tout = (split == 0 && transforms_abs && b0 + s < batch) ? transforms_abs + (size_t)(b0 + s) * 256 : nullptr;
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
            for (int i = 0; i < 12; ++i) sm.A[s][k][i] = a[i];
            sm.Gt[s][k][0] = gk[3]; sm.Gt[s][k][1] = gk[7]; sm.Gt[s][k][2] = gk[11];
            if (tout) {  // transforms_abs [16, 4, 4]: kTransfReorder is the identity (chain order)
                float4* o = reinterpret_cast<float4*>(tout + 16 * k);
                o[0] = make_float4(gk[0], gk[1], gk[2], gk[3]); o[1] = make_float4(gk[4], gk[5], gk[6], gk[7]);
                o[2] = make_float4(gk[8], gk[9], gk[10], gk[11]); o[3] = make_float4(0.f, 0.f, 0.f, 1.f);
            }
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
    LBS_MARK(3);
    // ---- blend shapes: v_posed[col] = template[col] + sum_k dirs[k][col] * coef[k], kS samples per load (packed pairs)
    const int ncol = nv * 3;
    for (int c0 = 0; c0 < ncol; c0 += kThreads) {
        const int c = c0 + tid;
        const bool on = c < ncol;
        float2 acc[kS / 2];
        if (on) {
            int lv = c / 3, d = c - 3 * lv;
            int gcol = (lv < nv_own ? v0 + lv : sm.extra[lv - nv_own]) * 3 + d;
            const float t = m.v_template[gcol];
#pragma unroll
            for (int j = 0; j < kS / 2; ++j) acc[j] = make_float2(t, t);
            auto fma_row = [&](float w, const float* crow) {
                const float2 w2 = make_float2(w, w);
#pragma unroll
                for (int j = 0; j < kS / 4; ++j) {
                    const float4 cc = *reinterpret_cast<const float4*>(crow + 4 * j);
                    acc[2 * j] = __ffma2_rn(w2, make_float2(cc.x, cc.y), acc[2 * j]);
                    acc[2 * j + 1] = __ffma2_rn(w2, make_float2(cc.z, cc.w), acc[2 * j + 1]);
                }
            };
            {
                float w[10];
#pragma unroll
                for (int k = 0; k < 10; ++k) w[k] = __ldg(m.shapedirs_t + (size_t)k * (kV * 3) + gcol);
#pragma unroll
                for (int k = 0; k < 10; ++k) fma_row(w[k], &sm.coef[k][0]);
            }
            static_assert(kMaxV * 3 >= 16 * 9, "vp covers R in the union (R is dead when vp is written)");
            static_assert(AB_MANO_POSE_FEAT % kAhead == 0, "the pose rows are fetched kAhead at a time");
#pragma unroll 1
            for (int k0 = 0; k0 < AB_MANO_POSE_FEAT; k0 += kAhead) {
                float w[kAhead];
#pragma unroll
                for (int k = 0; k < kAhead; ++k) w[k] = __ldg(m.posedirs_t + (size_t)(k0 + k) * (kV * 3) + gcol);
#pragma unroll
                for (int k = 0; k < kAhead; ++k) fma_row(w[k], &sm.coef[10 + k0 + k][0]);
            }
        }
        if (c0 == 0) __syncthreads();  // v_posed aliases R: every chain thread has finished reading it
        if (on) {
#pragma unroll
            for (int j = 0; j < kS / 2; ++j) { sm.vp[2 * j][c] = acc[j].x; sm.vp[2 * j + 1][c] = acc[j].y; }
        }
    }
    __syncthreads();
    LBS_MARK(4);
    // ---- skinning: x = (sum_k w_vk A_k) [v_posed; 1], in place; one (vertex, sample) pair per thread over the vertex's
    //      compacted bone list (ascending bone order: the same sum as a dense loop that skips zero weights)
    for (int i = tid; i < nv * kS; i += kThreads) {
        const int s = i / nv, lv = i - s * nv;
        const int n = sm.cn[lv];
        float T[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) T[j] = 0.0f;
        for (int q = 0; q < n; ++q) {
            const float w = sm.cw[lv][q];
            const float4* a = reinterpret_cast<const float4*>(sm.A[s][sm.cb[lv][q]]);
            const float4 a0 = a[0], a1 = a[1], a2 = a[2];
            T[0] += w * a0.x; T[1] += w * a0.y; T[2] += w * a0.z; T[3] += w * a0.w;
            T[4] += w * a1.x; T[5] += w * a1.y; T[6] += w * a1.z; T[7] += w * a1.w;
            T[8] += w * a2.x; T[9] += w * a2.y; T[10] += w * a2.z; T[11] += w * a2.w;
        }
        float* p = &sm.vp[s][3 * lv];
        const float x = p[0], y = p[1], z = p[2];
        p[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
        p[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
        p[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
    }
    __syncthreads();
    LBS_MARK(5);
    if (tid < kS * 3) {
        int s = tid / 3, d = tid % 3;
        float c = 0.0f;
        if (centre_chain >= 16) c = sm.vp[s][3 * (nv - 1) + d];
        else if (centre_chain >= 0) c = sm.Gt[s][centre_chain][d];
        sm.centre[s][d] = c;
    }
    __syncthreads();
    // ---- joints (split 0), from the un-centred values
    if (split == 0) {
        for (int i = tid; i < kS * 21; i += kThreads) {
            int s = i / 21, jn = i - 21 * s;
            if (b0 + s >= batch) continue;
            int src = kJointReorder[jn];
            float p[3];
#pragma unroll
            for (int d = 0; d < 3; ++d)
                p[d] = (src < 16 ? sm.Gt[s][src][d] : sm.vp[s][3 * (nv_own + src - 16) + d]) - sm.centre[s][d];
            const float* q = sm.post[s];
            float* o = joints + ((size_t)(b0 + s) * 21 + jn) * 3;
            o[0] = q[0] * p[0] + q[1] * p[1] + q[2] * p[2] + q[9];
            o[1] = q[3] * p[0] + q[4] * p[1] + q[5] * p[2] + q[10];
            o[2] = q[6] * p[0] + q[7] * p[1] + q[8] * p[2] + q[11];
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
    for (int i = tid; i < kS * nv_own * 3; i += kThreads) {
        const int s = i / (nv_own * 3), c = i - s * (nv_own * 3);
        if (b0 + s < batch) verts[((size_t)(b0 + s) * kV + v0) * 3 + c] = sm.vp[s][c];
    }
    LBS_MARK(6);
}

int launch_mano(const ab_mano_model* model, int batch, const float* pose, const float* betas, const float* post_rt,
                int center_idx, float* verts, float* joints, float* transforms_abs, cudaStream_t st) {
    dim3 grid(cdiv(batch, kS), kSplit);
    static std::atomic<bool> opted[64] = {};  // per device: function attributes belong to the device's context
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    if (!opted[dev & 63].load(std::memory_order_relaxed)) {
        AB_CUDA(cudaFuncSetAttribute(mano_lbs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ManoSmem)));
        // four CTAs of 42 KB per SM need more than the default shared-memory carve-out
        AB_CUDA(cudaFuncSetAttribute(mano_lbs_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        opted[dev & 63].store(true, std::memory_order_relaxed);
    }
    StageTimer tm(AB_STAGE_MANO_LBS, st);
    mano_lbs_kernel<<<grid, kThreads, sizeof(ManoSmem), st>>>(*model, batch, pose, betas, post_rt, center_idx, verts, joints,
                                                             transforms_abs);
    count_launch();
    return check_launch("mano_lbs_kernel");
}

}  // namespace ab

#ifdef AB_LBS_TRACE
extern "C" __attribute__((visibility("default"))) int ab_debug_lbs_trace(unsigned long long* host, int n) {
    return (int)cudaMemcpyFromSymbol(host, ab::g_lbs_trace, sizeof(unsigned long long) * 8 * (size_t)n);
}
#endif

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
