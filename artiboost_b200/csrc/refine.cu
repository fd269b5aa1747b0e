// Hand-object refiner and anatomical scramblers (sm_100a): the pieces of anakin/artiboost/refiner.py:150-285
// (HORefiner + _RefineNet, the shipped config's REFINER.TYPE "hand_obj") and scrambler.py:84-260 (random_2 / random_3)
// that sit between the pose-generator prelude and the final LBS launch.
//
//   chamfer_nn_kernel       point2point_signed (refiner.py:21-85): nearest object point of every hand vertex.  The
//                           reference calls the third-party chamfer_distance CUDA kernel on 778 x 10 000 points and
//                           materialises the rotated object cloud per sample (refiner.py:190-191); here the rotation
//                           is applied while the object tile is staged into shared memory, every thread scans the
//                           tile for two hand vertices held in registers, and the BatchNorm1d(778) of
//                           _RefineNet.forward (:266) is folded into the store.
//   linear_f32_kernel       the RefineNet MLP (ResBlock, :288-319) in fp32 on the FMA pipe: y = act(x W^T + b (+ res)).
//                           fp32 on purpose: the 1e-4 bar on vertex positions leaves no room for bf16 / tf32 operands
//                           across three refinement iterations, and the MLP is ~5 % of the refiner's arithmetic.
//   refine_encode_kernel    aa -> first two rotation-matrix columns ("CRot" 6D), refiner.py:253-257
//   refine_decode_kernel    CRot2rotmat + rotmat_to_aa (parms_decode, :88-107) + the rigid map of the next LBS launch
//   scramble_axis_kernel    manotorch AxisLayer + RandomScrambler2/3.forward
#include <atomic>

#include "mano_math.cuh"

namespace ab {

// ---------------------------------------------------------------------------------------------- nearest neighbour
constexpr int kNnThreads = 416;  // 13 warps x 2 queries = 832 slots for the 778 hand vertices
constexpr int kNnQ = 2;
constexpr int kNnTile = 2048;    // object points per shared-memory tile (32 KB of float4)
constexpr int kNnChunk = 8;      // points per running-minimum update

// d = fma(dz,dz, fma(dy,dy, dx*dx)): the arithmetic oracle/refiner.py restates (first minimum wins, like the
// sequential strict `<` scan of the chamfer_distance kernel).
__device__ __forceinline__ float nn_d2(float qx, float qy, float qz, float px, float py, float pz) {
    float dx = qx - px, dy = qy - py, dz = qz - pz;
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

// Blackwell packed fp32 (FADD2 / FMUL2 / FFMA2): two IEEE fp32 lanes per instruction, each rounded exactly like the
// scalar form above, so the scan below costs ~4 issue slots per (vertex, point) pair instead of ~7.5.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// y = R o with every operation individually rounded (matches the oracle's fp32 order)
__device__ __forceinline__ void nn_rotate(const float* R, bool rot, const float* o, float& px, float& py, float& pz) {
    float ox = o[0], oy = o[1], oz = o[2];
    if (rot) {
        px = __fadd_rn(__fadd_rn(__fmul_rn(R[0], ox), __fmul_rn(R[1], oy)), __fmul_rn(R[2], oz));
        py = __fadd_rn(__fadd_rn(__fmul_rn(R[3], ox), __fmul_rn(R[4], oy)), __fmul_rn(R[5], oz));
        pz = __fadd_rn(__fadd_rn(__fmul_rn(R[6], ox), __fmul_rn(R[7], oy)), __fmul_rn(R[8], oz));
    } else {
        px = ox; py = oy; pz = oz;
    }
}

__global__ void __launch_bounds__(kNnThreads)
chamfer_nn_kernel(int n_x, const float* __restrict__ x, int n_y, const float* __restrict__ y_points,
                  const int32_t* __restrict__ obj_id, const float* __restrict__ rot, int rot_stride,
                  unsigned long long* __restrict__ keys) {
    __shared__ __align__(16) float sx[kNnTile], sy[kNnTile], sz[kNnTile];  // SoA: one LDS.128 = 4 points of one axis
    const int b = blockIdx.y, tid = threadIdx.x;
    const int q0 = blockIdx.x * (kNnThreads * kNnQ);
    const float* yb = y_points + (size_t)(obj_id ? obj_id[b] : b) * n_y * 3;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (rot) {
        const int rs = rot_stride == 16 ? 4 : 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) R[3 * i + j] = rot[(size_t)b * rot_stride + rs * i + j];
    }
    float qx[kNnQ], qy[kNnQ], qz[kNnQ], best[kNnQ];
    f32x2 qx2[kNnQ], qy2[kNnQ], qz2[kNnQ];
    int bchunk[kNnQ];
#pragma unroll
    for (int k = 0; k < kNnQ; ++k) {
        int qi = q0 + tid + k * kNnThreads;
        const float* p = x + ((size_t)b * n_x + (qi < n_x ? qi : 0)) * 3;
        qx[k] = p[0]; qy[k] = p[1]; qz[k] = p[2];
        qx2[k] = pack2(qx[k], qx[k]); qy2[k] = pack2(qy[k], qy[k]); qz2[k] = pack2(qz[k], qz[k]);
        best[k] = __int_as_float(0x7f800000);
        bchunk[k] = 0;
    }
    {   // one tile of the cloud per CTA (blockIdx.z): the tiles of a sample meet again in the atomicMin below
        const int t0 = blockIdx.z * kNnTile;
        const int tn = min(kNnTile, n_y - t0);
        const int tn_pad = (tn + kNnChunk - 1) / kNnChunk * kNnChunk;
        __syncthreads();
        for (int i = tid; i < tn_pad; i += kNnThreads) {
            float px = 1e18f, py = 1e18f, pz = 1e18f;  // padding of the last chunk: never the minimum
            if (i < tn) nn_rotate(R, rot != nullptr, yb + (size_t)(t0 + i) * 3, px, py, pz);
            sx[i] = px; sy[i] = py; sz[i] = pz;
        }
        __syncthreads();
        for (int c = 0; c < tn_pad; c += kNnChunk) {
            float m[kNnQ];
#pragma unroll
            for (int k = 0; k < kNnQ; ++k) m[k] = __int_as_float(0x7f800000);
#pragma unroll
            for (int h = 0; h < kNnChunk; h += 4) {
                const float4 X = *reinterpret_cast<const float4*>(&sx[c + h]);
                const float4 Y = *reinterpret_cast<const float4*>(&sy[c + h]);
                const float4 Z = *reinterpret_cast<const float4*>(&sz[c + h]);
                const f32x2 x01 = pack2(X.x, X.y), x23 = pack2(X.z, X.w);
                const f32x2 y01 = pack2(Y.x, Y.y), y23 = pack2(Y.z, Y.w);
                const f32x2 z01 = pack2(Z.x, Z.y), z23 = pack2(Z.z, Z.w);
#pragma unroll
                for (int k = 0; k < kNnQ; ++k) {
                    f32x2 dx = sub2(qx2[k], x01), dy = sub2(qy2[k], y01), dz = sub2(qz2[k], z01);
                    f32x2 d01 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    dx = sub2(qx2[k], x23); dy = sub2(qy2[k], y23); dz = sub2(qz2[k], z23);
                    f32x2 d23 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
                    float d0, d1, d2, d3;
                    unpack2(d01, d0, d1);
                    unpack2(d23, d2, d3);
                    m[k] = fminf(fminf(m[k], fminf(d0, d1)), fminf(d2, d3));
                }
            }
#pragma unroll
            for (int k = 0; k < kNnQ; ++k)
                if (m[k] < best[k]) { best[k] = m[k]; bchunk[k] = t0 + c; }
        }
    }
    // the winning chunk is re-scanned from the source points for the first index that attains the minimum
#pragma unroll
    for (int k = 0; k < kNnQ; ++k) {
        int qi = q0 + tid + k * kNnThreads;
        if (qi >= n_x) continue;
        int bi = bchunk[k];
        for (int j = 0; j < kNnChunk; ++j) {
            int i = bchunk[k] + j;
            if (i >= n_y) break;
            float px, py, pz;
            nn_rotate(R, rot != nullptr, yb + (size_t)i * 3, px, py, pz);
            if (nn_d2(qx[k], qy[k], qz[k], px, py, pz) == best[k]) { bi = i; break; }
        }
        // d >= 0, so the float bits order like the values; the index in the low word makes the smallest index win ties
        atomicMin(keys + (size_t)b * n_x + qi, ((unsigned long long)__float_as_uint(best[k]) << 32) | (unsigned)bi);
    }
}

__global__ void chamfer_finalize_kernel(int total, int n_x, unsigned long long* __restrict__ keys,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        float* __restrict__ dist, long long dist_stride, int32_t* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = i / n_x, qi = i - b * n_x;
    const unsigned long long k = keys[i];
    float v = sqrtf(__uint_as_float((unsigned)(k >> 32)));
    if (scale) v = __fadd_rn(__fmul_rn(v, scale[qi]), shift[qi]);
    dist[(size_t)b * dist_stride + qi] = v;
    if (idx) idx[i] = (int32_t)(k & 0xffffffffull);
}

// ------------------------------------------------------------------- nearest neighbour over a static, grouped cloud
// The refiner's clouds are static per object (HORefiner.setup resamples them once), so the host sorts each cloud along a
// Morton curve into groups of 32 points with an axis-aligned box per group (object frame).  One CTA per sample:
//   phase 1  the whole sorted cloud is rotated into shared memory (y = R o, rounded per operation like the scan above),
//            SoA with a 36-float group stride; the boxes, and boxes of 8 consecutive groups, go beside it;
//   phase 2  one THREAD per hand vertex: lower bounds to the 8-group boxes, then to the groups inside the ones that can
//            matter; a seed group fixes a first minimum, every group whose bound does not exceed it is evaluated with the
//            packed-fp32 arithmetic of the scan (d = fma(dz,dz,fma(dy,dy,dx*dx)), bit-identical), typically 10-15 of 313;
//   phase 3  the group that holds the minimum is re-read for the smallest ORIGINAL index attaining it (ties between
//            groups -- exact duplicates -- are flagged and resolved by a second sweep), so indices match the scan's
//            first-minimum rule.
// The bounds are computed in the object frame (R^T q) with a relative + absolute slack that covers their own rounding;
// `rot` must be a rotation.  ~30x fewer pair evaluations than the scan, ~4x fewer instructions.
constexpr int kGrpThreads = 832;   // 26 warps: the 778 hand vertices of a sample
constexpr int kGrpStride = 36;     // floats per group and axis in shared memory (32 + 4: spreads the groups over banks)
constexpr int kGrpSuper = 8;       // groups per second-level box

__device__ __forceinline__ float box_lb(const float4& lo, const float2& hi, float ox, float oy, float oz) {
    const float dx = fmaxf(fmaxf(lo.x - ox, ox - lo.w), 0.f);   // lo = (lo.x, lo.y, lo.z, hi.x), hi = (hi.y, hi.z)
    const float dy = fmaxf(fmaxf(lo.y - oy, oy - hi.x), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - oz, oz - hi.y), 0.f);
    return dx * dx + dy * dy + dz * dz;
}

__device__ __forceinline__ float group_min(const float* gx, const float* gy, const float* gz, f32x2 qx2, f32x2 qy2, f32x2 qz2) {
    float m = __int_as_float(0x7f800000);
#pragma unroll
    for (int h = 0; h < 32; h += 4) {
        const float4 X = *reinterpret_cast<const float4*>(gx + h);
        const float4 Y = *reinterpret_cast<const float4*>(gy + h);
        const float4 Z = *reinterpret_cast<const float4*>(gz + h);
        f32x2 dx = sub2(qx2, pack2(X.x, X.y)), dy = sub2(qy2, pack2(Y.x, Y.y)), dz = sub2(qz2, pack2(Z.x, Z.y));
        const f32x2 d01 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
        dx = sub2(qx2, pack2(X.z, X.w)); dy = sub2(qy2, pack2(Y.z, Y.w)); dz = sub2(qz2, pack2(Z.z, Z.w));
        const f32x2 d23 = fma2(dz, dz, fma2(dy, dy, mul2(dx, dx)));
        float d0, d1, d2, d3;
        unpack2(d01, d0, d1);
        unpack2(d23, d2, d3);
        m = fminf(fminf(m, fminf(d0, d1)), fminf(d2, d3));
    }
    return m;
}

__global__ void __launch_bounds__(kGrpThreads)
chamfer_nn_grouped_kernel(int n_x, const float* __restrict__ x, int n_groups, const float* __restrict__ sorted_pts,
                          const int32_t* __restrict__ perm, const float* __restrict__ boxes,
                          const int32_t* __restrict__ obj_id, const float* __restrict__ rot, int rot_stride,
                          const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ dist,
                          long long dist_stride, int32_t* __restrict__ idx) {
    extern __shared__ float4 grp_smem[];
    const int NG = n_groups, NS = (n_groups + kGrpSuper - 1) / kGrpSuper;
    float* sx = reinterpret_cast<float*>(grp_smem);
    float* sy = sx + (size_t)NG * kGrpStride;
    float* sz = sy + (size_t)NG * kGrpStride;
    float4* blo = reinterpret_cast<float4*>(sz + (size_t)NG * kGrpStride);
    float4* slo = blo + NG;
    float2* bhi = reinterpret_cast<float2*>(slo + NS);
    float2* shi = bhi + NG;
    const int b = blockIdx.y, tid = threadIdx.x;
    const int obj = obj_id ? obj_id[b] : b;
    const int n_pts = NG * 32;
    const float* pts = sorted_pts + (size_t)obj * n_pts * 3;
    const int32_t* pm = perm + (size_t)obj * n_pts;
    const float* bx = boxes + (size_t)obj * 6 * NG;  // [6][NG]: lo.xyz, hi.xyz
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (rot) {
        const int rs = rot_stride == 16 ? 4 : 3;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) R[3 * i + j] = rot[(size_t)b * rot_stride + rs * i + j];
    }
    // ---- phase 1
    for (int j = tid; j < n_pts; j += kGrpThreads) {
        float px, py, pz;
        nn_rotate(R, rot != nullptr, pts + (size_t)j * 3, px, py, pz);
        const int o = (j >> 5) * kGrpStride + (j & 31);
        sx[o] = px; sy[o] = py; sz[o] = pz;
    }
    for (int g = tid; g < NG; g += kGrpThreads) {
        blo[g] = make_float4(bx[g], bx[NG + g], bx[2 * NG + g], bx[3 * NG + g]);
        bhi[g] = make_float2(bx[4 * NG + g], bx[5 * NG + g]);
    }
    __syncthreads();
    for (int s = tid; s < NS; s += kGrpThreads) {
        float4 lo = blo[s * kGrpSuper];
        float2 hi = bhi[s * kGrpSuper];
        for (int g = s * kGrpSuper + 1; g < min((s + 1) * kGrpSuper, NG); ++g) {
            const float4 l = blo[g];
            const float2 h = bhi[g];
            lo.x = fminf(lo.x, l.x); lo.y = fminf(lo.y, l.y); lo.z = fminf(lo.z, l.z); lo.w = fmaxf(lo.w, l.w);
            hi.x = fmaxf(hi.x, h.x); hi.y = fmaxf(hi.y, h.y);
        }
        slo[s] = lo;
        shi[s] = hi;
    }
    __syncthreads();
    // ---- phase 2
    // Vertices are dealt to threads sorted by their nearest 8-group box (a counting sort over <= 60 bins), so the threads
    // of a warp open mostly the same groups: MANO's vertex order alone leaves a warp with a handful of useful lanes per
    // evaluated group.
    __shared__ int bin_cnt[64], bin_start[64];
    __shared__ short order[kGrpThreads], order_ss[kGrpThreads];
    if (tid < 64) bin_cnt[tid] = 0;
    __syncthreads();
    int key = 63;  // threads past n_x sort to the end (NS <= 60)
    {
        const int q0 = blockIdx.x * kGrpThreads + tid;
        if (q0 < n_x) {
            const float* qp0 = x + ((size_t)b * n_x + q0) * 3;
            const float ax = qp0[0], ay = qp0[1], az = qp0[2];
            const float ux = R[0] * ax + R[3] * ay + R[6] * az, uy = R[1] * ax + R[4] * ay + R[7] * az,
                        uz = R[2] * ax + R[5] * ay + R[8] * az;
            float ms0 = __int_as_float(0x7f800000);
            key = 0;
            for (int s = 0; s < NS; ++s) {
                const float l = box_lb(slo[s], shi[s], ux, uy, uz);
                if (l < ms0) { ms0 = l; key = s; }
            }
        }
    }
    const int rank = atomicAdd(&bin_cnt[key], 1);
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int k = 0; k < 64; ++k) { bin_start[k] = run; run += bin_cnt[k]; }
    }
    __syncthreads();
    order[bin_start[key] + rank] = (short)tid;
    order_ss[bin_start[key] + rank] = (short)key;
    __syncthreads();
    const int qi = blockIdx.x * kGrpThreads + order[tid];
    if (qi >= n_x) return;
    const int ss = order_ss[tid];
    const float* qp = x + ((size_t)b * n_x + qi) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    const f32x2 qx2 = pack2(qx, qx), qy2 = pack2(qy, qy), qz2 = pack2(qz, qz);
    // the vertex in the object frame (R^T q), for the bounds only
    const float ox = R[0] * qx + R[3] * qy + R[6] * qz, oy = R[1] * qx + R[4] * qy + R[7] * qz,
                oz = R[2] * qx + R[5] * qy + R[8] * qz;
    float mg = __int_as_float(0x7f800000);
    int seed = ss * kGrpSuper;
    for (int g = ss * kGrpSuper; g < min((ss + 1) * kGrpSuper, NG); ++g) {
        const float l = box_lb(blo[g], bhi[g], ox, oy, oz);
        if (l < mg) { mg = l; seed = g; }
    }
    float best = group_min(sx + seed * kGrpStride, sy + seed * kGrpStride, sz + seed * kGrpStride, qx2, qy2, qz2);
    int best_g = seed;
    bool tie = false;
    float bound = best * 1.0001f + 1e-9f;  // slack over the rounding of the object-frame bounds
    for (int s = 0; s < NS; ++s) {
        if (!(box_lb(slo[s], shi[s], ox, oy, oz) <= bound)) continue;
        for (int g = s * kGrpSuper; g < min((s + 1) * kGrpSuper, NG); ++g) {
            if (g == seed || !(box_lb(blo[g], bhi[g], ox, oy, oz) <= bound)) continue;
            const float m = group_min(sx + g * kGrpStride, sy + g * kGrpStride, sz + g * kGrpStride, qx2, qy2, qz2);
            if (m < best) {
                best = m; best_g = g; tie = false;
                bound = best * 1.0001f + 1e-9f;
            } else if (m == best) {
                tie = true;  // an exact duplicate of the nearest point in another group
            }
        }
    }
    // ---- phase 3: smallest original index among the points that attain the minimum
    int bi = 0x7fffffff;
    for (int j = 0; j < 32; ++j) {
        const int o = best_g * kGrpStride + j;
        if (nn_d2(qx, qy, qz, sx[o], sy[o], sz[o]) == best) bi = min(bi, pm[best_g * 32 + j]);
    }
    if (tie) {
        for (int g = 0; g < NG; ++g) {
            if (g == best_g || !(box_lb(blo[g], bhi[g], ox, oy, oz) <= bound)) continue;
            for (int j = 0; j < 32; ++j) {
                const int o = g * kGrpStride + j;
                if (nn_d2(qx, qy, qz, sx[o], sy[o], sz[o]) == best) bi = min(bi, pm[g * 32 + j]);
            }
        }
    }
    float v = sqrtf(best);
    if (scale) v = __fadd_rn(__fmul_rn(v, scale[qi]), shift[qi]);
    dist[(size_t)b * dist_stride + qi] = v;
    if (idx) idx[(size_t)b * n_x + qi] = bi;
}

// ------------------------------------------------------------------------------------------------------ fp32 linear
// 32 x 64 output tile per CTA, 4 x 4 outputs per thread; the K loop is shared out over kLinKG groups of 128 threads
// (k-tiles g, g + KG, ...), whose partial tiles are added in a fixed order through shared memory: the MLP's GEMMs are
// small (M = batch, N <= 768), so the K split is what puts >= 16 warps on every SM, and the fixed order keeps the
// result bit-reproducible.
constexpr int kLinBM = 32, kLinBN = 64, kLinBK = 16, kLinKG = 4, kLinThreads = 128 * kLinKG;

__global__ void __launch_bounds__(kLinThreads)
linear_f32_kernel(int M, int N, int K, const float* __restrict__ x, long long ldx, const float* __restrict__ W,
                  long long ldw, const float* __restrict__ bias, const float* res, long long ldr, int act, float slope,
                  float* y, long long ldy) {
    // two stages of operand tiles (dynamic shared memory, 48 KB); after the K loop stage 0 carries the partial tiles of
    // groups 1.. ([group][element][thread])
    extern __shared__ __align__(16) float lin_smem[];
    constexpr int kStage = kLinKG * kLinBK * (kLinBM + kLinBN);
    static_assert((kLinKG - 1) * 16 * 128 <= kStage, "partial tiles must fit one stage of operand storage");
    float (*red)[16][128] = reinterpret_cast<float (*)[16][128]>(lin_smem);
    const int kg = threadIdx.x >> 7, tid = threadIdx.x & 127, tx = tid & 15, ty = tid >> 4;  // 16 x 8 threads per group
    const int m0 = blockIdx.y * kLinBM, n0 = blockIdx.x * kLinBN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // loaders: one row per thread, a contiguous run of k per thread, so the transposed shared-memory writes are
    // conflict-free (consecutive threads -> consecutive rows)
    const int ar = tid & 31, ak = (tid >> 5) * 4;   // A: 32 rows x 16 k, 4 k per thread
    const int br = tid & 63, bk = (tid >> 6) * 8;   // B: 64 rows x 16 k, 8 k per thread
    const bool a_ok = m0 + ar < M, b_ok = n0 + br < N;
    const float* ap = x + (size_t)(a_ok ? m0 + ar : 0) * ldx;
    const float* bp = W + (size_t)(b_ok ? n0 + br : 0) * ldw;
    const int n_tiles = (K + kLinBK - 1) / kLinBK;
    const int n_rounds = (n_tiles + kLinKG - 1) / kLinKG;  // same trip count for every group: the barriers are CTA-wide
    // the next round's operands are fetched into registers while the current round is multiplied out of shared memory
    float ra[4], rb[8];
    auto fetch = [&](int r) {
        const int k0 = (r * kLinKG + kg) * kLinBK;  // may lie past K for the last round: zero tiles
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + ak + i;
            ra[i] = (a_ok && k < K) ? ap[k] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + bk + i;
            rb[i] = (b_ok && k < K) ? bp[k] : 0.f;
        }
    };
    auto stage_a = [&](int st) { return reinterpret_cast<float (*)[kLinBK][kLinBM]>(lin_smem + st * kStage); };
    auto stage_b = [&](int st) {
        return reinterpret_cast<float (*)[kLinBK][kLinBN]>(lin_smem + st * kStage + kLinKG * kLinBK * kLinBM);
    };
    auto stash = [&](int st) {
        float (*As)[kLinBK][kLinBM] = stage_a(st);
        float (*Bs)[kLinBK][kLinBN] = stage_b(st);
#pragma unroll
        for (int i = 0; i < 4; ++i) As[kg][ak + i][ar] = ra[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[kg][bk + i][br] = rb[i];
    };
    fetch(0);
    stash(0);
    __syncthreads();
    for (int r = 0; r < n_rounds; ++r) {   // one barrier per round: round r + 1 is staged while round r is multiplied
        if (r + 1 < n_rounds) fetch(r + 1);
        float (*As)[kLinBK][kLinBM] = stage_a(r & 1);
        float (*Bs)[kLinBK][kLinBN] = stage_b(r & 1);
#pragma unroll
        for (int kk = 0; kk < kLinBK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kg][kk][ty * 4]);
            float4 w = *reinterpret_cast<const float4*>(&Bs[kg][kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        if (r + 1 < n_rounds) stash((r + 1) & 1);
        __syncthreads();
    }
    if (kg > 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) red[kg - 1][4 * i + j][tid] = acc[i][j];
    }
    __syncthreads();
    if (kg > 0) return;
#pragma unroll
    for (int g = 0; g < kLinKG - 1; ++g)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += red[g][4 * i + j][tid];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[n];
            if (res) v += res[(size_t)m * ldr + n];
            if (act == 1) v = v > 0.f ? v : v * slope;
            y[(size_t)m * ldy + n] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- 6D encode / decode
__global__ void refine_encode_kernel(int batch, const float* __restrict__ pose, const float* __restrict__ tsl,
                                     float* __restrict__ feat, long long ld) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * 16) return;
    int b = t >> 4, k = t & 15;
    const float* a = pose + (size_t)b * 48 + 3 * k;
    float R[9];
    rodrigues(a[0], a[1], a[2], R);
    float* o = feat + (size_t)b * ld + 6 * k;  // rotmat[..., :2].reshape(bs, -1): rows x the first two columns
    o[0] = R[0]; o[1] = R[1]; o[2] = R[3]; o[3] = R[4]; o[4] = R[6]; o[5] = R[7];
    if (k == 0) {
        float* tt = feat + (size_t)b * ld + 96;
        tt[0] = tsl[(size_t)b * 3]; tt[1] = tsl[(size_t)b * 3 + 1]; tt[2] = tsl[(size_t)b * 3 + 2];
    }
}

__global__ void refine_decode_kernel(int batch, const float* __restrict__ feat, long long ld,
                                     const float* __restrict__ rigid, int rigid_stride, const float* __restrict__ offset,
                                     float* __restrict__ pose_out, float* __restrict__ tsl_out,
                                     float* __restrict__ post_rt) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * 16) return;
    int b = t >> 4, k = t & 15;
    const float* c = feat + (size_t)b * ld + 6 * k;
    // CRot2rotmat (refiner.py:88-98): view(-1,3,2) -> columns (c0,c2,c4) and (c1,c3,c5); F.normalize eps 1e-12
    float a1[3] = {c[0], c[2], c[4]}, a2[3] = {c[1], c[3], c[5]};
    float n1 = fmaxf(sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]), 1e-12f);
    float b1[3] = {a1[0] / n1, a1[1] / n1, a1[2] / n1};
    float dot = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    float u[3] = {a2[0] - dot * b1[0], a2[1] - dot * b1[1], a2[2] - dot * b1[2]};
    float n2 = fmaxf(sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), 1e-12f);
    float b2[3] = {u[0] / n2, u[1] / n2, u[2] / n2};
    float b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
    float R[9] = {b1[0], b2[0], b3[0], b1[1], b2[1], b3[1], b1[2], b2[2], b3[2]};
    float aa[3];
    rotmat_to_aa(R, aa);
    float* po = pose_out + (size_t)b * 48 + 3 * k;
    po[0] = aa[0]; po[1] = aa[1]; po[2] = aa[2];
    if (k == 0) {
        const float* tt = feat + (size_t)b * ld + 96;
        float tv[3] = {tt[0], tt[1], tt[2]};
        if (tsl_out) { tsl_out[(size_t)b * 3] = tv[0]; tsl_out[(size_t)b * 3 + 1] = tv[1]; tsl_out[(size_t)b * 3 + 2] = tv[2]; }
        if (post_rt) {
            float* q = post_rt + (size_t)b * 12;
            if (rigid) {  // x' = Rf (x + t + offset)   (preprocessor.py:84-88)
                const float* F = rigid + (size_t)b * rigid_stride;
                int rs = rigid_stride == 16 ? 4 : 3;
                float sh[3] = {tv[0], tv[1], tv[2]};
                if (offset) { sh[0] += offset[(size_t)b * 3]; sh[1] += offset[(size_t)b * 3 + 1]; sh[2] += offset[(size_t)b * 3 + 2]; }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    q[3 * i] = F[rs * i]; q[3 * i + 1] = F[rs * i + 1]; q[3 * i + 2] = F[rs * i + 2];
                    q[9 + i] = F[rs * i] * sh[0] + F[rs * i + 1] * sh[1] + F[rs * i + 2] * sh[2];
                }
            } else {
                q[0] = 1.f; q[1] = 0.f; q[2] = 0.f; q[3] = 0.f; q[4] = 1.f; q[5] = 0.f; q[6] = 0.f; q[7] = 0.f; q[8] = 1.f;
                q[9] = tv[0]; q[10] = tv[1]; q[11] = tv[2];
            }
        }
    }
}

// ----------------------------------------------------------------------------------------- anatomical scramblers
// manotorch AxisLayer: back / up / left axes of the 15 articulated joints in MANO chain order, expressed in each
// joint's own frame.  21-keypoint ids of the joint and its child, chain order (index, middle, little, ring, thumb).
__device__ __constant__ const int kAxisJoint[15] = {5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3};

__device__ __forceinline__ void aa_compose(const float* a1, const float* a2, float* out) {  // axis_angle_op, scrambler.py:19-28
    float R1[9], R2[9], Rc[9];
    rodrigues(a1[0], a1[1], a1[2], R1);
    rodrigues(a2[0], a2[1], a2[2], R2);
    mat3_mul(R1, R2, Rc);
    rotmat_to_aa(Rc, out);
}

__device__ __forceinline__ void normalize3(float* v) {
    float n = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] /= n; v[1] /= n; v[2] /= n;
}

// one thread per sample: axes (15 small products), then 4 splay + 14 bend + 2 thumb compositions in the reference's
// order.  bend[B,14] holds the per-joint bend angles of chain joints (1..12, 14, 15) -- for random_2 the host expands
// its 5 per-finger draws with the interlink coefficients (scrambler.py:134-170), for random_3 they are the 14 draws.
__global__ void scramble_axis_kernel(int batch, const float* __restrict__ pose, const float* __restrict__ joints,
                                     const float* __restrict__ transf, const float* __restrict__ splay,
                                     const float* __restrict__ bend, const float* __restrict__ thumb,
                                     float* __restrict__ pose_out, float* __restrict__ axes_out) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const float* J = joints + (size_t)b * 63;
    const float* T = transf + (size_t)b * 256;
    float hp[48];
#pragma unroll 1
    for (int i = 0; i < 48; ++i) hp[i] = pose[(size_t)b * 48 + i];
    const int splay_joint[4] = {1, 4, 7, 10};
#pragma unroll 1
    for (int i = 0; i < 15; ++i) {
        int j = kAxisJoint[i];
        float d[3] = {J[3 * j] - J[3 * j + 3], J[3 * j + 1] - J[3 * j + 4], J[3 * j + 2] - J[3 * j + 5]};
        const float* G = T + 16 * (i + 1);  // transforms_abs[1 + i], rotation block transposed
        float bax[3] = {G[0] * d[0] + G[4] * d[1] + G[8] * d[2], G[1] * d[0] + G[5] * d[1] + G[9] * d[2],
                        G[2] * d[0] + G[6] * d[1] + G[10] * d[2]};
        float up[3] = {i < 12 ? 0.f : 1.f, 1.f, i < 12 ? 0.f : 1.f};
        float l[3] = {bax[1] * up[2] - bax[2] * up[1], bax[2] * up[0] - bax[0] * up[2], bax[0] * up[1] - bax[1] * up[0]};
        float u[3] = {l[1] * bax[2] - l[2] * bax[1], l[2] * bax[0] - l[0] * bax[2], l[0] * bax[1] - l[1] * bax[0]};
        normalize3(bax); normalize3(u); normalize3(l);
        if (axes_out) {
            float* ao = axes_out + ((size_t)b * 15 + i) * 9;
            ao[0] = bax[0]; ao[1] = bax[1]; ao[2] = bax[2]; ao[3] = u[0]; ao[4] = u[1]; ao[5] = u[2];
            ao[6] = l[0]; ao[7] = l[1]; ao[8] = l[2];
        }
        const int jc = i + 1;  // chain joint this axis belongs to
        float* h = hp + 3 * jc;
        if (jc == 13) {  // thumb base: bend about l, then splay about u, both pre-multiplied (scrambler.py:172-181)
            float t0 = thumb[(size_t)b * 2], t1 = thumb[(size_t)b * 2 + 1];
            float ab_[3] = {l[0] * t0, l[1] * t0, l[2] * t0}, as_[3] = {u[0] * t1, u[1] * t1, u[2] * t1}, tmp[3];
            aa_compose(ab_, h, tmp);
            aa_compose(as_, tmp, h);
            continue;
        }
        if (jc == splay_joint[0] || jc == splay_joint[1] || jc == splay_joint[2] || jc == splay_joint[3]) {
            float s = splay[(size_t)b * 4 + (jc - 1) / 3];  // post-multiplied: R(pose) R(splay)  (scrambler.py:124-128)
            float as_[3] = {u[0] * s, u[1] * s, u[2] * s}, tmp[3];
            aa_compose(h, as_, tmp);
            h[0] = tmp[0]; h[1] = tmp[1]; h[2] = tmp[2];
        }
        float g = bend[(size_t)b * 14 + (jc < 13 ? jc - 1 : jc - 2)];  // pre-multiplied: R(bend) R(pose)
        float ab_[3] = {l[0] * g, l[1] * g, l[2] * g}, tmp[3];
        aa_compose(ab_, h, tmp);
        h[0] = tmp[0]; h[1] = tmp[1]; h[2] = tmp[2];
    }
#pragma unroll 1
    for (int i = 0; i < 48; ++i) pose_out[(size_t)b * 48 + i] = hp[i];
}

}  // namespace ab

// ------------------------------------------------------------------------------------------------------- C ABI
extern "C" uint64_t ab_chamfer_nn_workspace_bytes(int batch, int n_x) {
    return batch > 0 && n_x > 0 ? (uint64_t)batch * n_x * 8 : 0;
}

extern "C" int ab_chamfer_nn(int batch, int n_x, const float* x, int n_y, const float* y_points, const int32_t* obj_id,
                             const float* rot, int rot_stride, const float* scale, const float* shift, float* dist,
                             int64_t dist_stride, int32_t* idx, void* ws, void* stream) {
    AB_REQUIRE(batch >= 0 && n_x >= 0 && n_y >= 0, "negative size");
    if (batch == 0 || n_x == 0) return AB_OK;
    AB_REQUIRE(n_y > 0, "empty target cloud: the nearest neighbour is undefined");
    AB_REQUIRE(batch <= 65535, "batch > 65535: split the call");
    AB_REQUIRE(x && y_points && dist && ws, "null pointer");
    AB_REQUIRE(((uintptr_t)ws & 7) == 0, "workspace must be 8-byte aligned");
    AB_REQUIRE(rot == nullptr || rot_stride == 9 || rot_stride == 16, "rot_stride must be 9 (3x3) or 16 (4x4 pose)");
    AB_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
    AB_REQUIRE(dist_stride >= n_x, "dist_stride < n_x");
    cudaStream_t st = (cudaStream_t)stream;
    const int splits = ab::cdiv(n_y, ab::kNnTile);
    AB_REQUIRE(splits <= 65535, "n_y too large");
    dim3 grid(ab::cdiv(n_x, ab::kNnThreads * ab::kNnQ), batch, splits);
    unsigned long long* keys = (unsigned long long*)ws;
    AB_CUDA(cudaMemsetAsync(keys, 0xFF, (size_t)batch * n_x * 8, st));
    {
        ab::StageTimer tm(AB_STAGE_CHAMFER, st);
        ab::chamfer_nn_kernel<<<grid, ab::kNnThreads, 0, st>>>(n_x, x, n_y, y_points, obj_id, rot, rot_stride, keys);
    }
    {
        ab::StageTimer tm(AB_STAGE_REFINE_MISC, st);
        const int total = batch * n_x;
        ab::chamfer_finalize_kernel<<<ab::cdiv(total, 256), 256, 0, st>>>(total, n_x, keys, scale, shift, dist,
                                                                         (long long)dist_stride, idx);
    }
    ab::count_launch(2);
    return ab::check_launch("chamfer_nn_kernel");
}

extern "C" int ab_chamfer_nn_grouped(int batch, int n_x, const float* x, int n_groups, const float* sorted_points,
                                     const int32_t* perm, const float* boxes, const int32_t* obj_id, const float* rot,
                                     int rot_stride, const float* scale, const float* shift, float* dist,
                                     int64_t dist_stride, int32_t* idx, void* stream) {
    AB_REQUIRE(batch >= 0 && n_x >= 0, "negative size");
    if (batch == 0 || n_x == 0) return AB_OK;
    AB_REQUIRE(n_groups > 0 && n_groups <= 480, "n_groups must be in 1..480 (the rotated cloud lives in shared memory)");
    AB_REQUIRE(batch <= 65535, "batch > 65535: split the call");
    AB_REQUIRE(x && sorted_points && perm && boxes && dist, "null pointer");
    AB_REQUIRE(rot == nullptr || rot_stride == 9 || rot_stride == 16, "rot_stride must be 9 (3x3) or 16 (4x4 pose)");
    AB_REQUIRE((scale == nullptr) == (shift == nullptr), "scale and shift go together");
    AB_REQUIRE(dist_stride >= n_x, "dist_stride < n_x");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_super = ab::cdiv(n_groups, ab::kGrpSuper);
    const size_t smem = (size_t)3 * n_groups * ab::kGrpStride * 4 + (size_t)(n_groups + n_super) * (16 + 8);
    static std::atomic<size_t> smem_set[64] = {};  // opt in to > 48 KB of dynamic shared memory (per device, once per size increase)
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    if (smem > smem_set[dev & 63].load(std::memory_order_relaxed)) {
        AB_CUDA(cudaFuncSetAttribute(ab::chamfer_nn_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev & 63].store(smem, std::memory_order_relaxed);
    }
    dim3 grid(ab::cdiv(n_x, ab::kGrpThreads), batch);
    {
        ab::StageTimer tm(AB_STAGE_CHAMFER, st);
        ab::chamfer_nn_grouped_kernel<<<grid, ab::kGrpThreads, smem, st>>>(n_x, x, n_groups, sorted_points, perm, boxes, obj_id, rot,
                                                                          rot_stride, scale, shift, dist, (long long)dist_stride, idx);
    }
    ab::count_launch();
    return ab::check_launch("chamfer_nn_grouped_kernel");
}

extern "C" int ab_linear_f32(int M, int N, int K, const float* x, int64_t ldx, const float* W, int64_t ldw,
                             const float* bias, const float* residual, int64_t ldr, int act, float slope, float* y,
                             int64_t ldy, void* stream) {
    AB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
    if (M == 0 || N == 0) return AB_OK;
    AB_REQUIRE(x && W && y, "null pointer");
    AB_REQUIRE(ldx >= K && ldw >= K && ldy >= N && (!residual || ldr >= N), "leading dimension too small");
    AB_REQUIRE(act == 0 || act == 1, "act: 0 none, 1 leaky relu");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(ab::cdiv(N, ab::kLinBN), ab::cdiv(M, ab::kLinBM));
    AB_REQUIRE(grid.y <= 65535, "M too large: split the call");
    {
        ab::StageTimer tm(AB_STAGE_LINEAR_F32, st);
        constexpr size_t smem = 2 * ab::kLinKG * ab::kLinBK * (ab::kLinBM + ab::kLinBN) * sizeof(float);  // 48 KB
        static std::atomic<bool> opted[64] = {};  // per device: function attributes belong to the device's context
        int dev = 0;
        AB_CUDA(cudaGetDevice(&dev));
        if (!opted[dev & 63].load(std::memory_order_relaxed)) {
            AB_CUDA(cudaFuncSetAttribute(ab::linear_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            opted[dev & 63].store(true, std::memory_order_relaxed);
        }
        ab::linear_f32_kernel<<<grid, ab::kLinThreads, smem, st>>>(M, N, K, x, (long long)ldx, W, (long long)ldw, bias,
                                                                  residual, (long long)ldr, act, slope, y, (long long)ldy);
    }
    ab::count_launch();
    return ab::check_launch("linear_f32_kernel");
}

extern "C" int ab_refine_encode(int batch, const float* pose, const float* tsl, float* feat, int64_t ld, void* stream) {
    AB_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(pose && tsl && feat && ld >= 99, "null pointer / ld < 99");
    cudaStream_t st = (cudaStream_t)stream;
    {
        ab::StageTimer tm(AB_STAGE_REFINE_MISC, st);
        ab::refine_encode_kernel<<<ab::cdiv(batch * 16, 128), 128, 0, st>>>(batch, pose, tsl, feat, (long long)ld);
    }
    ab::count_launch();
    return ab::check_launch("refine_encode_kernel");
}

extern "C" int ab_refine_decode(int batch, const float* feat, int64_t ld, const float* rigid, int rigid_stride,
                                const float* offset, float* pose_out, float* tsl_out, float* post_rt, void* stream) {
    AB_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(feat && pose_out && ld >= 99, "null pointer / ld < 99");
    AB_REQUIRE(rigid == nullptr || rigid_stride == 9 || rigid_stride == 16, "rigid_stride must be 9 or 16");
    AB_REQUIRE(rigid != nullptr || offset == nullptr, "offset without rigid");
    cudaStream_t st = (cudaStream_t)stream;
    {
        ab::StageTimer tm(AB_STAGE_REFINE_MISC, st);
        ab::refine_decode_kernel<<<ab::cdiv(batch * 16, 128), 128, 0, st>>>(batch, feat, (long long)ld, rigid, rigid_stride,
                                                                          offset, pose_out, tsl_out, post_rt);
    }
    ab::count_launch();
    return ab::check_launch("refine_decode_kernel");
}

extern "C" int ab_scramble_anatomical(int batch, const float* pose, const float* joints, const float* transforms_abs,
                                      const float* splay, const float* bend, const float* thumb, float* pose_out,
                                      float* axes_out, void* stream) {
    AB_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(pose && joints && transforms_abs && splay && bend && thumb && pose_out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    {
        ab::StageTimer tm(AB_STAGE_REFINE_MISC, st);
        ab::scramble_axis_kernel<<<ab::cdiv(batch, 64), 64, 0, st>>>(batch, pose, joints, transforms_abs, splay, bend,
                                                                    thumb, pose_out, axes_out);
    }
    ab::count_launch();
    return ab::check_launch("scramble_axis_kernel");
}
