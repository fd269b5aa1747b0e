// Small fp32 rotation / kinematic-chain helpers shared by the MANO LBS kernel and the pose-generator prelude.
#pragma once
#include "common.cuh"

namespace ab {

// MANO kinematic tree (anakin/postprocess/iknet/manolayer.py:210-249 level lists, flattened to parents).
__device__ __constant__ const int kManoParents[16] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};
// 21-keypoint order = [16 chain joints, 5 tips] permuted (iknet/manolayer.py:270).
__device__ __constant__ const int kJointReorder[21] = {0, 13, 14, 15, 16, 1, 2, 3, 17, 4, 5,
                                                        6, 18, 10, 11, 12, 19, 7, 8, 9, 20};
// fingertip vertex ids appended as joints 16..20 (iknet/manolayer.py:266).
__device__ __constant__ const int kTipVerts[5] = {745, 317, 444, 556, 673};
// transforms_abs order of manotorch's MANOOutput: the MANO chain order (wrist, index, middle, little, ring, thumb).
// manotorch is absent; the order is fixed by its consumers in the reference: AxisLayer pairs transforms_abs[:, 1:]
// with the chain-ordered keypoint list [5,6,7, 9,10,11, 17,18,19, 13,14,15, 1,2,3], and scrambler.py:124-181 indexes
// the resulting axes with chain-order pose indices (1,4,7,10 knuckles; 13 thumb base).
__device__ __constant__ const int kTransfReorder[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};

// Axis-angle -> rotation matrix (row-major 3x3), exact Rodrigues form.
__device__ __forceinline__ void rodrigues(float ax, float ay, float az, float* R) {
    float t2 = ax * ax + ay * ay + az * az;
    float a, b;  // R = I + a*K + b*K^2 with K = [aa]x (un-normalised)
    if (t2 < 1e-12f) {
        a = 1.0f - t2 * (1.0f / 6.0f);
        b = 0.5f - t2 * (1.0f / 24.0f);
    } else {
        float t = sqrtf(t2);
        float s, c;
        sincosf(t, &s, &c);
        a = s / t;
        // 1-cos(t) = 2 sin^2(t/2): no cancellation for small angles
        float sh = sinf(0.5f * t);
        b = 2.0f * sh * sh / t2;
    }
    R[0] = 1.0f - b * (ay * ay + az * az);
    R[1] = b * ax * ay - a * az;
    R[2] = b * ax * az + a * ay;
    R[3] = b * ax * ay + a * az;
    R[4] = 1.0f - b * (ax * ax + az * az);
    R[5] = b * ay * az - a * ax;
    R[6] = b * ax * az - a * ay;
    R[7] = b * ay * az + a * ax;
    R[8] = 1.0f - b * (ax * ax + ay * ay);
}

__device__ __forceinline__ void mat3_mul(const float* A, const float* B, float* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ void mat3_vec(const float* A, const float* v, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

__device__ __forceinline__ void mat3t_vec(const float* A, const float* v, float* o) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}

// Rotation matrix -> axis-angle through a unit quaternion (anakin/utils/transform.py:291-306 delegates to
// pytorch3d matrix_to_quaternion + quaternion_to_axis_angle; same candidate selection, fp32).
__device__ __forceinline__ void rotmat_to_aa(const float* m, float* aa) {
    float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[3], m11 = m[4], m12 = m[5], m20 = m[6], m21 = m[7], m22 = m[8];
    float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) qa[i] = sqrtf(fmaxf(qa[i], 0.0f));
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (qa[i] > qa[best]) best = i;
    float q[4];
    if (best == 0) {
        q[0] = qa[0] * qa[0]; q[1] = m21 - m12; q[2] = m02 - m20; q[3] = m10 - m01;
    } else if (best == 1) {
        q[0] = m21 - m12; q[1] = qa[1] * qa[1]; q[2] = m10 + m01; q[3] = m02 + m20;
    } else if (best == 2) {
        q[0] = m02 - m20; q[1] = m10 + m01; q[2] = qa[2] * qa[2]; q[3] = m12 + m21;
    } else {
        q[0] = m10 - m01; q[1] = m20 + m02; q[2] = m21 + m12; q[3] = qa[3] * qa[3];
    }
    float inv = 1.0f / (2.0f * fmaxf(qa[best], 0.1f));
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] *= inv;
    float n = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    float half = atan2f(n, q[0]);
    float ang = 2.0f * half;
    float s = fabsf(ang) < 1e-6f ? 0.5f - ang * ang * (1.0f / 48.0f) : sinf(half) / ang;
    aa[0] = q[1] / s;
    aa[1] = q[2] / s;
    aa[2] = q[3] / s;
}

// Forward kinematics of one sample. R [16][9] joint rotations, J [16][3] rest joints.
// Out: G [16][12] (row-major 3x4 global transforms, chain order).  Serial: 15 3x3 products, done by one thread.
__device__ __forceinline__ void mano_chain(const float* R, const float* J, float* G) {
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        int p = kManoParents[k];
        float* g = G + 12 * k;
        if (p < 0) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                g[4 * i] = R[3 * i]; g[4 * i + 1] = R[3 * i + 1]; g[4 * i + 2] = R[3 * i + 2];
                g[4 * i + 3] = J[i];
            }
        } else {
            const float* gp = G + 12 * p;
            const float* r = R + 9 * k;
            float rel[3] = {J[3 * k] - J[3 * p], J[3 * k + 1] - J[3 * p + 1], J[3 * k + 2] - J[3 * p + 2]};
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float a = gp[4 * i], b = gp[4 * i + 1], c = gp[4 * i + 2];
                g[4 * i] = a * r[0] + b * r[3] + c * r[6];
                g[4 * i + 1] = a * r[1] + b * r[4] + c * r[7];
                g[4 * i + 2] = a * r[2] + b * r[5] + c * r[8];
                g[4 * i + 3] = a * rel[0] + b * rel[1] + c * rel[2] + gp[4 * i + 3];
            }
        }
    }
}

}  // namespace ab
