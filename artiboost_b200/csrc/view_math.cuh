// ViewEngine.get_view / get_perspective_from_id / caculate_align_mat (anakin/artiboost/view_engine.py:17-86) for one
// perspective id: shared by ab_view_from_id (ccv.cu), the fused synthesis draw and the blacklist kernel (synth.cu).
#pragma once
#include "common.cuh"

namespace ab {

// ru, rth: U[0,1) jitter of the bin (u, theta); writes the row-major 3x3 rotation aligning +z with the view direction.
__device__ __forceinline__ void view_rotmat(int pid, int u_bins, int theta_bins, float ru, float rth, float* persp_rotmat) {
    const double kPi = 3.141592653589793;
    const int u_id = pid / theta_bins, th_id = pid % theta_bins;
    const double u_unit = 2.0 / u_bins, th_unit = (2.0 * kPi) / theta_bins;
    // Precision follows the reference with a 0-dim torch tensor persp_id (ovg_set.py:138,141): bin centres, jittered
    // u / theta, the direction vector and its normalisation are fp32 (view_engine.py:36-58); torch.rand(1) - 0.5 is
    // an fp32 subtraction and the product with the bin size a python float; the align matrix algebra is fp64.
    const float u_c = __fadd_rn((float)(-1.0 + u_unit / 2), __fmul_rn((float)u_id, (float)u_unit));
    const float th_c = __fadd_rn((float)(th_unit / 2), __fmul_rn((float)th_id, (float)th_unit));
    const float u_off = (float)((double)__fsub_rn(ru, 0.5f) * u_unit);
    const float th_off = (float)((double)__fsub_rn(rth, 0.5f) * th_unit);
    const float uf = fminf(fmaxf(__fadd_rn(u_c, u_off), -1.0f), 1.0f);
    const float thf = fminf(fmaxf(__fadd_rn(th_c, th_off), 0.0f), (float)(2.0 * kPi));
    const float sf = __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(uf, uf)));
    const float xf = __fmul_rn(sf, cosf(thf)), yf = __fmul_rn(sf, sinf(thf));
    const float nf = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(xf, xf), __fmul_rn(yf, yf)), __fmul_rn(uf, uf)));
    const double vx = (double)__fdiv_rn(xf, nf), vy = (double)__fdiv_rn(yf, nf), vz = (double)__fdiv_rn(uf, nf);
    double M[9];
    if (vz == -1.0 || vz == 1.0) {
        const double d = vz;
        M[0] = d; M[1] = 0; M[2] = 0; M[3] = 0; M[4] = d; M[5] = 0; M[6] = 0; M[7] = 0; M[8] = d;
    } else {
        // k = z cross v = (-vy, vx, 0);  I + [k]x + [k]x^2 / (1 + z.v)
        const double kx = -vy, ky = vx, inv = 1.0 / (1.0 + vz);
        M[0] = 1.0 - ky * ky * inv; M[1] = kx * ky * inv;       M[2] = ky;
        M[3] = kx * ky * inv;       M[4] = 1.0 - kx * kx * inv; M[5] = -kx;
        M[6] = -ky;                 M[7] = kx;                  M[8] = 1.0 - (kx * kx + ky * ky) * inv;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) persp_rotmat[j] = (float)M[j];
}

// rroll, rz: U[0,1) draws of the free in-plane roll and the camera distance.
__device__ __forceinline__ void view_from_id(int pid, int u_bins, int theta_bins, float z_min, float z_max, float ru,
                                             float rth, float rroll, float rz, float* persp_rotmat, float* f, float* z_offset) {
    const double kPi = 3.141592653589793;
    view_rotmat(pid, u_bins, theta_bins, ru, rth, persp_rotmat);
    const double roll = (double)rroll * (2.0 * kPi);
    const float c = (float)cos(roll), sn = (float)sin(roll);
    f[0] = c;  f[1] = -sn; f[2] = 0;  f[3] = 0;
    f[4] = sn; f[5] = c;   f[6] = 0;  f[7] = 0;
    f[8] = 0;  f[9] = 0;   f[10] = 1; f[11] = 0;
    f[12] = 0; f[13] = 0;  f[14] = 0; f[15] = 1;
    z_offset[0] = 0.0f;
    z_offset[1] = 0.0f;
    // torch Uniform.sample: low + rand * (high - low), three separately rounded fp32 operations
    z_offset[2] = __fadd_rn(z_min, __fmul_rn(rz, __fsub_rn(z_max, z_min)));
}

}  // namespace ab
