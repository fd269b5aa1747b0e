// Training-side kernels of the clasbased network (NHWC bf16 activations, fp32 statistics / parameter gradients):
// what torch.nn.BatchNorm2d / ReLU / MaxPool2d / autograd do around the convolutions in the reference's train step
// (anakin/models/resnet.py:72-152, simplebaseline.py:161-190, train/train_artiboost.py:91-96).
//
//   ab_col_stats          per-column sum / sum of squares of a [M,C] bf16 or fp32 matrix (BN batch statistics, bias grads)
//   ab_bn_finalize        mean / biased var -> (scale, shift) for the apply pass, saved mean / invstd, running-stat update
//   ab_bn_apply           y = relu(raw * scale + shift (+ residual))
//   ab_bn_bwd_reduce      dy' = dy * (y > 0);  sum(dy'), sum(dy' * xhat)   (= dbeta, dgamma)
//   ab_bn_bwd_apply       dx = gamma * invstd * (dy' - mean(dy') - xhat * mean(dy' * xhat));  optional copy of dy'
//   ab_maxpool3x3s2_bwd, ab_avgpool_bwd, ab_dilate2x (zero insertion for stride-2 data gradients),
//   ab_deconv4x4s2_gather (transpose of the col2im gather), ab_head_decode_bwd (softmax / soft-argmax backward),
//   ab_adam_step + ab_sumsq (fused multi-tensor Adam with gradient-norm clipping, train_artiboost.py:94-96)
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace ab {

__device__ __forceinline__ void bf16x8_to_float(const uint4& v, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(p[j]);
        f[2 * j] = t.x; f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 float_to_bf16x8(const float* f) {
    uint4 v;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return v;
}

constexpr int kRedThreads = 256;

// Column reduction skeleton, deterministic (no atomics): CTA b of gridDim.x visits row slices b, b + grid, ...; thread t
// owns column group g = t % C8 and row t / C8 of each slice (a warp reads whole contiguous rows).  The CTA's sums are
// combined in shared memory and stored as row b of the partial matrices part[a] ([gridDim.x, C] each); the small
// partial_sum_kernel below adds the rows up.
template <int NACC, class RowFn>
__device__ __forceinline__ void column_partial(int M, int C8, float* const* part, RowFn fn) {
    __shared__ float red[kRedThreads][8 * NACC + 1];
    const int t = threadIdx.x;
    const int tpr = min(C8, kRedThreads);       // threads per row
    const int rows_par = kRedThreads / tpr;     // rows visited in parallel
    const int g0 = t % tpr, rl = t / tpr;
    const int C = 8 * C8;
    for (int gbase = 0; gbase < C8; gbase += tpr) {   // every thread takes every pass (barriers inside)
        const int g = gbase + g0;
        float acc[8 * NACC];
#pragma unroll
        for (int j = 0; j < 8 * NACC; ++j) acc[j] = 0.0f;
        if (rl < rows_par && g < C8) {
#pragma unroll 4
            for (long long r = (long long)blockIdx.x * rows_par + rl; r < M; r += (long long)gridDim.x * rows_par) fn(r, g, acc);
        }
#pragma unroll
        for (int j = 0; j < 8 * NACC; ++j) red[t][j] = acc[j];
        __syncthreads();
        // the rows_par row-lanes of a column are added by tpr * 8 * NACC threads in parallel (a fixed order: k ascending),
        // consecutive threads on consecutive columns of the partial row
        for (int i = t; i < tpr * 8 * NACC; i += kRedThreads) {
            const int a = i / (tpr * 8), c = i - a * (tpr * 8);   // accumulator set, column within this pass
            const int gg = c >> 3, j = a * 8 + (c & 7);
            if (8 * gbase + c >= C) continue;
            float sum = red[gg][j];
            for (int k = 1; k < rows_par; ++k) sum += red[gg + k * tpr][j];
            part[a][(size_t)blockIdx.x * C + 8 * gbase + c] = sum;
        }
        __syncthreads();
    }
}

// out_a[c] = sum_p part_a[p, c]  for a = 0, 1 (part1 / out1 may be null).  Block = 8 columns x 128 partial rows
// (independent loads, unrolled: the kernel is pure latency).
__global__ void __launch_bounds__(1024)
partial_sum_kernel(const float* __restrict__ part0, const float* __restrict__ part1, int n_part, int C, float* out0, float* out1) {
    __shared__ float red[2][32][8];
    const int cx = threadIdx.x & 7, pr = threadIdx.x >> 3, c = blockIdx.x * 8 + cx;
    float s0 = 0.0f, s1 = 0.0f;
    if (c < C) {
#pragma unroll 4
        for (int p = pr; p < n_part; p += 128) {
            s0 += __ldg(part0 + (size_t)p * C + c);
            if (part1) s1 += __ldg(part1 + (size_t)p * C + c);
        }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane < 8) { red[0][wid][lane] = s0; red[1][wid][lane] = s1; }
    __syncthreads();
    if (threadIdx.x < 8 && c < C) {
        float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
        for (int w = 0; w < 32; ++w) { a0 += red[0][w][threadIdx.x]; a1 += red[1][w][threadIdx.x]; }
        out0[c] = a0;
        if (out1) out1[c] = a1;
    }
}

__global__ void __launch_bounds__(kRedThreads)
col_stats_bf16_kernel(const uint4* __restrict__ x, int M, int C8, long long ld8, float* psum, float* psumsq) {
    float* outs[2] = {psum, psumsq};
    if (psumsq) {
        column_partial<2>(M, C8, outs, [&](long long r, int g, float* acc) {
            float f[8];
            bf16x8_to_float(__ldg(x + r * ld8 + g), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] += f[j] * f[j]; }
        });
    } else {
        column_partial<1>(M, C8, outs, [&](long long r, int g, float* acc) {
            float f[8];
            bf16x8_to_float(__ldg(x + r * ld8 + g), f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        });
    }
}

__global__ void __launch_bounds__(kRedThreads)
col_stats_f32_kernel(const float4* __restrict__ x, int M, int C8, long long ld4, float* psum, float* psumsq) {
    float* outs[2] = {psum, psumsq};
    auto load = [&](long long r, int g, float* f) {
        const float4 a = __ldg(x + r * ld4 + 2 * g), b = __ldg(x + r * ld4 + 2 * g + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    };
    if (psumsq) {
        column_partial<2>(M, C8, outs, [&](long long r, int g, float* acc) {
            float f[8];
            load(r, g, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[j] += f[j]; acc[8 + j] += f[j] * f[j]; }
        });
    } else {
        column_partial<1>(M, C8, outs, [&](long long r, int g, float* acc) {
            float f[8];
            load(r, g, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        });
    }
}

// Batch statistics from the per-row-tile partial sums of the convolution epilogue (ab_gemm_bf16 / ab_conv_bf16_nhwc),
// then mean / biased variance -> folded (scale, shift), saved mean / invstd, running statistics.
// Block = 8 channels x 128 partial rows.
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psumsq, int n_part, int C, float count,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                   float* scale, float* shift, float* save_mean, float* save_invstd, float* running_mean,
                   float* running_var) {
    __shared__ float red[2][32][8];
    const int cx = threadIdx.x & 7, pr = threadIdx.x >> 3, c = blockIdx.x * 8 + cx;
    float s0 = 0.0f, s1 = 0.0f;
    if (c < C) {
#pragma unroll 4
        for (int p = pr; p < n_part; p += 128) {
            s0 += __ldg(psum + (size_t)p * C + c);
            s1 += __ldg(psumsq + (size_t)p * C + c);
        }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane < 8) { red[0][wid][lane] = s0; red[1][wid][lane] = s1; }
    __syncthreads();
    if (threadIdx.x >= 8 || c >= C) return;
    float sum = 0.0f, sumsq = 0.0f;
#pragma unroll
    for (int w = 0; w < 32; ++w) { sum += red[0][w][threadIdx.x]; sumsq += red[1][w][threadIdx.x]; }
    const float mean = sum / count;
    const float var = fmaxf(sumsq / count - mean * mean, 0.0f);  // biased, as used for normalisation
    const float invstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.0f, b = beta ? beta[c] : 0.0f;
    scale[c] = g * invstd;
    shift[c] = b - mean * g * invstd;
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    if (running_mean) {
        running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * mean;
        const float unbiased = count > 1.0f ? var * count / (count - 1.0f) : var;
        running_var[c] = (1.0f - momentum) * running_var[c] + momentum * unbiased;
    }
}

// Streaming skeleton shared by the BatchNorm apply / backward-apply passes: thread t owns column group t % C8 (its
// per-channel coefficients stay in registers) and walks rows blockIdx.x * rows_par + t / C8 (+ grid stride).
template <class Setup, class Body>
__device__ __forceinline__ void stream_rows(long long M, int C8, Setup setup, Body body) {
    const int t = threadIdx.x;
    const int tpr = min(C8, kRedThreads), rows_par = kRedThreads / tpr;
    const int g0 = t % tpr, rl = t / tpr;
    if (rl >= rows_par) return;
    for (int g = g0; g < C8; g += tpr) {
        setup(g);
#pragma unroll 4
        for (long long r = (long long)blockIdx.x * rows_par + rl; r < M; r += (long long)gridDim.x * rows_par) body(r * C8 + g);
    }
}

__device__ __forceinline__ void load8(const float* p, float* f) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

__global__ void __launch_bounds__(kRedThreads)
bn_apply_kernel(const uint4* __restrict__ raw, long long M, int C8, const float* __restrict__ scale,
                const float* __restrict__ shift, const uint4* __restrict__ residual, int relu, uint4* __restrict__ y) {
    float sc[8], sh[8];
    stream_rows(M, C8, [&](int g) { load8(scale + 8 * g, sc); load8(shift + 8 * g, sh); },
                [&](long long i) {
                    float f[8], r[8];
                    bf16x8_to_float(__ldg(raw + i), f);
                    if (residual) bf16x8_to_float(__ldg(residual + i), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float v = fmaf(f[j], sc[j], sh[j]);
                        if (residual) v += r[j];
                        f[j] = relu ? fmaxf(v, 0.0f) : v;
                    }
                    y[i] = float_to_bf16x8(f);
                });
}

// dy' = dy * (y > 0);  partial column sums of dy' and dy' * raw (the x-hat form follows in bn_bwd_finalize_kernel).
// Without a residual the forward output is y = relu(raw * scale + shift), so the mask follows from raw and the forward's
// (scale, shift) and y need not be read at all (y == nullptr): one activation pass less.
template <int MINB>
__global__ void __launch_bounds__(kRedThreads, MINB)
bn_bwd_reduce_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ raw, int M, int C8,
                     int relu, const float* __restrict__ fwd_scale, const float* __restrict__ fwd_shift, float* psum_dy,
                     float* psum_dy_x) {
    float* outs[2] = {psum_dy, psum_dy_x};
    const bool from_raw = relu && y == nullptr;
    float sc[8], sh[8];
    int cur_g = -1;
    column_partial<2>(M, C8, outs, [&](long long r, int g, float* acc) {
        float d[8], yy[8], x[8];
        bf16x8_to_float(__ldg(dy + r * C8 + g), d);
        bf16x8_to_float(__ldg(raw + r * C8 + g), x);
        if (from_raw) {
            if (g != cur_g) { load8(fwd_scale + 8 * g, sc); load8(fwd_shift + 8 * g, sh); cur_g = g; }
#pragma unroll
            for (int j = 0; j < 8; ++j) yy[j] = fmaf(x[j], sc[j], sh[j]);
        } else if (relu) {
            bf16x8_to_float(__ldg(y + r * C8 + g), yy);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dd = (relu && !(yy[j] > 0.0f)) ? 0.0f : d[j];
            acc[j] += dd;
            acc[8 + j] = fmaf(dd, x[j], acc[8 + j]);
        }
    });
}

// Adds the partial rows, then: dbeta = sum dy', dgamma = sum dy' * xhat = (sum dy'x - mean * sum dy') * invstd, written
// or accumulated into the parameter gradients, and the three per-channel coefficients of the apply pass
//   dx = gamma*invstd*(dy' - dbeta/M - xhat*dgamma/M) = A*dy' + B*raw + K,
//   A = gamma*invstd, B = -A*invstd*dgamma/M, K = A*(mean*invstd*dgamma - dbeta)/M.            coef: [3, C]
__global__ void __launch_bounds__(1024)
bn_bwd_finalize_kernel(const float* __restrict__ part0, const float* __restrict__ part1, int n_part, int C, float inv_count,
                       const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ invstd,
                       float* dgamma, float* dbeta, int accumulate, float* coef) {
    __shared__ float red[2][32][8];
    const int cx = threadIdx.x & 7, pr = threadIdx.x >> 3, c = blockIdx.x * 8 + cx;
    float s0 = 0.0f, s1 = 0.0f;
    if (c < C) {
#pragma unroll 4
        for (int p = pr; p < n_part; p += 128) {
            s0 += __ldg(part0 + (size_t)p * C + c);
            s1 += __ldg(part1 + (size_t)p * C + c);
        }
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 8);  s1 += __shfl_xor_sync(0xffffffffu, s1, 8);
    s0 += __shfl_xor_sync(0xffffffffu, s0, 16); s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane < 8) { red[0][wid][lane] = s0; red[1][wid][lane] = s1; }
    __syncthreads();
    if (threadIdx.x >= 8 || c >= C) return;
    float sdy = 0.0f, sdyx = 0.0f;
#pragma unroll
    for (int w = 0; w < 32; ++w) { sdy += red[0][w][threadIdx.x]; sdyx += red[1][w][threadIdx.x]; }
    const float mu = mean[c], is = invstd[c], gm = gamma ? gamma[c] : 1.0f;
    const float dg = (sdyx - mu * sdy) * is;
    if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
    if (dbeta) dbeta[c] = accumulate ? dbeta[c] + sdy : sdy;
    const float A = gm * is;
    coef[c] = A;
    coef[C + c] = -A * is * dg * inv_count;
    coef[2 * C + c] = A * (mu * is * dg - sdy) * inv_count;
}

template <int MINB>
__global__ void __launch_bounds__(kRedThreads, MINB)
bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, const uint4* __restrict__ raw, long long M,
                    int C8, const float* __restrict__ coef, int relu, const float* __restrict__ fwd_scale,
                    const float* __restrict__ fwd_shift, uint4* __restrict__ dx, uint4* __restrict__ dres) {
    float A[8], Bc[8], K[8], sc[8], sh[8];
    const int C = 8 * C8;
    const bool from_raw = relu && y == nullptr;  // mask from raw * scale + shift (see bn_bwd_reduce_kernel)
    stream_rows(M, C8, [&](int g) {
                    load8(coef + 8 * g, A); load8(coef + C + 8 * g, Bc); load8(coef + 2 * C + 8 * g, K);
                    if (from_raw) { load8(fwd_scale + 8 * g, sc); load8(fwd_shift + 8 * g, sh); }
                },
                [&](long long i) {
                    float d[8], yy[8], x[8], o[8];
                    bf16x8_to_float(__ldg(dy + i), d);
                    bf16x8_to_float(__ldg(raw + i), x);
                    if (from_raw) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) yy[j] = fmaf(x[j], sc[j], sh[j]);
                    } else if (relu) {
                        bf16x8_to_float(__ldg(y + i), yy);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float dd = (relu && !(yy[j] > 0.0f)) ? 0.0f : d[j];
                        d[j] = dd;
                        o[j] = fmaf(A[j], dd, fmaf(Bc[j], x[j], K[j]));
                    }
                    dx[i] = float_to_bf16x8(o);
                    if (dres) dres[i] = float_to_bf16x8(d);
                });
}

// eval-mode / frozen BN inside a training graph, or plain ReLU: dx = dy * (y > 0) * scale
__global__ void affine_relu_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, long long n8, int C8,
                                       const float* __restrict__ scale, int relu, uint4* __restrict__ dx,
                                       uint4* __restrict__ dres) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const int g = (int)(i % C8);
    float d[8], yy[8], o[8];
    bf16x8_to_float(__ldg(dy + i), d);
    if (relu) bf16x8_to_float(__ldg(y + i), yy);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        d[j] = (relu && !(yy[j] > 0.0f)) ? 0.0f : d[j];
        o[j] = d[j] * (scale ? scale[8 * g + j] : 1.0f);
    }
    dx[i] = float_to_bf16x8(o);
    if (dres) dres[i] = float_to_bf16x8(d);
}

// MaxPool2d(3, 2, 1) backward, gather form: an input pixel receives dy of every window (at most 4) whose recorded
// argmax tap (ky*3 + kx of the FIRST maximum in scan order -- torch's tie-break, written by the forward kernel) it is.
// One thread per 2 x 2 block of input pixels (rows 2j, 2j+1; columns 2i, 2i+1) and 8 channels: the block lies inside
// the four windows (j, i), (j, i+1), (j+1, i), (j+1, i+1) and no others, so idx / dy of each window are loaded once per
// block (8 loads instead of 18 per four pixels) and each (window, channel) tap is compared with the 4 / 2 / 2 / 1 block
// pixels the window covers.
__global__ void maxpool_bwd_kernel(const uint2* __restrict__ idx, const uint4* __restrict__ dy, int B, int H, int W, int C8,
                                   int Ho, int Wo, uint4* __restrict__ dx) {
    const int Hb = (H + 1) / 2, Wb = (W + 1) / 2;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Hb * Wb * C8;
    if (i >= total) return;
    const int g = (int)(i % C8);
    long long t = i / C8;
    const int bi = (int)(t % Wb);
    t /= Wb;
    const int bj = (int)(t % Hb);
    const int b = (int)(t / Hb);
    float acc[4][8];   // block pixels (0,0), (0,1), (1,0), (1,1)
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = 0.0f;
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
            const int oy = bj + wy, ox = bi + wx;
            if (oy >= Ho || ox >= Wo) continue;
            const long long o = (((long long)b * Ho + oy) * Wo + ox) * C8 + g;
            const uint2 id = __ldg(idx + o);
            float dv[8];
            bf16x8_to_float(__ldg(dy + o), dv);
            // window (oy, ox) covers input rows 2 oy - 1 .. 2 oy + 1: block row r is its tap row ky = 2 bj + r - (2 oy - 1) = r + 1 - 2 wy
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int ky = r + 1 - 2 * wy;
                if (ky < 0) continue;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int kx = c + 1 - 2 * wx;
                    if (kx < 0) continue;
                    const uint32_t me = (uint32_t)(ky * 3 + kx);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t tap = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xffu;
                        if (tap == me) acc[2 * r + c][j] += dv[j];
                    }
                }
            }
        }
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int iy = 2 * bj + r, ix = 2 * bi + c;
            if (iy < H && ix < W) dx[(((long long)b * H + iy) * W + ix) * C8 + g] = float_to_bf16x8(acc[2 * r + c]);
        }
}

__global__ void avgpool_bwd_kernel(const float* __restrict__ dmean, int B, int HW, int C8, uint4* __restrict__ dx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * HW * C8;
    if (i >= total) return;
    const int g = (int)(i % C8);
    const int b = (int)(i / ((long long)HW * C8));
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = dmean[(long long)b * C8 * 8 + 8 * g + j] / (float)HW;
    dx[i] = float_to_bf16x8(f);
}

// out[b, 2*oy, 2*ox, :] = in[b, oy, ox, :], zeros elsewhere; out is [B, H, W, C] (H, W of the stride-2 conv's input)
__global__ void dilate2x_kernel(const uint4* __restrict__ in, int B, int Ho, int Wo, int H, int W, int C8, uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * H * W * C8;
    if (i >= total) return;
    const int g = (int)(i % C8);
    long long t = i / C8;
    const int x = (int)(t % W);
    t /= W;
    const int y = (int)(t % H);
    const int b = (int)(t / H);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (!(x & 1) && !(y & 1) && (y >> 1) < Ho && (x >> 1) < Wo) v = __ldg(in + (((long long)b * Ho + (y >> 1)) * Wo + (x >> 1)) * C8 + g);
    out[i] = v;
}

// dycol[b, iy, ix, (ky, kx, co)] = dy[b, 2*iy - 1 + ky, 2*ix - 1 + kx, co] (0 outside): transpose of the col2im gather
__global__ void deconv_gather_kernel(const uint4* __restrict__ dy, int B, int H, int W, int C8, uint4* __restrict__ dycol) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * H * W * 16 * C8;
    if (i >= total) return;
    const int g = (int)(i % C8);
    long long t = i / C8;
    const int tap = (int)(t % 16);
    t /= 16;
    const int ix = (int)(t % W);
    t /= W;
    const int iy = (int)(t % H);
    const int b = (int)(t / H);
    const int oy = 2 * iy - 1 + (tap >> 2), ox = 2 * ix - 1 + (tap & 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (oy >= 0 && oy < 2 * H && ox >= 0 && ox < 2 * W) v = __ldg(dy + (((long long)b * 2 * H + oy) * 2 * W + ox) * C8 + g);
    dycol[i] = v;
}

// Backward of ab_head_decode w.r.t. the logits: p = softmax(l); u = sum p * w / W ...; dl_i = p_i * (g_i - sum_j p_j g_j)
// with g_i = du * w_i / W + dv * h_i / H + dd * d_i / D (the 1 / (1 + 1e-7) re-normalisation factor included).
constexpr int kDecodeThreads = 256;
__global__ void __launch_bounds__(kDecodeThreads)
head_decode_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dkp3d, int ncls, int D, int H, int W,
                       __nv_bfloat16* __restrict__ dlogits) {
    __shared__ float red[2][kDecodeThreads / 32];
    const int b = blockIdx.x / ncls, cls = blockIdx.x % ncls;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int HW = H * W, n = HW * D, ldc = ncls * D;
    const float* base = logits + (long long)b * HW * ldc + cls * D;
    __nv_bfloat16* obase = dlogits + (long long)b * HW * ldc + cls * D;
    const float renorm = 1.0f / (1.0f + 1e-7f);
    const float gu = dkp3d[((long long)b * ncls + cls) * 3] * renorm / (float)W;
    const float gv = dkp3d[((long long)b * ncls + cls) * 3 + 1] * renorm / (float)H;
    const float gd = dkp3d[((long long)b * ncls + cls) * 3 + 2] * renorm / (float)D;
    float mx = -INFINITY;
    for (int i = tid; i < n; i += kDecodeThreads) {
        const int p = i / D, d = i - p * D;
        mx = fmaxf(mx, base[(long long)p * ldc + d]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[0][wid] = mx;
    __syncthreads();
    mx = red[0][0];
#pragma unroll
    for (int w = 1; w < kDecodeThreads / 32; ++w) mx = fmaxf(mx, red[0][w]);
    __syncthreads();
    float s = 0.f, sg = 0.f;
    for (int i = tid; i < n; i += kDecodeThreads) {
        const int p = i / D, d = i - p * D;
        const int h = p / W, w = p - h * W;
        const float e = expf(base[(long long)p * ldc + d] - mx);
        s += e;
        sg += e * (gu * (float)w + gv * (float)h + gd * (float)d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        sg += __shfl_xor_sync(0xffffffffu, sg, o);
    }
    if (lane == 0) { red[0][wid] = s; red[1][wid] = sg; }
    __syncthreads();
    float S = 0.f, SG = 0.f;
    for (int w = 0; w < kDecodeThreads / 32; ++w) { S += red[0][w]; SG += red[1][w]; }
    const float inv = 1.0f / S, gbar = SG * inv;
    for (int i = tid; i < n; i += kDecodeThreads) {
        const int p = i / D, d = i - p * D;
        const int h = p / W, w = p - h * W;
        const float pr = expf(base[(long long)p * ldc + d] - mx) * inv;
        obase[(long long)p * ldc + d] = __float2bfloat16(pr * (gu * (float)w + gv * (float)h + gd * (float)d - gbar));
    }
}

// The same gradient in ONE sweep, given what the forward pass left behind: lse = max + log(sum exp) so that
// p_i = exp(l_i - lse), and kp3d itself, because sum_j p_j g_j = du * u + dv * v + dd * d (g is linear in the bin
// coordinates and kp3d holds their expectations with the same 1 / W, 1 / H, 1 / D and re-normalisation factors).
// D % 4 == 0: 16-byte loads, 8-byte stores.
__global__ void __launch_bounds__(kDecodeThreads)
head_decode_bwd_lse_kernel(const float* __restrict__ logits, const float* __restrict__ dkp3d, const float* __restrict__ kp3d,
                           const float* __restrict__ lse, int ncls, int D, int H, int W, __nv_bfloat16* __restrict__ dlogits) {
    const int b = blockIdx.x / ncls, cls = blockIdx.x % ncls;
    const int tid = threadIdx.x;
    const int HW = H * W, D4 = D / 4, n4 = HW * D4, ldc = ncls * D;
    const float* base = logits + (long long)b * HW * ldc + cls * D;
    __nv_bfloat16* obase = dlogits + (long long)b * HW * ldc + cls * D;
    const float renorm = 1.0f / (1.0f + 1e-7f);
    const float* dk = dkp3d + ((long long)b * ncls + cls) * 3;
    const float* kp = kp3d + ((long long)b * ncls + cls) * 3;
    const float gu = dk[0] * renorm / (float)W, gv = dk[1] * renorm / (float)H, gd = dk[2] * renorm / (float)D;
    const float gbar = dk[0] * kp[0] + dk[1] * kp[1] + dk[2] * kp[2];
    const float L = lse[(long long)b * ncls + cls];
    constexpr int U = 4;
    const int dq = kDecodeThreads % D4, dp = kDecodeThreads / D4;
    int p = tid / D4, q = tid - p * D4;
    for (int i0 = tid; i0 < n4; i0 += U * kDecodeThreads) {
        float4 v[U];
        int pp[U], qq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pp[u] = p; qq[u] = q;
            if (i0 + u * kDecodeThreads < n4) v[u] = __ldg(reinterpret_cast<const float4*>(base + (long long)p * ldc + 4 * q));
            q += dq; p += dp;
            if (q >= D4) { q -= D4; ++p; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * kDecodeThreads >= n4) break;
            const int h = pp[u] / W, w = pp[u] - h * W;
            const float g0 = gu * (float)w + gv * (float)h + gd * (float)(4 * qq[u]) - gbar;
            const float r0 = __expf(v[u].x - L) * g0, r1 = __expf(v[u].y - L) * (g0 + gd), r2 = __expf(v[u].z - L) * (g0 + 2.0f * gd),
                        r3 = __expf(v[u].w - L) * (g0 + 3.0f * gd);
            __nv_bfloat162 o01 = __floats2bfloat162_rn(r0, r1), o23 = __floats2bfloat162_rn(r2, r3);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&o01); pk.y = *reinterpret_cast<unsigned*>(&o23);
            *reinterpret_cast<uint2*>(obase + (long long)pp[u] * ldc + 4 * qq[u]) = pk;
        }
    }
}

// ---- optimiser: sum of squares of the flat gradient (for clip_grad_norm_) and the fused Adam update
// Two passes in a fixed order (no atomics: the clip coefficient, and with it the whole step, is bit-reproducible):
// CTA b leaves its partial in out[1 + b], sumsq_finish_kernel adds the partials up into out[0].
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* out) {
    __shared__ float red[32];
    float s = 0.0f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += g[i] * g[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[1 + blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) sumsq_finish_kernel(float* out, int n_part) {
    __shared__ float red[8];
    float s = 0.0f;
    for (int i = threadIdx.x; i < n_part; i += 256) s += out[1 + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = t;
    }
}

// torch.optim.Adam (no amsgrad, weight_decay added to the gradient) on one flat fp32 buffer.  grad_sumsq (device scalar):
// total squared norm of the gradient; the clip coefficient min(1, max_norm / (norm + 1e-6)) of clip_grad_norm_ is
// applied on the fly (max_norm <= 0: no clipping).  state = {step count, 1 - beta1^t, 1 - beta2^t} lives on the device
// and is advanced by adam_tick_kernel, so a captured CUDA graph of the step replays with the right bias corrections.
__global__ void adam_tick_kernel(float* state, float beta1, float beta2) {
    const float t = state[0] + 1.0f;
    state[0] = t;
    state[1] = 1.0f - powf(beta1, t);
    state[2] = 1.0f - powf(beta2, t);
}

__global__ void adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                            long long n4, float lr, float beta1, float beta2, float eps, float weight_decay,
                            const float* __restrict__ state, const float* __restrict__ grad_sumsq, float max_norm,
                            float grad_scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float bias1 = state[1], rsb2 = rsqrtf(state[2]);
    float clip = 1.0f;
    if (max_norm > 0.0f) clip = fminf(1.0f, max_norm / (sqrtf(*grad_sumsq) * grad_scale + 1e-6f));
    const float gs = grad_scale * clip, step = lr / bias1;
    float4 P = p[i], G = __ldg(g + i), M = m[i], V = v[i];
    float* pp = &P.x; float* gg = &G.x; float* mm = &M.x; float* vv = &V.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float gi = gg[j] * gs + weight_decay * pp[j];
        mm[j] = beta1 * mm[j] + (1.0f - beta1) * gi;
        vv[j] = beta2 * vv[j] + (1.0f - beta2) * gi * gi;
        pp[j] -= step * mm[j] / (sqrtf(vv[j]) * rsb2 + eps);
    }
    p[i] = P; m[i] = M; v[i] = V;
}

static inline unsigned nblk(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace ab

using namespace ab;

#define AB_LAUNCH_END(name)      \
    count_launch();              \
    return check_launch(name)

static inline bool bn_tuned() {
    static const bool v = !(getenv("AB_BN_TUNE") && atoi(getenv("AB_BN_TUNE")) == 0);
    return v;
}
static inline int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
static inline int stat_parts(int M, int C8) {
    const int rows_par = kRedThreads / min(C8, kRedThreads);
    return max(1, min(AB_STAT_PARTS, (M + 4 * rows_par - 1) / (4 * rows_par)));
}

extern "C" int ab_col_stats(const void* x, int is_f32, int M, int C, int64_t ld, float* sum, float* sumsq, float* ws, void* stream) {
    AB_REQUIRE(M >= 0 && C > 0 && C % 8 == 0 && ld % 8 == 0, "bad shape (C and ld multiples of 8)");
    AB_REQUIRE(x && sum && ws, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (M == 0) {
        AB_CUDA(cudaMemsetAsync(sum, 0, sizeof(float) * C, st));
        if (sumsq) AB_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(float) * C, st));
        return AB_OK;
    }
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    const int parts = stat_parts(M, C / 8);
    float* p0 = ws;
    float* p1 = sumsq ? ws + (size_t)AB_STAT_PARTS * C : nullptr;
    if (is_f32) col_stats_f32_kernel<<<parts, kRedThreads, 0, st>>>((const float4*)x, M, C / 8, ld / 4, p0, p1);
    else col_stats_bf16_kernel<<<parts, kRedThreads, 0, st>>>((const uint4*)x, M, C / 8, ld / 8, p0, p1);
    partial_sum_kernel<<<nblk(C, 8), 1024, 0, st>>>(p0, p1, parts, C, sum, sumsq);
    count_launch(2);
    return check_launch("col_stats_kernel");
}

extern "C" int ab_bn_finalize(const float* sum_part, const float* sumsq_part, int n_part, int C, float count, const float* gamma,
                              const float* beta, float eps, float momentum, float* scale, float* shift, float* save_mean,
                              float* save_invstd, float* running_mean, float* running_var, void* stream) {
    AB_REQUIRE(C > 0 && count > 0 && n_part > 0, "bad shape");
    AB_REQUIRE(sum_part && sumsq_part && scale && shift && save_mean && save_invstd, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_BN_FINALIZE, st);
    bn_finalize_kernel<<<nblk(C, 8), 1024, 0, st>>>(sum_part, sumsq_part, n_part, C, count, gamma, beta, eps, momentum, scale,
                                                    shift, save_mean, save_invstd, running_mean, running_var);
    AB_LAUNCH_END("bn_finalize_kernel");
}

static inline unsigned stream_grid(long long M, int C8) {
    const int rows_par = kRedThreads / min(C8, kRedThreads);
    return (unsigned)max(1ll, min((long long)148 * 8, (M + rows_par - 1) / rows_par));
}

extern "C" int ab_bn_apply(const void* raw, int64_t M, int C, const float* scale, const float* shift, const void* residual,
                           int relu, void* y, void* stream) {
    AB_REQUIRE(M >= 0 && C > 0 && C % 8 == 0, "bad shape");
    if (M == 0) return AB_OK;
    AB_REQUIRE(raw && scale && shift && y, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_BN_APPLY, st);
    bn_apply_kernel<<<stream_grid(M, C / 8), kRedThreads, 0, st>>>((const uint4*)raw, M, C / 8, scale, shift, (const uint4*)residual,
                                                                   relu, (uint4*)y);
    AB_LAUNCH_END("bn_apply_kernel");
}

extern "C" int ab_bn_bwd_reduce(const void* dy, const void* y, const void* raw, int M, int C, const float* gamma,
                                const float* mean, const float* invstd, int relu, float* dgamma, float* dbeta, int accumulate,
                                float* coef, float* ws, const float* fwd_scale, const float* fwd_shift, void* stream) {
    AB_REQUIRE(M > 0 && C > 0 && C % 8 == 0, "bad shape");
    AB_REQUIRE(dy && raw && mean && invstd && coef && ws && (!relu || y || (fwd_scale && fwd_shift)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_BN_BWD_REDUCE, st);
    // 3 CTAs per SM (80 registers) in ONE wave of persistent CTAs: with the compiler's free choice (98 registers, 2 CTAs per
    // SM) the 8-per-SM grid ran as four waves, each ending in its own shared-memory reduction (AB_BN_TUNE=0: that form)
    const bool tuned = bn_tuned();
    const int parts = tuned ? min(stat_parts(M, C / 8), 3 * sm_count()) : stat_parts(M, C / 8);
    float* p0 = ws;
    float* p1 = ws + (size_t)AB_STAT_PARTS * C;
    if (tuned)
        bn_bwd_reduce_kernel<3><<<parts, kRedThreads, 0, st>>>((const uint4*)dy, (const uint4*)y, (const uint4*)raw, M, C / 8, relu,
                                                               fwd_scale, fwd_shift, p0, p1);
    else
        bn_bwd_reduce_kernel<1><<<parts, kRedThreads, 0, st>>>((const uint4*)dy, (const uint4*)y, (const uint4*)raw, M, C / 8, relu,
                                                               fwd_scale, fwd_shift, p0, p1);
    bn_bwd_finalize_kernel<<<nblk(C, 8), 1024, 0, st>>>(p0, p1, parts, C, 1.0f / (float)M, gamma, mean, invstd, dgamma, dbeta,
                                                       accumulate, coef);
    count_launch(2);
    return check_launch("bn_bwd_reduce_kernel");
}

extern "C" int ab_bn_bwd_apply(const void* dy, const void* y, const void* raw, int64_t M, int C, const float* coef, int relu,
                               void* dx, void* dres, const float* fwd_scale, const float* fwd_shift, void* stream) {
    AB_REQUIRE(M >= 0 && C > 0 && C % 8 == 0, "bad shape");
    if (M == 0) return AB_OK;
    AB_REQUIRE(dy && raw && coef && dx && (!relu || y || (fwd_scale && fwd_shift)), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_BN_BWD_APPLY, st);
    if (bn_tuned())
        bn_bwd_apply_kernel<3><<<min(stream_grid(M, C / 8), 3u * (unsigned)sm_count()), kRedThreads, 0, st>>>((const uint4*)dy, (const uint4*)y, (const uint4*)raw, M,
                                                                              C / 8, coef, relu, fwd_scale, fwd_shift, (uint4*)dx, (uint4*)dres);
    else
        bn_bwd_apply_kernel<1><<<stream_grid(M, C / 8), kRedThreads, 0, st>>>((const uint4*)dy, (const uint4*)y, (const uint4*)raw, M,
                                                                              C / 8, coef, relu, fwd_scale, fwd_shift, (uint4*)dx, (uint4*)dres);
    AB_LAUNCH_END("bn_bwd_apply_kernel");
}

extern "C" int ab_affine_relu_bwd(const void* dy, const void* y, int64_t M, int C, const float* scale, int relu, void* dx,
                                  void* dres, void* stream) {
    AB_REQUIRE(M >= 0 && C > 0 && C % 8 == 0, "bad shape");
    if (M == 0) return AB_OK;
    AB_REQUIRE(dy && dx && (!relu || y), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    const long long n8 = M * (C / 8);
    affine_relu_bwd_kernel<<<nblk(n8, 256), 256, 0, st>>>((const uint4*)dy, (const uint4*)y, n8, C / 8, scale, relu, (uint4*)dx,
                                                          (uint4*)dres);
    AB_LAUNCH_END("affine_relu_bwd_kernel");
}

extern "C" int ab_maxpool3x3s2_bwd(const void* idx, const void* dy, int B, int H, int W, int C, void* dx, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(idx && dy && dx, "null pointer");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    maxpool_bwd_kernel<<<nblk((long long)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8), 256), 256, 0, st>>>((const uint2*)idx, (const uint4*)dy, B, H, W, C / 8,
                                                                                  Ho, Wo, (uint4*)dx);
    AB_LAUNCH_END("maxpool_bwd_kernel");
}

extern "C" int ab_avgpool_bwd(const float* dmean, int B, int HW, int C, void* dx, void* stream) {
    AB_REQUIRE(B >= 0 && HW > 0 && C > 0 && C % 8 == 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(dmean && dx, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    avgpool_bwd_kernel<<<nblk((long long)B * HW * (C / 8), 256), 256, 0, st>>>(dmean, B, HW, C / 8, (uint4*)dx);
    AB_LAUNCH_END("avgpool_bwd_kernel");
}

extern "C" int ab_dilate2x(const void* in, int B, int Ho, int Wo, int H, int W, int C, void* out, void* stream) {
    AB_REQUIRE(B >= 0 && Ho > 0 && Wo > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(in && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    dilate2x_kernel<<<nblk((long long)B * H * W * (C / 8), 256), 256, 0, st>>>((const uint4*)in, B, Ho, Wo, H, W, C / 8, (uint4*)out);
    AB_LAUNCH_END("dilate2x_kernel");
}

extern "C" int ab_deconv4x4s2_gather(const void* dy, int B, int H, int W, int C, void* dycol, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(dy && dycol, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_TRAIN_ELEMENTWISE, st);
    deconv_gather_kernel<<<nblk((long long)B * H * W * 16 * (C / 8), 256), 256, 0, st>>>((const uint4*)dy, B, H, W, C / 8, (uint4*)dycol);
    AB_LAUNCH_END("deconv_gather_kernel");
}

extern "C" int ab_head_decode_bwd(const float* logits, const float* dkp3d, const float* kp3d, const float* lse, int B, int ncls,
                                  int D, int H, int W, void* dlogits, void* stream) {
    AB_REQUIRE(B >= 0 && ncls > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(logits && dkp3d && dlogits, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_HEAD_DECODE, st);
    AB_REQUIRE((kp3d == nullptr) == (lse == nullptr), "kp3d and lse come together (both from ab_head_decode) or not at all");
    if (lse && D % 4 == 0 && D / 4 <= kDecodeThreads && ((uintptr_t)logits & 15) == 0 && ((uintptr_t)dlogits & 7) == 0)
        head_decode_bwd_lse_kernel<<<B * ncls, kDecodeThreads, 0, st>>>(logits, dkp3d, kp3d, lse, ncls, D, H, W, (__nv_bfloat16*)dlogits);
    else
        head_decode_bwd_kernel<<<B * ncls, kDecodeThreads, 0, st>>>(logits, dkp3d, ncls, D, H, W, (__nv_bfloat16*)dlogits);
    AB_LAUNCH_END("head_decode_bwd_kernel");
}

extern "C" int ab_sumsq(const float* g, int64_t n, float* out, void* stream) {
    AB_REQUIRE(n >= 0, "negative size");
    if (n == 0) return AB_OK;
    AB_REQUIRE(g && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_OPTIMIZER, st);
    const unsigned parts = (unsigned)min((long long)AB_SUMSQ_PARTS, (long long)nblk(n, 256));
    sumsq_kernel<<<parts, 256, 0, st>>>(g, n, out);
    sumsq_finish_kernel<<<1, 256, 0, st>>>(out, (int)parts);
    count_launch(1);
    AB_LAUNCH_END("sumsq_kernel");
}

extern "C" int ab_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float* state, const float* grad_sumsq, float max_norm, float grad_scale,
                            void* stream) {
    AB_REQUIRE(n >= 0 && n % 4 == 0, "n must be a non-negative multiple of 4 (pad the flat buffer)");
    if (n == 0) return AB_OK;
    AB_REQUIRE(p && g && m && v && state && (max_norm <= 0.0f || grad_sumsq), "null pointer");
    AB_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "buffers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_OPTIMIZER, st);
    adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2);
    adam_kernel<<<nblk(n / 4, 256), 256, 0, st>>>((float4*)p, (const float4*)g, (float4*)m, (float4*)v, n / 4, lr, beta1, beta2, eps,
                                                  weight_decay, state, grad_sumsq, max_norm, grad_scale);
    count_launch(2);
    return check_launch("adam_kernel");
}
