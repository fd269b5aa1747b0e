// Data-movement kernels around the tensor-core contraction (gemm_tc.cu) for the clasbased network, NHWC bf16.
//
//   ab_image_to_nhwc        image f32 [B,C,H,W] -> bf16 [B,H,W,Cp] (Cp >= C, zero padded)   (resnet.py:200 input)
//   ab_im2col_nhwc          [B,H,W,C] -> rows [B*Ho*Wo, Kp], K order (ky, kx, c), zero padded to Kp
//                           (the A operand of conv = GEMM; resnet.py:154, conv3x3 / 1x1 stride-2 downsample)
//   ab_maxpool3x3s2_nhwc    nn.MaxPool2d(3, 2, 1)                                             (resnet.py:157,203)
//   ab_avgpool_nhwc         x.mean(3).mean(2)                                                 (resnet.py:219-221)
//   ab_deconv4x4s2_col2im   gather form of ConvTranspose2d(k4, s2, p1) after the GEMM X . W -> [B*H*W, 16*Cout],
//                           fused with the BatchNorm affine + ReLU that follow it          (simplebaseline.py:161-172)
//   ab_head_decode          softmax over D*H*W per class, confidence, re-normalisation, 3-D soft-argmax
//                           (simplebaseline.py:16-71,182-190) in ONE pass over the logits
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace ab {

__global__ void image_to_nhwc_kernel(const float* __restrict__ img, int B, int C, int H, int W, int Cp,
                                     __nv_bfloat16* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over B*H*W
    const long long n = (long long)B * H * W;
    if (i >= n) return;
    const int b = (int)(i / ((long long)H * W));
    const long long hw = i - (long long)b * H * W;
    __nv_bfloat16* o = out + i * Cp;
    if (Cp == 4) {   // the RGB image padded to 4 channels: one 8-byte store per pixel
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = c < C ? __ldg(img + ((long long)b * C + c) * H * W + hw) : 0.0f;
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(v[2], v[3]);
        *reinterpret_cast<uint2*>(o) = pk;
        return;
    }
    for (int c = 0; c < Cp; ++c)
        o[c] = __float2bfloat16(c < C ? img[((long long)b * C + c) * H * W + hw] : 0.0f);
}

// one thread per (output row, tap, 8-channel group) when C % 8 == 0
__global__ void im2col_vec8_kernel(const uint4* __restrict__ in, int B, int H, int W, int C8, int kh, int kw, int stride,
                                   int pad, int Ho, int Wo, int Kp8, uint4* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * Kp8;
    if (i >= total) return;
    const int k8 = (int)(i % Kp8);
    const long long m = i / Kp8;
    uint4 v = make_uint4(0, 0, 0, 0);
    const int tap = k8 / C8;
    if (tap < kh * kw) {
        const int c8 = k8 - tap * C8;
        const int ky = tap / kw, kx = tap - ky * kw;
        const int ox = (int)(m % Wo);
        const long long t = m / Wo;
        const int oy = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(in + (((long long)b * H + iy) * W + ix) * C8 + c8);
    }
    out[i] = v;
}

// generic element-wise variant (conv1: C = 3 or 4)
__global__ void im2col_scalar_kernel(const __nv_bfloat16* __restrict__ in, int B, int H, int W, int C, int kh, int kw,
                                     int stride, int pad, int Ho, int Wo, int Kp, __nv_bfloat16* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * Kp;
    if (i >= total) return;
    const int k = (int)(i % Kp);
    const long long m = i / Kp;
    __nv_bfloat16 v = __float2bfloat16(0.0f);
    const int tap = k / C;
    if (tap < kh * kw) {
        const int c = k - tap * C;
        const int ky = tap / kw, kx = tap - ky * kw;
        const int ox = (int)(m % Wo);
        const long long t = m / Wo;
        const int oy = (int)(t % Ho);
        const int b = (int)(t / Ho);
        const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[(((long long)b * H + iy) * W + ix) * C + c];
    }
    out[i] = v;
}

// C == 4 (the RGB image padded to 4 channels: 8 bytes per pixel).  One thread per (output row, filter row ky): the kw
// pixels of a filter row are contiguous in NHWC, copied as kw 8-byte loads / stores; the last filter row's thread also
// zero-fills the K padding.
__global__ void im2col_c4_kernel(const uint2* __restrict__ in, int B, int H, int W, int kh, int kw, int stride, int pad,
                                 int Ho, int Wo, int Kp4, uint2* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * kh;
    if (i >= total) return;
    const int ky = (int)(i % kh);
    const long long m = i / kh;
    const int ox = (int)(m % Wo);
    const long long t = m / Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int iy = oy * stride - pad + ky, ix0 = ox * stride - pad;
    uint2* o = out + m * Kp4 + ky * kw;
    const bool row_ok = iy >= 0 && iy < H;
    const uint2* src = in + ((long long)b * H + (row_ok ? iy : 0)) * W;
    for (int kx = 0; kx < kw; ++kx) {
        const int ix = ix0 + kx;
        o[kx] = (row_ok && ix >= 0 && ix < W) ? __ldg(src + ix) : make_uint2(0, 0);
    }
    if (ky == kh - 1)
        for (int k = kh * kw; k < Kp4; ++k) out[m * Kp4 + k] = make_uint2(0, 0);
}

// The same matrix written with fully coalesced 16-byte stores: one thread per pair of taps (2 x 8 bytes) of an output row,
// consecutive threads on consecutive 16-byte words of the row-major matrix (the per-(row, ky) form above scatters 8-byte
// stores 56 bytes apart: 1.9 TB/s on the 838 MB stem matrix).  blockIdx.y = (image, output row); the two divisions left
// per thread (by the words per row and by kw) are multiply-shifts with host-made constants, exact for operands < 2^16.
// The input is small (67 MB) and every pixel is read ~12 times: it stays in L1 / L2.
__global__ void im2col_c4_pairs_kernel(const uint2* __restrict__ in, int H, int W, int kh, int kw, int stride, int pad,
                                       int Ho, int Wo, int Kp8, unsigned mul_kp8, unsigned mul_kw, uint4* __restrict__ out) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;   // (ox, q) within the output row
    if (idx >= (unsigned)(Wo * Kp8)) return;
    const unsigned ox = (idx * mul_kp8) >> 16, q = idx - ox * (unsigned)Kp8;
    const int oy = blockIdx.y % Ho, b = blockIdx.y / Ho;
    const int taps = kh * kw;
    const uint2* img = in + (long long)b * H * W;
    uint2 v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const unsigned tap = 2 * q + h;
        const unsigned ky = (tap * mul_kw) >> 16, kx = tap - ky * (unsigned)kw;
        const int iy = oy * stride - pad + (int)ky, ix = (int)ox * stride - pad + (int)kx;
        v[h] = ((int)tap < taps && iy >= 0 && iy < H && ix >= 0 && ix < W) ? __ldg(img + iy * W + ix) : make_uint2(0, 0);
    }
    out[(long long)blockIdx.y * (Wo * Kp8) + idx] = make_uint4(v[0].x, v[0].y, v[1].x, v[1].y);
}

// Packed bf16 copies of a convolution filter [Cout, Cin, kh, kw] (fp32 parameter layout), one launch per layer:
//   wp [Cout, Kp]           forward / implicit-GEMM operand, K order (ky, kx, ci) with ci padded to cin_pad, zero tail
//   wd [Cin, kh*kw*Cout]    data-gradient operand: taps flipped, (ky', kx', co) order           (optional)
__global__ void pack_conv_filters_kernel(const float* __restrict__ w, int Cout, int Cin, int kh, int kw, int cin_pad, int Kp,
                                         __nv_bfloat16* __restrict__ wp, __nv_bfloat16* __restrict__ wd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int taps = kh * kw;
    const long long n_wp = (long long)Cout * Kp, n_wd = wd ? (long long)Cin * taps * Cout : 0;
    if (i < n_wp) {
        const int co = (int)(i / Kp), k = (int)(i - (long long)co * Kp);
        const int tap = k / cin_pad, ci = k - tap * cin_pad;
        float v = 0.0f;
        if (tap < taps && ci < Cin) v = w[((size_t)co * Cin + ci) * taps + tap];
        wp[i] = __float2bfloat16(v);
    } else if (i < n_wp + n_wd) {
        const long long j = i - n_wp;
        const int co = (int)(j % Cout);
        const long long r = j / Cout;
        const int tap = (int)(r % taps), ci = (int)(r / taps);
        wd[j] = __float2bfloat16(w[((size_t)co * Cin + ci) * taps + (taps - 1 - tap)]);  // flip(ky) & flip(kx) = reversed tap
    }
}

// idx (optional, training): uint8 per output element = ky*3 + kx of the FIRST maximum in scan order (torch's tie-break),
// consumed by ab_maxpool3x3s2_bwd.
// AFFINE: the input is a RAW convolution output and every tap goes through y = bf16(relu(raw * scale[c] + shift[c])) first
// (the training-mode BatchNorm + ReLU of the stem, the same fp32 operations and bf16 rounding as ab_bn_apply): the
// normalised activation -- 268 MB at batch 128 -- is never written or read.
template <bool AFFINE>
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, int B, int H, int W, int C8, int Ho, int Wo,
                                    uint4* __restrict__ out, uint2* __restrict__ idx, const float* __restrict__ scale,
                                    const float* __restrict__ shift) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)B * Ho * Wo * C8;
    if (i >= total) return;
    const int c8 = (int)(i % C8);
    long long t = i / C8;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[8];
    uint32_t tap[8];
    float sc[8], sh[8];
    if (AFFINE) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + 8 * c8 + j); sh[j] = __ldg(shift + 8 * c8 + j); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; tap[j] = 4; }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
            if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
            const uint4 v = __ldg(in + (((long long)b * H + iy) * W + ix) * C8 + c8);
            const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float2 f = __bfloat1622float2(pv[j]);
                if (AFFINE) {   // relu(fma) rounded to bf16, as the separate BatchNorm pass stores it
                    f = __bfloat1622float2(__floats2bfloat162_rn(fmaxf(fmaf(f.x, sc[2 * j], sh[2 * j]), 0.0f),
                                                                 fmaxf(fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]), 0.0f)));
                }
                if (f.x > best[2 * j]) { best[2 * j] = f.x; tap[2 * j] = ky * 3 + kx; }
                if (f.y > best[2 * j + 1]) { best[2 * j + 1] = f.y; tap[2 * j + 1] = ky * 3 + kx; }
            }
        }
    uint4 o;
    __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) po[j] = __floats2bfloat162_rn(best[2 * j], best[2 * j + 1]);  // exact: values are bf16
    out[i] = o;
    if (idx) idx[i] = make_uint2(tap[0] | (tap[1] << 8) | (tap[2] << 16) | (tap[3] << 24),
                                 tap[4] | (tap[5] << 8) | (tap[6] << 16) | (tap[7] << 24));
}

// [B, HW, C] bf16 -> mean over HW: fp32 [B, C] and bf16 [B, C].  One thread per (b, c); consecutive threads read
// consecutive channels.
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ in, int B, int HW, int C, float* __restrict__ out_f32,
                               __nv_bfloat16* __restrict__ out_bf16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    float s = 0.0f;
    for (int p = 0; p < HW; ++p) s += __bfloat162float(in[((long long)b * HW + p) * C + c]);
    s /= (float)HW;
    if (out_f32) out_f32[i] = s;
    if (out_bf16) out_bf16[i] = __float2bfloat16(s);
}

// ConvTranspose2d(k=4, s=2, p=1): oy = 2*iy - 1 + ky.  ycol [B*H*W, 16*Cout] fp32 with column order (ky, kx, co).
// One thread per (b, oy, ox, 4 output channels): sums its 4 contributing taps, applies scale/bias (+ReLU), writes bf16.
__global__ void deconv_col2im_kernel(const float4* __restrict__ ycol, int B, int H, int W, int Cout4,
                                     const float* __restrict__ scale, const float* __restrict__ bias, int relu,
                                     __nv_bfloat16* __restrict__ out, float* __restrict__ out_raw) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int Ho = 2 * H, Wo = 2 * W;
    const long long total = (long long)B * Ho * Wo * Cout4;
    if (i >= total) return;
    const int c4 = (int)(i % Cout4);
    long long t = i / Cout4;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int iy0 = (oy + 1) >> 1, ix0 = (ox + 1) >> 1;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
        const int iy = iy0 - dy, ky = oy + 1 - 2 * iy;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int ix = ix0 - dx, kx = ox + 1 - 2 * ix;
            if (ix < 0 || ix >= W) continue;
            const float4 v = __ldg(ycol + ((((long long)b * H + iy) * W + ix) * 16 + (ky * 4 + kx)) * Cout4 + c4);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    if (out_raw) reinterpret_cast<float4*>(out_raw)[i] = acc;
    if (out) {
        const int c = 4 * c4;
        float y[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            y[j] = fmaf(y[j], scale ? scale[c + j] : 1.0f, bias ? bias[c + j] : 0.0f);
            if (relu) y[j] = fmaxf(y[j], 0.0f);
        }
        __nv_bfloat162* o = reinterpret_cast<__nv_bfloat162*>(out + i * 4);
        o[0] = __floats2bfloat162_rn(y[0], y[1]);
        o[1] = __floats2bfloat162_rn(y[2], y[3]);
    }
}

// logits fp32 [B, H*W, ncls*D] (channel = cls*D + d, i.e. NHWC of the reference's [B, ncls*D, H, W]).
// One CTA per (b, cls).  p = softmax over (d,h,w); confd = max p; p /= (sum p + 1e-7); u = sum p*w/W, v = sum p*h/H,
// d = sum p*d/D.  Two sweeps over the class's D*H*W logits (max, then exp-sums); they stay in L1/L2.
constexpr int kDecodeThreads = 256;
__global__ void __launch_bounds__(kDecodeThreads)
head_decode_kernel(const float* __restrict__ logits, int ncls, int D, int H, int W, float* __restrict__ kp3d,
                   float* __restrict__ confd) {
    __shared__ float red[4][kDecodeThreads / 32];
    const int b = blockIdx.x / ncls, cls = blockIdx.x % ncls;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int HW = H * W, n = HW * D, ldc = ncls * D;
    const float* base = logits + (long long)b * HW * ldc + cls * D;
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
    float s = 0.f, su = 0.f, sv = 0.f, sd = 0.f;
    for (int i = tid; i < n; i += kDecodeThreads) {
        const int p = i / D, d = i - p * D;
        const int h = p / W, w = p - h * W;
        const float e = expf(base[(long long)p * ldc + d] - mx);
        s += e; su += e * (float)w; sv += e * (float)h; sd += e * (float)d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        su += __shfl_xor_sync(0xffffffffu, su, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sd += __shfl_xor_sync(0xffffffffu, sd, o);
    }
    if (lane == 0) { red[0][wid] = s; red[1][wid] = su; red[2][wid] = sv; red[3][wid] = sd; }
    __syncthreads();
    if (tid == 0) {
        float S = 0.f, SU = 0.f, SV = 0.f, SD = 0.f;
        for (int w = 0; w < kDecodeThreads / 32; ++w) { S += red[0][w]; SU += red[1][w]; SV += red[2][w]; SD += red[3][w]; }
        // softmax p_i = e_i / S: sum p = 1 up to rounding; the reference divides by (sum p + 1e-7) once more
        const float inv = 1.0f / S;
        const float renorm = 1.0f / (1.0f + 1e-7f);
        float* o = kp3d + ((long long)b * ncls + cls) * 3;
        o[0] = SU * inv * renorm / (float)W;
        o[1] = SV * inv * renorm / (float)H;
        o[2] = SD * inv * renorm / (float)D;
        confd[(long long)b * ncls + cls] = inv;  // max p = exp(0) / S
    }
}

// The same decode in ONE sweep for D % 4 == 0 (the shipped head: D = 28): 16-byte loads of four consecutive depth bins,
// running maximum with rescaling (online softmax), four loads in flight per thread.  Also leaves lse = max + log(sum exp)
// per (b, cls), with which the backward pass needs a single sweep as well (ab_head_decode_bwd).
struct DecodeAcc { float m, s, su, sv, sd; };
__device__ __forceinline__ void acc_rescale(DecodeAcc& a, float m_new) {
    if (m_new > a.m) {
        const float sc = __expf(a.m - m_new);  // exp(-inf) = 0 on the first element
        a.s *= sc; a.su *= sc; a.sv *= sc; a.sd *= sc;
        a.m = m_new;
    }
}
__device__ __forceinline__ void acc_merge(DecodeAcc& a, const DecodeAcc& b) {
    const float m = fmaxf(a.m, b.m);
    const float fa = a.m == -INFINITY ? 0.0f : __expf(a.m - m), fb = b.m == -INFINITY ? 0.0f : __expf(b.m - m);
    a.s = a.s * fa + b.s * fb; a.su = a.su * fa + b.su * fb; a.sv = a.sv * fa + b.sv * fb; a.sd = a.sd * fa + b.sd * fb;
    a.m = m;
}

__global__ void __launch_bounds__(kDecodeThreads)
head_decode_online_kernel(const float* __restrict__ logits, int ncls, int D, int H, int W, float* __restrict__ kp3d,
                          float* __restrict__ confd, float* __restrict__ lse) {
    __shared__ DecodeAcc red[kDecodeThreads / 32];
    const int b = blockIdx.x / ncls, cls = blockIdx.x % ncls;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int HW = H * W, D4 = D / 4, n4 = HW * D4, ldc = ncls * D;
    const float* base = logits + (long long)b * HW * ldc + cls * D;
    DecodeAcc a = {-INFINITY, 0.f, 0.f, 0.f, 0.f};
    constexpr int U = 4;
    const int dq = kDecodeThreads % D4, dp = kDecodeThreads / D4;
    int p = tid / D4, q = tid - p * D4;
    for (int i0 = tid; i0 < n4; i0 += U * kDecodeThreads) {
        float4 v[U];
        int pp[U], qq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            pp[u] = p; qq[u] = q;
            v[u] = i0 + u * kDecodeThreads < n4 ? __ldg(reinterpret_cast<const float4*>(base + (long long)p * ldc + 4 * q))
                                                : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            q += dq; p += dp;
            if (q >= D4) { q -= D4; ++p; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + u * kDecodeThreads >= n4) break;
            const int h = pp[u] / W, w = pp[u] - h * W;
            const float d0 = (float)(4 * qq[u]);
            acc_rescale(a, fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w)));
            const float e0 = __expf(v[u].x - a.m), e1 = __expf(v[u].y - a.m), e2 = __expf(v[u].z - a.m), e3 = __expf(v[u].w - a.m);
            const float es = (e0 + e1) + (e2 + e3);
            a.s += es; a.su += es * (float)w; a.sv += es * (float)h;
            a.sd += es * d0 + (e1 + 2.0f * e2 + 3.0f * e3);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        DecodeAcc bb;
        bb.m = __shfl_xor_sync(0xffffffffu, a.m, o); bb.s = __shfl_xor_sync(0xffffffffu, a.s, o);
        bb.su = __shfl_xor_sync(0xffffffffu, a.su, o); bb.sv = __shfl_xor_sync(0xffffffffu, a.sv, o);
        bb.sd = __shfl_xor_sync(0xffffffffu, a.sd, o);
        // merge in lane order so both partners compute the same bits
        if (lane & o) { DecodeAcc t = bb; acc_merge(t, a); a = t; } else acc_merge(a, bb);
    }
    if (lane == 0) red[wid] = a;
    __syncthreads();
    if (tid == 0) {
        DecodeAcc t = red[0];
        for (int w = 1; w < kDecodeThreads / 32; ++w) acc_merge(t, red[w]);
        const float inv = 1.0f / t.s;
        const float renorm = 1.0f / (1.0f + 1e-7f);
        float* o = kp3d + ((long long)b * ncls + cls) * 3;
        o[0] = t.su * inv * renorm / (float)W;
        o[1] = t.sv * inv * renorm / (float)H;
        o[2] = t.sd * inv * renorm / (float)D;
        confd[(long long)b * ncls + cls] = inv;
        if (lse) lse[(long long)b * ncls + cls] = t.m + logf(t.s);
    }
}

static inline unsigned blocks_for(long long n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace ab

using namespace ab;

extern "C" int ab_image_to_nhwc(const float* image, int B, int C, int H, int W, int Cp, void* out, void* stream) {
    AB_REQUIRE(B >= 0 && C > 0 && H > 0 && W > 0 && Cp >= C, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(image && out, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    image_to_nhwc_kernel<<<blocks_for((long long)B * H * W, 256), 256, 0, st>>>(image, B, C, H, W, Cp, (__nv_bfloat16*)out);
    count_launch();
    return check_launch("image_to_nhwc_kernel");
}

extern "C" int ab_im2col_nhwc(const void* in, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int Kp,
                              void* out, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "bad shape");
    AB_REQUIRE(Kp >= kh * kw * C && Kp % 8 == 0, "Kp must be a multiple of 8 and >= kh*kw*C");
    if (B == 0) return AB_OK;
    AB_REQUIRE(in && out, "null pointer");
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    AB_REQUIRE(Ho > 0 && Wo > 0, "empty output");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_IM2COL, st);
    if (C % 8 == 0) {
        AB_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "16-byte alignment required");
        const long long total = (long long)B * Ho * Wo * (Kp / 8);
        im2col_vec8_kernel<<<blocks_for(total, 256), 256, 0, st>>>((const uint4*)in, B, H, W, C / 8, kh, kw, stride, pad, Ho,
                                                                   Wo, Kp / 8, (uint4*)out);
    } else if (C == 4 && Kp % 4 == 0) {
        AB_REQUIRE(((uintptr_t)in & 7) == 0 && ((uintptr_t)out & 7) == 0, "8-byte alignment required");
        static const bool pairs = !(getenv("AB_IM2COL_PAIRS") && atoi(getenv("AB_IM2COL_PAIRS")) == 0);
        // multiply-shift division: floor(n / d) == (n * ceil(2^16 / d)) >> 16 for n * (d - 1) < 2^16 ... checked below
        const int Kp8 = Kp / 8;
        auto magic_ok = [](unsigned d, unsigned n_max) {
            const unsigned m = (65536u + d - 1) / d;
            for (unsigned n = 0; n <= n_max; ++n) if (((n * m) >> 16) != n / d) return false;
            return true;
        };
        static int ok_kp8 = -1, ok_for_kp8 = 0, ok_for_wo = 0, ok_for_kw = 0;
        if (ok_kp8 < 0 || ok_for_kp8 != Kp8 || ok_for_wo != Wo || ok_for_kw != kw) {
            ok_kp8 = (Wo * Kp8 < 65536 && (long long)B * Ho <= 65535 && magic_ok((unsigned)Kp8, (unsigned)(Wo * Kp8)) &&
                      magic_ok((unsigned)kw, (unsigned)(2 * Kp8 + 1))) ? 1 : 0;
            ok_for_kp8 = Kp8; ok_for_wo = Wo; ok_for_kw = kw;
        }
        if (pairs && Kp % 8 == 0 && ((uintptr_t)out & 15) == 0 && ok_kp8 == 1) {
            const dim3 grid((unsigned)((Wo * Kp8 + 255) / 256), (unsigned)(B * Ho));
            im2col_c4_pairs_kernel<<<grid, 256, 0, st>>>((const uint2*)in, H, W, kh, kw, stride, pad, Ho, Wo, Kp8,
                                                         (65536u + Kp8 - 1) / Kp8, (65536u + kw - 1) / kw, (uint4*)out);
        } else {
            const long long total = (long long)B * Ho * Wo * kh;
            im2col_c4_kernel<<<blocks_for(total, 256), 256, 0, st>>>((const uint2*)in, B, H, W, kh, kw, stride, pad, Ho, Wo, Kp / 4,
                                                                     (uint2*)out);
        }
    } else {
        const long long total = (long long)B * Ho * Wo * Kp;
        im2col_scalar_kernel<<<blocks_for(total, 256), 256, 0, st>>>((const __nv_bfloat16*)in, B, H, W, C, kh, kw, stride,
                                                                     pad, Ho, Wo, Kp, (__nv_bfloat16*)out);
    }
    count_launch();
    return check_launch("im2col_kernel");
}

extern "C" int ab_pack_conv_filters(const float* w, int Cout, int Cin, int kh, int kw, int cin_pad, int Kp, void* wp, void* wd,
                                    void* stream) {
    AB_REQUIRE(Cout > 0 && Cin > 0 && kh > 0 && kw > 0 && cin_pad >= Cin && Kp >= kh * kw * cin_pad, "bad shape");
    AB_REQUIRE(w && wp, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    const long long n = (long long)Cout * Kp + (wd ? (long long)Cin * kh * kw * Cout : 0);
    pack_conv_filters_kernel<<<blocks_for(n, 256), 256, 0, st>>>(w, Cout, Cin, kh, kw, cin_pad, Kp, (__nv_bfloat16*)wp,
                                                                 (__nv_bfloat16*)wd);
    count_launch();
    return check_launch("pack_conv_filters_kernel");
}

extern "C" int ab_maxpool3x3s2_nhwc(const void* in, int B, int H, int W, int C, void* out, void* idx, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape (C must be a multiple of 8)");
    if (B == 0) return AB_OK;
    AB_REQUIRE(in && out, "null pointer");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    maxpool3x3s2_kernel<false><<<blocks_for((long long)B * Ho * Wo * (C / 8), 256), 256, 0, st>>>((const uint4*)in, B, H, W, C / 8,
                                                                                                  Ho, Wo, (uint4*)out, (uint2*)idx, nullptr, nullptr);
    count_launch();
    return check_launch("maxpool3x3s2_kernel");
}

extern "C" int ab_maxpool3x3s2_affine_nhwc(const void* raw, int B, int H, int W, int C, const float* scale, const float* shift,
                                           void* out, void* idx, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "bad shape (C must be a multiple of 8)");
    if (B == 0) return AB_OK;
    AB_REQUIRE(raw && out && scale && shift, "null pointer");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    maxpool3x3s2_kernel<true><<<blocks_for((long long)B * Ho * Wo * (C / 8), 256), 256, 0, st>>>((const uint4*)raw, B, H, W, C / 8,
                                                                                                 Ho, Wo, (uint4*)out, (uint2*)idx, scale, shift);
    count_launch();
    return check_launch("maxpool3x3s2_kernel");
}

extern "C" int ab_avgpool_nhwc(const void* in, int B, int HW, int C, float* out_f32, void* out_bf16, void* stream) {
    AB_REQUIRE(B >= 0 && HW > 0 && C > 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(in && (out_f32 || out_bf16), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    avgpool_kernel<<<blocks_for((long long)B * C, 128), 128, 0, st>>>((const __nv_bfloat16*)in, B, HW, C, out_f32,
                                                                      (__nv_bfloat16*)out_bf16);
    count_launch();
    return check_launch("avgpool_kernel");
}

extern "C" int ab_deconv4x4s2_col2im(const float* ycol, int B, int H, int W, int Cout, const float* scale,
                                     const float* bias, int relu, void* out_bf16, float* out_raw, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && Cout > 0 && Cout % 4 == 0, "bad shape (Cout must be a multiple of 4)");
    if (B == 0) return AB_OK;
    AB_REQUIRE(ycol && (out_bf16 || out_raw), "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_ELEMENTWISE, st);
    const long long total = (long long)B * 4 * H * W * (Cout / 4);
    deconv_col2im_kernel<<<blocks_for(total, 256), 256, 0, st>>>((const float4*)ycol, B, H, W, Cout / 4, scale, bias, relu,
                                                                 (__nv_bfloat16*)out_bf16, out_raw);
    count_launch();
    return check_launch("deconv_col2im_kernel");
}

extern "C" int ab_head_decode(const float* logits, int B, int ncls, int D, int H, int W, float* kp3d, float* confd,
                              float* lse, void* stream) {
    AB_REQUIRE(B >= 0 && ncls > 0 && D > 0 && H > 0 && W > 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(logits && kp3d && confd, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    StageTimer tm(AB_STAGE_HEAD_DECODE, st);
    AB_REQUIRE(!lse || D % 4 == 0, "lse is produced by the vector path: D must be a multiple of 4");
    if (D % 4 == 0 && ((uintptr_t)logits & 15) == 0 && D / 4 <= kDecodeThreads)
        head_decode_online_kernel<<<B * ncls, kDecodeThreads, 0, st>>>(logits, ncls, D, H, W, kp3d, confd, lse);
    else
        head_decode_kernel<<<B * ncls, kDecodeThreads, 0, st>>>(logits, ncls, D, H, W, kp3d, confd);
    count_launch();
    return check_launch("head_decode_kernel");
}
