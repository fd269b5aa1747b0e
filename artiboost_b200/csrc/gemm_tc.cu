// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   D[M,N] = epilogue( A[M,K] . B[N,K]^T )          A, B bf16 row-major with K contiguous ("TN"), fp32 accumulate
//   epilogue: y = acc * scale[n] + bias[n] (+ residual[m,n]) -> optional ReLU -> bf16 or fp32 store
//             (folded BatchNorm / conv bias / residual add / ReLU of anakin/models/resnet.py:72-152 fused into
//              the producing GEMM), optional per-column sum / sum-of-squares for training-mode BatchNorm.
//
// This is the contraction behind every convolution, transposed convolution and linear layer of the clasbased
// network: conv = im2col rows x filter matrix, see conv.cu.
//
// Structure (one 128 x BN output tile per CTA, 2 CTAs co-resident per SM so one tile's epilogue overlaps the
// other's main loop):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D tiles of A (128 x 64) and B (BN x 64), 128-byte swizzle, into a
//            STAGES-deep shared-memory ring; completion by mbarrier transaction bytes.  Out-of-range rows / K tail
//            are zero-filled by the TMA unit, so M, N, K need no padding (K only a multiple of 8 for the 16-byte pitch).
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) four
//            times per stage from shared-memory descriptors; tcgen05.commit releases the stage / signals the epilogue.
//   warp 2   TMEM allocator (BN fp32 columns x 128 lanes).
//   warps 4-7 epilogue: tcgen05.ld 32x32b (one accumulator row per thread), fused math, 16-byte global stores.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <atomic>

#include "common.cuh"

namespace ab {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 256;

struct GemmEpilogue {
    void* D;                 // bf16 or fp32 [M, ldd]
    long long ldd;
    int out_fp32;
    const float* scale;      // [N] or null (1)
    const float* bias;       // [N] or null (0)
    const __nv_bfloat16* residual;  // [M, ldr] or null
    long long ldr;
    int relu;
    float* col_sum;          // [ceil(M/128), N] or null: per-row-tile column sums of the PRE-epilogue accumulator
    float* col_sumsq;        // [ceil(M/128), N] or null   (plain stores, no atomics: deterministic; see ab_bn_finalize)
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_im2col_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c, int w, int h, int n,
                                                   uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// Implicit-GEMM geometry: the A operand is the NHWC activation itself, fetched by TMA in im2col mode -- one
// (filter tap, 64-channel slice) per k-block, 128 consecutive output pixels per tile, padding zero-filled by the
// TMA unit.  K index of the packed filter matrix = tap * C + channel (pack_conv_weight order).
struct ConvGeom {
    int Ho, Wo, stride, pad, kw, cblocks;  // cblocks = C / 64
};

// K-major operand tile in shared memory, rows of 64 bf16 (128 B), 128-byte swizzle, 8-row atoms of 1024 B:
// start address >> 4 | LBO (ignored for swizzled K-major) = 1 | SBO = 1024 B >> 4 | version 1 (Blackwell) | SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_k_sw128(const void* smem_tile) {
    return (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN, int STAGES>
struct GemmSmem {
    __nv_bfloat16 a[STAGES][kBM * kBK];
    __nv_bfloat16 b[STAGES][BN * kBK];
    uint64_t full[STAGES], empty[STAGES], tmem_full;
    uint32_t tmem_base;
    float stat[2][4][BN];    // per-epilogue-warp column sums / sums of squares of this tile
};

// Column sums of a 32 x 16 fragment held one row per lane: butterfly that halves the columns a lane owns at every
// step (8 + 4 + 2 + 1 + 1 shuffles instead of 16 x 5).  The result is the sum of column
// 8*bit4(lane) + 4*bit3(lane) + 2*bit2(lane) + bit1(lane), present on both lanes of each even/odd pair.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2];
    bool hi = lane & 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (hi ? v[j + 8] : v[j]) + __shfl_xor_sync(0xffffffffu, hi ? v[j] : v[j + 8], 16);
    hi = lane & 8;
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = (hi ? a[j + 4] : a[j]) + __shfl_xor_sync(0xffffffffu, hi ? a[j] : a[j + 4], 8);
    hi = lane & 4;
#pragma unroll
    for (int j = 0; j < 2; ++j) c[j] = (hi ? b[j + 2] : b[j]) + __shfl_xor_sync(0xffffffffu, hi ? b[j] : b[j + 2], 4);
    hi = lane & 2;
    float d = (hi ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, hi ? c[0] : c[1], 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&w)[8]) {   // 256-bit store (PTX ISA 8.8, sm_100)
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}
__device__ __forceinline__ void ld_global_v8(const void* p, uint32_t (&w)[8]) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
}

// 32-byte row stores need 32-byte aligned rows: base and pitch of D (and of the residual)
__device__ __forceinline__ bool epilogue_wide_ok(const GemmEpilogue& ep) {
    const int esz = ep.out_fp32 ? 4 : 2;
    return (((uintptr_t)ep.D | (uintptr_t)(ep.ldd * esz)) & 31) == 0 &&
           (!ep.residual || (((uintptr_t)ep.residual | (uintptr_t)(ep.ldr * 2)) & 31) == 0);
}

// Fused epilogue of one accumulator fragment (one output row, 16 consecutive columns from `col`) held by a lane:
// scale / bias / residual / ReLU, bf16 or fp32 store.  When the whole fragment is inside N and the rows are 32-byte
// aligned the lane writes its 16 columns with ONE 32-byte store (st.global.v8.b32): a lane's two 16-byte stores went to
// the same sector of 32 different lines, i.e. two passes of the store path per line (1x1 convolution 64 -> 256 at
// 64 x 64, batch 128: 147 -> 97 us).  Otherwise two groups of 8 columns (N is a multiple of 8).
__device__ __forceinline__ void epilogue_frag16(const GemmEpilogue& ep, bool wide, long long row, int col, int N, const float (&v)[16]) {
    if (wide && col + 16 <= N) {
        float y[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float sc = ep.scale ? __ldg(ep.scale + col + j) : 1.0f;
            const float bi = ep.bias ? __ldg(ep.bias + col + j) : 0.0f;
            y[j] = fmaf(v[j], sc, bi);
        }
        if (ep.residual) {
            uint32_t rr[8];
            ld_global_v8(ep.residual + (size_t)row * ep.ldr + col, rr);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&rr[j]));
                y[2 * j] += f.x; y[2 * j + 1] += f.y;
            }
        }
        if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) y[j] = fmaxf(y[j], 0.0f);
        }
        if (ep.out_fp32) {
            float* o = (float*)ep.D + (size_t)row * ep.ldd + col;
            uint32_t w0[8], w1[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { w0[j] = __float_as_uint(y[j]); w1[j] = __float_as_uint(y[8 + j]); }
            st_global_v8(o, w0);
            st_global_v8(o + 8, w1);
        } else {
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
                pk[j] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            st_global_v8((__nv_bfloat16*)ep.D + (size_t)row * ep.ldd + col, pk);
        }
        return;
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cc = col + 8 * h;
        if (cc >= N) break;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sc = ep.scale ? __ldg(ep.scale + cc + j) : 1.0f;
            const float bi = ep.bias ? __ldg(ep.bias + cc + j) : 0.0f;
            y[j] = fmaf(v[8 * h + j], sc, bi);
        }
        if (ep.residual) {
            const uint4 rr = *reinterpret_cast<const uint4*>(ep.residual + (size_t)row * ep.ldr + cc);
            const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __bfloat1622float2(rp[j]);
                y[2 * j] += f.x; y[2 * j + 1] += f.y;
            }
        }
        if (ep.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.0f);
        }
        if (ep.out_fp32) {
            float4* o = reinterpret_cast<float4*>((float*)ep.D + (size_t)row * ep.ldd + cc);
            o[0] = make_float4(y[0], y[1], y[2], y[3]);
            o[1] = make_float4(y[4], y[5], y[6], y[7]);
        } else {
            uint4 pk;
            __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
            for (int j = 0; j < 4; ++j) pp[j] = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
            *reinterpret_cast<uint4*>((__nv_bfloat16*)ep.D + (size_t)row * ep.ldd + cc) = pk;
        }
    }
}

template <int BN, int STAGES, bool IM2COL>
__global__ void __launch_bounds__(kGemmThreads)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K,
                    const GemmEpilogue ep, const ConvGeom cg) {
    extern __shared__ uint8_t smem_raw[];
    auto& sm = *reinterpret_cast<GemmSmem<BN, STAGES>*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int num_k = (K + kBK - 1) / kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int base_w = 0, base_h = 0, img = 0;
            if (IM2COL) {  // first output pixel of this tile -> top-left input pixel of its filter window
                const int m0 = tile_m * kBM, per = cg.Ho * cg.Wo;
                img = m0 / per;
                const int r = m0 - img * per, oy = r / cg.Wo, ox = r - oy * cg.Wo;
                base_w = ox * cg.stride - cg.pad;
                base_h = oy * cg.stride - cg.pad;
            }
            int cb = 0, kx = 0, ky = 0;   // (channel block, tap) of the k-block, advanced without divisions
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&sm.empty[s], ph ^ 1);
                mbar_expect_tx(&sm.full[s], (kBM + BN) * kBK * 2);
                if (IM2COL) {
                    tma_load_im2col_4d(sm.a[s], &tmA, &sm.full[s], cb * kBK, base_w, base_h, img, (uint16_t)kx, (uint16_t)ky);
                    if (++cb == cg.cblocks) { cb = 0; if (++kx == cg.kw) { kx = 0; ++ky; } }
                } else {
                    tma_load_2d(sm.a[s], &tmA, &sm.full[s], kb * kBK, tile_m * kBM);
                }
                tma_load_2d(sm.b[s], &tmB, &sm.full[s], kb * kBK, tile_n * BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(&sm.full[s], ph);
                tc_fence_after();
                const uint64_t ad = umma_desc_k_sw128(sm.a[s]), bd = umma_desc_k_sw128(sm.b[s]);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)  // +32 bytes (16 bf16) along K inside the swizzled row: +2 in the address field
                    umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                umma_commit(&sm.empty[s]);  // frees the stage when these MMAs have read it
            }
            umma_commit(&sm.tmem_full);
        }
    } else if (warp >= 4) {
        mbar_wait(&sm.tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const bool wide = epilogue_wide_ok(ep);
        const int row = tile_m * kBM + q * 32 + lane;
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            const int col = tile_n * BN + c0;
            if (col >= N) break;  // warp-uniform
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (ep.col_sum) {  // training-mode BatchNorm statistics of the raw convolution output
                float s1[16], s2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { s1[j] = row_ok ? v[j] : 0.0f; s2[j] = s1[j] * s1[j]; }
                const float t1 = warp_colsum16(s1, lane), t2 = warp_colsum16(s2, lane);
                if (!(lane & 1)) {
                    const int cj = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    sm.stat[0][q][cj] = t1;
                    sm.stat[1][q][cj] = t2;
                }
            }
            if (!row_ok) continue;
            epilogue_frag16(ep, wide, row, col, N, v);
        }
        if (ep.col_sum) {  // combine the four epilogue warps, one plain store per (row tile, column)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = threadIdx.x - 128; c < BN; c += 128) {
                const int col = tile_n * BN + c;
                if (col < N) {
                    ep.col_sum[(size_t)tile_m * N + col] = (sm.stat[0][0][c] + sm.stat[0][1][c]) + (sm.stat[0][2][c] + sm.stat[0][3][c]);
                    ep.col_sumsq[(size_t)tile_m * N + col] = (sm.stat[1][0][c] + sm.stat[1][1][c]) + (sm.stat[1][2][c] + sm.stat[1][3][c]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, BN);
}


// ------------------------------------------------------------------------------------------ persistent form
// The same contraction with a static tile scheduler: gridDim.x CTAs (two per SM) walk the output tiles t = blockIdx.x,
// blockIdx.x + gridDim.x, ... (tile_n fastest, so the CTAs that share an A tile run at the same time).  The TMA producer and
// the MMA issuer run ahead across tile boundaries through the same shared-memory ring; the accumulator is double buffered
// in TMEM (2 x BN columns) and each buffer has its own epilogue group of four warps, so the TMEM loads / row-strided
// stores of tile i overlap the loads and MMAs of tile i + 1.  What this buys: the one-tile-per-CTA kernel pays barrier
// set-up, TMEM allocation, the first load's latency and the whole epilogue once per tile with nothing of its own to
// overlap them -- for the short-K contractions (1x1 convolutions of ResNet-50: K = 64 ... 512, one to eight k-blocks)
// that is most of a CTA's life.
constexpr int kGemmPThreads = 128 + 2 * 128;   // warps 0-3: TMA / MMA / TMEM roles; warps 4-7 and 8-11: the two epilogue groups

template <int BN, int STAGES>
struct GemmPSmem {
    __nv_bfloat16 a[STAGES][kBM * kBK];
    __nv_bfloat16 b[STAGES][BN * kBK];
    uint64_t full[STAGES], empty[STAGES], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    float stat[2][2][4][BN];   // [epilogue group][sum, sum of squares][warp][column]
};

__device__ __forceinline__ void mbar_arrive1(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int BN, int STAGES, bool IM2COL>
__global__ void __launch_bounds__(kGemmPThreads)
gemm_bf16_tn_persistent_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N,
                               int K, const GemmEpilogue ep, const ConvGeom cg, int tiles_n, int n_tiles) {
    extern __shared__ uint8_t smem_raw[];
    auto& sm = *reinterpret_cast<GemmPSmem<BN, STAGES>*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_k = (K + kBK - 1) / kBK;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], 4); }   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            int kbg = 0;   // k-blocks issued so far, over all tiles: ring position
            for (int it = 0; it < my_tiles; ++it) {
                const int t = (int)blockIdx.x + it * (int)gridDim.x;
                const int tile_m = t / tiles_n, tile_n = t - tile_m * tiles_n;
                int base_w = 0, base_h = 0, img = 0;
                if (IM2COL) {
                    const int m0 = tile_m * kBM, per = cg.Ho * cg.Wo;
                    img = m0 / per;
                    const int r = m0 - img * per, oy = r / cg.Wo, ox = r - oy * cg.Wo;
                    base_w = ox * cg.stride - cg.pad;
                    base_h = oy * cg.stride - cg.pad;
                }
                int cb = 0, kx = 0, ky = 0;   // (channel block, tap) of the k-block, advanced without divisions
                for (int kb = 0; kb < num_k; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    mbar_wait(&sm.empty[s], ((kbg / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&sm.full[s], (kBM + BN) * kBK * 2);
                    if (IM2COL) {
                        tma_load_im2col_4d(sm.a[s], &tmA, &sm.full[s], cb * kBK, base_w, base_h, img, (uint16_t)kx, (uint16_t)ky);
                        if (++cb == cg.cblocks) { cb = 0; if (++kx == cg.kw) { kx = 0; ++ky; } }
                    } else {
                        tma_load_2d(sm.a[s], &tmA, &sm.full[s], kb * kBK, tile_m * kBM);
                    }
                    tma_load_2d(sm.b[s], &tmB, &sm.full[s], kb * kBK, tile_n * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            int kbg = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int ac = it & 1;
                mbar_wait(&sm.acc_empty[ac], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem + (uint32_t)(ac * BN);
                for (int kb = 0; kb < num_k; ++kb, ++kbg) {
                    const int s = kbg % STAGES;
                    mbar_wait(&sm.full[s], (kbg / STAGES) & 1);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_k_sw128(sm.a[s]), bd = umma_desc_k_sw128(sm.b[s]);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) umma_bf16(acc, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&sm.empty[s]);
                }
                umma_commit(&sm.acc_full[ac]);
            }
        }
    } else if (warp >= 4) {
        const int q = warp & 3, grp = (warp - 4) >> 2;
        const bool wide = epilogue_wide_ok(ep);
        for (int it = grp; it < my_tiles; it += 2) {
            const int t = (int)blockIdx.x + it * (int)gridDim.x;
            const int tile_m = t / tiles_n, tile_n = t - tile_m * tiles_n;
            const int row = tile_m * kBM + q * 32 + lane;
            const bool row_ok = row < M;
            mbar_wait(&sm.acc_full[grp], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                const int col = tile_n * BN + c0;
                if (col >= N) break;  // warp-uniform
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * BN + c0), r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (ep.col_sum) {
                    float s1[16], s2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { s1[j] = row_ok ? v[j] : 0.0f; s2[j] = s1[j] * s1[j]; }
                    const float t1 = warp_colsum16(s1, lane), t2 = warp_colsum16(s2, lane);
                    if (!(lane & 1)) {
                        const int cj = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        sm.stat[grp][0][q][cj] = t1;
                        sm.stat[grp][1][q][cj] = t2;
                    }
                }
                if (!row_ok) continue;
                epilogue_frag16(ep, wide, row, col, N, v);
            }
            // the accumulator has been read (tcgen05.wait::ld inside tmem_ld16): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive1(&sm.acc_empty[grp]);
            if (ep.col_sum) {
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                for (int c = (threadIdx.x & 127); c < BN; c += 128) {
                    const int col = tile_n * BN + c;
                    if (col < N) {
                        ep.col_sum[(size_t)tile_m * N + col] = (sm.stat[grp][0][0][c] + sm.stat[grp][0][1][c]) + (sm.stat[grp][0][2][c] + sm.stat[grp][0][3][c]);
                        ep.col_sumsq[(size_t)tile_m * N + col] = (sm.stat[grp][1][0][c] + sm.stat[grp][1][1][c]) + (sm.stat[grp][1][2][c] + sm.stat[grp][1][3][c]);
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the group's next tile rewrites the slots
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 2 * BN);
}

// --------------------------------------------------------------------------- halo-resident 3x3 stride-1 convolution
// The implicit GEMM above fetches the activation once per filter tap (TMA im2col): nine times through L2.  For the
// 3x3 / stride 1 / pad 1 convolutions (every block convolution of ResNet-18/34 and their data gradients) the nine A
// operands of a tile are ONE shared-memory window read at nine row offsets:
//   * outputs are enumerated in the width-padded raster of an image, Wp = W + 2 columns per row (the last two of a row
//     are dummies that are computed and dropped: 3 % at W = 64, 20 % at W = 8); a tile is 128 consecutive padded pixels;
//   * the input rows the tile touches are loaded ONCE per 64-channel block by a tiled 4-D TMA box [64 ch, Wp, R, 1]
//     starting at x = -1, y = y0 - 1: the zero padding of the convolution is the TMA unit's out-of-bounds fill, and the
//     box lands as R * Wp pixel rows of 128 B (SWIZZLE_128B);
//   * padded output pixel q of the tile needs, for tap (ky, kx), box row q + ky * Wp + kx: the A operand of a tap is the
//     same 128-row window shifted by ky * Wp + kx rows, i.e. the same UMMA descriptor with its start address advanced by
//     that many 128-byte rows.  tcgen05 applies the 128-byte swizzle to the absolute shared-memory address, so a start
//     address that is not 1024-byte aligned reads the TMA-written tile correctly with the base-offset field left at zero
//     (probed on B200 for shifts 0..79 with tools/umma_shift_test.py; setting the field breaks it).
// L2 -> SM traffic per tile: one activation window (42 KB at W = 64) + the filters, instead of nine 16 KB tiles + the
// filters.  Epilogue as above (scale / bias / residual / ReLU, bf16 or fp32 store, per-tile BatchNorm partial sums over
// the real pixels); partial-sum rows are per (image, tile): ab_conv_stat_rows().
struct HaloGeom {
    int H, W, Wp, R, tiles_per_img, cblocks, a_stages;
    unsigned a_bytes;   // bytes of one activation window, rounded up to 1024
    unsigned mul_wp, mul_tpi;   // exact multiply-shift division by Wp / tiles_per_img over the ranges the kernels use (halo_geom)
};

// floor(n / d) as (n * ceil(2^20 / d)) >> 20: the role threads of the persistent kernel divide once or twice per tile, and a
// runtime integer division is ~100 cycles on the one thread that issues a tile's 36 MMAs
__device__ __forceinline__ int div_wp(const HaloGeom& hg, int n) { return (int)(((unsigned long long)(unsigned)n * hg.mul_wp) >> 20); }
__device__ __forceinline__ int div_tpi(const HaloGeom& hg, int n) { return (int)(((unsigned long long)(unsigned)n * hg.mul_tpi) >> 20); }

__device__ __forceinline__ void tma_load_tile_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int kHaloBStages = 3;
template <int BN>
struct HaloSmemTail {   // behind the activation windows
    __nv_bfloat16 b[kHaloBStages][BN * kBK];
    uint64_t a_full[2], a_empty[2], b_full[kHaloBStages], b_empty[kHaloBStages], tmem_full;
    uint32_t tmem_base;
    float stat[2][4][BN];
};

template <int BN>
__global__ void __launch_bounds__(kGemmThreads)
conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, int N, int C,
                    const GemmEpilogue ep, const HaloGeom hg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    auto& sm = *reinterpret_cast<HaloSmemTail<BN>*>(base + (size_t)hg.a_stages * hg.a_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int img = blockIdx.x / hg.tiles_per_img, t = blockIdx.x - img * hg.tiles_per_img, tile_n = blockIdx.y;
    const int q_start = t * kBM, y0 = q_start / hg.Wp, q0 = q_start - y0 * hg.Wp;
    const int CB = hg.cblocks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], 1); }
        for (int s = 0; s < kHaloBStages; ++s) { mbar_init(&sm.b_full[s], 1); mbar_init(&sm.b_empty[s], 1); }
        mbar_init(&sm.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int cb = 0; cb < CB; ++cb) {
                const int as = cb % hg.a_stages;
                mbar_wait(&sm.a_empty[as], ((cb / hg.a_stages) & 1) ^ 1);
                mbar_expect_tx(&sm.a_full[as], (uint32_t)(hg.R * hg.Wp) * 128u);
                tma_load_tile_4d(base + (size_t)as * hg.a_bytes, &tmX, &sm.a_full[as], cb * kBK, -1, y0 - 1, img);
                for (int tap = 0; tap < 9; ++tap) {
                    const int kb = cb * 9 + tap, s = kb % kHaloBStages;
                    mbar_wait(&sm.b_empty[s], ((kb / kHaloBStages) & 1) ^ 1);
                    mbar_expect_tx(&sm.b_full[s], BN * kBK * 2);
                    tma_load_2d(sm.b[s], &tmB, &sm.b_full[s], tap * C + cb * kBK, tile_n * BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            for (int cb = 0; cb < CB; ++cb) {
                const int as = cb % hg.a_stages;
                mbar_wait(&sm.a_full[as], (cb / hg.a_stages) & 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(base + (size_t)as * hg.a_bytes) + (uint32_t)q0 * 128u;
                for (int tap = 0; tap < 9; ++tap) {
                    const int kb = cb * 9 + tap, s = kb % kHaloBStages;
                    const int ky = tap / 3, kx = tap - 3 * ky;
                    mbar_wait(&sm.b_full[s], (kb / kHaloBStages) & 1);
                    tc_fence_after();
                    const uint32_t start = a0 + (uint32_t)(ky * hg.Wp + kx) * 128u;   // the window, shifted by whole 128-byte rows
                    const uint64_t ad = (uint64_t)((start >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
                    const uint64_t bd = umma_desc_k_sw128(sm.b[s]);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
                    umma_commit(&sm.b_empty[s]);
                }
                umma_commit(&sm.a_empty[as]);
            }
            umma_commit(&sm.tmem_full);
        }
    } else if (warp >= 4) {
        mbar_wait(&sm.tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const bool wide = epilogue_wide_ok(ep);
        const int qq = q0 + q * 32 + lane, yl = qq / hg.Wp, xp = qq - yl * hg.Wp, y = y0 + yl;
        const bool row_ok = xp < hg.W && y < hg.H;
        const long long row = ((long long)img * hg.H + y) * hg.W + xp;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
            const int col = tile_n * BN + c0;
            if (col >= N) break;  // warp-uniform
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
            if (ep.col_sum) {
                float s1[16], s2[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) { s1[j] = row_ok ? v[j] : 0.0f; s2[j] = s1[j] * s1[j]; }
                const float t1 = warp_colsum16(s1, lane), t2 = warp_colsum16(s2, lane);
                if (!(lane & 1)) {
                    const int cj = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                    sm.stat[0][q][cj] = t1;
                    sm.stat[1][q][cj] = t2;
                }
            }
            if (!row_ok) continue;
            epilogue_frag16(ep, wide, row, col, N, v);
        }
        if (ep.col_sum) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = threadIdx.x - 128; c < BN; c += 128) {
                const int col = tile_n * BN + c;
                if (col < N) {
                    ep.col_sum[(size_t)blockIdx.x * N + col] = (sm.stat[0][0][c] + sm.stat[0][1][c]) + (sm.stat[0][2][c] + sm.stat[0][3][c]);
                    ep.col_sumsq[(size_t)blockIdx.x * N + col] = (sm.stat[1][0][c] + sm.stat[1][1][c]) + (sm.stat[1][2][c] + sm.stat[1][3][c]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, BN);
}

// Persistent form for one 64-channel block and one 64-wide output tile (C = 64, Cout <= 64: the layer1 convolutions and
// their data gradients).  One CTA per SM walks the tiles of the whole batch: the nine filter tiles (72 KB) are loaded
// ONCE and stay in shared memory, activation windows are double buffered (the TMA unit fetches tile i + 1 while tile i is
// multiplied), and so are the accumulators (2 x 64 TMEM columns, one epilogue group of four warps each: the epilogue of tile
// i runs beside the MMAs of tile i + 1).  L2 -> SM traffic per tile: the 42 KB window alone.  Measured at C = Cout = 64, 64 x 64, batch 128 (tools/time_conv.py):
// im2col ring 95.6 us, one halo tile per CTA 75.3 us, this kernel 58.6 us (659 TFLOP/s).  Probed and neutral: three windows in
// flight (60.6), four accumulators with four epilogue groups (57.9), one TMA box per input row (58.8), L2 prefetch of later
// windows (59), 1024-byte aligned operand starts (59.1), shared-memory staged coalesced stores (67.5), four interleaved
// accumulator chains per tile summed by the epilogue (61.6); without the window loads 56.9 us, without the MMAs 37.7 us.  What
// remains is ~110 cycles per 128 x 64 x 16 tcgen05.mma against 32 nominal -- this pool's cuBLAS sustains 62 % of the nominal
// bf16 rate (MEASURED_PEAKS.json), i.e. ~52 cycles for this shape, so the kernel stands at 47 % of the measured peak.
constexpr int kHaloPA = 2;   // activation windows in flight (2 x 42 KB + 72 KB of filters at W = 64)
constexpr int kHaloAcc = 2;  // accumulators (64 TMEM columns each), each drained by its own epilogue group
struct HaloPersistSmemTail {
    __nv_bfloat16 b[9][64 * kBK];
    uint64_t b_full, a_full[kHaloPA], a_empty[kHaloPA], acc_full[kHaloAcc], acc_empty[kHaloAcc];
    uint32_t tmem_base;
    float stat[kHaloAcc][2][4][64];   // [epilogue group][sum, sum of squares][warp][column]
};
constexpr int kHaloPThreads = 128 + 128 * kHaloAcc;   // warps 0-3: TMA / MMA / TMEM roles; then one epilogue group of 4 warps per accumulator

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kHaloPThreads)
conv3x3_halo_persistent_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmB, int N,
                               int n_tiles, const GemmEpilogue ep, const HaloGeom hg) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    auto& sm = *reinterpret_cast<HaloPersistSmemTail*>(base + kHaloPA * (size_t)hg.a_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        mbar_init(&sm.b_full, 1);
        for (int s = 0; s < kHaloPA; ++s) { mbar_init(&sm.a_full[s], 1); mbar_init(&sm.a_empty[s], 1); }
        for (int s = 0; s < kHaloAcc; ++s) { mbar_init(&sm.acc_full[s], 1); mbar_init(&sm.acc_empty[s], 4); }   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, 64 * kHaloAcc);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&sm.b_full, 9 * 64 * kBK * 2);
            for (int tap = 0; tap < 9; ++tap) tma_load_2d(sm.b[tap], &tmB, &sm.b_full, tap * kBK, 0);   // C = 64: K offset tap * 64
            for (int it = 0; it < my_tiles; ++it) {
                const int g = (int)blockIdx.x + it * (int)gridDim.x;
                const int img = div_tpi(hg, g), t = g - img * hg.tiles_per_img;
                const int y0 = div_wp(hg, t * kBM), s = it % kHaloPA;
                mbar_wait(&sm.a_empty[s], ((it / kHaloPA) & 1) ^ 1);
                mbar_expect_tx(&sm.a_full[s], (uint32_t)(hg.R * hg.Wp) * 128u);
                tma_load_tile_4d(base + (size_t)s * hg.a_bytes, &tmX, &sm.a_full[s], 0, -1, y0 - 1, img);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
            mbar_wait(&sm.b_full, 0);
            for (int it = 0; it < my_tiles; ++it) {
                const int g = (int)blockIdx.x + it * (int)gridDim.x;
                const int t = g - div_tpi(hg, g) * hg.tiles_per_img;
                const int q_start = t * kBM, y0 = div_wp(hg, q_start), q0 = q_start - y0 * hg.Wp, s = it % kHaloPA, ac = it % kHaloAcc;
                mbar_wait(&sm.a_full[s], (it / kHaloPA) & 1);
                mbar_wait(&sm.acc_empty[ac], ((it / kHaloAcc) & 1) ^ 1);   // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t a0 = smem_u32(base + (size_t)s * hg.a_bytes) + (uint32_t)q0 * 128u;
                const uint32_t acc = tmem + (uint32_t)(ac * 64);
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap - 3 * ky;
                    const uint32_t start = a0 + (uint32_t)(ky * hg.Wp + kx) * 128u;
                    const uint64_t ad = (uint64_t)((start >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
                    const uint64_t bd = umma_desc_k_sw128(sm.b[tap]);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) umma_bf16(acc, ad + 2 * k, bd + 2 * k, idesc, (tap | k) != 0);
                }
                umma_commit(&sm.a_empty[s]);
                umma_commit(&sm.acc_full[ac]);
            }
        }
    } else if (warp >= 4) {
        // epilogue group g drains accumulator g (tiles g, g + kHaloAcc, ...): a tile's TMEM loads and row-strided stores take
        // several times its MMA time, so several tiles must be in their epilogue at once to keep the tensor pipe fed
        const int q = warp & 3, grp = (warp - 4) >> 2;
        const bool wide = epilogue_wide_ok(ep);
        for (int it = grp; it < my_tiles; it += kHaloAcc) {
            const int g = (int)blockIdx.x + it * (int)gridDim.x;
            const int img = div_tpi(hg, g), t = g - img * hg.tiles_per_img;
            const int q_start = t * kBM, y0 = div_wp(hg, q_start), q0 = q_start - y0 * hg.Wp, s = grp;
            const int qq = q0 + q * 32 + lane, yl = div_wp(hg, qq), xp = qq - yl * hg.Wp, y = y0 + yl;
            const bool row_ok = xp < hg.W && y < hg.H;
            const long long row = ((long long)img * hg.H + y) * hg.W + xp;
            mbar_wait(&sm.acc_full[s], (it / kHaloAcc) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 16) {
                if (c0 >= N) break;  // warp-uniform
                uint32_t r[16];
                tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 64 + c0), r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                if (ep.col_sum) {
                    float s1[16], s2[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) { s1[j] = row_ok ? v[j] : 0.0f; s2[j] = s1[j] * s1[j]; }
                    const float t1 = warp_colsum16(s1, lane), t2 = warp_colsum16(s2, lane);
                    if (!(lane & 1)) {
                        const int cj = c0 + ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        sm.stat[grp][0][q][cj] = t1;
                        sm.stat[grp][1][q][cj] = t2;
                    }
                }
                if (!row_ok) continue;
                epilogue_frag16(ep, wide, row, c0, N, v);
            }
            // the accumulator has been read (tcgen05.wait::ld inside tmem_ld16): hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.acc_empty[s]);
            if (ep.col_sum) {
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
                for (int c = (threadIdx.x & 127); c < 64; c += 128) {
                    if (c < N) {
                        ep.col_sum[(size_t)g * N + c] = (sm.stat[grp][0][0][c] + sm.stat[grp][0][1][c]) + (sm.stat[grp][0][2][c] + sm.stat[grp][0][3][c]);
                        ep.col_sumsq[(size_t)g * N + c] = (sm.stat[grp][1][0][c] + sm.stat[grp][1][1][c]) + (sm.stat[grp][1][2][c] + sm.stat[grp][1][3][c]);
                    }
                }
                asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");   // the group's next tile rewrites the slots
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 64 * kHaloAcc);
}

// --------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 tensor map of a row-major [rows, cols] matrix with pitch ld (elements): box = 64 cols x box_rows, 128B swizzle
static int make_map(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return AB_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AB_ERR_ARG; }
    return AB_OK;
}

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeIm2colFn)p;
    }
    return fn;
}

// im2col-mode map over an NHWC bf16 activation [B,H,W,C]: 64 channels x 128 output pixels per load.  The pixel box
// corners bound the TOP-LEFT pixel of the filter window: [-pad, W - 1 + pad - (kw - 1)] (the CUTLASS fprop convention).
static int make_im2col_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int kh, int kw, int stride, int pad) {
    EncodeIm2colFn enc = get_encode_im2col();
    if (!enc) { set_error("cuTensorMapEncodeIm2col unavailable"); return AB_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper, (cuuint32_t)kBK,
                     (cuuint32_t)kBM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed (%d)", (int)r); return AB_ERR_ARG; }
    // driver workaround carried by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors must not set bit 21 of word 1
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (size_t)B * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1llu << 21);
    return AB_OK;
}

// One tile per CTA when the grid is at most one wave of co-resident CTAs (two tiles per SM); the persistent form above beyond that
// (AB_GEMM_PERSISTENT=0: never).
template <int BN, int STAGES, bool IM2COL>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, int M, int N, int K, const GemmEpilogue& ep,
                       const ConvGeom& cg, cudaStream_t st) {
    static const int persist = getenv("AB_GEMM_PERSISTENT") ? atoi(getenv("AB_GEMM_PERSISTENT")) : 1;
    static std::atomic<int> sm_count[64] = {};
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    int sms = sm_count[dev & 63].load(std::memory_order_relaxed);
    if (sms == 0) {
        AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        sm_count[dev & 63].store(sms, std::memory_order_relaxed);
    }
    const int tiles_m = cdiv(M, kBM), tiles_n = cdiv(N, BN);
    const long long n_tiles = (long long)tiles_m * tiles_n;
    StageTimer tm(IM2COL ? AB_STAGE_CONV_IMPLICIT : AB_STAGE_GEMM, st);
    static const int min_per_sm = getenv("AB_GEMM_PERSIST_MIN") ? atoi(getenv("AB_GEMM_PERSIST_MIN")) : 2;   // tiles per SM beyond which the persistent form runs (4: 10.34, 2: 10.22, 1: 10.21 ms per step)
    if (persist && n_tiles > (long long)min_per_sm * sms && n_tiles < (1ll << 31)) {
        const size_t smem = sizeof(GemmPSmem<BN, STAGES>) + 1024;
        static std::atomic<bool> configured[64] = {};
        if (!configured[dev & 63].load(std::memory_order_relaxed)) {
            AB_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_persistent_kernel<BN, STAGES, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured[dev & 63].store(true, std::memory_order_relaxed);
        }
        gemm_bf16_tn_persistent_kernel<BN, STAGES, IM2COL><<<2 * sms, kGemmPThreads, smem, st>>>(ta, tb, M, N, K, ep, cg, tiles_n, (int)n_tiles);
        count_launch();
        return check_launch("gemm_bf16_tn_persistent_kernel");
    }
    const size_t smem = sizeof(GemmSmem<BN, STAGES>) + 1024;
    static bool configured = false;
    if (!configured) {
        AB_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, STAGES, IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(tiles_m, tiles_n);
    gemm_bf16_tn_kernel<BN, STAGES, IM2COL><<<grid, kGemmThreads, smem, st>>>(ta, tb, M, N, K, ep, cg);
    count_launch();
    return check_launch("gemm_bf16_tn_kernel");
}

int gemm_bf16_tn(int M, int N, int K, const void* A, long long lda, const void* B, long long ldb, const GemmEpilogue& ep,
                 cudaStream_t st) {
    // tile N: the smallest of {64, 128} that keeps the tile count low; BN = 64 lets 2 CTAs share an SM
    const int bn = (N <= 64) ? 64 : 128;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, M, K, lda, kBM);
    if (rc) return rc;
    rc = make_map(&tb, B, N, K, ldb, bn);
    if (rc) return rc;
    const ConvGeom cg{};
    return bn == 64 ? launch_gemm<64, 4, false>(ta, tb, M, N, K, ep, cg, st) : launch_gemm<128, 3, false>(ta, tb, M, N, K, ep, cg, st);
}

// Convolution as implicit GEMM: x bf16 NHWC [B,H,W,C] (C % 64 == 0), w packed [Cout, kh*kw*C] (K order ky, kx, c).
int conv_bf16_implicit(const void* x, int B, int H, int W, int C, const void* w, int Cout, int kh, int kw, int stride, int pad,
                       const GemmEpilogue& ep, cudaStream_t st) {
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    const int M = B * Ho * Wo, K = kh * kw * C;
    const int bn = (Cout <= 64) ? 64 : 128;
    CUtensorMap ta, tb;
    int rc = make_im2col_map(&ta, x, B, H, W, C, kh, kw, stride, pad);
    if (rc) return rc;
    rc = make_map(&tb, w, Cout, K, K, bn);
    if (rc) return rc;
    const ConvGeom cg{Ho, Wo, stride, pad, kw, C / kBK};
    return bn == 64 ? launch_gemm<64, 4, true>(ta, tb, M, Cout, K, ep, cg, st) : launch_gemm<128, 3, true>(ta, tb, M, Cout, K, ep, cg, st);
}


// ---- halo-resident 3x3: geometry, tensor map (tiled mode, box [64 ch, W + 2, R, 1]) and launch
// Used where it measured faster than the im2col pipeline on B200 (tools/time_conv.py, batch 128): output tiles of 64
// channels, i.e. the layer1 convolutions and their data gradients (C = Cout = 64 at 64 x 64: 95.6 -> 75.3 us).  With 128-wide
// output tiles the im2col ring is already at 0.7-1.0 PFLOP/s and the window's start-up latency and dummy columns lose
// (C = Cout = 128 at 32 x 32: 52.6 -> 74.9 us).  AB_CONV_HALO=0 turns it off, =2 forces it for every eligible geometry.
static bool halo_eligible(int H, int W, int C, int Cout, int kh, int kw, int stride, int pad) {
    static const int mode = getenv("AB_CONV_HALO") ? atoi(getenv("AB_CONV_HALO")) : 1;
    if (!(mode && kh == 3 && kw == 3 && stride == 1 && pad == 1 && C % 64 == 0 && W + 2 <= 256 && W >= 2 && H >= 1)) return false;
    return mode == 2 || Cout <= 64;
}

static HaloGeom halo_geom(int H, int W, int C) {
    HaloGeom g;
    g.H = H; g.W = W; g.Wp = W + 2;
    g.tiles_per_img = cdiv(H * g.Wp, kBM);
    g.R = cdiv(3 * g.Wp + kBM + 1, g.Wp);           // rows y0 - 1 .. : window start q0 <= Wp - 1, + 2 Wp + 2 of taps, + 128 pixels
    g.cblocks = C / kBK;
    g.a_stages = g.cblocks > 1 ? 2 : 1;
    g.a_bytes = (unsigned)(((size_t)g.R * g.Wp * 128 + 1023) / 1024 * 1024);
    g.mul_wp = (unsigned)(((1ull << 20) + g.Wp - 1) / g.Wp);
    g.mul_tpi = (unsigned)(((1ull << 20) + g.tiles_per_img - 1) / g.tiles_per_img);
    return g;
}

static int make_halo_map(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, const HaloGeom& g) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return AB_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBK, (cuuint32_t)g.Wp, (cuuint32_t)g.R, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (halo box) failed (%d)", (int)r); return AB_ERR_ARG; }
    return AB_OK;
}

template <int BN>
static int launch_halo(const CUtensorMap& tx, const CUtensorMap& tb, int B, int Cout, int C, const GemmEpilogue& ep,
                       const HaloGeom& g, cudaStream_t st) {
    const size_t smem = (size_t)g.a_stages * g.a_bytes + sizeof(HaloSmemTail<BN>) + 1024;
    static std::atomic<size_t> smem_set[64] = {};  // per device, raised when a geometry needs more
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    if (smem_set[dev & 63].load(std::memory_order_relaxed) < smem) {
        AB_CUDA(cudaFuncSetAttribute(conv3x3_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev & 63].store(smem, std::memory_order_relaxed);
    }
    dim3 grid(B * g.tiles_per_img, cdiv(Cout, BN));
    StageTimer tm(AB_STAGE_CONV_IMPLICIT, st);
    conv3x3_halo_kernel<BN><<<grid, kGemmThreads, smem, st>>>(tx, tb, Cout, C, ep, g);
    count_launch();
    return check_launch("conv3x3_halo_kernel");
}

static int launch_halo_persistent(const CUtensorMap& tx, const CUtensorMap& tb, int B, int Cout, const GemmEpilogue& ep,
                                  const HaloGeom& g, cudaStream_t st) {
    const size_t smem = kHaloPA * (size_t)g.a_bytes + sizeof(HaloPersistSmemTail) + 1024;
    static std::atomic<size_t> smem_set[64] = {};
    static std::atomic<int> sm_count[64] = {};
    int dev = 0;
    AB_CUDA(cudaGetDevice(&dev));
    if (smem_set[dev & 63].load(std::memory_order_relaxed) < smem) {
        AB_CUDA(cudaFuncSetAttribute(conv3x3_halo_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[dev & 63].store(smem, std::memory_order_relaxed);
    }
    int sms = sm_count[dev & 63].load(std::memory_order_relaxed);
    if (sms == 0) {
        AB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        sm_count[dev & 63].store(sms, std::memory_order_relaxed);
    }
    const int n_tiles = B * g.tiles_per_img;
    StageTimer tm(AB_STAGE_CONV_IMPLICIT, st);
    conv3x3_halo_persistent_kernel<<<min(n_tiles, sms), kHaloPThreads, smem, st>>>(tx, tb, Cout, n_tiles, ep, g);
    count_launch();
    return check_launch("conv3x3_halo_persistent_kernel");
}

int conv3x3_halo(const void* x, int B, int H, int W, int C, const void* w, int Cout, const GemmEpilogue& ep, cudaStream_t st) {
    const HaloGeom g = halo_geom(H, W, C);
    static const int persist = getenv("AB_CONV_HALO_PERSISTENT") ? atoi(getenv("AB_CONV_HALO_PERSISTENT")) : 1;
    // the persistent kernel's multiply-shift divisions must be exact: n / Wp for n < tiles_per_img * 128 + 128, and
    // g / tiles_per_img for g < B * tiles_per_img
    auto magic_exact = [](unsigned d, unsigned mul, unsigned long long n_max) {
        if (n_max >= (1ull << 24)) return false;
        for (unsigned long long n = 0; n <= n_max; ++n) if (((n * mul) >> 20) != n / d) return false;
        return true;
    };
    static int cached_key[4] = {0, 0, 0, 0}, cached_ok = 0;   // (Wp, tiles_per_img, B) -> exactness, checked once per geometry
    if (cached_key[0] != g.Wp || cached_key[1] != g.tiles_per_img || cached_key[2] != B) {
        cached_ok = magic_exact((unsigned)g.Wp, g.mul_wp, (unsigned long long)g.tiles_per_img * kBM + kBM) &&
                    magic_exact((unsigned)g.tiles_per_img, g.mul_tpi, (unsigned long long)B * g.tiles_per_img);
        cached_key[0] = g.Wp; cached_key[1] = g.tiles_per_img; cached_key[2] = B;
    }
    const bool use_persistent = persist && cached_ok && C == kBK && Cout <= 64 &&
                                kHaloPA * (size_t)g.a_bytes + sizeof(HaloPersistSmemTail) + 1024 <= 227 * 1024;
    const int bn = (Cout <= 64) ? 64 : 128;
    CUtensorMap tx, tb;
    int rc = make_halo_map(&tx, x, B, H, W, C, g);
    if (rc) return rc;
    rc = make_map(&tb, w, Cout, 9 * C, 9 * C, bn);
    if (rc) return rc;
    // one channel block, one output tile, and the windows in flight + the nine filter tiles within the 227 KB of an SM
    if (use_persistent) return launch_halo_persistent(tx, tb, B, Cout, ep, g, st);
    return bn == 64 ? launch_halo<64>(tx, tb, B, Cout, C, ep, g, st) : launch_halo<128>(tx, tb, B, Cout, C, ep, g, st);
}

// ------------------------------------------------------------------------------------------------- weight gradient
// D[Mo, No] += sum_p G[p, mo] * X[p, no]      (fp32 atomic accumulation, split over the pixel axis)
//
//   G = dY   bf16 [P, Mo]  (Mo = Cout contiguous)
//   X        bf16 [P, No]  (No contiguous), or the NHWC activation read through TMA im2col: No = taps * C, P = B*Ho*Wo
//
// The reduction runs over the ROWS of both operands, so both are MN-major for the tensor core: tiles are loaded as
// [64 rows (k) x 64 contiguous elements] TMA boxes with 128-byte swizzle -- byte-identical to the forward pass's A tile
// -- and described to tcgen05 with a_major = b_major = MN (instruction-descriptor bits 15/16), LBO = 8 KB between the
// two 64-wide MN atoms of a 128-wide tile, SBO = 1 KB between 8-row groups, +2 KB per K=16 step.
// Where element (row, col) of the [Mo, No] product lands: rows / columns are split as (hi, lo) = (i / div, i % div) and
//   offset = row_hi * s_rh + row_lo * s_rl + col_hi * s_ch + col_lo * s_cl,   skipped unless col_lo < cl_valid,
// so the gradient is accumulated straight into the parameter's own layout (conv [Cout,Cin,kh,kw], deconv
// [Cin,Cout,kh,kw], linear [N,K]) with no packed intermediate.
struct WgradOutMap {
    int rdiv, cdiv, cl_valid;
    long long s_rh, s_rl, s_ch, s_cl;
};

constexpr int kWgStages = 3;
struct WgradSmem {
    __nv_bfloat16 a[kWgStages][2][64 * 64];
    __nv_bfloat16 b[kWgStages][2][64 * 64];
    uint64_t full[kWgStages], empty[kWgStages], tmem_full;
    uint32_t tmem_base;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(const void* smem_tile) {
    return (uint64_t)((smem_u32(smem_tile) >> 4) & 0x3FFF) | (512ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <bool IM2COL>
__global__ void __launch_bounds__(kGemmThreads)
wgrad_bf16_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, int P, int Mo, int No,
                  float* __restrict__ ws, int kblocks_per_split, const ConvGeom cg, int taps) {
    extern __shared__ uint8_t smem_raw[];
    auto& sm = *reinterpret_cast<WgradSmem*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int total_kb = (P + 63) / 64;
    const int kb0 = blockIdx.z * kblocks_per_split;
    const int num_k = min(kblocks_per_split, total_kb - kb0);
    if (num_k <= 0) return;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmG) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmX) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        mbar_init(&sm.tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&sm.tmem_base, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            // second halves of the tiles may fall outside Mo / No: plain TMA zero-fills them, im2col boxes are skipped
            int nb_valid = 2;
            if (IM2COL) nb_valid = min(2, taps * cg.cblocks - tile_n * 2);
            // everything that does not change along the pixel axis is worked out once: the (tap, channel block) of the two
            // column halves, and the first pixel block's (image, row, column), advanced by 64 pixels per k-block without a
            // division (the producer is ONE thread: five integer divisions per k-block were on the critical path of the ring)
            int h_cb[2] = {0, 0}, h_kx[2] = {0, 0}, h_ky[2] = {0, 0};
            int img = 0, oy = 0, ox = 0;
            if (IM2COL) {
                for (int h = 0; h < nb_valid; ++h) {
                    const int nblk = tile_n * 2 + h, tap = nblk / cg.cblocks;
                    h_cb[h] = nblk - tap * cg.cblocks;
                    h_ky[h] = tap / cg.kw;
                    h_kx[h] = tap - h_ky[h] * cg.kw;
                }
                const int per = cg.Ho * cg.Wo, p0 = kb0 * 64;
                img = p0 / per;
                const int r = p0 - img * per;
                oy = r / cg.Wo;
                ox = r - oy * cg.Wo;
            }
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % kWgStages;
                const uint32_t ph = (kb / kWgStages) & 1;
                const int p0 = (kb0 + kb) * 64;
                mbar_wait(&sm.empty[s], ph ^ 1);
                mbar_expect_tx(&sm.full[s], (2 + (IM2COL ? nb_valid : 2)) * 64 * 64 * 2);
                tma_load_2d(sm.a[s][0], &tmG, &sm.full[s], tile_m * 128, p0);
                tma_load_2d(sm.a[s][1], &tmG, &sm.full[s], tile_m * 128 + 64, p0);
                if (IM2COL) {
                    const int base_w = ox * cg.stride - cg.pad, base_h = oy * cg.stride - cg.pad;
                    for (int h = 0; h < nb_valid; ++h)
                        tma_load_im2col_4d(sm.b[s][h], &tmX, &sm.full[s], h_cb[h] * 64, base_w, base_h, img, (uint16_t)h_kx[h], (uint16_t)h_ky[h]);
                    ox += 64;
                    while (ox >= cg.Wo) { ox -= cg.Wo; if (++oy == cg.Ho) { oy = 0; ++img; } }
                } else {
                    tma_load_2d(sm.b[s][0], &tmX, &sm.full[s], tile_n * 128, p0);
                    tma_load_2d(sm.b[s][1], &tmX, &sm.full[s], tile_n * 128 + 64, p0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D fp32, A/B bf16, both MN-major, N = 128, M = 128
            constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) |
                                       ((uint32_t)(128 >> 4) << 24);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % kWgStages;
                const uint32_t ph = (kb / kWgStages) & 1;
                mbar_wait(&sm.full[s], ph);
                tc_fence_after();
                const uint64_t ad = umma_desc_mn_sw128(sm.a[s][0]), bd = umma_desc_mn_sw128(sm.b[s][0]);
#pragma unroll
                for (int k = 0; k < 4; ++k)  // 16 reduction rows = 16 * 128 B = 2 KB: +128 in the address field
                    umma_bf16(tmem, ad + 128 * k, bd + 128 * k, idesc, (kb | k) != 0);
                umma_commit(&sm.empty[s]);
            }
            umma_commit(&sm.tmem_full);
        }
    } else if (warp >= 4) {
        mbar_wait(&sm.tmem_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int row = tile_m * 128 + q * 32 + lane;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 16) {
            const int col = tile_n * 128 + c0;
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            // this split's partial tile goes to the workspace [split][mt*128][nt*128] with plain 16-byte stores (whole
            // tile, padding included: nothing to zero, no atomics); wgrad_reduce_kernel adds the splits up
            float* o = ws + ((size_t)blockIdx.z * gridDim.x * 128 + row) * ((size_t)gridDim.y * 128) + col;
            if (((uintptr_t)ws & 31) == 0) {   // two 32-byte stores per lane (rows are 512-byte multiples apart)
                uint32_t w0[8], w1[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { w0[j] = r[j]; w1[j] = r[8 + j]; }
                st_global_v8(o, w0);
                st_global_v8(o + 8, w1);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    reinterpret_cast<float4*>(o)[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                  __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 128);
}

// D(row, col) += sum over splits of ws[split][row][col], placed through the output map.  One thread per element;
// consecutive threads read consecutive columns (coalesced), single writer per destination (deterministic).
__global__ void wgrad_reduce_kernel(const float* __restrict__ ws, int splits, int Mo, int No, int Mp, int Np, float* __restrict__ D,
                                    const WgradOutMap om) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)Mo * No) return;
    const int row = (int)(i / No), col = (int)(i - (long long)row * No);
    const int chi = col / om.cdiv, clo = col - chi * om.cdiv;
    if (clo >= om.cl_valid) return;
    const float* p = ws + (size_t)row * Np + col;
    float acc = 0.0f;
#pragma unroll 4
    for (int z = 0; z < splits; ++z) acc += __ldg(p + (size_t)z * Mp * Np);
    float* d = D + (long long)(row / om.rdiv) * om.s_rh + (long long)(row % om.rdiv) * om.s_rl + (long long)chi * om.s_ch +
               (long long)clo * om.s_cl;
    *d += acc;
}

// The same for the convolution parameter layout [Cout, C, taps] (columns arrive as (tap, c)): one CTA per (row, 32 channels),
// thread (zg, tap, cc) adds every ZG-th split of 32 consecutive columns per tap (coalesced; ZG groups of threads so that a
// layer with few rows still has enough loads in flight: the kernel is pure latency), the groups are added in a fixed order,
// the tile is transposed through shared memory and written as 32 * taps consecutive floats of the destination.
__global__ void wgrad_reduce_conv_kernel(const float* __restrict__ ws, int splits, int C, int taps, int Mp, int Np, int zgroups,
                                         float* __restrict__ D) {
    extern __shared__ float tile[];  // [zgroups][32 * taps]
    const int row = blockIdx.x, c0 = blockIdx.y * 32, t = threadIdx.x;
    const int per = 32 * taps;
    const int zg = t / per, tt = t - zg * per;
    const int tap = tt >> 5, cc = tt & 31;
    const float* p = ws + (size_t)row * Np + (size_t)tap * C + c0 + cc;
    float acc = 0.0f;
    if (c0 + cc < C) {
#pragma unroll 4
        for (int z = zg; z < splits; z += zgroups) acc += __ldg(p + (size_t)z * Mp * Np);
    }
    tile[zg * per + cc * taps + tap] = acc;
    __syncthreads();
    const int n = min(32, C - c0) * taps;
    if (t < n) {
        float sum = tile[t];
        for (int g = 1; g < zgroups; ++g) sum += tile[g * per + t];
        D[((size_t)row * C + c0) * taps + t] += sum;
    }
}

// Pixel-axis split: CTAs are 2 per SM, so cost ~ waves(tiles * s) * (k-blocks per split + epilogue), minimised over s.
static int wgrad_splits(int P, int Mo, int No, int* per_out) {
    const int tiles = cdiv(Mo, 128) * cdiv(No, 128), total_kb = cdiv(P, 64);
    int best = 1;
    long long best_cost = -1;
    for (int sp = 1; sp <= min(total_kb, 1024); ++sp) {
        const int per = cdiv(total_kb, sp);
        if (cdiv(total_kb, per) != sp) continue;  // only split counts that leave no empty split
        const long long cost = (long long)cdiv(tiles * sp, 2 * 148) * (per + 12);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
    }
    *per_out = cdiv(total_kb, best);
    return best;
}

size_t wgrad_workspace_bytes(int P, int Mo, int No) {
    if (P <= 0) return 0;
    int per;
    const int sp = wgrad_splits(P, Mo, No, &per);
    return (size_t)sp * cdiv(Mo, 128) * 128 * cdiv(No, 128) * 128 * sizeof(float);
}

// box of 64 contiguous elements x 64 rows over a row-major [rows, cols] bf16 matrix
static int make_map_mn(CUtensorMap* m, const void* ptr, long long rows, long long cols, long long ld) {
    EncodeTiledFn enc = get_encode();
    if (!enc) { set_error("cuTensorMapEncodeTiled unavailable"); return AB_ERR_UNSUPPORTED; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AB_ERR_ARG; }
    return AB_OK;
}

static int make_im2col_map_px(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                              int pixels) {
    EncodeIm2colFn enc = get_encode_im2col();
    if (!enc) { set_error("cuTensorMapEncodeIm2col unavailable"); return AB_ERR_UNSUPPORTED; }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (kw - 1), pad - (kh - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, lower, upper, 64u,
                     (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeIm2col failed (%d)", (int)r); return AB_ERR_ARG; }
    int drv = 0;
    cudaDriverGetVersion(&drv);
    if (drv <= 13010 && (size_t)B * H * W * C * 2 < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1llu << 21);
    return AB_OK;
}

template <bool IM2COL>
static int launch_wgrad(const CUtensorMap& tg, const CUtensorMap& tx, int P, int Mo, int No, float* D, const WgradOutMap& om,
                        float* ws, const ConvGeom& cg, int taps, cudaStream_t st, int conv_param_C = 0) {
    const size_t smem = sizeof(WgradSmem) + 1024;
    static bool configured = false;
    if (!configured) {
        AB_CUDA(cudaFuncSetAttribute(wgrad_bf16_kernel<IM2COL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int mt = cdiv(Mo, 128), nt = cdiv(No, 128);
    int per;
    const int splits = wgrad_splits(P, Mo, No, &per);
    StageTimer tm(AB_STAGE_WGRAD, st);
    wgrad_bf16_kernel<IM2COL><<<dim3(mt, nt, splits), kGemmThreads, smem, st>>>(tg, tx, P, Mo, No, ws, per, cg, taps);
    if (conv_param_C > 0 && taps <= 32) {
        // split groups only where the grid alone is too small to keep the loads in flight (layer1: 128 CTAs)
        const int ctas = Mo * cdiv(conv_param_C, 32);
        const int zg = ctas >= 296 ? 1 : max(1, min(min(3, splits), 1024 / (32 * taps)));
        wgrad_reduce_conv_kernel<<<dim3(Mo, cdiv(conv_param_C, 32)), 32 * taps * zg, 32 * taps * zg * sizeof(float), st>>>(
            ws, splits, conv_param_C, taps, mt * 128, nt * 128, zg, D);
    } else {
        const long long n = (long long)Mo * No;
        wgrad_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws, splits, Mo, No, mt * 128, nt * 128, D, om);
    }
    count_launch(2);
    return check_launch("wgrad_bf16_kernel");
}

#ifdef AB_UMMA_SHIFT_TEST
// Debug build only (tools/umma_shift_test.py): does a K-major SWIZZLE_128B operand whose start address is `shift` rows
// (128 B each) into a TMA-written tile read the right data, and does it need the descriptor's base-offset field?
__global__ void __launch_bounds__(256)
umma_shift_test_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int shift,
                       int use_base_offset, int a_rows) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __nv_bfloat16* a = reinterpret_cast<__nv_bfloat16*>(base);                       // a_rows x 64
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(base + 256 * 128);           // 64 x 64
    uint64_t* full = reinterpret_cast<uint64_t*>(base + 256 * 128 + 64 * 128);
    uint64_t* done = full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 1 && lane == 0) { mbar_init(full, 1); mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 2) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0 && lane == 0) {
        mbar_expect_tx(full, (a_rows + 64) * 128);
        tma_load_2d(a, &tmA, full, 0, 0);
        tma_load_2d(b, &tmB, full, 0, 0);
    } else if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        mbar_wait(full, 0);
        tc_fence_after();
        const uint32_t start = smem_u32(a) + (uint32_t)shift * 128u;
        uint64_t ad = (uint64_t)((start >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
        if (use_base_offset) ad |= (uint64_t)((start >> 7) & 7) << 49;
        const uint64_t bd = umma_desc_k_sw128(b);
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc, k != 0);
        umma_commit(done);
    } else if (warp >= 4) {
        mbar_wait(done, 0);
        tc_fence_after();
        const int q = warp & 3, row = q * 32 + lane;
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
            for (int j = 0; j < 16; ++j) out[row * 64 + c0 + j] = __uint_as_float(r[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem, 64);
}
#endif

}  // namespace ab

#ifdef AB_UMMA_SHIFT_TEST
extern "C" __attribute__((visibility("default"))) int ab_debug_umma_shift(const void* A, int a_rows, const void* B, float* out,
                                                                          int shift, int use_base_offset, void* stream) {
    using namespace ab;
    CUtensorMap ta, tb;
    int rc = make_map(&ta, A, a_rows, 64, 64, a_rows);
    if (rc) return rc;
    rc = make_map(&tb, B, 64, 64, 64, 64);
    if (rc) return rc;
    const size_t smem = 256 * 128 + 64 * 128 + 64 + 1024;
    AB_CUDA(cudaFuncSetAttribute(umma_shift_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_shift_test_kernel<<<1, 256, smem, (cudaStream_t)stream>>>(ta, tb, out, shift, use_base_offset, a_rows);
    return check_launch("umma_shift_test_kernel");
}
#endif

extern "C" int ab_gemm_bf16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                            int out_fp32, const float* scale, const float* bias, const void* residual, int64_t ldr,
                            int relu, float* col_sum, float* col_sumsq, void* stream) {
    AB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative size");
    if (M == 0 || N == 0) return AB_OK;
    AB_REQUIRE(K > 0, "K must be positive");
    AB_REQUIRE(A && B && D, "null matrix");
    AB_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, "K, lda, ldb must be multiples of 8 (16-byte rows for TMA)");
    AB_REQUIRE(N % 8 == 0 && ldd % 8 == 0 && (!residual || ldr % 8 == 0), "N, ldd, ldr must be multiples of 8");
    AB_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)D & 15) == 0 &&
                   ((uintptr_t)residual & 15) == 0, "matrices must be 16-byte aligned");
    AB_REQUIRE((col_sum == nullptr) == (col_sumsq == nullptr), "col_sum and col_sumsq go together");
    ab::GemmEpilogue ep;
    ep.D = D; ep.ldd = ldd; ep.out_fp32 = out_fp32; ep.scale = scale; ep.bias = bias;
    ep.residual = (const __nv_bfloat16*)residual; ep.ldr = ldr; ep.relu = relu; ep.col_sum = col_sum; ep.col_sumsq = col_sumsq;
    return ab::gemm_bf16_tn(M, N, K, A, lda, B, ldb, ep, (cudaStream_t)stream);
}

extern "C" int ab_conv_bf16_nhwc(const void* x, int B, int H, int W, int C, const void* w_packed, int Cout, int kh, int kw,
                                 int stride, int pad, void* D, int64_t ldd, int out_fp32, const float* scale,
                                 const float* bias, const void* residual, int64_t ldr, int relu, float* col_sum,
                                 float* col_sumsq, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(C % 64 == 0, "implicit-GEMM convolution needs C % 64 == 0 (use ab_im2col_nhwc + ab_gemm_bf16 otherwise)");
    AB_REQUIRE(kh <= 16 && kw <= 16 && pad < kh && pad < kw, "filter / padding out of range");
    AB_REQUIRE(x && w_packed && D, "null pointer");
    AB_REQUIRE(Cout % 8 == 0 && ldd % 8 == 0 && (!residual || ldr % 8 == 0), "Cout, ldd, ldr must be multiples of 8");
    AB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)D & 15) == 0 &&
                   ((uintptr_t)residual & 15) == 0, "tensors must be 16-byte aligned");
    AB_REQUIRE((col_sum == nullptr) == (col_sumsq == nullptr), "col_sum and col_sumsq go together");
    AB_REQUIRE((long long)B * H * W < (1ll << 31), "too many pixels");
    ab::GemmEpilogue ep;
    ep.D = D; ep.ldd = ldd; ep.out_fp32 = out_fp32; ep.scale = scale; ep.bias = bias;
    ep.residual = (const __nv_bfloat16*)residual; ep.ldr = ldr; ep.relu = relu; ep.col_sum = col_sum; ep.col_sumsq = col_sumsq;
    if (ab::halo_eligible(H, W, C, Cout, kh, kw, stride, pad))
        return ab::conv3x3_halo(x, B, H, W, C, w_packed, Cout, ep, (cudaStream_t)stream);
    return ab::conv_bf16_implicit(x, B, H, W, C, w_packed, Cout, kh, kw, stride, pad, ep, (cudaStream_t)stream);
}

extern "C" int ab_conv_stat_rows(int B, int H, int W, int C, int Cout, int kh, int kw, int stride, int pad) {
    if (B <= 0 || H <= 0 || W <= 0 || kh <= 0 || kw <= 0 || stride <= 0) return 0;
    if (ab::halo_eligible(H, W, C, Cout, kh, kw, stride, pad)) return B * ab::halo_geom(H, W, C).tiles_per_img;
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    return ab::cdiv(B * Ho * Wo, ab::kBM);
}

extern "C" uint64_t ab_wgrad_workspace_bytes(int P, int Mo, int No) { return (uint64_t)ab::wgrad_workspace_bytes(P, Mo, No); }

extern "C" int ab_wgrad_bf16(int P, int Mo, int No, const void* G, int64_t ldg, const void* X, int64_t ldx, float* D,
                             const ab_wgrad_map* map, void* ws, void* stream) {
    AB_REQUIRE(P >= 0 && Mo > 0 && No > 0, "bad shape");
    if (P == 0) return AB_OK;
    AB_REQUIRE(G && X && D && map && ws, "null pointer");
    AB_REQUIRE(((uintptr_t)ws & 15) == 0, "workspace must be 16-byte aligned");
    AB_REQUIRE(map->row_div > 0 && map->col_div > 0 && map->col_lo_valid > 0, "bad output map");
    AB_REQUIRE(ldg % 8 == 0 && ldx % 8 == 0 && Mo <= ldg && No <= ldx, "ldg, ldx must be multiples of 8 (16-byte rows) and cover Mo, No");
    AB_REQUIRE(((uintptr_t)G & 15) == 0 && ((uintptr_t)X & 15) == 0, "operands must be 16-byte aligned");
    CUtensorMap tg, tx;
    int rc = ab::make_map_mn(&tg, G, P, Mo, ldg);
    if (rc) return rc;
    rc = ab::make_map_mn(&tx, X, P, No, ldx);
    if (rc) return rc;
    const ab::WgradOutMap om{map->row_div, map->col_div, map->col_lo_valid, map->s_row_hi, map->s_row_lo, map->s_col_hi, map->s_col_lo};
    return ab::launch_wgrad<false>(tg, tx, P, Mo, No, D, om, (float*)ws, ab::ConvGeom{}, 0, (cudaStream_t)stream);
}

extern "C" int ab_conv_wgrad_bf16_nhwc(const void* x, int B, int H, int W, int C, const void* dy, int Cout, int kh, int kw,
                                       int stride, int pad, float* dw, int param_layout, void* ws, void* stream) {
    AB_REQUIRE(B >= 0 && H > 0 && W > 0 && C > 0 && Cout > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "bad shape");
    if (B == 0) return AB_OK;
    AB_REQUIRE(C % 64 == 0 && Cout % 8 == 0, "implicit weight gradient needs C % 64 == 0 and Cout % 8 == 0");
    AB_REQUIRE(x && dy && dw && ws, "null pointer");
    AB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)ws & 15) == 0, "tensors must be 16-byte aligned");
    const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
    const int P = B * Ho * Wo, taps = kh * kw;
    CUtensorMap tg, tx;
    int rc = ab::make_map_mn(&tg, dy, P, Cout, Cout);
    if (rc) return rc;
    rc = ab::make_im2col_map_px(&tx, x, B, H, W, C, kh, kw, stride, pad, 64);
    if (rc) return rc;
    const ab::ConvGeom cg{Ho, Wo, stride, pad, kw, C / 64};
    // columns are (tap, c): packed [Cout, taps*C] keeps them; the parameter layout [Cout, C, kh, kw] swaps them
    const ab::WgradOutMap om = param_layout ? ab::WgradOutMap{1 << 30, C, C, 0, (long long)C * taps, 1, taps}
                                            : ab::WgradOutMap{1 << 30, C, C, 0, (long long)C * taps, C, 1};
    return ab::launch_wgrad<true>(tg, tx, P, Cout, taps * C, dw, om, (float*)ws, cg, taps, (cudaStream_t)stream, param_layout ? C : 0);
}
