// One-launch draw of a batch of synthesis inputs (sm_100a), the cached CCV distribution, and the back-of-hand blacklist.
//
//   ab_ccv_cdf        the fp64 inclusive prefix sum of the flat weight map, once per epoch (the map only changes in
//                     ArtiBoostLoader.step_eval, anakin/artiboost/artiboost_loader.py:292-340)
//   ab_synth_draw     per sample, one thread: categorical cell draw + row_col_calc + occurrence count
//                     (ovg_set.py:104-132,161-178), view (view_engine.py:17-86), grasp lookup (grasp_engine.py:47-53),
//                     the `random` scrambler's noise (scrambler.py:65-81) and the renderer's per-view draws -- hand
//                     texture, light intensity, background crop (utils/renderer.py:102-104,125-136).  The reference spreads
//                     these over np.random / random / torch.rand in three processes; here one counter-based Philox stream
//                     (seed, sample index, call offset) feeds all of them, so a batch is reproducible from (seed, offset)
//                     alone and the whole synthesis front end is a single launch instead of ~60 small ones.
//   ab_ccv_blacklist  _construct_blacklist_map (artiboost_loader.py:415-500): cells whose view shows the back of the hand
//                     (th_sgn < -0.8), one thread per cell instead of a Python loop over the CCV space.
#include <curand_kernel.h>

#include "mano_math.cuh"
#include "view_math.cuh"

namespace ab {

void launch_ccv_cdf(const float* w, int n, double* cdf, cudaStream_t st);

constexpr int kDrawFloats = AB_SYNTH_UNIFORMS;  // 32

// Uniforms of one sample, all U[0,1):
//   0        cell of the CCV space
//   1..4     view: u jitter, theta jitter, in-plane roll, camera distance
//   5..24    ten Box-Muller pairs -> 20 normals: 0..2 translation noise, 3..18 joint-angle noise (19 unused)
//   25, 26   hand texture, light intensity
//   27..30   background: image, crop height, crop y0, crop x0
//   31       unused
__global__ void __launch_bounds__(128)
synth_draw_kernel(const ab_synth_space sp, const double* __restrict__ cdf, int n, unsigned long long seed,
                  unsigned long long offset, const float* __restrict__ uniforms_in, int32_t* __restrict__ obj_id,
                  int32_t* __restrict__ persp_id, int32_t* __restrict__ grasp_id, int32_t* __restrict__ occurrence,
                  float* __restrict__ hand_pose, float* __restrict__ hand_shape, float* __restrict__ hand_tsl,
                  float* __restrict__ persp_rotmat, float* __restrict__ free_transf, float* __restrict__ z_offset,
                  float* __restrict__ noise_tsl, float* __restrict__ noise_angle, int32_t* __restrict__ hand_tex,
                  float* __restrict__ light, int32_t* __restrict__ bg_sel, float* __restrict__ uniforms_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float u[kDrawFloats];
    if (uniforms_in) {
#pragma unroll
        for (int k = 0; k < kDrawFloats / 4; ++k) {
            const float4 r = reinterpret_cast<const float4*>(uniforms_in)[(size_t)i * (kDrawFloats / 4) + k];
            u[4 * k] = r.x; u[4 * k + 1] = r.y; u[4 * k + 2] = r.z; u[4 * k + 3] = r.w;
        }
    } else {
        curandStatePhilox4_32_10_t st;
        curand_init(seed, (unsigned long long)i, offset, &st);
#pragma unroll
        for (int k = 0; k < kDrawFloats / 4; ++k) {
            const float4 r = curand_uniform4(&st);  // (0, 1]  ->  [0, 1)
            u[4 * k] = 1.0f - r.x; u[4 * k + 1] = 1.0f - r.y; u[4 * k + 2] = 1.0f - r.z; u[4 * k + 3] = 1.0f - r.w;
        }
    }
    if (uniforms_out) {
#pragma unroll
        for (int k = 0; k < kDrawFloats / 4; ++k)
            reinterpret_cast<float4*>(uniforms_out)[(size_t)i * (kDrawFloats / 4) + k] = make_float4(u[4 * k], u[4 * k + 1], u[4 * k + 2], u[4 * k + 3]);
    }
    // ---- cell: inverse CDF (first index with cdf > u * total), unflattened like row_col_calc
    const int n_cells = sp.n_obj * sp.n_persp * sp.n_grasp;
    const double target = (double)u[0] * cdf[n_cells - 1];
    int lo = 0, hi = n_cells;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] > target) hi = mid; else lo = mid + 1;
    }
    const int flat = min(lo, n_cells - 1);
    const int o = flat / (sp.n_persp * sp.n_grasp), p = (flat / sp.n_grasp) % sp.n_persp, g = flat % sp.n_grasp;
    obj_id[i] = o; persp_id[i] = p; grasp_id[i] = g;
    if (occurrence) atomicAdd(&occurrence[flat], 1);
    // ---- view
    view_from_id(p, sp.u_bins, sp.theta_bins, sp.z_min, sp.z_max, u[1], u[2], u[3], u[4], persp_rotmat + (size_t)i * 9,
                 free_transf + (size_t)i * 16, z_offset + (size_t)i * 3);
    // ---- grasp
    const float* row = sp.grasp_table + ((size_t)o * sp.n_grasp + g) * 61;
    for (int k = 0; k < 48; ++k) hand_pose[(size_t)i * 48 + k] = __ldg(row + k);
    for (int k = 0; k < 10; ++k) hand_shape[(size_t)i * 10 + k] = __ldg(row + 48 + k);
    for (int k = 0; k < 3; ++k) hand_tsl[(size_t)i * 3 + k] = __ldg(row + 58 + k);
    // ---- scrambler noise: N(0, sigma) by Box-Muller on the uniform pairs
    float nrm[20];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const float r = sqrtf(-2.0f * logf(1.0f - u[5 + 2 * k]));  // 1 - u in (0, 1]
        float s, c;
        sincosf(6.283185307179586f * u[6 + 2 * k], &s, &c);
        nrm[2 * k] = r * c;
        nrm[2 * k + 1] = r * s;
    }
    if (noise_tsl)
        for (int k = 0; k < 3; ++k) noise_tsl[(size_t)i * 3 + k] = nrm[k] * sp.tsl_sigma;
    if (noise_angle)
        for (int k = 0; k < 16; ++k) noise_angle[(size_t)i * 16 + k] = nrm[3 + k] * sp.pose_sigma;
    // ---- renderer draws
    if (hand_tex) hand_tex[i] = min((int)(u[25] * (float)sp.n_hand_tex), sp.n_hand_tex - 1);
    if (light) light[i] = sp.light_lo + u[26] * (sp.light_hi - sp.light_lo);
    if (bg_sel) {
        int32_t* b = bg_sel + (size_t)i * 5;
        if (sp.n_bg <= 0) {
            b[0] = -1; b[1] = b[2] = b[3] = b[4] = 0;
        } else {
            // the crop rule of renderer.py:125-136 for backgrounds 1.5 x the frame: height drawn in [H, bg_h], width in proportion
            const int bid = min((int)(u[27] * (float)sp.n_bg), sp.n_bg - 1);
            const int ch = sp.height + min((int)(u[28] * (float)(sp.bg_h - sp.height + 1)), sp.bg_h - sp.height);
            const int cw = min((int)(((long long)ch * sp.width) / sp.height), sp.bg_w);
            const int y0 = min((int)(u[29] * (float)(sp.bg_h - ch + 1)), sp.bg_h - ch);
            const int x0 = min((int)(u[30] * (float)(sp.bg_w - cw + 1)), sp.bg_w - cw);
            b[0] = bid; b[1] = x0; b[2] = y0; b[3] = cw; b[4] = ch;
        }
    }
}

// th_sgn = (Rv^T Rw back) . z with Rw the wrist rotation of the grasp and back = (1, 0.2, 0) / |.|
// (artiboost_loader.py:482-492); rand2 [cells, 2] = the (u, theta) jitter get_view draws for the cell, NULL = bin centre.
__global__ void ccv_blacklist_kernel(const ab_synth_space sp, const float* __restrict__ rand2, float threshold,
                                     uint8_t* __restrict__ out, float* __restrict__ th_out) {
    const int n_cells = sp.n_obj * sp.n_persp * sp.n_grasp;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int o = i / (sp.n_persp * sp.n_grasp), p = (i / sp.n_grasp) % sp.n_persp, g = i % sp.n_grasp;
    const float* row = sp.grasp_table + ((size_t)o * sp.n_grasp + g) * 61;
    float Rw[9], Rv[9];
    rodrigues(__ldg(row), __ldg(row + 1), __ldg(row + 2), Rw);
    view_rotmat(p, sp.u_bins, sp.theta_bins, rand2 ? rand2[2 * (size_t)i] : 0.5f, rand2 ? rand2[2 * (size_t)i + 1] : 0.5f, Rv);
    const float inv = 0.9805806756909202f;  // 1 / |(1, 0.2, 0)|
    const float back[3] = {inv, 0.2f * inv, 0.0f};
    float w[3];
    mat3_vec(Rw, back, w);
    const float th = Rv[2] * w[0] + Rv[5] * w[1] + Rv[8] * w[2];
    out[i] = th < threshold ? 1 : 0;
    if (th_out) th_out[i] = th;
}

static int check_space(const ab_synth_space* sp) {
    if (!sp) return -1;
    if (sp->n_obj <= 0 || sp->n_persp <= 0 || sp->n_grasp <= 0 || (int64_t)sp->n_obj * sp->n_persp * sp->n_grasp >= (1ll << 30)) return -1;
    if (sp->u_bins <= 0 || sp->theta_bins <= 0 || sp->u_bins * sp->theta_bins != sp->n_persp || !sp->grasp_table) return -1;
    return 0;
}

}  // namespace ab

extern "C" int ab_ccv_cdf(const float* weight_map, int n_cells, double* cdf, void* stream) {
    AB_REQUIRE(n_cells > 0 && n_cells < (1 << 30), "bad CCV space size");
    AB_REQUIRE(weight_map && cdf, "null pointer");
    ab::StageTimer tm(AB_STAGE_CCV, (cudaStream_t)stream);
    ab::launch_ccv_cdf(weight_map, n_cells, cdf, (cudaStream_t)stream);
    ab::count_launch();
    return ab::check_launch("ab_ccv_cdf");
}

extern "C" int ab_synth_draw(const ab_synth_space* space, const double* cdf, int n, uint64_t seed, uint64_t offset,
                             const float* uniforms, int32_t* obj_id, int32_t* persp_id, int32_t* grasp_id,
                             int32_t* occurrence, float* hand_pose, float* hand_shape, float* hand_tsl, float* persp_rotmat,
                             float* camera_free_transf, float* z_offset, float* noise_tsl, float* noise_angle,
                             int32_t* hand_tex, float* light, int32_t* bg_sel, float* uniforms_out, void* stream) {
    AB_REQUIRE(ab::check_space(space) == 0, "bad CCV space (sizes, u_bins * theta_bins == n_persp, grasp table)");
    AB_REQUIRE(n >= 0, "negative n");
    if (n == 0) return AB_OK;
    AB_REQUIRE(cdf && obj_id && persp_id && grasp_id && hand_pose && hand_shape && hand_tsl && persp_rotmat &&
                   camera_free_transf && z_offset, "null pointer");
    AB_REQUIRE((!uniforms || ((uintptr_t)uniforms & 15) == 0) && (!uniforms_out || ((uintptr_t)uniforms_out & 15) == 0),
               "uniform arrays must be 16-byte aligned");
    AB_REQUIRE(!bg_sel || space->n_bg <= 0 || (space->width > 0 && space->height > 0 && space->bg_h >= space->height &&
                                               space->bg_w >= space->width), "backgrounds must be at least as large as the frame");
    AB_REQUIRE(!hand_tex || space->n_hand_tex > 0, "n_hand_tex must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    ab::StageTimer tm(AB_STAGE_SYNTH_DRAW, st);
    ab::synth_draw_kernel<<<ab::cdiv(n, 128), 128, 0, st>>>(*space, cdf, n, seed, offset, uniforms, obj_id, persp_id, grasp_id,
                                                           occurrence, hand_pose, hand_shape, hand_tsl, persp_rotmat,
                                                           camera_free_transf, z_offset, noise_tsl, noise_angle, hand_tex, light,
                                                           bg_sel, uniforms_out);
    ab::count_launch();
    return ab::check_launch("ab_synth_draw");
}

extern "C" int ab_ccv_blacklist(const ab_synth_space* space, const float* rand2, float threshold, uint8_t* blacklist,
                                float* th_sgn, void* stream) {
    AB_REQUIRE(ab::check_space(space) == 0, "bad CCV space (sizes, u_bins * theta_bins == n_persp, grasp table)");
    AB_REQUIRE(blacklist, "null pointer");
    const int n_cells = space->n_obj * space->n_persp * space->n_grasp;
    cudaStream_t st = (cudaStream_t)stream;
    ab::StageTimer tm(AB_STAGE_CCV, st);
    ab::ccv_blacklist_kernel<<<ab::cdiv(n_cells, 256), 256, 0, st>>>(*space, rand2, threshold, blacklist, th_sgn);
    ab::count_launch();
    return ab::check_launch("ab_ccv_blacklist");
}
