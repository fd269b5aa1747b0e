// CCV-space sampler and view engine on device (sm_100a).
//
//   ab_ccv_sample    OVGSet.update's Categorical draw + row_col_calc + occurrence count map
//                    (anakin/artiboost/ovg_set.py:104-132,161-178)
//   ab_view_from_id  ViewEngine.get_view / get_perspective_from_id / caculate_align_mat
//                    (anakin/artiboost/view_engine.py:17-86)
//
// The categorical draw is an inverse-CDF search over a float64 inclusive prefix sum of the flat weight map.  Every
// partial sum of fp32 weights in [2^-27, 2^22] is exact in fp64, so the parallel scan equals the sequential one
// bit for bit and the draw is integer-exact against the CPU oracle for the same uniforms.
#include "view_math.cuh"

namespace ab {

constexpr int kScanThreads = 1024;

__global__ void __launch_bounds__(kScanThreads)
ccv_cdf_kernel(const float* __restrict__ w, int n, double* __restrict__ cdf) {
    __shared__ double warp_tot[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = cdiv(n, kScanThreads);
    const int lo = min(n, tid * per), hi = min(n, lo + per);
    double local = 0.0;
    for (int i = lo; i < hi; ++i) local += (double)w[i];
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        double t = warp_tot[lane];
        double s = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += u;
        }
        warp_tot[lane] = s - t;  // exclusive
    }
    __syncthreads();
    double run = warp_tot[wid] + (incl - local);
    for (int i = lo; i < hi; ++i) {
        run += (double)w[i];
        cdf[i] = run;
    }
}

__global__ void ccv_draw_kernel(const double* __restrict__ cdf, int n_cells, int n_persp, int n_grasp,
                                const float* __restrict__ u, int n, int32_t* __restrict__ obj_id,
                                int32_t* __restrict__ persp_id, int32_t* __restrict__ grasp_id,
                                int32_t* __restrict__ occurrence) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double target = (double)u[i] * cdf[n_cells - 1];
    int lo = 0, hi = n_cells;  // first index with cdf[idx] > target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] > target) hi = mid; else lo = mid + 1;
    }
    int flat = min(lo, n_cells - 1);
    obj_id[i] = flat / (n_persp * n_grasp);
    persp_id[i] = (flat / n_grasp) % n_persp;
    grasp_id[i] = flat % n_grasp;
    if (occurrence) atomicAdd(&occurrence[flat], 1);
}

__global__ void view_kernel(const int32_t* __restrict__ persp_id, int n, int u_bins, int theta_bins, float z_min,
                            float z_max, const float* __restrict__ rand4, float* __restrict__ persp_rotmat,
                            float* __restrict__ free_transf, float* __restrict__ z_offset) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = reinterpret_cast<const float4*>(rand4)[i];
    view_from_id(persp_id[i], u_bins, theta_bins, z_min, z_max, r.x, r.y, r.z, r.w, persp_rotmat + (size_t)i * 9,
                 free_transf + (size_t)i * 16, z_offset + (size_t)i * 3);
}

void launch_ccv_cdf(const float* w, int n, double* cdf, cudaStream_t st) { ccv_cdf_kernel<<<1, kScanThreads, 0, st>>>(w, n, cdf); }

}  // namespace ab

extern "C" int ab_ccv_sample(const float* weight_map, int n_obj, int n_persp, int n_grasp, const float* uniforms,
                             int n, double* cdf_ws, int32_t* obj_id, int32_t* persp_id, int32_t* grasp_id,
                             int32_t* occurrence, void* stream) {
    AB_REQUIRE(n_obj > 0 && n_persp > 0 && n_grasp > 0, "empty CCV space");
    AB_REQUIRE((int64_t)n_obj * n_persp * n_grasp < (1ll << 30), "CCV space too large");
    AB_REQUIRE(n >= 0, "negative n");
    if (n == 0) return AB_OK;
    AB_REQUIRE(weight_map && uniforms && cdf_ws && obj_id && persp_id && grasp_id, "null pointer");
    const int n_cells = n_obj * n_persp * n_grasp;
    cudaStream_t st = (cudaStream_t)stream;
    ab::StageTimer tm(AB_STAGE_CCV, st);
    ab::ccv_cdf_kernel<<<1, ab::kScanThreads, 0, st>>>(weight_map, n_cells, cdf_ws);
    ab::ccv_draw_kernel<<<ab::cdiv(n, 256), 256, 0, st>>>(cdf_ws, n_cells, n_persp, n_grasp, uniforms, n, obj_id,
                                                         persp_id, grasp_id, occurrence);
    ab::count_launch(2);
    return ab::check_launch("ab_ccv_sample");
}

extern "C" int ab_view_from_id(const int32_t* persp_id, int n, int u_bins, int theta_bins, float z_min, float z_max,
                               const float* rand4, float* persp_rotmat, float* camera_free_transf, float* z_offset,
                               void* stream) {
    AB_REQUIRE(n >= 0 && u_bins > 0 && theta_bins > 0, "bad sizes");
    if (n == 0) return AB_OK;
    AB_REQUIRE(persp_id && rand4 && persp_rotmat && camera_free_transf && z_offset, "null pointer");
    AB_REQUIRE(((uintptr_t)rand4 & 15) == 0, "rand4 must be 16-byte aligned");
    ab::StageTimer tm(AB_STAGE_VIEW, (cudaStream_t)stream);
    ab::view_kernel<<<ab::cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(persp_id, n, u_bins, theta_bins, z_min, z_max,
                                                                       rand4, persp_rotmat, camera_free_transf,
                                                                       z_offset);
    ab::count_launch();
    return ab::check_launch("ab_view_from_id");
}
