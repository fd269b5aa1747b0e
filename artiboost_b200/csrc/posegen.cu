// Pose generator (sm_100a): PreProcessorPoseGenerator.forward (anakin/artiboost/preprocessor.py:20-99) with the
// `random` scrambler (scrambler.py:65-81) and NullRefine (refiner.py:131-147).
//
// The reference runs three full MANO forwards per batch.  Only the last one needs vertices: the first is read for
// the root rotation and keypoint 9 (a chain joint with the root as parent), the second only feeds the anatomical
// scramblers.  So the work is one thread-per-sample prelude (rotations, object pose, translation fix-up,
// scrambling) that emits the final pose and a rigid map, followed by ONE fused LBS launch (mano.cu) that applies
// that rigid map in its store.
#include "mano_math.cuh"

namespace ab {

int launch_mano(const ab_mano_model* model, int batch, const float* pose, const float* betas, const float* post_rt,
                int center_idx, float* verts, float* joints, float* transforms_abs, cudaStream_t st);

__global__ void posegen_prelude_kernel(ab_mano_model m, int batch, const float* __restrict__ hand_pose,
                                       const float* __restrict__ hand_shape, const float* __restrict__ hand_tsl,
                                       const float* __restrict__ persp_rotmat, const float* __restrict__ free_transf,
                                       const float* __restrict__ z_offset, const float* __restrict__ noise_tsl,
                                       const float* __restrict__ noise_angle, float* __restrict__ obj_pose,
                                       float* __restrict__ pose_out, float* __restrict__ post_rt,
                                       float* __restrict__ tsl_out, float* __restrict__ cso_out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const float* pose = hand_pose + (size_t)b * 48;
    float beta[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) beta[k] = hand_shape ? hand_shape[(size_t)b * 10 + k] : 0.0f;
    // rest joints 0 (root = MANO rotation centre, preprocessor.py:55) and 4 (keypoint 9 of the 21-order)
    float J0[3], J4[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float a0 = m.j_template[d], a4 = m.j_template[12 + d];
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            a0 += m.j_shapedirs[d * 10 + k] * beta[k];
            a4 += m.j_shapedirs[(12 + d) * 10 + k] * beta[k];
        }
        J0[d] = a0;
        J4[d] = a4;
    }
    float tsl[3] = {hand_tsl[(size_t)b * 3], hand_tsl[(size_t)b * 3 + 1], hand_tsl[(size_t)b * 3 + 2]};
    float R0[9];
    rodrigues(pose[0], pose[1], pose[2], R0);
    // keypoint 9 of MANO forward #1 (+ hand_tsl): G4.t = R0 (J4 - J0) + J0     (preprocessor.py:25-28)
    float rel[3] = {J4[0] - J0[0], J4[1] - J0[1], J4[2] - J0[2]}, j9[3];
    mat3_vec(R0, rel, j9);
#pragma unroll
    for (int d = 0; d < 3; ++d) j9[d] += J0[d] + tsl[d];
    float Rv[9], Rvi[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) Rv[j] = persp_rotmat[(size_t)b * 9 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) Rvi[3 * i + j] = Rv[3 * j + i];
    // object pose = camera_free . [Rv^T | z_offset - Rv^T j9 / 2]                (preprocessor.py:37-40)
    float op[3], cso[3];
    mat3_vec(Rvi, j9, op);
#pragma unroll
    for (int d = 0; d < 3; ++d) cso[d] = z_offset[(size_t)b * 3 + d] - op[d] / 2.0f;
    const float* F = free_transf + (size_t)b * 16;
    float* O = obj_pose + (size_t)b * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float f0 = F[4 * i], f1 = F[4 * i + 1], f2 = F[4 * i + 2], f3 = F[4 * i + 3];
#pragma unroll
        for (int j = 0; j < 3; ++j) O[4 * i + j] = f0 * Rvi[j] + f1 * Rvi[3 + j] + f2 * Rvi[6 + j];
        O[4 * i + 3] = f0 * cso[0] + f1 * cso[1] + f2 * cso[2] + f3;
    }
    // hand root rotation in the view frame                                      (preprocessor.py:49-52)
    float R1m[9], aa1[3], R1[9];
    mat3_mul(Rvi, R0, R1m);
    rotmat_to_aa(R1m, aa1);
    rodrigues(aa1[0], aa1[1], aa1[2], R1);
    // translation fix-up for MANO's off-origin rotation centre                  (preprocessor.py:55-60)
    float r0c[3], r1c[3], t0[3], nt[3];
    mat3_vec(R0, J0, r0c);
    mat3_vec(R1, J0, r1c);
#pragma unroll
    for (int d = 0; d < 3; ++d) t0[d] = (J0[d] - r0c[d]) + tsl[d];
    mat3_vec(Rvi, t0, nt);
#pragma unroll
    for (int d = 0; d < 3; ++d) nt[d] -= (J0[d] - r1c[d]);
    // scrambler `random`: per-joint angle noise along the same axis, translation noise  (scrambler.py:73-79)
    float* po = pose_out + (size_t)b * 48;
#pragma unroll 1
    for (int k = 0; k < 16; ++k) {
        float a[3];
        if (k == 0) { a[0] = aa1[0]; a[1] = aa1[1]; a[2] = aa1[2]; }
        else { a[0] = pose[3 * k]; a[1] = pose[3 * k + 1]; a[2] = pose[3 * k + 2]; }
        if (noise_angle) {
            float n = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            float den = fmaxf(n, 1e-7f);
            float ang = n + noise_angle[(size_t)b * 16 + k];
            a[0] = a[0] / den * ang; a[1] = a[1] / den * ang; a[2] = a[2] / den * ang;
        }
        po[3 * k] = a[0]; po[3 * k + 1] = a[1]; po[3 * k + 2] = a[2];
    }
    if (noise_tsl) {
#pragma unroll
        for (int d = 0; d < 3; ++d) nt[d] += noise_tsl[(size_t)b * 3 + d];
    }
    if (tsl_out) {  // the refiner's inputs when it is not the fused NullRefine (preprocessor.py:76-81)
#pragma unroll
        for (int d = 0; d < 3; ++d) { tsl_out[(size_t)b * 3 + d] = nt[d]; cso_out[(size_t)b * 3 + d] = cso[d]; }
    }
    // rigid map applied by the LBS store: x' = Rf (x + nt + cso)               (refiner.py:139-140, preprocessor.py:84-88)
    float* q = post_rt + (size_t)b * 12;
    float sh[3] = {nt[0] + cso[0], nt[1] + cso[1], nt[2] + cso[2]};
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        q[3 * i] = F[4 * i]; q[3 * i + 1] = F[4 * i + 1]; q[3 * i + 2] = F[4 * i + 2];
        q[9 + i] = F[4 * i] * sh[0] + F[4 * i + 1] * sh[1] + F[4 * i + 2] * sh[2];
    }
}

}  // namespace ab

extern "C" uint64_t ab_pose_generate_workspace_bytes(int batch) {
    return (uint64_t)(batch > 0 ? batch : 0) * (48 + 12) * sizeof(float);
}

extern "C" int ab_pose_generate(const ab_mano_model* model, int batch, const float* hand_pose, const float* hand_shape,
                                const float* hand_tsl, const float* persp_rotmat, const float* camera_free_transf,
                                const float* z_offset, const float* noise_tsl, const float* noise_angle,
                                float* final_obj_pose, float* final_hand_verts, float* final_joints, void* ws,
                                void* stream) {
    AB_REQUIRE(model && model->v_template && model->shapedirs_t && model->posedirs_t && model->j_template &&
                   model->j_shapedirs && model->weights, "null model array");
    AB_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(hand_pose && hand_tsl && persp_rotmat && camera_free_transf && z_offset, "null input");
    AB_REQUIRE(final_obj_pose && final_hand_verts && final_joints && ws, "null output / workspace");
    AB_REQUIRE((noise_tsl == nullptr) == (noise_angle == nullptr), "noise_tsl and noise_angle go together");
    cudaStream_t st = (cudaStream_t)stream;
    float* pose2 = (float*)ws;
    float* post = pose2 + (size_t)batch * 48;
    {
    ab::StageTimer tm(AB_STAGE_POSEGEN_PRELUDE, st);
    ab::posegen_prelude_kernel<<<ab::cdiv(batch, 64), 64, 0, st>>>(*model, batch, hand_pose, hand_shape, hand_tsl,
                                                                  persp_rotmat, camera_free_transf, z_offset,
                                                                  noise_tsl, noise_angle, final_obj_pose, pose2, post,
                                                                  nullptr, nullptr);
    }
    ab::count_launch();
    int rc = ab::check_launch("posegen_prelude_kernel");
    if (rc) return rc;
    // NullRefine decodes with betas=None (refiner.py:138)
    return ab::launch_mano(model, batch, pose2, nullptr, post, -1, final_hand_verts, final_joints, nullptr, st);
}

/* The prelude on its own, for the refiners / scramblers that are not fused into ab_pose_generate. */
extern "C" int ab_pose_prelude(const ab_mano_model* model, int batch, const float* hand_pose, const float* hand_shape,
                               const float* hand_tsl, const float* persp_rotmat, const float* camera_free_transf,
                               const float* z_offset, const float* noise_tsl, const float* noise_angle,
                               float* final_obj_pose, float* pose_out, float* tsl_out, float* cam_sys_offset,
                               float* post_rt, void* stream) {
    AB_REQUIRE(model && model->j_template && model->j_shapedirs, "null model array");
    AB_REQUIRE(batch >= 0, "negative batch");
    if (batch == 0) return AB_OK;
    AB_REQUIRE(hand_pose && hand_tsl && persp_rotmat && camera_free_transf && z_offset, "null input");
    AB_REQUIRE(final_obj_pose && pose_out && tsl_out && cam_sys_offset && post_rt, "null output");
    AB_REQUIRE((noise_tsl == nullptr) == (noise_angle == nullptr), "noise_tsl and noise_angle go together");
    cudaStream_t st = (cudaStream_t)stream;
    {
    ab::StageTimer tm(AB_STAGE_POSEGEN_PRELUDE, st);
    ab::posegen_prelude_kernel<<<ab::cdiv(batch, 64), 64, 0, st>>>(*model, batch, hand_pose, hand_shape, hand_tsl,
                                                                  persp_rotmat, camera_free_transf, z_offset,
                                                                  noise_tsl, noise_angle, final_obj_pose, pose_out, post_rt,
                                                                  tsl_out, cam_sys_offset);
    }
    ab::count_launch();
    return ab::check_launch("posegen_prelude_kernel");
}
