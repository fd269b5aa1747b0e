/* artiboost_b200 -- C ABI of the B200-native ArtiBoost synthesis + clasbased-network hot path.
 *
 * The reference (lixiny/ArtiBoost) has no FFI: its boundaries are Python classes.  Each
 * entry point below names the reference interface whose arithmetic it replaces; the Python drop-ins in
 * artiboost_b200/ (ManoLayer, PreProcessorPoseGenerator, Renderer / RendererProvider, IntegralDeconvHead ...)
 * call these through ctypes.  INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / CUDA types in signatures.  `stream` is a cudaStream_t passed as void*
 *    (NULL = legacy default stream).  All data pointers are DEVICE pointers unless the name ends in `_host`.
 *  - nothing allocates: the caller owns every buffer, workspaces are sized by the *_workspace_bytes() queries.
 *  - all calls are asynchronous on `stream` and re-entrant across streams.
 *  - return 0 on success, <0 for an argument error, >0 for a CUDA error code; ab_last_error() gives the text
 *    (thread-local).
 *  - row-major everywhere; matrices are [rows][cols]; 4x4 poses are row-major with translation in column 3.
 */
#ifndef ARTIBOOST_B200_H
#define ARTIBOOST_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define AB_API __attribute__((visibility("default")))
#else
#define AB_API
#endif

#define AB_OK 0
#define AB_ERR_ARG (-1)
#define AB_ERR_UNSUPPORTED (-2)

#define AB_MANO_VERTS 778
#define AB_MANO_JOINTS 16
#define AB_MANO_KEYPOINTS 21
#define AB_MANO_POSE_FEAT 135

AB_API int ab_version(void);
AB_API const char* ab_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches counter) */
AB_API uint64_t ab_launch_count(void);

/* Per-stage device timing for bench.py's roofline line: when enabled, every kernel launch of this library is
 * bracketed by CUDA events on the launching stream; ab_profile_collect waits for them and returns the summed
 * milliseconds and launch counts per stage id (AB_STAGE_*).  Off by default.                                  */
#define AB_STAGE_RASTER_BIN 0
#define AB_STAGE_RASTER_TILE 1
#define AB_STAGE_SYNTH_DRAW 2
#define AB_STAGE_MANO_LBS 3
#define AB_STAGE_POSEGEN_PRELUDE 4
#define AB_STAGE_CCV 5
#define AB_STAGE_VIEW 6
#define AB_STAGE_GEMM 7
#define AB_STAGE_IM2COL 8
#define AB_STAGE_ELEMENTWISE 9
#define AB_STAGE_HEAD_DECODE 10
#define AB_STAGE_CONV_IMPLICIT 11
#define AB_STAGE_WGRAD 12
#define AB_STAGE_TRAIN_ELEMENTWISE 13
#define AB_STAGE_OPTIMIZER 14
#define AB_STAGE_BN_APPLY 15
#define AB_STAGE_BN_BWD_REDUCE 16
#define AB_STAGE_BN_BWD_APPLY 17
#define AB_STAGE_BN_FINALIZE 18
#define AB_STAGE_AUGMENT 19
#define AB_STAGE_CHAMFER 20
#define AB_STAGE_LINEAR_F32 21
#define AB_STAGE_REFINE_MISC 22
#define AB_STAGE_TAIL_LOSS 23
#define AB_STAGE_COUNT 24
AB_API int ab_profile_enable(int on);
AB_API int ab_profile_collect(double* ms_per_stage, int64_t* launches_per_stage, int n_stages);

/* ------------------------------------------------------------------------------------------------ MANO LBS
 * Replaces manotorch.ManoLayer.forward as called at anakin/artiboost/preprocessor.py:25,62,
 * anakin/artiboost/refiner.py:138 and anakin/artiboost/grasp_engine.py:90-95,149-155
 * (algorithm: anakin/postprocess/iknet/manolayer.py:182-276).
 * Model constants are device arrays prepared once by the host (artiboost_b200/manolayer.py):              */
typedef struct {
    const float* v_template;    /* [778*3]                                   */
    const float* shapedirs_t;   /* [10][778*3]   k-major                     */
    const float* posedirs_t;    /* [135][778*3]  k-major                     */
    const float* j_template;    /* [16*3]     = J_regressor . v_template     */
    const float* j_shapedirs;   /* [16*3][10] = J_regressor . shapedirs      */
    const float* weights;       /* [778][16]                                 */
} ab_mano_model;

/* pose [B,48] axis-angle (root first), betas [B,10] or NULL (zeros, the NullRefine case refiner.py:138),
 * post_rt [B,12] or NULL: optional rigid map applied to verts and joints, x' = R x + t, R = post_rt[0:9] row-major.
 * center_idx <0 => none.  Outputs: verts [B,778,3], joints [B,21,3] (MANO 21-keypoint order),
 * transforms_abs [B,16,4,4] or NULL.                                                                       */
AB_API int ab_mano_forward(const ab_mano_model* model, int batch, const float* pose, const float* betas, const float* post_rt,
                    int center_idx, float* verts, float* joints, float* transforms_abs, void* stream);

/* ------------------------------------------------------------------------------------------ CCV-space sampler
 * Replaces OVGSet.update's Categorical draw + row_col_calc (anakin/artiboost/ovg_set.py:104-132,161-170):
 * inverse-CDF draw of n flat cells from weight_map[n_cells] with caller-supplied uniforms u[n] in [0,1);
 * cdf_ws is a float64[n_cells] workspace.  Outputs int32 obj/persp/grasp ids and (optional, may be NULL)
 * occurrence counts int32[n_cells] (ovg_set.py:172-178), which the caller zeroes.                           */
AB_API int ab_ccv_sample(const float* weight_map, int n_obj, int n_persp, int n_grasp, const float* uniforms, int n,
                  double* cdf_ws, int32_t* obj_id, int32_t* persp_id, int32_t* grasp_id, int32_t* occurrence,
                  void* stream);

/* Replaces ViewEngine.get_view (anakin/artiboost/view_engine.py:17-86).  rand4 [n,4] = U[0,1) draws for
 * (u jitter, theta jitter, in-plane roll, z).  Outputs persp_rotmat [n,9], camera_free_transf [n,16], z_offset [n,3]. */
AB_API int ab_view_from_id(const int32_t* persp_id, int n, int u_bins, int theta_bins, float z_min, float z_max,
                    const float* rand4, float* persp_rotmat, float* camera_free_transf, float* z_offset, void* stream);

/* ------------------------------------------------------------------------- fused synthesis draw (one launch)
 * Everything ArtiBoostLoader.generate_render_cache + RenderedDataset.prepare_essential + Renderer.__call__ DRAW for a
 * batch of views (anakin/artiboost/artiboost_loader.py:352-387, ovg_set.py:104-159, view_engine.py:17-86,
 * grasp_engine.py:47-53, scrambler.py:65-81, utils/renderer.py:102-104,125-136), from one counter-based Philox stream:
 * sample i of a call reads subsequence i of (seed, offset); a batch is reproducible from (seed, offset) alone, and
 * successive calls pass offset += AB_SYNTH_UNIFORMS / 4.
 *   ab_ccv_cdf: the fp64 inclusive prefix sum of weight_map[n_cells] -- once per epoch, the map only changes in step_eval.
 *   ab_synth_draw: n samples.  `uniforms` [n, AB_SYNTH_UNIFORMS] (16-byte aligned) replaces the Philox stream when
 *     given (tests; layout at the top of csrc/synth.cu); `uniforms_out` (optional) receives the uniforms used.
 *     Outputs as ab_ccv_sample (+ occurrence counts, accumulated), ab_view_from_id, the grasp table rows (hand_pose [n,48],
 *     hand_shape [n,10], hand_tsl [n,3]), noise_tsl [n,3] / noise_angle [n,16] = N(0, sigma) (optional), hand_tex [n],
 *     light [n] = U[light_lo, light_hi), bg_sel [n,5] = {bg id, x0, y0, crop_w, crop_h} (optional).
 *   ab_ccv_blacklist: _construct_blacklist_map (artiboost_loader.py:415-500): blacklist u8 [n_cells] = th_sgn < threshold
 *     (reference: -0.8); rand2 [n_cells, 2] = the (u, theta) jitter of get_view for every cell, NULL = bin centres;
 *     th_sgn f32 [n_cells] optional.                                                                                */
#define AB_SYNTH_UNIFORMS 32
typedef struct {
    int32_t n_obj, n_persp, n_grasp;   /* CCV space                                                         */
    int32_t u_bins, theta_bins;        /* VIEW_ENGINE PERSP_U_BINS / PERSP_THETA_BINS, n_persp = their product */
    float z_min, z_max;                /* CAMERA_Z_RANGE                                                    */
    const float* grasp_table;          /* device [n_obj, n_grasp, 61]: hand_pose 48 | hand_shape 10 | hand_tsl 3 */
    float tsl_sigma, pose_sigma;       /* SCRAMBLER HAND_TSL_SIGMA / HAND_POSE_SIGMA (yaml:42-45)            */
    int32_t n_hand_tex;                /* renderer.py:102                                                    */
    float light_lo, light_hi;          /* renderer.py:103-104: U(1, 5)                                       */
    int32_t n_bg, bg_h, bg_w, width, height; /* backgrounds and frame size for the crop rule (renderer.py:125-136) */
} ab_synth_space;
AB_API int ab_ccv_cdf(const float* weight_map, int n_cells, double* cdf, void* stream);
AB_API int ab_synth_draw(const ab_synth_space* space, const double* cdf, int n, uint64_t seed, uint64_t offset,
                  const float* uniforms, int32_t* obj_id, int32_t* persp_id, int32_t* grasp_id, int32_t* occurrence,
                  float* hand_pose, float* hand_shape, float* hand_tsl, float* persp_rotmat, float* camera_free_transf,
                  float* z_offset, float* noise_tsl, float* noise_angle, int32_t* hand_tex, float* light, int32_t* bg_sel,
                  float* uniforms_out, void* stream);
AB_API int ab_ccv_blacklist(const ab_synth_space* space, const float* rand2, float threshold, uint8_t* blacklist,
                     float* th_sgn, void* stream);

/* ---------------------------------------------------------------------------------------------- pose generator
 * Replaces PreProcessorPoseGenerator.forward (anakin/artiboost/preprocessor.py:20-99) with the `random`
 * scrambler (scrambler.py:65-81; noise_tsl [B,3] and noise_angle [B,16] are pre-scaled N(0,sigma) draws, NULL = no
 * scrambling) and NullRefine (refiner.py:131-147).
 * Inputs as in the synth_extend dict: hand_pose [B,48], hand_shape [B,10], hand_tsl [B,3], persp_rotmat [B,9],
 * camera_free_transf [B,16], z_offset [B,3].  Outputs final_obj_pose [B,16], final_hand_verts [B,778,3],
 * final_joints [B,21,3].  ws: float workspace of ab_pose_generate_workspace_bytes(B) bytes.                   */
AB_API uint64_t ab_pose_generate_workspace_bytes(int batch);
AB_API int ab_pose_generate(const ab_mano_model* model, int batch, const float* hand_pose, const float* hand_shape,
                     const float* hand_tsl, const float* persp_rotmat, const float* camera_free_transf,
                     const float* z_offset, const float* noise_tsl, const float* noise_angle, float* final_obj_pose,
                     float* final_hand_verts, float* final_joints, void* ws, void* stream);

/* The per-sample prelude of the pose generator on its own (preprocessor.py:20-74: object pose, view-frame hand pose,
 * translation fix-up, optional `random` scrambling), for the refiners / scramblers that are not fused into
 * ab_pose_generate.  Outputs final_obj_pose [B,16], pose_out [B,48] (scrambler_res["hand_pose"]), tsl_out [B,3]
 * (scrambler_res["hand_tsl"]), cam_sys_offset [B,3] (preprocessor.py:38) and post_rt [B,12], the rigid map
 * x' = Rf (x + tsl + cam_sys_offset) of preprocessor.py:84-88 in ab_mano_forward's post_rt layout.             */
AB_API int ab_pose_prelude(const ab_mano_model* model, int batch, const float* hand_pose, const float* hand_shape,
                    const float* hand_tsl, const float* persp_rotmat, const float* camera_free_transf,
                    const float* z_offset, const float* noise_tsl, const float* noise_angle, float* final_obj_pose,
                    float* pose_out, float* tsl_out, float* cam_sys_offset, float* post_rt, void* stream);

/* ---------------------------------------------------------------------------- hand-object refiner (REFINER hand_obj)
 * Replaces point2point_signed (anakin/artiboost/refiner.py:21-85) as HORefiner / _RefineNet call it (:193,:264):
 * for every x[b,i] the nearest point of the sample's object cloud and the Euclidean distance to it (the third-party
 * chamfer_distance kernel + gather + norm of the reference; first minimum wins).  The cloud of sample b is
 * y_points[obj_id[b]] (obj_id NULL: y_points[b]), [*, n_y, 3]; rot (NULL or [B,rot_stride], rot_stride 9 = 3x3,
 * 16 = the rotation block of a 4x4 pose) rotates it first: y = R o, the `verts_object` of refiner.py:190-191, never
 * materialised.  Output dist[b*dist_stride + i] = |x - y_nn| (x scale[i] + shift[i] when given: the eval-mode
 * BatchNorm1d(778) of refiner.py:266 folded in), idx [B,n_x] (may be NULL) = index of the nearest point.
 * ws: ab_chamfer_nn_workspace_bytes(batch, n_x) bytes, 8-byte aligned (per-vertex 64-bit keys through which the CTAs that
 * scan different tiles of one cloud combine their minima).                                                        */
AB_API uint64_t ab_chamfer_nn_workspace_bytes(int batch, int n_x);
AB_API int ab_chamfer_nn(int batch, int n_x, const float* x, int n_y, const float* y_points, const int32_t* obj_id,
                  const float* rot, int rot_stride, const float* scale, const float* shift, float* dist,
                  int64_t dist_stride, int32_t* idx, void* ws, void* stream);

/* The same search over a STATIC cloud that the host has grouped once (artiboost_b200/artiboost/refiner.py
 * build_nn_groups): sorted_points [n_obj, 32 n_groups, 3] = the cloud sorted along a Morton curve (padded by repeating
 * its last point), perm [n_obj, 32 n_groups] = original index of every sorted point, boxes [n_obj, 6, n_groups] =
 * per-group axis-aligned box (lo.x, lo.y, lo.z, hi.x, hi.y, hi.z rows) in the cloud's own frame.  One CTA per sample
 * rotates the cloud into shared memory, one thread per vertex evaluates only the groups whose box can hold the nearest
 * point; distances and indices are bit-identical to ab_chamfer_nn (same arithmetic per evaluated pair, ties to the
 * smallest original index).  rot must be a rotation (the bounds are taken in the cloud's frame).  n_groups <= 480.  */
AB_API int ab_chamfer_nn_grouped(int batch, int n_x, const float* x, int n_groups, const float* sorted_points,
                          const int32_t* perm, const float* boxes, const int32_t* obj_id, const float* rot,
                          int rot_stride, const float* scale, const float* shift, float* dist, int64_t dist_stride,
                          int32_t* idx, void* stream);

/* fp32 linear layer of the RefineNet MLP (nn.Linear + folded BatchNorm1d + LeakyReLU + ResBlock skip,
 * refiner.py:288-319): y[M,N] = act(x[M,K] W[N,K]^T + bias[N] (+ residual[M,N])), act 0 none / 1 leaky relu(slope).
 * Leading dimensions in elements; y may alias residual.                                                          */
AB_API int ab_linear_f32(int M, int N, int K, const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                  const float* residual, int64_t ldr, int act, float slope, float* y, int64_t ldy, void* stream);

/* refiner.py:253-257: feat[b, 0:96] = first two columns of the 16 joint rotation matrices of pose [B,48] (row-major
 * 3x2 per joint), feat[b, 96:99] = tsl [B,3].  ld = row stride of feat in floats.                                 */
AB_API int ab_refine_encode(int batch, const float* pose, const float* tsl, float* feat, int64_t ld, void* stream);

/* parms_decode (refiner.py:88-107): CRot2rotmat + rotmat_to_aa of feat[b, 0:96] -> pose_out [B,48]; feat[b, 96:99] ->
 * tsl_out [B,3] (may be NULL); post_rt [B,12] (may be NULL) = the rigid map of the LBS launch that follows:
 * x' = x + t when rigid is NULL, x' = R (x + t + offset) otherwise (rigid [B,rigid_stride], 9 or 16; offset [B,3]
 * or NULL) -- preprocessor.py:84-88 fused into the refiner's last MANO forward.                                   */
AB_API int ab_refine_decode(int batch, const float* feat, int64_t ld, const float* rigid, int rigid_stride,
                     const float* offset, float* pose_out, float* tsl_out, float* post_rt, void* stream);

/* RandomScrambler2 / RandomScrambler3 (anakin/artiboost/scrambler.py:84-260) over manotorch's AxisLayer: per-joint
 * back / up / left axes from joints [B,21,3] and transforms_abs [B,16,4,4] (MANO chain order), then
 * pose[j] <- aa(R(pose[j]) R(u * splay)) for the four knuckles, pose[j] <- aa(R(l * bend) R(pose[j])) for the 14
 * bending joints, and the thumb base about l then u.  splay [B,4], bend [B,14] (chain joints 1..12, 14, 15; the host
 * expands random_2's five per-finger draws with the interlink coefficients 1 / 1.1 / 0.9), thumb [B,2], all
 * pre-scaled N(0, sigma) draws.  pose_out [B,48]; axes_out [B,15,3,3] (b,u,l rows) or NULL.                       */
AB_API int ab_scramble_anatomical(int batch, const float* pose, const float* joints, const float* transforms_abs,
                           const float* splay, const float* bend, const float* thumb, float* pose_out, float* axes_out,
                           void* stream);

/* --------------------------------------------------------------------------------------------------- rasteriser
 * Replaces Renderer.__call__ (anakin/utils/renderer.py:101-123) over pyrender's OffscreenRenderer
 * (anakin/utils/frender_utils.py:179-205), batched: one call renders `batch` hand+object views.
 * Rule set (pixel-exact on seg/coverage/depth-bits against oracle/raster.c): see DESIGN.md "Raster rules".
 *
 * Mesh patches.  The rasteriser walks a mesh as patches of <= 32 faces over <= 32 vertices (one warp per patch: a lane
 * per vertex, then a lane per face), each with a bounding sphere and a normal cone so that back-facing patches are
 * dropped and the others are placed into 64x64 pixel tiles before any of their vertices is fetched.  This replaces
 * the per-object VBO + GL culling / clipping of pyrender (renderer.py:79-93).  ab_build_patches_host is HOST code
 * (no device needed): it fills caller-owned HOST arrays of `capacity` = ab_patch_capacity(n_faces) patches for ONE
 * mesh; the caller concatenates the first *n_patches rows of every mesh, uploads them and fills ab_patch_table.
 *   pos   f32 [capacity][32][4]  xyz of the lane's vertex + 1.0 (0.0: lane unused); may be NULL for a mesh whose
 *                                 vertices change per view (the hand: the kernel gathers hand_verts through `vid`)
 *   vid   i32 [capacity][32]     mesh-local vertex index of the lane, -1 unused
 *   face  u32 [capacity][32]     patch-local corner lanes a | b << 8 | c << 16, 0xFFFFFFFF unused
 *   prim  i32 [capacity][32]     ORIGINAL mesh-local face index (the z-test tie-break id), -1 unused
 *   bound f32 [capacity][12]     sphere centre xyz, radius | cone axis xyz, min cos to the axis | max perimeter / area,
 *                                 max 1 / (2 area) over the faces (for the snapping margin of the cone test) | #verts, #faces
 * faces: [n_faces][face_stride] i32, the first three entries of a row are used.                              */
typedef struct {
    int32_t n_mesh;
    const int32_t* patch_off_host; /* HOST [n_mesh+1] prefix offsets (in patches) of every mesh                 */
    const float* pos;              /* device [P][32][4] or NULL                                                  */
    const int32_t* vid;            /* device [P][32]                                                             */
    const uint32_t* face;          /* device [P][32]                                                             */
    const int32_t* prim;           /* device [P][32]                                                             */
    const float* bound;            /* device [P][12]                                                             */
} ab_patch_table;
AB_API int ab_patch_capacity(int n_faces);
AB_API int ab_build_patches_host(const float* verts, int n_verts, const int32_t* faces, int n_faces, int face_stride,
                          int capacity, float* pos, int32_t* vid, uint32_t* face, int32_t* prim, float* bound,
                          int32_t* n_patches);

typedef struct {
    int32_t n_obj;              /* number of object meshes                                              */
    const float* obj_verts;     /* [sum V,3] canonical (bbox-centred) vertices, all objects concatenated  */
    const int32_t* obj_faces;   /* [sum F,4] per-object-local vertex ids, 4th component unused (16-byte rows) */
    const uint8_t* obj_colors;  /* [sum V,4] RGBA vertex colours                                        */
    const int32_t* obj_vert_off_host; /* HOST [n_obj+1] prefix offsets into obj_verts / obj_colors       */
    const int32_t* obj_face_off_host; /* HOST [n_obj+1] prefix offsets into obj_faces                    */
    int32_t n_hand_verts, n_hand_faces, n_hand_tex;
    const int32_t* hand_faces;  /* [n_hand_faces,4]                                                     */
    const uint8_t* hand_colors; /* [n_hand_tex, n_hand_verts, 4]                                        */
    const uint8_t* bgs;         /* [n_bg, bg_h, bg_w, bg_channels] or NULL                              */
    int32_t n_bg, bg_h, bg_w;
    int32_t bg_channels;        /* 3 (RGB; 0 means 3) or 4 (RGBX, 4-byte aligned: one 32-bit load per pixel) */
    const ab_patch_table* obj_patches;  /* n_mesh = n_obj, with `pos` (NULL allowed when n_obj == 0)     */
    const ab_patch_table* hand_patches; /* n_mesh = 1, `pos` unused                                      */
} ab_scene;

typedef struct {
    int32_t width, height;
    float fx, fy, cx, cy;       /* renderer.py:76 IntrinsicsCamera(K[0,0], K[1,1], K[0,2], K[1,2])       */
    float znear;                /* 0.05 (pyrender default)                                               */
    int32_t cull_backface;
    float ambient, diffuse;     /* ambient 0.8 (renderer.py:77)                                          */
    int32_t bg_r, bg_g, bg_b;   /* flat background when no bg image is selected (bg_color 0.5 -> 128)    */
} ab_camera;

/* Per-view inputs: hand_verts [B,778,3] camera space, hand_tex [B] texture id (renderer.py:102), obj_id [B]
 * (<0 = CONST.DUMMY: hand only), obj_pose [B,16], light [B] point-light intensity (renderer.py:103-104),
 * bg_sel [B,5] = {bg id (<0 none), x0, y0, crop_w, crop_h} or NULL.
 * obj_id_host: HOST copy of obj_id or NULL.  When given, the binning launch is sized to the largest selected object;
 * when NULL to the largest object in the scene (no host sync either way).
 * Outputs: rgba u8[B,H,W,4], depth f32[B,H,W] (0 = background), seg u8[B,H,W] (0 bg, 1 hand, 2 object); any may
 * be NULL.  ws: workspace of ab_render_workspace_bytes() bytes (per-tile patch lists for `chunk` views); a batch
 * larger than `chunk` is rendered as consecutive groups of `chunk` views on `stream`.  Two kernels per group:
 * raster_bin_kernel (cone cull + tile placement of every patch) and raster_tile_kernel (one CTA per 64x64 tile:
 * z-buffer in shared memory, vertex transform + exact integer coverage per patch, shading and background fill; the
 * only HBM traffic is the per-view inputs, the patch lists and the output image).                           */
AB_API uint64_t ab_render_workspace_bytes(const ab_scene* scene, const ab_camera* cam, int chunk);
AB_API int ab_render_batch(const ab_scene* scene, const ab_camera* cam, int batch, int chunk, const float* hand_verts,
                    const int32_t* hand_tex, const int32_t* obj_id, const int32_t* obj_id_host, const float* obj_pose,
                    const float* light, const int32_t* bg_sel, uint8_t* rgba, float* depth, uint8_t* seg, void* ws,
                    void* stream);

/* ------------------------------------------------------------------------------------- tensor-core contraction
 * The contraction behind every conv / deconv / linear layer of the clasbased network (anakin/models/resnet.py:154-221,
 * simplebaseline.py:152-190, mlp.py:11-25), which the reference hands to cuDNN / cuBLAS through torch.nn.
 * D[M,N] = epilogue(A[M,K] . B[N,K]^T): A, B bf16 row-major (K contiguous, pitches lda / ldb in elements), fp32
 * accumulation in tensor memory (tcgen05).  epilogue: y = acc * scale[n] + bias[n] (+ residual[m,n]) (ReLU) -> bf16
 * (out_fp32 = 0) or fp32 store with pitch ldd.  scale / bias / residual may be NULL.  col_sum / col_sumsq (both or
 * neither): fp32 [ceil(M/128), N] matrices; row t receives the per-column sum / sum of squares of the raw accumulator
 * over output rows 128t .. 128t+127 (the batch statistics of training-mode BatchNorm, summed by ab_bn_finalize).
 * Plain stores, no atomics: nothing to zero, bit-reproducible.
 * K, N, lda, ldb, ldd, ldr multiples of 8; pointers 16-byte aligned.                                           */
AB_API int ab_gemm_bf16(int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                        int out_fp32, const float* scale, const float* bias, const void* residual, int64_t ldr, int relu,
                        float* col_sum, float* col_sumsq, void* stream);

/* Convolution as implicit GEMM (anakin/models/resnet.py:72-152 conv3x3 / 1x1 stride-2 downsample): x bf16 NHWC
 * [B,H,W,C] with C % 64 == 0, w_packed bf16 [Cout, kh*kw*C] with K order (ky, kx, c).  The A operand is fetched by TMA
 * in im2col mode (one filter tap x 64 channels per k-block, padding zero-filled by the TMA unit), so no im2col matrix
 * is ever materialised.  D [B*Ho*Wo, Cout] and the epilogue arguments as in ab_gemm_bf16.                     */
AB_API int ab_conv_bf16_nhwc(const void* x, int B, int H, int W, int C, const void* w_packed, int Cout, int kh, int kw,
                             int stride, int pad, void* D, int64_t ldd, int out_fp32, const float* scale, const float* bias,
                             const void* residual, int64_t ldr, int relu, float* col_sum, float* col_sumsq, void* stream);
/* Rows of the col_sum / col_sumsq partial matrices ab_conv_bf16_nhwc writes for this geometry (one row per output tile):
 * ceil(B*Ho*Wo / 128) for the im2col path, B * ceil(H * (W + 2) / 128) for the halo-resident 3x3 / stride 1 / pad 1
 * path (Cout <= 64; tiles run over the width-padded raster of each image).  ab_bn_finalize takes the count as n_part.            */
AB_API int ab_conv_stat_rows(int B, int H, int W, int C, int Cout, int kh, int kw, int stride, int pad);

/* Weight gradients (the wgrad half of loss.backward() at train/train_artiboost.py:91-93, cuDNN wgrad in the
 * reference).  D(row, col) += sum_p G[p,row] * X[p,col], added into the caller's gradient buffer (which already holds
 * whatever was accumulated before, like autograd's .grad).  The pixel axis is split across CTAs; each split stores its
 * partial tile in the caller-owned workspace ws (ab_wgrad_workspace_bytes(P, Mo, No); for a convolution P = B*Ho*Wo,
 * Mo = Cout, No = kh*kw*C) and a second kernel adds the splits: no atomics, bit-reproducible.
 * G = dY bf16 [P,Mo], X bf16 [P,No] (ab_wgrad_bf16), or X = the NHWC activation read through TMA im2col
 * (ab_conv_wgrad_bf16_nhwc, columns (ky,kx,c), C % 64 == 0; param_layout = 1 writes the nn.Conv2d layout
 * [Cout, C, kh, kw], 0 the packed [Cout, kh*kw*C]).  ab_wgrad_map places (row, col) anywhere: with (hi, lo) =
 * (i / div, i % div), offset = row_hi*s_row_hi + row_lo*s_row_lo + col_hi*s_col_hi + col_lo*s_col_lo, and columns with
 * col_lo >= col_lo_valid (channel padding) are dropped -- so conv, transposed-conv and linear gradients land directly
 * in the parameter's own layout.  The reduction runs over operand rows: both operands are MN-major for tcgen05;
 * the pixel axis is split across CTAs.                                                                          */
typedef struct ab_wgrad_map {
    int32_t row_div, col_div, col_lo_valid;
    int64_t s_row_hi, s_row_lo, s_col_hi, s_col_lo;
} ab_wgrad_map;
AB_API uint64_t ab_wgrad_workspace_bytes(int P, int Mo, int No);
AB_API int ab_wgrad_bf16(int P, int Mo, int No, const void* G, int64_t ldg, const void* X, int64_t ldx, float* D,
                         const ab_wgrad_map* map, void* ws, void* stream);
AB_API int ab_conv_wgrad_bf16_nhwc(const void* x, int B, int H, int W, int C, const void* dy, int Cout, int kh, int kw,
                                   int stride, int pad, float* dw, int param_layout, void* ws, void* stream);

/* ------------------------------------------------------------------- data movement around the contraction (NHWC bf16)
 * Activations are bf16 NHWC ([B,H,W,C], C contiguous) between layers; the reference keeps fp32 NCHW and lets cuDNN
 * pick layouts (anakin/models/resnet.py:199-221).
 * ab_image_to_nhwc: f32 [B,C,H,W] -> bf16 [B,H,W,Cp], channels C..Cp-1 zero.
 * ab_im2col_nhwc:   -> rows [B*Ho*Wo, Kp] with K order (ky, kx, c), columns kh*kw*C..Kp-1 zero; Ho = (H+2*pad-kh)/stride+1.
 * ab_maxpool3x3s2_nhwc: nn.MaxPool2d(3, 2, 1) (resnet.py:157); idx (optional, u8 [B,Ho,Wo,C]) receives the argmax tap
 *   ky*3+kx of each output (first maximum in scan order) for ab_maxpool3x3s2_bwd.  ab_avgpool_nhwc: mean over H*W (resnet.py:219).
 * ab_deconv4x4s2_col2im: ycol f32 [B*H*W, 16*Cout] (column order ky, kx, co) = X . W of a ConvTranspose2d(4, 2, 1)
 *   (simplebaseline.py:161-170) -> out[b, 2H, 2W, Cout] = sum of the 4 contributing taps, * scale + bias, ReLU, bf16;
 *   out_raw (optional) receives the un-normalised f32 sums for training-mode batch statistics.
 * ab_head_decode: logits f32 [B, H*W, ncls*D] (channel = cls*D + d) -> kp3d f32 [B,ncls,3] = (u,v,d) in [0,1),
 *   confd f32 [B,ncls]: IntegralDeconvHead.forward after the final conv (simplebaseline.py:182-190).  One sweep over the
 *   logits when D % 4 == 0 (online softmax).  lse (optional, needs D % 4 == 0) f32 [B,ncls] = log-sum-exp per class, what
 *   ab_head_decode_bwd needs to run in one sweep too.                                                            */
/* ab_pack_conv_filters: bf16 operand copies of an nn.Conv2d weight f32 [Cout,Cin,kh,kw] in one launch: wp [Cout,Kp]
 *   (K order (ky,kx,ci), ci padded to cin_pad, zero tail: the forward / implicit-GEMM filter matrix) and, optionally,
 *   wd [Cin, kh*kw*Cout] (taps flipped, (ky,kx,co) order: the filter matrix of the data gradient).               */
AB_API int ab_pack_conv_filters(const float* w, int Cout, int Cin, int kh, int kw, int cin_pad, int Kp, void* wp, void* wd,
                                void* stream);
AB_API int ab_image_to_nhwc(const float* image, int B, int C, int H, int W, int Cp, void* out, void* stream);
AB_API int ab_im2col_nhwc(const void* in, int B, int H, int W, int C, int kh, int kw, int stride, int pad, int Kp,
                          void* out, void* stream);
AB_API int ab_maxpool3x3s2_nhwc(const void* in, int B, int H, int W, int C, void* out, void* idx, void* stream);
/* The same pooling over y = relu(raw * scale[c] + shift[c]) evaluated on the fly from the RAW convolution output (training-mode
 * BatchNorm + ReLU of the stem, resnet.py:154-157: bn1, relu, maxpool): fp32 fma, ReLU, bf16 rounding per tap as ab_bn_apply
 * stores it, so outputs and argmax taps equal ab_bn_apply followed by ab_maxpool3x3s2_nhwc; the activation is never stored. */
AB_API int ab_maxpool3x3s2_affine_nhwc(const void* raw, int B, int H, int W, int C, const float* scale, const float* shift,
                                       void* out, void* idx, void* stream);
AB_API int ab_avgpool_nhwc(const void* in, int B, int HW, int C, float* out_f32, void* out_bf16, void* stream);
AB_API int ab_deconv4x4s2_col2im(const float* ycol, int B, int H, int W, int Cout, const float* scale, const float* bias,
                                 int relu, void* out_bf16, float* out_raw, void* stream);
AB_API int ab_head_decode(const float* logits, int B, int ncls, int D, int H, int W, float* kp3d, float* confd, float* lse,
                          void* stream);

/* ------------------------------------------------------------------------------------------ training-side kernels
 * What nn.BatchNorm2d (training mode), ReLU, MaxPool2d and autograd do around the convolutions in the reference's
 * train step (anakin/models/resnet.py:72-152, simplebaseline.py:161-190, train/train_artiboost.py:91-96).
 * Activations / activation gradients bf16 [M,C] (NHWC rows), statistics and parameter gradients fp32.
 * Column reductions are deterministic: CTAs write partial rows into a caller-owned workspace ws of
 *   2 * AB_STAT_PARTS * C floats and a second small kernel adds them up; outputs are overwritten, never accumulated.
 * ab_col_stats: sum[c] = sum_r x[r,c], sumsq[c] = sum_r x^2 (sumsq may be NULL).
 * ab_bn_finalize: adds the n_part partial rows ([n_part, C] each; n_part = ceil(M/128) after a convolution, 1 after
 *   ab_col_stats), then mean, biased var -> scale = gamma*invstd, shift = beta - mean*scale, saved mean / invstd, and the
 *   running-stat update running = (1-momentum)*running + momentum*{mean, unbiased var} (running_* may be NULL).
 * ab_bn_apply: y = relu?(raw*scale + shift (+ residual)).
 * ab_bn_bwd_reduce: dy' = dy*(y>0) when relu; dbeta = sum dy', dgamma = sum dy'*xhat, stored (accumulate = 0) or added
 *   (accumulate = 1) into dgamma / dbeta (either may be NULL), plus coef f32 [3, C] = the per-channel coefficients of
 * ab_bn_bwd_apply: dx = gamma*invstd*(dy' - dbeta/M - xhat*dgamma/M) = coef0*dy' + coef1*raw + coef2; dres (optional)
 *   receives dy' for the residual branch.  Both backward passes: with y == NULL and the forward's folded (fwd_scale,
 *   fwd_shift) from ab_bn_finalize the ReLU mask is raw*scale + shift > 0 (layers without a residual input), which
 *   saves reading y.
 * ab_affine_relu_bwd: dx = dy*(y>0)*scale for frozen / eval-mode BatchNorm.
 * ab_dilate2x: zero insertion, turns the data gradient of a stride-2 conv into a stride-1 conv of the dilated dy.
 * ab_deconv4x4s2_gather: dycol[b,iy,ix,(ky,kx,co)] = dy[b,2iy-1+ky,2ix-1+kx,co], the transpose of ab_deconv4x4s2_col2im.
 * ab_head_decode_bwd: gradient of ab_head_decode's kp3d w.r.t. the logits (bf16 [B*H*W, ncls*D]).  With kp3d and lse
 *   (both as written by ab_head_decode; D % 4 == 0) it is a single sweep over the logits; with both NULL, three sweeps.
 * ab_sumsq / ab_adam_step: clip_grad_norm_(max_norm) + torch.optim.Adam on one flat fp32 parameter buffer;
 *   grad_scale multiplies the gradient first (1/world_size after a sum all-reduce).  state f32 [3] = {step count,
 *   1 - beta1^t, 1 - beta2^t} lives on the device (zero it once) and is advanced by the call itself, so a captured
 *   CUDA graph of the training step replays with the right bias corrections.  n % 4 == 0.
 *   ab_sumsq: out is f32 [1 + AB_SUMSQ_PARTS]; out[0] receives the sum (out[1..] hold the per-CTA partials of a
 *   fixed-order two-pass reduction: no atomics, the result is bit-reproducible and needs no zeroing).            */
#define AB_SUMSQ_PARTS 1184
#define AB_STAT_PARTS 1184
AB_API int ab_col_stats(const void* x, int is_f32, int M, int C, int64_t ld, float* sum, float* sumsq, float* ws, void* stream);
AB_API int ab_bn_finalize(const float* sum_part, const float* sumsq_part, int n_part, int C, float count, const float* gamma,
                          const float* beta, float eps, float momentum, float* scale, float* shift, float* save_mean,
                          float* save_invstd, float* running_mean, float* running_var, void* stream);
AB_API int ab_bn_apply(const void* raw, int64_t M, int C, const float* scale, const float* shift, const void* residual,
                       int relu, void* y, void* stream);
AB_API int ab_bn_bwd_reduce(const void* dy, const void* y, const void* raw, int M, int C, const float* gamma,
                            const float* mean, const float* invstd, int relu, float* dgamma, float* dbeta, int accumulate,
                            float* coef, float* ws, const float* fwd_scale, const float* fwd_shift, void* stream);
AB_API int ab_bn_bwd_apply(const void* dy, const void* y, const void* raw, int64_t M, int C, const float* coef, int relu,
                           void* dx, void* dres, const float* fwd_scale, const float* fwd_shift, void* stream);
AB_API int ab_affine_relu_bwd(const void* dy, const void* y, int64_t M, int C, const float* scale, int relu, void* dx,
                              void* dres, void* stream);
AB_API int ab_maxpool3x3s2_bwd(const void* idx, const void* dy, int B, int H, int W, int C, void* dx, void* stream);
AB_API int ab_avgpool_bwd(const float* dmean, int B, int HW, int C, void* dx, void* stream);
AB_API int ab_dilate2x(const void* in, int B, int Ho, int Wo, int H, int W, int C, void* out, void* stream);
AB_API int ab_deconv4x4s2_gather(const void* dy, int B, int H, int W, int C, void* dycol, void* stream);
AB_API int ab_head_decode_bwd(const float* logits, const float* dkp3d, const float* kp3d, const float* lse, int B, int ncls, int D,
                              int H, int W, void* dlogits,
                              void* stream);
AB_API int ab_sumsq(const float* g, int64_t n, float* out, void* stream);
AB_API int ab_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                        float weight_decay, float* state, const float* grad_sumsq, float max_norm, float grad_scale, void* stream);

/* ------------------------------------------------------------------------ clasbased tail + criterion (training)
 * ab_tail_losses: the HybridBaseline tail (anakin/models/hybridbaseline.py:41-96: uvd -> xyz utils/transform.py:512-546,
 *   6D -> rotation :578-598, box corners, corner projection) and the Criterion of the clasbased configs
 *   (anakin/criterions/criterion.py:57-67 over jointloss.py:14-67, ordinal.py:75-306, symcornerloss.py:18-108), forward and
 *   gradient in one launch.  Inputs f32: kp3d [B,22,3] (21 joints + box root, the head's soft-argmax), rot6d [B,6],
 *   root_joint [B,3], cam_intr [B,3,3], corners_can [B,8,3], targets joints_3d [B,21,3] / corners_3d [B,8,3] (root relative),
 *   joints_vis [B,21], corners_vis [B,8].  Random draws of the ordinal losses are inputs (drawn by the caller in the
 *   criterion's order): vv_hand [n_views_hand,3], jp [n_pairs_joint,2] joint pairs, pp [n_pairs_part,2] part pairs (parts are
 *   numbered joint - 1), vv_scene [n_views_scene,3], hp [n_pairs_scene,2] (joint, corner).  SymCornerLoss: sym_R
 *   [n_obj,n_sym,3,3], sym_t [n_obj,n_sym,3] (metres), obj_idx [B] (1-based), obj_transf [B,4,4].  A weight of 0 disables a
 *   term (its inputs may be NULL); weights are criterion lambda x the loss's own lambda.
 *   Outputs f32: the seven tensors of the reference's forward (joints_abs [B,21,3], corners_abs [B,8,3], joints_rel,
 *   corners_rel (minus joint center_idx), uvd2d [B,30,3], boxroot [B,1,3], rotmat [B,3,3]); parts [8] = {joints_3d_loss,
 *   corners_3d_loss, joint_ord_loss, part_ord_loss, scene_ord_loss, sym_corners_3d_loss, 0, weighted total};
 *   d_kp3d [B,22,3], d_rot6d [B,6] = d total / d input.  Fixed-order reductions, no atomics (bit-reproducible).
 *   ws: ab_tail_losses_workspace_bytes(B).                                                                        */
typedef struct {
    int32_t batch, center_idx;
    float inp_w, inp_h;          /* DATA_PRESET.IMAGE_SIZE: scales kp3d's u, v                            */
    float img_w, img_h;          /* network input width / height: normalises the projected corners       */
    float depth_range;           /* 0.4 (utils/transform.py:516)                                          */
    float w_joints, w_corners, w_joint_ord, w_part_ord, w_scene_ord, w_sym;
    int32_t n_views_hand, n_pairs_joint, n_pairs_part, n_views_scene, n_pairs_scene, n_sym, sym_ho3d;
} ab_tail_cfg;
AB_API uint64_t ab_tail_losses_workspace_bytes(int batch);
AB_API int ab_tail_losses(const ab_tail_cfg* cfg, const float* kp3d, const float* rot6d, const float* root_joint,
                          const float* cam_intr, const float* corners_can, const float* joints_3d, const float* corners_3d,
                          const float* joints_vis, const float* corners_vis, const float* vv_hand, const int32_t* jp,
                          const int32_t* pp, const float* vv_scene, const int32_t* hp, const float* sym_R, const float* sym_t,
                          const int32_t* obj_idx, const float* obj_transf, float* joints_abs, float* corners_abs,
                          float* joints_rel, float* corners_rel, float* uvd2d, float* boxroot, float* rotmat, float* parts,
                          float* d_kp3d, float* d_rot6d, void* ws, void* stream);

/* ------------------------------------------------------------------------------------------- crop / augment
 * RenderedDataset.__getitem__ for a batch of rendered views (anakin/artiboost/rendered_dataset.py:127-133,155-274;
 * utils/transform.py:425-470; utils/img_augment.py:6-80; datasets/hodata.py:161-186), which the reference runs per
 * sample on the CPU with PIL inside the DataLoader workers.  rgba u8 [B,raw_h,raw_w,4] (the rasteriser's output),
 * joints f32 [B,21,3] and obj_pose f32 [B,4,4] in camera space, corners_can f32 [B,8,3] -> image f32 [B,3,out_h,out_w]
 * (x/255 - 0.5) and the transformed annotations with the reference's sample keys.  Every random draw is an input:
 * draws f32 [B, AB_AUG_DRAWS] = {centre jitter x, y in U(-1,1); scale jitter ~ N(0, scale_jit/3); cos, sin of the in-plane
 * rotation; GaussianBlur radius (<= 0.1 in the reference); brightness, contrast, saturation factors; hue factor} and
 * order i32 [B,4] = execution order of the colour operations as indices into {brightness, saturation, hue, contrast}.
 * Byte-exact Pillow arithmetic (blur, enhancers, HSV round trip, NEAREST AFFINE warp): see oracle/augment.py.
 * affine / inv_affine (optional, f32 [B,6]) return the forward / inverse 2x3 matrices; status (optional, device i32):
 * bit 0 = a blur radius outside the supported range (box radius >= 1), bit 1 = warp outside Pillow's fixed-point range.
 * ws: ab_augment_workspace_bytes(cfg, B), 256-byte aligned.                                                        */
#define AB_AUG_DRAWS 10
typedef struct ab_augment_cfg {
    int32_t raw_w, raw_h, out_w, out_h;
    int32_t center_idx;     /* DATA_PRESET.CENTER_IDX */
    int32_t crop_model;     /* 0 root_obj, 1 hand_obj, 2 hand (DATA_PRESET.CROP_MODEL) */
    int32_t full_image;     /* DATA_PRESET.FULL_IMAGE */
    int32_t aug;            /* cfg_dataset.AUG */
    float bbox_expand_ratio, center_jit, scale_jit;
    float K[9];             /* render camera intrinsics, row-major */
} ab_augment_cfg;
AB_API uint64_t ab_augment_workspace_bytes(const ab_augment_cfg* cfg, int B);
AB_API int ab_crop_augment(const ab_augment_cfg* cfg, int B, const uint8_t* rgba, const float* joints, const float* obj_pose,
                           const float* corners_can, const float* draws, const int32_t* order, float* image, float* cam_intr,
                           float* root_joint, float* joints_3d, float* joints_2d, float* joints_vis, float* corners_3d,
                           float* corners_2d, float* corners_vis, float* obj_transf, float* affine, float* inv_affine,
                           int32_t* status, void* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif
