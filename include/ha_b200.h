/*
 * ha_b200.h — C ABI of the B200-native cross-view pose-refinement engine.
 *
 * One shared library (libha_b200.so, sm_100a only) exports exactly the entry points below.
 * They are what a binding of the reference's hot path would call; each one names the
 * reference interface (file:line under the upstream tree) it replaces.  Plain pointers and
 * sizes only — no torch / C++ types.  All `float*` / `void*` data arguments are DEVICE
 * pointers unless the name ends in `_host`.  Every function is asynchronous on `stream`
 * (a cudaStream_t passed as void*), allocates nothing, keeps no mutable global state, and
 * returns 0 on success or a negative HA_E* code (ha_error_string() describes it).
 *
 * Feature layout in HBM: NHWC fp32, i.e. [B][H][W][C] with C contiguous, so that one
 * bilinear tap of one pixel is one contiguous run of 4*C bytes (128-bit loads).
 *
 * There are no environment knobs: the one validation switch (HaLmParams.kernel_variant) is an
 * explicit per-call argument.
 */
#ifndef HA_B200_H_
#define HA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HA_MAX_LEVELS 4

enum {
  HA_OK = 0,
  HA_EINVAL = -1,      /* bad argument (shape / alignment / unsupported channel count)      */
  HA_ENOSPACE = -2,    /* workspace too small                                              */
  HA_ECUDA = -3,       /* a CUDA runtime / driver call failed (see ha_last_cuda_error)     */
  HA_EUNSUPPORTED = -4, /* device is not sm_100                                             */
  HA_ECOMM = -5        /* NCCL missing or an NCCL call failed (see ha_last_cuda_error)      */
};

/* geometry functors of the satellite->ground warp */
enum {
  HA_GEOM_KITTI = 0,   /* models_kitti.py:700-801  LM_S2GP.grd2cam2world2sat               */
  HA_GEOM_FORD = 1,    /* models_ford.py:173-264   LM_S2GP_Ford.cam2body2world2sat         */
  HA_GEOM_G2SP = 2,    /* models_kitti.py:54-160   LM_G2SP.get_warp_sat2real + seq_warp_real2camera:
                          ground features warped onto the satellite plane, residual over the
                          whole satellite map, LM_update of :333-379 (no renormalisation)     */
  HA_GEOM_G2SP_NN = 3  /* models_kitti.py:289-331  LM_G2SP.inplane_grd_to_map (--proj nn): the ground
                          features (VGGUnet_G2S) already live on a square map; the warp is an in-plane
                          rotation about its centre plus a shift in pixels; LM_update as for G2SP   */
};

/* update rule applied by a step to the reduced sums (args.Optimizer; the default everywhere is LM) */
enum {
  HA_OPT_LM = 0,    /* models_kitti.py:939-1041 / models_ford.py:380-466 / G2SP :333-379                          */
  HA_OPT_SGD = 1,   /* LM_S2GP.SGD_update (models_kitti.py:1056-1084): pose -= 0.01 * J^T (2 r), r = s - g with the
                       L2-normalised features (HaLevel.scale = the U-Net's 1/||x||, or NULL for data that is already
                       normalised), all three components, dof = 3, unweighted, no reset draws                  */
  HA_OPT_ADAM = 2,  /* LM_S2GP.ADAM_update (:1086-1124): the same gradient through Adam moments kept in the workspace */
  HA_OPT_GN = 3     /* LM_S2GP_Ford.GN_update (models_ford.py:534-598): s / ||s|| (no clamp), g as is, undamped
                       inverse of J^T W J, reset draws as in LM                                                 */
};

/* device status word bits.  ha_lm_run / ha_lm_step CLEAR *status on entry (on the stream) and OR
 * bits into it; the host reads it ONCE after the loop, never per step (this replaces the
 * reference's per-step host syncs jacobian.py:172,200 / models_kitti.py:1037). */
enum {
  HA_STATUS_NO_INRANGE = 1u, /* some step saw no in-range sample point in the WHOLE batch: the
                                condition of the reference's `assert torch.sum(mask) > 0`
                                (jacobian.py:172).  The host mirrors raise AssertionError.     */
  HA_STATUS_NAN_POSE = 2u,   /* a pose became NaN (the reference prints, models_kitti.py:1037) */
  HA_STATUS_RESET = 4u,      /* an out-of-range shift was re-drawn (models_kitti.py:1030)      */
  HA_STATUS_SAMPLE_EMPTY = 8u, /* some SAMPLE had no in-range point in some step (its H is 0 and
                                its pose does not move; the reference carries on silently)     */
  HA_STATUS_TIMEOUT = 16u    /* a step of ha_lm_run's chained launches gave up waiting for a sample's previous step
                                (never expected; the poses are then invalid).  No reference analogue.  */
};

/* One pyramid level of one branch. */
typedef struct {
  const float* data;   /* [B][H][W][C] fp32 NHWC                                           */
  const float* scale;  /* [B] per-sample multiplier (1/||x||, VGG.py:511-514 L2_norm applied
                          lazily) or NULL for 1                                            */
  int32_t C, H, W;
} HaLevel;

/* Parameters of the LM loop = the fields of the reference's `args` Namespace that the path
 * reads (train_kitti.py:439-476) plus per-level geometry constants. */
typedef struct {
  int32_t geometry;          /* HA_GEOM_*                                                  */
  int32_t n_levels;          /* pyramid levels used by the loop (<= HA_MAX_LEVELS)          */
  int32_t n_iters;           /* args.N_iters                                                */
  int32_t level_first;       /* 0: for iter: for level (models_kitti.py:1176-1180);
                                1: for level: for iter (:1349-1353)                        */
  int32_t dof;               /* 3: (su,sv,theta); 2: shifts only (rotation_range==0);
                                1: theta only (both shift ranges 0) (models_kitti.py:954)  */
  int32_t using_weight;      /* W = grd_conf (models_kitti.py:994-998)                      */
  int32_t use_hessian;       /* damping * diag(H) instead of damping * I (:1005-1010)       */
  int32_t batch;             /* B: samples in every per-sample array of the call             */
  float rotation_range;      /* degrees                                                    */
  float shift_range_lat;     /* metres                                                     */
  float shift_range_lon;     /* metres                                                     */
  float damping[3];          /* resolved lambda per DOF column (:958-966)                   */
  float meter_per_pixel[HA_MAX_LEVELS]; /* satellite metres per pixel at each level, already
                                rounded to fp32 the way the reference's python double is   */
  float inv_meter_per_pixel[HA_MAX_LEVELS]; /* fp32(1/mpp) with mpp in double (models_kitti.py:795) */
  float sat_center[HA_MAX_LEVELS];      /* A/2 (KITTI) or A//2 (Ford, G2SP)                  */
  int32_t ori_grd_h, ori_grd_w;         /* G2SP: size of the ground image `left_camera_k` refers to
                                           (models_kitti.py:111-114); ignored otherwise              */
  int32_t kernel_variant;               /* 0 = default kernels; ha_lm_run chains its step launches for the S2GP geometries
                                           (programmatic stream serialization + per-sample flags: a sample's next step
                                           starts as soon as ITS previous step is solved); 1 = register-staged validation
                                           kernel for the S2GP geometries (same algorithm, other schedule); 2 = the
                                           default step kernel with plain stream-ordered launches also in ha_lm_run.
                                           The parity tests hold all three to the same bar.  No reference analogue.  */
  int32_t optimizer;                    /* HA_OPT_*: the update rule of a step (args.Optimizer)       */
  int32_t full_height;                  /* 0: residual over the bottom half of the ground image (args.proj == 'geo',
                                           models_kitti.py:1194-1199); 1: over the whole image (any other proj:
                                           the polar ground table of models_kitti.py:684-698, :1200-1205)        */
  int32_t adam_level_mult;              /* HA_OPT_ADAM: t = iter * adam_level_mult + level (the reference multiplies
                                           by args.level, models_kitti.py:1241)                                  */
  int32_t adam_iter;                    /* HA_OPT_ADAM with ha_lm_step: the iteration index of this step (ha_lm_run
                                           counts for itself); the moments live in the caller's workspace         */
  float adam_beta1, adam_beta2;         /* HA_OPT_ADAM: args.beta1 / args.beta2 (train_kitti.py:480-481)         */
  int32_t reserved;                     /* must be 0                                                 */
} HaLmParams;

/* ---- library ---------------------------------------------------------------------- */
int ha_version(void);                      /* ABI version, currently 3                     */
const char* ha_error_string(int code);
const char* ha_last_cuda_error(void);      /* text of the last CUDA failure on this thread  */
int ha_device_check(int device);           /* HA_OK iff `device` is compute capability 10.x */
unsigned long long ha_launch_count(void);  /* kernels launched by this library since it was loaded
                                              (statistics for bench.py's gpu_launches; no reference analogue) */

/* ---- layout helpers (host wrappers use them at the boundary; reference is NCHW) ---- */
int ha_nchw_to_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream);
int ha_nhwc_to_nchw(const float* src, float* dst, int B, int C, int H, int W, void* stream);

/* ---- fused LM step: warp + residual + analytic Jacobian + J^T J / J^T r + 3x3 solve ----
 * Replaces, per (iteration, level): project_map_to_grd (models_kitti.py:803-937 /
 * models_ford.py:266-378) -> jacobian.grid_sample (jacobian.py:138-205) -> LM_update
 * (models_kitti.py:939-1041 / models_ford.py:380-466).
 *
 * ground_table: [H][W][4] fp32 = (x, y, z, mask) of the ground-plane lift in the camera
 *   frame (models_kitti.py:655-682 / models_ford.py:110-155); only rows H/2.. are read.
 * grd_conf:  [B][H][W] fp32 or NULL (required iff using_weight).
 * extrinsics: Ford: [B][12] = R_FL row-major (9) then T_FL (3); G2SP: [B][9] = left_camera_k
 *            row-major (models_kitti.py:381); NULL for KITTI S2GP.  ground_table may be NULL for G2SP.
 * pose:      [B][3] = (shift_u, shift_v, theta) normalised units, updated IN PLACE.
 * reset_uv:  [2][B] uniform(-1,1) draws for this step (models_kitti.py:1028-1029) or NULL
 *            (required iff dof == 3).
 * stats:     NULL or [B][HA_STATS] fp32 diagnostics of this step (see HA_STAT_*).
 * status:    device uint32: cleared on entry, then OR-ed with HA_STATUS_* bits.
 */
#define HA_STATS 24
enum { HA_STAT_H = 0 /*9: row-major J~^T W J~*/, HA_STAT_GRAD = 9 /*3: J~^T W r*/, HA_STAT_SAT_NORM = 12,
       HA_STAT_GRD_NORM = 13, HA_STAT_RES_SQ = 14, HA_STAT_DELTA = 15 /*3*/, HA_STAT_N_INRANGE = 18,
       HA_STAT_JTG = 19 /*3: the J~^T W g~ part of GRAD (GRAD = J~^T W s~ - J~^T W g~), kept for the backward pass*/,
       HA_STAT_RESET_MASK = 22 /*bit 0: shift_u was re-drawn in this step, bit 1: shift_v*/ };

size_t ha_lm_workspace_bytes(int B);
int ha_lm_step(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* grd_conf,
               const float* ground_table, const float* extrinsics, float* pose, const float* reset_uv,
               float* stats, uint32_t* status, void* ws, size_t ws_bytes, void* stream);

/* ---- the whole LM loop (forward_iter_first / forward_level_first, mode='test') ----------
 * Replaces models_kitti.py:1176-1283 / :1349-1459 and models_ford.py:652-825 / :868-1000.
 * Launches n_iters * n_levels dependent steps back to back on `stream`, no host sync.
 * sat/grd:   arrays of n_levels HaLevel;  grd_conf / ground_tables: arrays of n_levels ptrs.
 * reset_uv:  [n_steps][2][B] in EXECUTION order, or NULL when dof != 3.
 * pose:      [B][3] initial pose in, final pose out.
 * traj:      [B][n_iters][n_levels][3] pose after every step (the reference's
 *            shift_us / shift_vs / headings stacks, models_kitti.py:1281-1283).
 * stats:     NULL or [n_iters][n_levels][B][HA_STATS].
 */
int ha_lm_run(const HaLmParams* p, const HaLevel* sat, const HaLevel* grd, const float* const* grd_conf,
              const float* const* ground_tables, const float* extrinsics, float* pose,
              const float* reset_uv, float* traj, float* stats, uint32_t* status, void* ws, size_t ws_bytes,
              void* stream);

/* ---- Optimizer 'NN' (LM_S2GP.NN_update, models_kitti.py:1043-1054 + RNNs.NNrefine, RNNs.py:98-126): the one ablation that
 * needs the MATERIALISED residual.  A step = ha_lm_residual (relu = 1: the leading ReLU of NNrefine.linear_k) ->
 * ha_conv3x3_nhwc (linear_k's Conv2d(C, 64) with bias) -> ha_nn_pose_update (spatial mean, the 64-16-3 mapping with its
 * ReLUs and Tanh, pose += delta for all three components, no reset, no RNG draw).
 * ha_lm_residual: out [B][n_px][C] fp32 = (relu of) sat_proj - grd over the residual pixels (the bottom half of the ground
 *   image, or all of it with full_height), features scaled by HaLevel.scale, masks as in models_kitti.py:927,1191.
 * ha_nn_pose_update: x [B][n_px][64] fp32; w0 [16][64], b0 [16], w1 [3][16], b1 [3] (torch Linear layouts, device);
 *   pose [B][3] updated in place; traj_step = &traj[0][it][lv][0] (or NULL), traj_stride floats between samples. */
int ha_lm_residual(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* ground_table,
                   const float* extrinsics, const float* pose, int relu, float* out, void* stream);
int ha_nn_pose_update(const float* x, int B, int n_px, const float* w0, const float* b0, const float* w1, const float* b1, float* pose,
                      float* traj_step, int traj_stride, uint32_t* status, void* stream);

/* ---- backward of ONE fused LM step (training; first slice of SURVEY.md section 8 f-1) -----------------
 * The adjoint of ha_lm_step for the S2GP geometries with all three degrees of freedom and unweighted residuals
 * (the defaults of train_kitti.py / train_ford.py): what `loss.backward()` (train_kitti.py:365) computes through
 * project_map_to_grd -> jacobian.grid_sample -> LM_update of one (iteration, level).  Features must be the
 * L2-normalised ones (HaLevel.scale == NULL); the U-Net backward stays with the caller.
 * pose_in:   [B][3] the pose the forward step started from.
 * stats:     [B][HA_STATS] the forward step's diagnostics (ha_lm_run / ha_lm_step with stats != NULL).
 * gpose_out: [B][3] adjoint of the step's output pose;  gpose_in: [B][3] adjoint of pose_in (written).
 * gsat / ggrd: gradients w.r.t. the satellite / ground features of this level, same layout as the features,
 *            ACCUMULATED into (zero them before the first step of a backward sweep).
 * glambda:   [B][3] adjoint of the damping columns, accumulated.
 */
size_t ha_lm_backward_workspace_bytes(int B);
int ha_lm_step_backward(const HaLmParams* p, int level, const HaLevel* sat, const HaLevel* grd, const float* ground_table,
                        const float* extrinsics, const float* pose_in, const float* stats, const float* gpose_out,
                        float* gpose_in, float* gsat, float* ggrd, float* glambda, void* ws, size_t ws_bytes, void* stream);

/* ---- pose loss: loss_func with loss_method 0 (models_ford.py:1041-1093; KITTI imports it, models_kitti.py:16) ----
 * traj: [B][n_iters][n_levels][3] device (shift_u, shift_v, theta);  gt: [B][3] device, same component order.
 * coe3_host: HOST array of the three loss coefficients in that order.
 * err:  [n_iters][n_levels][3] device = mean_b |traj - gt|;  loss: device scalar = mean_{n,l} sum_k coe[k] err[n][l][k].
 * The rest of the reference's 13-tuple (decreases, last-iteration values) are differences / slices of `err`.
 * Backward: gtraj = (gerr[n][l][k] + gloss coe[k] / (n_iters n_levels)) sign(traj - gt) / B; gerr / gloss may be NULL. */
int ha_pose_loss(const float* traj, const float* gt, int B, int n_iters, int n_levels, const float* coe3_host, float* err,
                 float* loss, void* stream);
int ha_pose_loss_backward(const float* traj, const float* gt, int B, int n_iters, int n_levels, const float* coe3_host,
                          const float* gerr, const float* gloss, float* gtraj, void* stream);

/* ---- VGG16 U-Net feature extractor (VGG.py:13-203, estimate_depth off) ----------------- */
/* Weights, packed by ha_vgg_pack_weights from the reference's state-dict tensors (OIHW fp32,
 * DEVICE pointers). */
#define HA_VGG_N_CONV 17
/* order: conv0 conv2 conv5 conv7 conv10 conv12 conv14 | dec1.1 dec1.3 dec2.1 dec2.3 dec3.1
 * dec3.3 | conf0 conf1 conf2 conf3 ; bias[i] may be NULL (decoder + conf convs have none) */
typedef struct {
  const float* weight[HA_VGG_N_CONV];
  const float* bias[HA_VGG_N_CONV];
} HaVggStateDict;

/* precision of the tensor-core convolutions */
enum {
  HA_CONV_FP32_SIMT = 0,   /* CUDA-core fp32 direct convolution (validation path)            */
  HA_CONV_F16X3 = 1,       /* tcgen05 kind::f16, fp16 hi/lo split, 3 MMAs, fp32-grade result */
  HA_CONV_F16 = 2,         /* tcgen05 kind::f16 single pass (fp16 operands, fp32 accumulate) */
  HA_CONV_F16X3_1CTA = 3   /* HA_CONV_F16X3 on the single-CTA kernels only (cta_group::1): the validation twin of
                              the CTA-pair (cta_group::2) schedule HA_CONV_F16X3 uses where the shape allows      */
};

size_t ha_vgg_packed_weight_bytes(void);
int ha_vgg_pack_weights(const HaVggStateDict* sd, void* packed, size_t packed_bytes, void* stream);

/* out_feat[l] : [B][H/2^(3-l)][W/2^(3-l)][C_l] fp32 NHWC, raw (not L2-normalised), C_l =
 * 256,128,64,16; out_scale[l] : [B] = 1/max(||feat||_2, 1e-12) (VGG.py:511-514), or NULL (the
 * array or an entry) to skip the norm — the S2GP LM step renormalises, so the scale cancels there
 * (models_kitti.py:982-989) and the S2GP eval path does not ask for it;
 * out_conf[l] : [B][H_l][W_l] fp32 = sigmoid(-sigmoid(conv(relu(feat)))) (VGG.py:160-163)
 * or NULL to skip the confidence heads.  n_levels = 3 (level 3) or 4 (level 4). */
size_t ha_vgg_workspace_bytes(int B, int H, int W, int n_levels, int precision);
int ha_vgg_forward(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels,
                   int precision, float* const* out_feat, float* const* out_scale, float* const* out_conf,
                   void* ws, size_t ws_bytes, void* stream);

/* VGGUnet_G2S (VGG.py:206-345; ground branch of LM_G2SP --proj nn): same weights, encoder and workspace as ha_vgg_forward, but
 * the decoders run on the maps folded from [h, w] to [2h, w/2] (a re-interpretation of the row-major pixel order, VGG.py:283-299).
 * out_feat[l] holds the same number of elements as for ha_vgg_forward; read it as [B][2 h_l][w_l / 2][C_l].  out_conf[0] is
 * [B][h_0][w_0] (taken from the un-folded x15, VGG.py:326), out_conf[l >= 1] are [B][2 h_l][w_l / 2].  Tensor-core precisions
 * only; W % 128 == 0, H % 32 == 0. */
int ha_vgg_g2s_forward(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels, int precision,
                       float* const* out_feat, float* const* out_scale, float* const* out_conf, void* ws, size_t ws_bytes,
                       void* stream);

/* ---- training: the same U-Net, keeping what its backward pass needs, and that backward pass ----------
 * Replaces what torch autograd does for VGGUnet.forward under `loss.backward()` (train_kitti.py:365, VGG.py:121-203).
 * ha_vgg_forward_train = ha_vgg_forward in HA_CONV_F16X3 precision whose workspace afterwards holds the post-ReLU input of
 *   every convolution (fp16 hi / lo planes) and the raw conv outputs in front of the three max-pools; keep `ws` untouched
 *   until ha_vgg_backward has run.
 * ha_vgg_backward: g_feat[l] = gradient w.r.t. out_feat[l] of the forward (fp32 NHWC, all n_levels required, zeros where a
 *   level is unused) -> grads->weight[i] ([Cout][Cin][3][3] fp32, torch layout) and grads->bias[i] ([Cout]) (HaVggGrads) for the 13
 *   feature convolutions (entries may be NULL to skip; the confidence heads are not differentiated: they only matter with
 *   using_weight, which keeps the torch path).  sd = the raw OIHW weights (device), img_nchw = the forward's input.
 *   Data gradients run on the forward's tcgen05 convolution kernels with flipped weights, weight gradients on a tcgen05
 *   split-K GEMM over the pixels (csrc/vgg_backward.cu).  n_levels = 3 (level 3 / -1 / 2 models); W % 64 == 0, H % 32 == 0. */
typedef struct {             /* where ha_vgg_backward writes: same order as HaVggStateDict, entries may be NULL */
  float* weight[HA_VGG_N_CONV];
  float* bias[HA_VGG_N_CONV];
} HaVggGrads;
size_t ha_vgg_train_workspace_bytes(int B, int H, int W, int n_levels);
int ha_vgg_forward_train(const void* packed_weights, const float* img_nchw, int B, int H, int W, int n_levels,
                         float* const* out_feat, float* const* out_scale, float* const* out_conf, void* ws, size_t ws_bytes,
                         void* stream);
size_t ha_vgg_backward_workspace_bytes(int B, int H, int W, int n_levels);
int ha_vgg_backward(const HaVggStateDict* sd, const float* img_nchw, int B, int H, int W, int n_levels, void* fwd_ws,
                    const float* const* g_feat, const HaVggGrads* grads, void* ws, size_t ws_bytes, void* stream);

/* ---- one 3x3 / pad 1 / stride 1 convolution layer (the building block of ha_vgg_forward) ----
 * Replaces a single nn.Conv2d call of VGG.py:123-155.  fp32 NHWC in ([B][H][W][cin]) and out
 * ([B][H][W][cout], bias added, no activation); weights in torch OIHW layout, device pointers.
 * Tensor-core precisions need cin % 8 == 0, cout in {16, 32, 64, 128, 256}, W % 16 == 0, H % 8 == 0. */
size_t ha_conv3x3_workspace_bytes(int cin, int cout, int B, int H, int W);
int ha_conv3x3_nhwc(const float* in_nhwc, int cin, const float* w_oihw, const float* bias, float* out_nhwc, int cout,
                    int B, int H, int W, int precision, void* ws, size_t ws_bytes, void* stream);

/* ---- backward of one 3x3 / pad 1 convolution layer (the twin of ha_conv3x3_nhwc; building blocks of ha_vgg_backward) ----
 * x [B][H][W][cin], dy [B][H][W][cout] fp32 NHWC, w OIHW -> dx [B][H][W][cin] (or NULL), dw [cout][cin][3][3] (or NULL),
 * db [cout] (or NULL).  f16x3 precision on tcgen05 (fp32-grade).  cin, cout multiples of 64 (dx: cin in {64, 128, 256}). */
size_t ha_conv3x3_backward_workspace_bytes(int cin, int cout, int B, int H, int W);
int ha_conv3x3_backward_nhwc(const float* x_nhwc, int cin, const float* w_oihw, const float* dy_nhwc, int cout, int B, int H, int W,
                             float* dx_nhwc, float* dw_oihw, float* db, void* ws, size_t ws_bytes, void* stream);

/* ---- input pipeline on the device --------------------------------------------------------
 * Replaces the per-sample PIL / torchvision preparation of dataLoader/KITTI_dataset.py:128-157, :256-288 and
 * dataLoader/Ford_dataset.py:178-209, stage by stage and bit for bit (Pillow's libImaging arithmetic; every stage
 * rounds to uint8, so the stages stay separate launches).  Images are uint8 RGB, [B][H][W][3] packed
 * (src_pixel_bytes = 3, what PNG decoders deliver) or [B][H][W][4] RGBX (src_pixel_bytes = 4, what these functions write).
 *
 * ha_img_affine_u8: Image.transform(size, Image.AFFINE, data, resample) of a batch, one 6-vector per image.
 *   coef:     DEVICE [B][6] float64 = PIL's `data` (destination pixel centre -> source point), as Image.rotate builds
 *             it (round(cos, 15), ...) or (1, 0, tx, 0, 1, ty) for the shifts.
 *   resample: 0 = NEAREST (Image.rotate's default: 16.16 fixed point, Geometry.c affine_fixed), 2 = BILINEAR
 *             (float64, clamped neighbours, truncation; Geometry.c bilinear_filter32RGB); zero fill outside.
 *   dst_rgbx: [B][H][W][4] uint8, or NULL when dst_chw is given.
 *   dst_chw:  NULL, or [B][3][crop_side][crop_side] fp32: TF.center_crop(crop_side) + ToTensor (x / 255) of the
 *             result, written directly (KITTI_dataset.py:150-155, Ford_dataset.py:206-207).
 * ha_img_resize_to_tensor: transforms.Resize([out_h, out_w]) (PIL's antialiased bilinear resample, Resample.c 8bpc:
 *   horizontal then vertical pass in 22-bit fixed point) + ToTensor: [B][H][W] uint8 -> [B][3][out_h][out_w] fp32
 *   (grdimage_transform, KITTI_dataset.py:299-302, Ford_dataset.py:151-154).  Equal sizes = ToTensor alone. */
int ha_img_affine_u8(const uint8_t* src, int src_pixel_bytes, uint8_t* dst_rgbx, float* dst_chw, int crop_side, int B, int H,
                     int W, const double* coef, int resample, void* stream);
size_t ha_img_resize_workspace_bytes(int B, int H, int W, int out_h, int out_w);
int ha_img_resize_to_tensor(const uint8_t* src, int src_pixel_bytes, int B, int H, int W, int out_h, int out_w, float* dst_chw,
                            void* ws, size_t ws_bytes, void* stream);

/* ---- multi-GPU: the ONE collective of the path (the reference has no distributed code) ----
 * Samples are independent, so ranks own contiguous batch shards and exchange nothing until the
 * end: one all-gather of the final (shift_u, shift_v, theta) poses.  NCCL (libnccl.so.2) is bound
 * at run time, from the copy already loaded in the process if there is one (torch's), so a
 * single-GPU consumer needs no NCCL at all.
 * ha_comm_unique_id: rank 0 fills `id128_host` (HA_COMM_ID_BYTES bytes, host memory) and ships it
 *   to the other ranks by any out-of-band means (torch.distributed store, MPI, a file).
 * ha_comm_init: collective over all ranks; `device` is the CUDA ordinal of this rank.
 * ha_pose_allgather: local [n_local][3] fp32 -> all [world * n_local][3] fp32, rank-major, on
 *   `stream`; in place when local == all + rank * n_local * 3 (ha_lm_run can write its final poses
 *   straight into the gather buffer).
 */
#define HA_COMM_ID_BYTES 128
int ha_comm_unique_id(void* id128_host);
int ha_comm_init(void** comm, int world, int rank, const void* id128_host, int device);
int ha_comm_destroy(void* comm);
int ha_pose_allgather(void* comm, const float* local, float* all, int n_local, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HA_B200_H_ */
