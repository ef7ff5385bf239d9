/*
 * neo360_b200 -- C ABI of the B200-native NeO-360 ray-marching hot path.
 *
 * The reference (zubair-irshad/NeO-360) is pure Python; it has no FFI / operator registry.  The seam
 * this library sits behind is the Python call `self.model(rays, randomized, white_bkgd, near, far,
 * out_depth)` made by models/neo360/model.py:725-732, 841-843, 882-884.
 * `neo360_b200/renderer.py` mirrors that call and binds the entry points below through ctypes;
 * INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless marked HOST; the caller (PyTorch)
 *     owns all buffers, the library borrows them for the duration of the call;
 *   - the only library-owned object is the opaque NeoScene (re-laid-out feature maps + packed weights),
 *     released with neo_scene_free;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*); no call synchronises except
 *     neo_scene_create (once per scene) and neo_check_async;
 *   - return value 0 = ok, negative = NeoStatus; neo_last_error() gives the message (per thread);
 *     nothing throws across the ABI.  One caller thread per process / GPU.
 */
#ifndef NEO360_B200_H
#define NEO360_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    NEO_OK = 0,
    NEO_ERR_INVALID = -1,   /* bad argument */
    NEO_ERR_CUDA = -2,      /* CUDA runtime error, see neo_last_error */
    NEO_ERR_WORKSPACE = -3, /* workspace too small */
    NEO_ERR_GEOMETRY = -4,  /* a ray misses the unit sphere: the reference asserts (helper.py:271,426) */
    NEO_ERR_UNSUPPORTED = -5
} NeoStatus;

/* arithmetic of the density/colour MLP */
typedef enum {
    NEO_PREC_FP32 = 0, /* CUDA-core fp32, reference formulation; tightest parity (debug / validation) */
    NEO_PREC_TC = 1    /* tcgen05 tensor cores, 16-bit operands, fp32 accumulate (the fast path) */
} NeoPrecision;

/* One NeRFPPMLP (models/neo360/model.py:37-158).  nn.Linear layout: weight (out,in) row-major, bias (out). */
typedef struct {
    int in_ch;              /* 3 = fg (63-d pos-enc), 4 = bg (84-d) */
    const float *w0, *b0;   /* pts_linears.0   (128, 63|84 + 512 + 128) */
    const float *w1, *b1;   /* pts_linears.1   (128,128) */
    const float *w2, *b2;   /* pts_linears.2   (128,128) */
    const float *w3, *b3;   /* pts_linears.3   (128, 128 + 63|84 + 512 + 128) */
    const float *wb, *bb;   /* bottleneck_layer (128,128) */
    const float *wsig, *bsig; /* density_layer (1,128) */
    const float *wv0, *bv0; /* views_linear.0  (64, 128+27) */
    const float *wv1, *bv1; /* views_linear.1  (64,64) */
    const float *wrgb, *brgb; /* rgb_layer     (3,64) */
} NeoMLPParams;

/* What the (out-of-scope) encoder produced for one scene + the source cameras.
 * Replaces: encoder outputs consumed by index_grid (encoder_tp_fusion_conv.py:122-209) and
 * SpatialEncoder.index (encoder_pn.py:101-152); rays["src_poses"|"src_focal"|"src_c"|"src_imgs"] of
 * NeRF_TP.forward (model.py:266-274). */
typedef struct {
    int nv;                       /* number of source views (reference: 3) */
    int plane_h, plane_w;         /* tri-plane size (reference: 120 x 160) */
    int world_ch;                 /* 128 */
    int lat_h, lat_w, local_ch;   /* pixel-aligned latent (reference: H/2, W/2, 512) */
    int img_w, img_h;             /* src_imgs.shape[-1], [-2] (model.py:267-269) */
    const float* planes_xz;       /* (nv, world_ch, plane_h, plane_w) NCHW */
    const float* planes_xy;
    const float* planes_yz;
    const float* latent;          /* (nv, local_ch, lat_h, lat_w) NCHW */
    const float* src_poses;       /* (nv,4,4) camera-to-world */
    const float* src_focal;       /* (nv,)  only [0] is used (model.py:242) */
    const float* src_c;           /* (nv,2) only [0] is used (model.py:244) */
} NeoSceneDesc;

typedef struct NeoScene NeoScene;

/* Build the per-scene state: channel-last feature maps, R^T / -R^T t per view, MLP weights packed for
 * the selected precision.  mlps[4] = {fg_coarse, bg_coarse, fg_fine, bg_fine} (model.py:215-237).
 * `precision_mask` is a bit-or of (1<<NEO_PREC_FP32) | (1<<NEO_PREC_TC): which paths to prepare; 0 = cameras and grid geometry only
 * (enough for neo_index_maps / neo_index_maps_bwd; rendering such a scene returns NEO_ERR_INVALID). */
int neo_scene_create(const NeoSceneDesc* desc, const NeoMLPParams mlps[4], int precision_mask,
                     NeoScene** out, void* stream);
void neo_scene_free(NeoScene* scene);
/* bytes of device memory held by the scene */
size_t neo_scene_bytes(const NeoScene* scene);
/* neo_scene_free keeps the device blocks of a destroyed scene (per device, up to 6 GB in total) for the next scene of the same shape, so
 * that a scene change costs its kernels and not cudaMalloc / cudaFree.  This returns every kept block to the driver. */
void neo_release_cached(void);

/* rays of one call: rays["rays_o"|"rays_d"|"viewdirs"] (nerds360_ae.py:1007-1023). */
typedef struct {
    int n_rays;
    int chunk;               /* the reference's --chunk (opt.py:195-200): rays are conditioned on the view
                                direction of ray ((b*N+s) mod B) of their own chunk (quirk Q1, model.py:358-360);
                                chunk <= 0 means one chunk = n_rays (what a direct model(...) call does) */
    const float* rays_o;     /* (n_rays,3) */
    const float* rays_d;     /* (n_rays,3) */
    const float* viewdirs;   /* (n_rays,3) */
    const int* ray_order;    /* optional (n_rays) permutation used only for locality: the TC kernel's g-th tile covers rays
                                ray_order[32g..32g+31] (e.g. 8x4 pixel blocks of a frame); results are unchanged.  NULL = identity */
} NeoRays;

typedef struct {
    int n_coarse;            /* NeRF_TP.num_coarse_samples (level 0 evaluates n_coarse+1 points, quirk Q6) */
    int n_fine;              /* NeRF_TP.num_fine_samples   (level 1 evaluates n_coarse+1+n_fine points) */
    int white_bkgd;          /* only honoured when out_depth == 0 (model.py:501,519 vs 551,560) */
    int out_depth;           /* 1: eval tuple, 0: train tuple */
    int precision;           /* NeoPrecision */
    /* randomized=True: uniforms the reference would draw with torch.rand (helper.py:50,199); NULL = deterministic */
    const float* u_fg0;      /* (n_rays, n_coarse+1) */
    const float* u_bg0;      /* (n_rays, n_coarse+1) */
    const float* u_fg1;      /* (n_rays, n_fine) */
    const float* u_bg1;      /* (n_rays, n_fine) */
} NeoCfg;

/* Per level l in {0,1}; N_l = n_coarse+1 (+ n_fine).  NULL pointers are skipped.
 * eval tuple  (model.py:525-527): comp_rgb, fg_rgb, bg_rgb, fg_acc, bg_lambda, depth
 * train tuple (model.py:577-579): comp_rgb, fg_w, bg_w, fg_sdist, bg_sdist, bg_acc */
typedef struct {
    float* comp_rgb[2];   /* (n_rays,3) */
    float* fg_rgb[2];     /* (n_rays,3) */
    float* bg_rgb[2];     /* (n_rays,3) */
    float* fg_acc[2];     /* (n_rays)   */
    float* bg_lambda[2];  /* (n_rays,1) */
    float* depth[2];      /* (n_rays)   */
    float* bg_acc[2];     /* (n_rays)   */
    float* fg_w[2];       /* (n_rays,N_l) */
    float* bg_w[2];       /* (n_rays,N_l) */
    float* fg_sdist[2];   /* (n_rays,N_l) */
    float* bg_sdist[2];   /* (n_rays,N_l) */
    /* optional debug taps (parity tests): per-sample values of level l */
    float* fg_t[2];       /* (n_rays,N_l) */
    float* bg_s[2];       /* (n_rays,N_l) */
    float* fg_sigma[2];   /* (n_rays,N_l) */
    float* bg_sigma[2];   /* (n_rays,N_l) */
    float* fg_rgb_s[2];   /* (n_rays,N_l,3) */
    float* bg_rgb_s[2];   /* (n_rays,N_l,3) */
} NeoOut;

/* Workspace (device) the caller must provide to neo_render_fwd for n_rays rays. */
size_t neo_render_workspace_bytes(int n_rays, const NeoCfg* cfg);

/* NeRF_TP.forward with the encoder hoisted (model.py:266-581 minus :272-274).  Replaces the call at
 * model.py:725-732 / 841-843 / 882-884. */
int neo_render_fwd(const NeoScene* scene, const NeoRays* rays, const NeoCfg* cfg, NeoOut* out,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Synchronise `stream` and report deferred device-side errors (the reference's two asserts,
 * helper.py:271,426, are host syncs; here they are a flag checked on demand). */
int neo_check_async(const NeoScene* scene, void* stream);

/* ---- stage-level entry points (mirror the reference helpers one to one; used by the parity tests) ---- */

/* datasets/ray_utils.py:84-104 + 133-176: pixel grid + c2w (3,4 row-major, device) -> rays. */
int neo_get_rays(int H, int W, float focal, const float* c2w, float* rays_o, float* viewdirs, float* rays_d,
                 float* radii, void* stream);
/* datasets/nerds360_ae.py:730-748 (train __getitem__): `n` sampled pixels of `n_views` target views.  pix_inds (n) int64 index the
 * flattened (n_views, H, W) stack exactly as the reference's `torch.randint(0, T*H*W)` does; c2w (n_views,3,4); images
 * (n_views,H,W,3) fp32 or NULL.  Outputs (n,3)/(n) as neo_get_rays, target (n,3) = images[pix]; any output may be NULL.  Each ray is
 * bit-identical to the same pixel of neo_get_rays.  Out-of-range indices raise *err_flag (device int) = NEO_ERR_INVALID. */
int neo_sample_rays(int n, const long long* pix_inds, int n_views, int H, int W, float focal, const float* c2w,
                    const float* images, float* rays_o, float* viewdirs, float* rays_d, float* radii, float* target,
                    int* err_flag, void* stream);
/* models/neo360/helper.py:253-273 */
int neo_intersect_sphere(const float* rays_o, const float* rays_d, int n_rays, float* far, int* err_flag,
                         void* stream);
/* models/neo360/helper.py:24-75.  in_sphere=1: t (n,N+1), pts (n,N+1,3).  in_sphere=0: t=s (n,N+1) descending,
 * pts (n,N+1,4), pts_linear (n,N+1,3).  u_rand NULL = deterministic. */
int neo_sample_along_rays(const float* rays_o, const float* rays_d, const float* far, int n_rays, int num_samples,
                          int in_sphere, float far_uncontracted, const float* u_rand, float* t_vals, float* pts,
                          float* pts_linear, void* stream);
/* models/neo360/helper.py:218-249 (sorted_piecewise_constant_pdf 174-215 inside): bins=mids(t_old), weights[1:-1]. */
int neo_sample_pdf(const float* rays_o, const float* rays_d, const float* far, const float* t_old,
                   const float* weights, int n_rays, int n_old, int num_samples, int in_sphere,
                   float far_uncontracted, const float* u_rand, float* t_vals, float* pts, float* pts_linear,
                   void* stream);
/* models/neo360/helper.py:128-171.  rgb (n,N,3), sigma (n,N), t (n,N). */
int neo_volumetric_rendering(const float* rgb, const float* sigma, const float* t_vals, const float* rays_d,
                             const float* far, int n_rays, int N, int white_bkgd, int in_sphere, float* comp_rgb,
                             float* acc, float* weights, float* bg_lambda, float* depth, void* stream);
/* The same two lookups over CALLER-OWNED channel-last maps (nv,H,W,C) fp32 of any channel count C % 4 == 0, with the scene's cameras and
 * grid geometry (spatial sizes = the scene's latent / plane sizes).  Used by the training path, which looks up maps projected through the
 * current first / skip layer weights (linearity of encoder_tp_fusion_conv.py:122-209 and encoder_pn.py:101-152).  latent_cl or the three
 * planes may be NULL (then that output is skipped).  out_local / out_world: (nv*M, C), rows ordered (view, point). */
int neo_index_maps(const NeoScene* scene, const float* pts, int M, int C, const float* latent_cl, const float* xz_cl, const float* xy_cl,
                   const float* yz_cl, float* out_local, float* out_world, void* stream);
/* Backward of neo_index_maps: scatter-add of the row gradients (nv*M, C) into zero-initialised channel-last gradient maps. */
int neo_index_maps_bwd(const NeoScene* scene, const float* pts, int M, int C, const float* g_local, const float* g_world,
                       float* g_latent_cl, float* g_xz_cl, float* g_xy_cl, float* g_yz_cl, void* stream);
/* encoder_tp_fusion_conv.py:122-209: pts (M,3) world -> (nv*M,128), rows ordered (view, point). */
int neo_index_grid(const NeoScene* scene, const float* pts, int M, float* out, void* stream);
/* model.py:239-264 (get_local_feats): pts (M,3) world -> (nv*M,512). */
int neo_index_local(const NeoScene* scene, const float* pts, int M, float* out, void* stream);
/* Output side (models/interface.py:53-61, LitModel.psnr_each): *out_sum (device double) = sum_i (clip(pred_i,0,1) - clip(gt_i,0,1))^2 over n
 * floats; PSNR = -10 log10(out_sum / n). */
int neo_clipped_sq_err(const float* pred, const float* gt, long long n, double* out_sum, void* stream);

/* ---- backward of the hand-written stages (training: models/neo360/model.py:697-820 differentiates through this path) ----
 * The field's backward is split: the lookups' scatter and the compositing backward are the entry points below; the dense layers of
 * NeRFPPMLP are differentiated by the host framework (plain library GEMMs) in neo360_b200/training.py.  Sample positions carry no
 * gradient (the reference detaches them, helper.py:225). */
/* d(volumetric_rendering)/d(rgb, sigma): upstream gradients of comp_rgb (n,3), acc (n), weights (n,N), bg_lambda (n), depth (n) -- any
 * may be NULL -- -> d_rgb (n,N,3), d_sigma (n,N).  Same rgb / sigma / t / rays_d / far as the forward call. */
int neo_volumetric_rendering_bwd(const float* rgb, const float* sigma, const float* t_vals, const float* rays_d, const float* far,
                                 int n_rays, int N, int white_bkgd, int in_sphere, const float* g_comp_rgb, const float* g_acc,
                                 const float* g_weights, const float* g_bg_lambda, const float* g_depth, float* d_rgb, float* d_sigma,
                                 void* stream);
/* d(index_grid)/d(planes): g_out (nv*M,128) -> ACCUMULATES into channel-last gradient maps (nv, plane_h, plane_w, 128) x3 (zeroed by the caller). */
int neo_index_grid_bwd(const NeoScene* scene, const float* pts, int M, const float* g_out, float* g_planes_xz, float* g_planes_xy,
                       float* g_planes_yz, void* stream);
/* d(get_local_feats)/d(latent): g_out (nv*M,512) -> ACCUMULATES into the channel-last gradient map (nv, lat_h, lat_w, 512). */
int neo_index_local_bwd(const NeoScene* scene, const float* pts, int M, const float* g_out, float* g_latent, void* stream);
/* `predict` (model.py:343-407) for one branch of one level: t/s (n,N) -> rgb (n,N,3), sigma (n,N).
 * mlp_index in 0..3 = {fg_coarse,bg_coarse,fg_fine,bg_fine}; is_bg selects the NeRF++ background parametrisation. */
int neo_field_eval(const NeoScene* scene, const NeoRays* rays, const float* far, const float* t_vals, int N,
                   int mlp_index, int precision, float* rgb, float* sigma, void* stream);

/* ---- vanilla two-level NeRF: models/vanilla_nerf/model.py:44-216 ---- */
typedef struct {
    const float* w[8];        /* pts_linears.{0..7}.weight: (256,63) (256,256)x4 (256,319) (256,256)x2 */
    const float* b[8];
    const float *wb, *bb;     /* bottleneck_layer (256,256) */
    const float *wsig, *bsig; /* density_layer (1,256) */
    const float *wv0, *bv0;   /* views_linear.0 (128, 256+27) */
    const float *wrgb, *brgb; /* rgb_layer (3,128) */
} NeoVanillaMLPParams;
typedef struct NeoVanilla NeoVanilla;
typedef struct {
    int n_coarse, n_fine, white_bkgd;
    float near_plane, far_plane;   /* the scalar near / far the caller passes to NeRF.forward (model.py:154) */
    const float* u0;               /* randomized: (n_rays, n_coarse+1) uniforms, helper.py:438; NULL = deterministic */
    const float* u1;               /* randomized: (n_rays, n_fine) uniforms, helper.py:587 */
    int precision;                 /* NeoPrecision: NEO_PREC_FP32 = fused fp32 CUDA-core field kernel (tight parity), NEO_PREC_TC = NeRFMLP layer by layer
                                      on tcgen05 (csrc/gemm_tc.cu), fp16 weights / activations */
} NeoVanillaCfg;
typedef struct {
    float* comp_rgb[2];  /* (n_rays,3)   model.py:214 returns (comp_rgb, acc, depth) per level */
    float* acc[2];       /* (n_rays) */
    float* depth[2];     /* (n_rays) */
    float* t[2];         /* optional debug taps: (n_rays,N_l) */
    float* sigma[2];
    float* rgb_s[2];     /* (n_rays,N_l,3) */
    float* weights[2];
} NeoVanillaOut;
/* mlps[2] = {coarse_mlp, fine_mlp} */
int neo_vanilla_create(const NeoVanillaMLPParams mlps[2], NeoVanilla** out, void* stream);
void neo_vanilla_free(NeoVanilla* v);
size_t neo_vanilla_workspace_bytes(int n_rays, const NeoVanillaCfg* cfg);
/* NeRF.forward (models/vanilla_nerf/model.py:154-216); rays->chunk is ignored (no cross-ray coupling in this model) */
int neo_vanilla_render_fwd(const NeoVanilla* v, const NeoRays* rays, const NeoVanillaCfg* cfg, NeoVanillaOut* out,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- Mip-NeRF 360: models/mipnerf360/model.py:30-365 ---- */
typedef struct {
    int depth, width;          /* PropMLP: 4 x 256 (density only), NeRFMLP: 8 x 1024 (model.py:176-195) */
    const float* basis;        /* pos_basis_t (3,21) */
    const float* w[8];         /* pts_linear.{i}.weight: (width,504), (width,width)..., layer 5 of a depth-8 MLP is (width, width+504) */
    const float* b[8];
    const float *wsig, *bsig;  /* density_layer (1,width) */
    const float *wb, *bb;      /* bottleneck_layer (256,width)      -- NULL for a PropMLP (disable_rgb) */
    const float *wv0, *bv0;    /* views_linear.0 (128, 256+27) */
    const float *wrgb, *brgb;  /* rgb_layer (3,128) */
} NeoMipMLPParams;
typedef struct {
    int n_prop, n_nerf;            /* MipNeRF360.num_prop_samples / num_nerf_samples (two proposal levels + one NeRF level) */
    float near_plane, far_plane;   /* near / far passed to MipNeRF360.forward (model.py:236) */
    float train_frac;              /* anneals the proposal logits (model.py:288-292) */
    const float* jitter[3];        /* randomized: one (n_rays) uniform per level (single_jitter, helper.py:357-363); NULL = deterministic */
    int precision;                 /* NeoPrecision: NEO_PREC_FP32 = fp32 CUDA-core SGEMM chain (tight parity), NEO_PREC_TC = fp16 activations / weights,
                                      every dense layer on tcgen05 (csrc/gemm_tc.cu) */
} NeoMipCfg;
typedef struct {                   /* per level: renderings[l]["rgb"], ray_history[l]{"density","rgb","sdist","weights"} (model.py:359-365) */
    float* rgb[3];       /* (n_rays,3) */
    float* density[3];   /* (n_rays,n_l) */
    float* rgb_s[3];     /* (n_rays,n_l,3) -- zeros for the proposal levels */
    float* sdist[3];     /* (n_rays,n_l+1) */
    float* weights[3];   /* (n_rays,n_l) */
} NeoMipOut;
size_t neo_mip_workspace_bytes(int n_rays, const NeoMipCfg* cfg, int nerf_width);
/* MipNeRF360.forward (models/mipnerf360/model.py:236-365); mlps[3] = {PropMLP, PropMLP, NeRFMLP}; radii (n_rays) */
int neo_mip_render_fwd(const NeoMipMLPParams mlps[3], const float* rays_o, const float* rays_d, const float* viewdirs,
                       const float* radii, int n_rays, const NeoMipCfg* cfg, NeoMipOut* out, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- tri-plane builder, dense part: models/neo360/encoder_tp_fusion_conv.py:472-597 between the ResNet feature
 * extractor and the floor-plan conv stacks (both stay in the host framework).  64^3 world grid x nv views: latent lookup, DepthPillarEncoder
 * 518->512->512->512, three pillar aggregators (513->512->1, softmax along one grid axis), softmax-weighted pillar sums.  Every dense layer
 * runs on tcgen05 (csrc/gemm_tc.cu, fp16 weights / activations, fp32 accumulation).  nn.Linear layout (out,in) fp32 device pointers. ---- */
typedef struct {
    const float* fc_w[3];   /* depth_fc.common_branch.0 (512,518), depth_fc.common_branch.2 (512,512), depth_fc.depth_encoder (512,512) */
    const float* fc_b[3];
    const float *agg_xz_w0, *agg_xz_b0, *agg_xz_w1, *agg_xz_b1;   /* pillar_aggregator_xz.{0,2}: (512,513), (512), (1,512), (1) */
    const float *agg_yz_w0, *agg_yz_b0, *agg_yz_w1, *agg_yz_b1;
    const float *agg_xy_w0, *agg_xy_b0, *agg_xy_w1, *agg_xy_b1;
} NeoGridEncoderParams;
size_t neo_grid_encoder_workspace_bytes(int nv, int lat_h, int lat_w);
/* latent (nv,512,lat_h,lat_w) NCHW = SpatialEncoder output; src_poses (nv,4,4) camera-to-world; focal / (cx,cy) = src_focal[0] / src_c[0].
 * Outputs: the three pillar-aggregated floor plans (nv,512,64,64) NCHW that feed floorplan_convnet_{xz,xy,yz}. */
int neo_grid_encoder_dense(const NeoGridEncoderParams* params, const float* latent, int nv, int lat_h, int lat_w, int img_w, int img_h,
                           const float* src_poses, float focal, float cx, float cy, float* floor_xz, float* floor_xy, float* floor_yz,
                           void* workspace, size_t workspace_bytes, void* stream);

/* bench support: CUDA events around every field-kernel launch on the launching stream + launch accounting.
 * neo_profile(1) resets and enables, neo_profile(0) resets and disables; neo_profile_read synchronises. */
int neo_profile(int enable);
int neo_profile_read(float* field_ms, int* n_field, unsigned long long* launches, double* points);

/* Self-test of the tcgen05 building blocks of the NEO_PREC_TC path (TMEM-resident A operand, 128B-swizzled K-major
 * operand tiles in K-major and MN-major form, TMEM loads): X (128,128), W (128,128), Wn (80,128) fp32 device ->
 * out1 = out3 (128,128) = W X^T (B operand K-major / MN-major), out2 = out4 (128,80) = X Wn^T (A operand K-major / MN-major),
 * computed with fp16 operands / fp32 accumulation. */
int neo_tc_selftest(const float* X, const float* W, const float* Wn, float* out1, float* out2, float* out3, float* out4,
                    void* stream);
/* Self-test of the no-swizzle K-major operand descriptor (identity A operand): outa / outb (128,128) = X^T under the two readings
 * of the descriptor's LBO/SBO fields (outa is the one the library uses). */
int neo_tc_selftest_transpose(const float* X, float* outa, float* outb, void* stream);
/* Self-test of the texel-window MMA of the NEO_PREC_TC field kernel (the bilinear lookups on the tensor pipe): `texels` (H*W,256)
 * fp32 texel-major stands for one projected map, (ox,oy) is the top-left texel of a 4x4 window (may lie partly or wholly outside
 * the map: zero fill), wt (64,16) the tap weights of 64 points over the window's 16 texels (y-major).  The window is staged by one
 * TMA box load (cp.async.bulk.tensor, 128B swizzle) and multiplied on tcgen05:
 * out0[c][p] = sum_k texel(oy + k/4, ox + k%4)[c] * wt[p][k],  out3 the same for channels 128..255; both (128,64) fp32. */
int neo_tc_selftest_window(const float* texels, int H, int W, int ox, int oy, const float* wt, float* out0, float* out3, void* stream);
/* Stage-level entry point of the tensor-core dense layer used by the wide MLPs (csrc/gemm_tc.cu: 2-D TMA tile loads, tcgen05 MMAs,
 * double-buffered TMEM accumulators): A (M,K), W (N,K) fp32 device, bias (N) or NULL -> out (M,N) fp32 = act(fp16(A) . fp16(W)^T + bias)
 * rounded to fp16.  K % 64 == 0, N % 64 == 0.  Synchronises the stream (allocates its fp16 staging buffers). */
int neo_tc_dense(const float* A, const float* W, const float* bias, long long M, int N, int K, int relu, float* out, void* stream);
/* Host-side view of the TC kernel's encoding-column layout (csrc/field_tc.cu enc_col<>): for in_ch = 3|4 and operand column
 * `col` in [0, KE = 64|96) returns the reference's positional-encoding index (helper.py:121-125 order) in [0, 21*in_ch),
 * -1 for the constant-one (bias) column, -2 for a zero padding column, -3 for invalid arguments.  Pure host code (no GPU). */
int neo_tc_enc_column(int in_ch, int col);

/* Diagnostics only, NOT for production callers: per-CTA cycle accounting of the NEO_PREC_TC field kernel into a caller-zeroed
 * device array of (#SMs x 64) int64 that must outlive every render issued while it is set; NULL disables (a separate
 * instantiation of the kernel carries the timers; production launches have none).  neo_scene_create resets it to NULL.
 * Slots: tools/tc_debug.py. */
int neo_tc_debug(long long* buf);

/* "" or a description of the mbarrier wait that timed out inside the NEO_PREC_TC field kernel (the kernel bounds every wait and
 * traps instead of hanging; the waiter's identity is recorded in host-mapped memory, which survives the failed context). */
const char* neo_tc_trap_info(void);
const char* neo_last_error(void);
/* "neo360_b200 <version> sm_100a" */
const char* neo_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NEO360_B200_H */
