// Shared device helpers for the NeO-360 hot path (sm_100a).  All math is fp32 and follows the order of
// operations of the reference's eager PyTorch ops (separate mul / add kernels => no FMA contraction) wherever
// a value feeds a discrete decision (sample positions, CDF brackets); see SURVEY.md Appendix A.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <vector>
#include "../../include/neo360_b200.h"

namespace neo {

constexpr int kMaxViews = 8;
constexpr int kWorldCh = 128;
constexpr int kLocalCh = 512;
constexpr int kHidden = 128;
constexpr int kDirEnc = 27;   // deg_view = 4  -> 3 + 3*2*4
constexpr int kPosDeg = 10;   // max_deg_point

// R^T and -(R^T t) of one source camera (models/neo360/util.py:52-70)
struct ViewXform {
    float rt[9];
    float tr[3];
    float pad[4];
};

struct SceneDev {
    int nv, plane_h, plane_w, lat_h, lat_w, img_w, img_h;
    float focal, cx, cy;          // src_focal[0], src_c[0]   (model.py:242-244)
    float lat_scale_x, lat_scale_y;  // latent_scaling / image_size (encoder_pn.py:119,204-206)
    const ViewXform* views;       // device, nv entries
    // channel-last fp32 copies (exact path): (nv, H, W, C)
    const float* planes_cl[3];    // xz, xy, yz
    const float* latent_cl;
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define NEO_CUDA(call)                                              \
    do {                                                            \
        cudaError_t _e = (call);                                    \
        if (_e != cudaSuccess) return neo::cuda_fail(_e, #call);    \
    } while (0)

#define NEO_LAUNCH_CHECK(name)                                      \
    do {                                                            \
        cudaError_t _e = cudaGetLastError();                        \
        if (_e != cudaSuccess) return neo::cuda_fail(_e, name);     \
    } while (0)

// ---- rounding-explicit arithmetic (mimics separate eager kernels) ----
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float dot3_(const float* a, const float* b) {
    return add_(add_(mul_(a[0], b[0]), mul_(a[1], b[1])), mul_(a[2], b[2]));
}

// torch.linspace(0, 1, steps)[i]  (ATen RangeFactories: symmetric evaluation around the midpoint)
__device__ __forceinline__ float linspace01(int i, int steps) {
    if (steps == 1) return 0.f;
    float step = __fdiv_rn(1.0f, (float)(steps - 1));
    int half = steps / 2;
    return (i < half) ? mul_(step, (float)i) : sub_(1.0f, mul_(step, (float)(steps - i - 1)));
}

// Per-ray constants of the NeRF++ parametrisation (helper.py:253-273 and 401-436).
struct RayGeom {
    float o[3], d[3];
    float far;        // intersect_sphere
    float rho;        // |p_mid|
    float phi;        // asin(rho)
    float psph[3];    // sphere hit point
    float axis[3];    // normalised rotation axis
    float check;      // 1 - |p_mid|^2  (must be >= 0)
};

__device__ __forceinline__ void ray_geom(const float* __restrict__ o, const float* __restrict__ d, RayGeom& g,
                                         bool need_bg) {
    g.o[0] = o[0]; g.o[1] = o[1]; g.o[2] = o[2];
    g.d[0] = d[0]; g.d[1] = d[1]; g.d[2] = d[2];
    float dd = dot3_(g.d, g.d);
    float d1 = __fdiv_rn(-dot3_(g.d, g.o), dd);
    float p[3] = {add_(g.o[0], mul_(d1, g.d[0])), add_(g.o[1], mul_(d1, g.d[1])), add_(g.o[2], mul_(d1, g.d[2]))};
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(dd));
    float p2 = dot3_(p, p);
    g.check = sub_(1.0f, p2);
    g.far = add_(d1, mul_(__fsqrt_rn(sub_(1.0f, p2)), inv));
    if (need_bg) {
        // depth2pts_outside uses norm(p_mid) and rho*rho instead of the squared sum (helper.py:422-428)
        g.rho = __fsqrt_rn(p2);
        float d2 = mul_(__fsqrt_rn(sub_(1.0f, mul_(g.rho, g.rho))), inv);
        float s = add_(d1, d2);
        for (int i = 0; i < 3; ++i) g.psph[i] = add_(g.o[i], mul_(s, g.d[i]));
        float ax[3] = {sub_(mul_(g.o[1], g.psph[2]), mul_(g.o[2], g.psph[1])),
                       sub_(mul_(g.o[2], g.psph[0]), mul_(g.o[0], g.psph[2])),
                       sub_(mul_(g.o[0], g.psph[1]), mul_(g.o[1], g.psph[0]))};
        float an = __fsqrt_rn(dot3_(ax, ax));
        for (int i = 0; i < 3; ++i) g.axis[i] = __fdiv_rn(ax[i], an);
        g.phi = asinf(g.rho);
    }
}

__device__ __forceinline__ void fg_point(const RayGeom& g, float t, float* x) {
    for (int i = 0; i < 3; ++i) x[i] = add_(g.o[i], mul_(t, g.d[i]));
}

// bg: s = inverse radius.  xhat = depth2pts_outside (unit vector), lin = o + (far(1-s) + far_unc*s) d (quirk Q2)
__device__ __forceinline__ void bg_point(const RayGeom& g, float s, float far_unc, float* xhat, float* lin) {
    float theta = asinf(mul_(g.rho, s));
    float ang = sub_(g.phi, theta);
    float ca = cosf(ang), sa = sinf(ang);
    const float* a = g.axis;
    const float* p = g.psph;
    float cr[3] = {sub_(mul_(a[1], p[2]), mul_(a[2], p[1])), sub_(mul_(a[2], p[0]), mul_(a[0], p[2])),
                   sub_(mul_(a[0], p[1]), mul_(a[1], p[0]))};
    float ap = dot3_(a, p);
    float omc = sub_(1.0f, ca);
    float q[3];
    for (int i = 0; i < 3; ++i) q[i] = add_(add_(mul_(p[i], ca), mul_(cr[i], sa)), mul_(mul_(a[i], ap), omc));
    float qn = add_(__fsqrt_rn(dot3_(q, q)), 1e-10f);
    for (int i = 0; i < 3; ++i) xhat[i] = __fdiv_rn(q[i], qn);
    if (lin) {
        float tl = add_(mul_(g.far, sub_(1.0f, s)), mul_(far_unc, s));
        for (int i = 0; i < 3; ++i) lin[i] = add_(g.o[i], mul_(tl, g.d[i]));
    }
}

__device__ __forceinline__ void to_camera(const ViewXform& v, const float* x, float* c) {
    c[0] = fmaf(v.rt[2], x[2], fmaf(v.rt[1], x[1], v.rt[0] * x[0])) + v.tr[0];
    c[1] = fmaf(v.rt[5], x[2], fmaf(v.rt[4], x[1], v.rt[3] * x[0])) + v.tr[1];
    c[2] = fmaf(v.rt[8], x[2], fmaf(v.rt[7], x[1], v.rt[6] * x[0])) + v.tr[2];
}
__device__ __forceinline__ void rotate_to_camera(const ViewXform& v, const float* x, float* c) {
    c[0] = fmaf(v.rt[2], x[2], fmaf(v.rt[1], x[1], v.rt[0] * x[0]));
    c[1] = fmaf(v.rt[5], x[2], fmaf(v.rt[4], x[1], v.rt[3] * x[0]));
    c[2] = fmaf(v.rt[8], x[2], fmaf(v.rt[7], x[1], v.rt[6] * x[0]));
}

// grid_sample(align_corners=True, zeros) tap set: indices (clamped) and weights (0 when out of range)
struct Taps {
    int idx[4];     // y*W + x of nw, ne, sw, se (valid even when weight is 0)
    float w[4];
};
__device__ __forceinline__ void bilinear_taps(float gx, float gy, int W, int H, Taps& t) {
    float ix = ((gx + 1.f) / 2.f) * (float)(W - 1);
    float iy = ((gy + 1.f) / 2.f) * (float)(H - 1);
    float x0f = floorf(ix), y0f = floorf(iy);
    float fx = ix - x0f, fy = iy - y0f;      // (ix - ix_nw)
    float gx1 = (x0f + 1.f) - ix, gy1 = (y0f + 1.f) - iy;  // (ix_se - ix)
    // NaN / huge coordinates: every tap is out of range -> 0
    bool finite = (ix == ix) && (iy == iy) && fabsf(ix) < 1e9f && fabsf(iy) < 1e9f;
    int x0 = finite ? (int)x0f : -2, y0 = finite ? (int)y0f : -2;
    int x1 = x0 + 1, y1 = y0 + 1;
    bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
    bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
    int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
    int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
    t.idx[0] = cy0 * W + cx0; t.w[0] = (vx0 & vy0) ? gx1 * gy1 : 0.f;
    t.idx[1] = cy0 * W + cx1; t.w[1] = (vx1 & vy0) ? fx * gy1 : 0.f;
    t.idx[2] = cy1 * W + cx0; t.w[2] = (vx0 & vy1) ? gx1 * fy : 0.f;
    t.idx[3] = cy1 * W + cx1; t.w[3] = (vx1 & vy1) ? fx * fy : 0.f;
}

// projection (util.py:92-111) + latent grid coords (encoder_pn.py:116-120)
__device__ __forceinline__ void local_grid_coords(const SceneDev& sc, const float* c, float& gx, float& gy) {
    float z = c[2] + 1e-9f;
    float u = (-c[0] / z) * sc.focal + sc.cx;
    float v = (-c[1] / z) * (-sc.focal) + sc.cy;
    gx = u * sc.lat_scale_x - 1.0f;
    gy = v * sc.lat_scale_y - 1.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- host-side launchers shared between translation units ----
struct MLPFp32 {   // weights transposed to (in, out) for coalesced reads by output-neuron threads
    int in_ch, enc_dim, in_dim;   // 3|4, 63|84, enc+512+128
    const float *w0t, *b0, *w1t, *b1, *w2t, *b2, *w3t, *b3, *wbt, *bb, *wsig, *bsig, *wv0t, *bv0, *wv1t, *bv1, *wrgb, *brgb;
};

}  // namespace neo

struct NeoScene {
    neo::SceneDev dev;
    NeoSceneDesc desc;
    int precision_mask;
    neo::MLPFp32 mlp32[4];
    void* tc_state;            // opaque state of the tensor-core path (field_tc.cu)
    int* err_flag;             // device int
    std::vector<std::pair<void*, size_t>> allocations;
    size_t bytes;
};

namespace neo {
// scene.cu
int scene_alloc_bytes(NeoScene* sc, void** p, size_t bytes);
// Device blocks of destroyed scenes (and the scene builder's temporaries) are kept, per device and up to a cap, for the next scene of
// the same shape: a scene change then costs its kernels, not cudaMalloc / cudaFree (which are slow, erratic and device-synchronising).
int pool_alloc(void** p, size_t bytes);
void pool_release(void* p, size_t bytes);
// sampling.cu
int launch_far(const float* o, const float* d, int n, float* far, int* err, cudaStream_t s);
int launch_sample_coarse(const float* o, const float* d, const float* far, int n, int num_samples, int in_sphere,
                         float far_unc, const float* u_rand, float* t, float* pts, float* pts_lin, cudaStream_t s);
int launch_resample(const float* o, const float* d, const float* far, const float* t_old, const float* w, int n,
                    int n_old, int m, int in_sphere, float far_unc, const float* u_rand, float* t, float* pts,
                    float* pts_lin, cudaStream_t s);
int launch_composite(const float* rgb, const float* sigma, const float* t, const float* d, const float* far, int n,
                     int N, int white, int in_sphere, float* comp, float* acc, float* w, float* lam, float* depth,
                     cudaStream_t s);
int launch_composite_bwd(const float* rgb, const float* sigma, const float* t, const float* d, const float* far, int n, int N, int white,
                         int in_sphere, const float* g_comp, const float* g_acc, const float* g_w, const float* g_lam, const float* g_depth,
                         float* d_rgb, float* d_sigma, cudaStream_t s);
int launch_combine(int n, int N, const float* fg_c, const float* bg_c, const float* lam, const float* fg_depth,
                   const float* bg_depth, const float* fg_t, const float* bg_s, float* comp, float* depth,
                   float* fg_sdist, float* bg_sdist, cudaStream_t s);
int launch_clipped_sq_err(const float* a, const float* b, long long n, double* out, cudaStream_t s);
int launch_sample_rays(int n, const long long* pix, int T, int H, int W, float focal, const float* c2w, const float* images,
                       float* o, float* vd, float* rd, float* radii, float* target, int* err, cudaStream_t s);
int launch_get_rays(int H, int W, float focal, const float* c2w, float* o, float* vd, float* rd, float* radii,
                    cudaStream_t s);
// field_fp32.cu
int launch_field_fp32(const NeoScene* sc, const NeoRays* rays, const float* far, const float* t, int N, int mlp_index,
                      float* rgb, float* sigma, cudaStream_t s);
int launch_index_grid(const NeoScene* sc, const float* pts, int M, float* out, cudaStream_t s);
int launch_index_local(const NeoScene* sc, const float* pts, int M, float* out, cudaStream_t s);
int launch_index_maps(const NeoScene* sc, const float* pts, int M, int C, const float* lat, const float* xz, const float* xy, const float* yz,
                      float* out_local, float* out_world, cudaStream_t s);
int launch_index_maps_bwd(const NeoScene* sc, const float* pts, int M, int C, const float* g_local, const float* g_world, float* g_lat, float* g_xz,
                          float* g_xy, float* g_yz, cudaStream_t s);
int launch_index_bwd(const NeoScene* sc, const float* pts, int M, int local, const float* g_out, float* g_lat, float* g_xz, float* g_xy,
                     float* g_yz, cudaStream_t s);
// field_tc.cu
int tc_scene_create(NeoScene* sc, const NeoMLPParams mlps[4], cudaStream_t s);
void tc_scene_free(NeoScene* sc);
int launch_field_tc(const NeoScene* sc, const NeoRays* rays, const float* far, const float* t, int N, int mlp_index,
                    float* rgb, float* sigma, cudaStream_t s);
}  // namespace neo
