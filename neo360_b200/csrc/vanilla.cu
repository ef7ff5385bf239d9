// Vanilla two-level NeRF renderer of the reference, fp32 CUDA cores, reference formulation.
// Reference: models/vanilla_nerf/model.py:44-216 (NeRFMLP 8x256 with skip after layer 4, NeRF.forward),
// models/vanilla_nerf/helper.py:415-442 (sampling), 445-449 (pos-enc), 521-559 (compositing), 567-616 (inverse CDF).
// Sampling marches along `viewdirs`, compositing scales by |rays_d| (quirk Q15); depth gets nan_to_num(inf) (quirk Q10).
#include "common.cuh"
#include <cuda_fp16.h>

struct NeoVanilla {
    struct Mlp {
        const float* wt[8];   // transposed (in,out)
        const float* b[8];
        const float *wbt, *bb, *wsig, *bsig, *wv0t, *bv0, *wrgb, *brgb;
        // tensor-core path (NEO_PREC_TC): nn.Linear layout (out, in padded to a multiple of 64) in fp16
        const void* w16[8];   // (256, 64) | (256,256) x4 | (256, 256+64) | (256,256) x2
        const void *wb16, *wv016;   // (256,256), (128, 256+64)
    } mlp[2];
    std::vector<void*> allocations;
    size_t bytes;
};

namespace neo {
namespace van {

constexpr int kW = 256, kEnc = 63, kP = 8, kThreads = 256, kCond = 128;

// helper.py:415-442, deterministic or jittered; points = o + t * viewdirs
__global__ void sample_kernel(const float* __restrict__ o, const float* __restrict__ vd, int n, int ns, float near, float far,
                              const float* __restrict__ u_rand, float* __restrict__ t_out) {
    int steps = ns + 1;
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)n * steps) return;
    int b = (int)(gid / steps), k = (int)(gid % steps);
    auto tv = [&](int kk) { float u = linspace01(kk, steps); return add_(mul_(near, sub_(1.0f, u)), mul_(far, u)); };
    float t = tv(k);
    if (u_rand) {
        float lower = (k > 0) ? mul_(0.5f, add_(t, tv(k - 1))) : t;
        float upper = (k < ns) ? mul_(0.5f, add_(tv(k + 1), t)) : t;
        t = add_(lower, mul_(sub_(upper, lower), u_rand[(long long)b * steps + k]));
    }
    t_out[gid] = t;
}

template <int ROWS>
__device__ __forceinline__ void dense(const float* __restrict__ Wt, int ldw, int K, const float* __restrict__ A, int lda, float* acc, int j) {
    int k = 0;
    for (; k + 4 <= K; k += 4) {
        float w0 = __ldg(Wt + (size_t)(k + 0) * ldw + j), w1 = __ldg(Wt + (size_t)(k + 1) * ldw + j);
        float w2 = __ldg(Wt + (size_t)(k + 2) * ldw + j), w3 = __ldg(Wt + (size_t)(k + 3) * ldw + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 a = *reinterpret_cast<const float4*>(A + r * lda + k);
            acc[r] = fmaf(a.w, w3, fmaf(a.z, w2, fmaf(a.y, w1, fmaf(a.x, w0, acc[r]))));
        }
    }
    for (; k < K; ++k) {
        float w0 = __ldg(Wt + (size_t)k * ldw + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(A[r * lda + k], w0, acc[r]);
    }
}

__global__ void __launch_bounds__(kThreads) field_kernel(NeoVanilla::Mlp m, const float* __restrict__ rays_o,
                                                         const float* __restrict__ viewdirs, const float* __restrict__ tvals,
                                                         int n_rays, int N, float* __restrict__ rgb_out, float* __restrict__ sigma_out) {
    __shared__ __align__(16) float X[kP][64];
    __shared__ __align__(16) float Ha[kP][kW + 4];
    __shared__ __align__(16) float Hb[kP][kW + 4];
    __shared__ __align__(16) float Dn[kP][28];
    __shared__ float xp[kP][3];
    __shared__ float red[8][kP];
    const int j = threadIdx.x;
    const long long total = (long long)n_rays * N, tile0 = (long long)blockIdx.x * kP;
    if (j < kP) {
        long long gp = min(tile0 + j, total - 1);
        int b = (int)(gp / N);
        float t = tvals[gp];
        for (int c = 0; c < 3; ++c) xp[j][c] = add_(rays_o[3 * b + c], mul_(t, viewdirs[3 * b + c]));   // cast_rays along viewdirs
    }
    __syncthreads();
    for (int e = j; e < kP * 64; e += kThreads) {
        int p = e / 64, c = e % 64;
        float val = 0.f;
        if (c < 3) val = xp[p][c];
        else if (c < kEnc) {
            int q = c - 3;
            bool shifted = q >= 30;
            if (shifted) q -= 30;
            float xb = mul_(xp[p][q % 3], (float)(1 << (q / 3)));
            val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
        }
        X[p][c] = val;
    }
    for (int e = j; e < kP * 28; e += kThreads) {
        int p = e / 28, c = e % 28;
        long long gp = min(tile0 + p, total - 1);
        const float* d = viewdirs + 3 * (gp / N);
        float val = 0.f;
        if (c < 3) val = d[c];
        else if (c < kDirEnc) {
            int q = c - 3;
            bool shifted = q >= 12;
            if (shifted) q -= 12;
            float xb = mul_(d[q % 3], (float)(1 << (q / 3)));
            val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
        }
        Dn[p][c] = val;
    }
    __syncthreads();
    float acc[kP];
    auto init = [&](const float* b) { float v = __ldg(b + j); for (int r = 0; r < kP; ++r) acc[r] = v; };
    auto relu_to = [&](float (*H)[kW + 4]) { for (int r = 0; r < kP; ++r) H[r][j] = fmaxf(acc[r], 0.f); };
    float (*src)[kW + 4] = Ha;
    float (*dst)[kW + 4] = Hb;
    // layer 0
    init(m.b[0]); dense<kP>(m.wt[0], kW, kEnc, &X[0][0], 64, acc, j); relu_to(Ha); __syncthreads();
    for (int l = 1; l < 8; ++l) {
        init(m.b[l]);
        dense<kP>(m.wt[l], kW, kW, &src[0][0], kW + 4, acc, j);
        if (l == 5) dense<kP>(m.wt[l] + (size_t)kW * kW, kW, kEnc, &X[0][0], 64, acc, j);      // cat([h, inputs]) after layer 4
        relu_to(dst);
        __syncthreads();
        float (*tmp)[kW + 4] = src; src = dst; dst = tmp;
    }
    // src = h7.  density head
    {
        float ws = __ldg(m.wsig + j);
        for (int p = 0; p < kP; ++p) {
            float part = warp_sum(src[p][j] * ws);
            if ((j & 31) == 0) red[j >> 5][p] = part;
        }
    }
    // bottleneck -> dst
    init(m.bb); dense<kP>(m.wbt, kW, kW, &src[0][0], kW + 4, acc, j);
    for (int r = 0; r < kP; ++r) dst[r][j] = acc[r];
    __syncthreads();
    if (j < kP) {
        long long gp = tile0 + j;
        if (gp < total) {
            float raw = __ldg(m.bsig);
            for (int w = 0; w < 8; ++w) raw += red[w][j];
            float xs = raw - 1.0f;
            sigma_out[gp] = xs > 20.f ? xs : log1pf(expf(xs));
        }
    }
    // view branch: [bottleneck | dir_enc] -> 128 relu  (result into src rows, first 128 columns)
    float a2[kP];
    if (j < kCond) {
        float v = __ldg(m.bv0 + j);
        for (int r = 0; r < kP; ++r) a2[r] = v;
        dense<kP>(m.wv0t, kCond, kW, &dst[0][0], kW + 4, a2, j);
        dense<kP>(m.wv0t + (size_t)kW * kCond, kCond, kDirEnc, &Dn[0][0], 28, a2, j);
    }
    __syncthreads();
    if (j < kCond) for (int r = 0; r < kP; ++r) src[r][j] = fmaxf(a2[r], 0.f);
    __syncthreads();
    if (j < kP * 3) {
        int p = j / 3, c = j % 3;
        long long gp = tile0 + p;
        if (gp < total) {
            float a = __ldg(m.brgb + c);
            for (int k = 0; k < kCond; ++k) a = fmaf(src[p][k], __ldg(m.wrgb + c * kCond + k), a);
            rgb_out[gp * 3 + c] = (1.f / (1.f + expf(-a))) * 1.002f - 0.001f;
        }
    }
}

// ---- tensor-core path: positional encodings as fp16 rows (63 -> 64, 27 -> 64 zero padded), activations of the heads ----
// One thread per sample row: the point and its 10 octaves (one sincosf per coordinate + exact angle doubling: error <= 2^9 ulp = 3e-5,
// far below the fp16 the row is stored in), then the direction encoding of its ray (4 octaves).  Rows are 128 bytes; a lane owns a row,
// so each row is assembled in a per-warp shared-memory tile ([32 rows][128 B], 16-byte pieces XOR-swizzled by row) and written out as
// whole lines, 4 rows per store instruction (the one-thread-per-element version spent 15 % of the frame here).
__device__ __forceinline__ void put_h(unsigned char* stage, int lane, int c, float v) {
    const int byte = c * 2;
    *reinterpret_cast<__half*>(stage + lane * 128 + ((((byte >> 4) ^ (lane & 7)) << 4) | (byte & 15))) = __float2half_rn(v);
}
__device__ __forceinline__ void flush_rows(const unsigned char* stage, int lane, __half* __restrict__ dst, long long ld, long long row0, long long M) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), p = lane & 7;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + rr * 128 + ((p ^ (rr & 7)) << 4));
        if (row0 + rr < M) *reinterpret_cast<uint4*>(dst + (row0 + rr) * ld + p * 8) = val;
    }
    __syncwarp();
}
__global__ void __launch_bounds__(256) enc16_kernel(const float* __restrict__ rays_o, const float* __restrict__ viewdirs, const float* __restrict__ tvals,
                                                    long long M, int N, __half* __restrict__ X16, long long ldx, __half* __restrict__ D16, long long ldd) {
    __shared__ __align__(16) unsigned char stage_all[8][32 * 128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* stage = stage_all[warp];
    const long long row0 = ((long long)blockIdx.x * 8 + warp) * 32, m = row0 + lane;
    const bool live = m < M;
    const int b = live ? (int)(m / N) : 0;
    const float t = live ? tvals[m] : 0.f;
    float d[3], x[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) { d[j] = viewdirs[3 * b + j]; x[j] = add_(rays_o[3 * b + j], mul_(t, d[j])); }
    // point encoding: [x (3) | sin(x 2^k) k-major (30) | cos (30) | 0]
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        put_h(stage, lane, j, x[j]);
        float sn, cs;
        sincosf(x[j], &sn, &cs);
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            put_h(stage, lane, 3 + 3 * k + j, sn);
            put_h(stage, lane, 33 + 3 * k + j, cs);
            const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
            sn = s2; cs = c2;
        }
    }
    put_h(stage, lane, 63, 0.f);
    flush_rows(stage, lane, X16, ldx, row0, M);
    // direction encoding: [d (3) | sin (12) | cos (12) | 0 ... 0]
#pragma unroll
    for (int p = 3; p < 8; ++p) *reinterpret_cast<uint4*>(stage + lane * 128 + ((p ^ (lane & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    for (int c = 27; c < 32; ++c) put_h(stage, lane, c, 0.f);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        put_h(stage, lane, j, d[j]);
        float sn, cs;
        sincosf(d[j], &sn, &cs);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            put_h(stage, lane, 3 + 3 * k + j, sn);
            put_h(stage, lane, 15 + 3 * k + j, cs);
            const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
            sn = s2; cs = c2;
        }
    }
    flush_rows(stage, lane, D16, ldd, row0, M);
}
__global__ void head_act_kernel(const float* __restrict__ raw_sigma, const float* __restrict__ raw_rgb, long long M, float* __restrict__ sigma,
                                float* __restrict__ rgb) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * 4) return;
    const long long m = i >> 2;
    const int c = (int)(i & 3);
    if (c == 3) { const float xs = raw_sigma[m] - 1.0f; sigma[m] = xs > 20.f ? xs : log1pf(expf(xs)); }
    else rgb[m * 3 + c] = (1.f / (1.f + expf(-raw_rgb[m * 3 + c]))) * 1.002f - 0.001f;
}

}  // namespace van
// csrc/gemm_tc.cu, csrc/mip.cu
int gemm_f16(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* C, long long ldc, long long M, int N, int K,
             int relu, cudaStream_t s);
int f32_to_f16_pad(const float* in, long long rows, int cols_in, long long ld_in, void* out, int cols_out, long long ld_out, cudaStream_t s);
int launch_rowdot_f16(const void* H, long long ld, int K, const float* W, const float* b, int N, long long M, float* out, cudaStream_t s);
}  // namespace neo

using namespace neo;

namespace {
__global__ void transpose2(const float* __restrict__ w, float* __restrict__ wt, int out_f, int in_f) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_f * in_f) return;
    wt[(size_t)(idx % in_f) * out_f + idx / in_f] = w[idx];
}
struct Carver2 { float* base; size_t used; float* take(size_t n) { n = (n + 63) & ~size_t(63); float* p = base ? base + used : nullptr; used += n; return p; } };
struct WSV { float *t0, *w0, *t1, *w1, *sig, *rgb; void *A16[2], *B16, *V16; float *rawd, *rawc; };
constexpr int kLdA = 256 + 64, kLdB = 256 + 64;      // activation rows: [h (256) | padded encoding (64)], [bottleneck (256) | padded direction encoding (64)]
size_t carve2(Carver2& c, int n, int N0, int N1, WSV& w, int precision) {
    w.t0 = c.take((size_t)n * N0); w.w0 = c.take((size_t)n * N0); w.t1 = c.take((size_t)n * N1); w.w1 = c.take((size_t)n * N1);
    w.sig = c.take((size_t)n * N1); w.rgb = c.take((size_t)n * N1 * 3);
    if (precision == NEO_PREC_TC) {
        const size_t M = (size_t)n * N1;
        for (int i = 0; i < 2; ++i) w.A16[i] = c.take((M * kLdA + 1) / 2);
        w.B16 = c.take((M * kLdB + 1) / 2);
        w.V16 = c.take((M * 128 + 1) / 2);
        w.rawd = c.take(M); w.rawc = c.take(M * 3);
    }
    return c.used * sizeof(float);
}
int check(const NeoVanillaCfg* c) {
    if (!c || c->n_coarse < 3 || c->n_fine < 1 || c->n_coarse > 4096 || c->n_fine > 4096) { set_error("vanilla: bad sample counts"); return NEO_ERR_INVALID; }
    if (c->precision != NEO_PREC_FP32 && c->precision != NEO_PREC_TC) { set_error("vanilla: bad precision %d", c->precision); return NEO_ERR_INVALID; }
    return NEO_OK;
}
}  // namespace

extern "C" int neo_vanilla_create(const NeoVanillaMLPParams mlps[2], NeoVanilla** out, void* stream) {
    if (!mlps || !out) { set_error("neo_vanilla_create: null argument"); return NEO_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    NeoVanilla* v = new NeoVanilla();
    v->bytes = 0;
    auto fail = [&](int rc) { neo_vanilla_free(v); return rc; };
    auto dup = [&](const float* src, size_t n, const float** dst, int out_f, int in_f) -> int {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, n * sizeof(float));
        if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(vanilla)");
        v->allocations.push_back(q);
        v->bytes += n * sizeof(float);
        if (out_f > 0) transpose2<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(src, (float*)q, out_f, in_f);
        else { e = cudaMemcpyAsync(q, src, n * sizeof(float), cudaMemcpyDeviceToDevice, s); if (e != cudaSuccess) return cuda_fail(e, "copy"); }
        *dst = (const float*)q;
        return NEO_OK;
    };
    for (int i = 0; i < 2; ++i) {
        const NeoVanillaMLPParams& p = mlps[i];
        NeoVanilla::Mlp& m = v->mlp[i];
        int rc;
        for (int l = 0; l < 8; ++l) {
            int in_f = l == 0 ? 63 : (l == 5 ? 319 : 256);
            if ((rc = dup(p.w[l], (size_t)256 * in_f, &m.wt[l], 256, in_f))) return fail(rc);
            if ((rc = dup(p.b[l], 256, &m.b[l], 0, 0))) return fail(rc);
        }
        if ((rc = dup(p.wb, 256 * 256, &m.wbt, 256, 256))) return fail(rc);
        if ((rc = dup(p.bb, 256, &m.bb, 0, 0))) return fail(rc);
        if ((rc = dup(p.wsig, 256, &m.wsig, 0, 0))) return fail(rc);
        if ((rc = dup(p.bsig, 1, &m.bsig, 0, 0))) return fail(rc);
        if ((rc = dup(p.wv0, 128 * 283, &m.wv0t, 128, 283))) return fail(rc);
        if ((rc = dup(p.bv0, 128, &m.bv0, 0, 0))) return fail(rc);
        if ((rc = dup(p.wrgb, 3 * 128, &m.wrgb, 0, 0))) return fail(rc);
        if ((rc = dup(p.brgb, 3, &m.brgb, 0, 0))) return fail(rc);
        // fp16 images for the tensor-core path: K padded to a multiple of 64; the skip layer's [h | inputs] columns land at [0,256) | [256,319)
        auto pack16 = [&](const float* src, int rows, int c0, int ncols, int ld_in, void* dst, int off, int pad_cols, int ld_out) -> int {
            return f32_to_f16_pad(src + c0, rows, ncols, ld_in, (__half*)dst + off, pad_cols, ld_out, s);
        };
        auto alloc16 = [&](size_t halves, const void** dst) -> int {
            void* q = nullptr;
            cudaError_t e = cudaMalloc(&q, halves * 2);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(vanilla fp16)");
            v->allocations.push_back(q);
            v->bytes += halves * 2;
            *dst = q;
            return NEO_OK;
        };
        for (int l = 0; l < 8; ++l) {
            const int in_f = l == 0 ? 63 : (l == 5 ? 319 : 256), kp = l == 0 ? 64 : (l == 5 ? 320 : 256);
            if ((rc = alloc16((size_t)256 * kp, &m.w16[l]))) return fail(rc);
            if (l == 5) {
                if ((rc = pack16(p.w[l], 256, 0, 256, in_f, (void*)m.w16[l], 0, 256, kp))) return fail(rc);
                if ((rc = pack16(p.w[l], 256, 256, 63, in_f, (void*)m.w16[l], 256, 64, kp))) return fail(rc);
            } else if ((rc = pack16(p.w[l], 256, 0, in_f, in_f, (void*)m.w16[l], 0, kp, kp))) return fail(rc);
        }
        if ((rc = alloc16((size_t)256 * 256, &m.wb16))) return fail(rc);
        if ((rc = pack16(p.wb, 256, 0, 256, 256, (void*)m.wb16, 0, 256, 256))) return fail(rc);
        if ((rc = alloc16((size_t)128 * 320, &m.wv016))) return fail(rc);
        if ((rc = pack16(p.wv0, 128, 0, 256, 283, (void*)m.wv016, 0, 256, 320))) return fail(rc);
        if ((rc = pack16(p.wv0, 128, 256, 27, 283, (void*)m.wv016, 256, 64, 320))) return fail(rc);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(cuda_fail(e, "neo_vanilla_create sync"));
    *out = v;
    return NEO_OK;
}

extern "C" void neo_vanilla_free(NeoVanilla* v) {
    if (!v) return;
    for (void* p : v->allocations) cudaFree(p);
    delete v;
}

extern "C" size_t neo_vanilla_workspace_bytes(int n_rays, const NeoVanillaCfg* cfg) {
    if (n_rays <= 0 || check(cfg)) return 0;
    Carver2 c{nullptr, 0};
    WSV w;
    return carve2(c, n_rays, cfg->n_coarse + 1, cfg->n_coarse + 1 + cfg->n_fine, w, cfg->precision);
}

extern "C" int neo_vanilla_render_fwd(const NeoVanilla* v, const NeoRays* rays, const NeoVanillaCfg* cfg, NeoVanillaOut* out,
                                      void* workspace, size_t workspace_bytes, void* stream) {
    if (!v || !rays || !out) { set_error("neo_vanilla_render_fwd: null argument"); return NEO_ERR_INVALID; }
    int rc = check(cfg);
    if (rc) return rc;
    if (rays->n_rays <= 0 || !rays->rays_o || !rays->rays_d || !rays->viewdirs) { set_error("neo_vanilla_render_fwd: empty rays"); return NEO_ERR_INVALID; }
    cudaStream_t s = (cudaStream_t)stream;
    const int n = rays->n_rays, N0 = cfg->n_coarse + 1, N1 = N0 + cfg->n_fine;
    Carver2 c{reinterpret_cast<float*>(workspace), 0};
    WSV w;
    size_t need = carve2(c, n, N0, N1, w, cfg->precision);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes); return NEO_ERR_WORKSPACE; }
    for (int lvl = 0; lvl < 2; ++lvl) {
        const int N = lvl ? N1 : N0;
        float* t = lvl ? w.t1 : w.t0;
        float* wt = lvl ? w.w1 : w.w0;
        if (lvl == 0) {
            long long total = (long long)n * N0;
            van::sample_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(rays->rays_o, rays->viewdirs, n, cfg->n_coarse, cfg->near_plane,
                                                                               cfg->far_plane, cfg->u0, t);
            NEO_LAUNCH_CHECK("vanilla sample_kernel");
        } else {
            // sample_pdf: bins = mids(t), weights[1:-1]; same ascending-bin inverse CDF + merge as the NeO-360 foreground branch
            if ((rc = launch_resample(rays->rays_o, rays->viewdirs, nullptr, w.t0, w.w0, n, N0, cfg->n_fine, 1, 0.f, cfg->u1, t, nullptr, nullptr, s))) return rc;
        }
        long long total = (long long)n * N;
        if (cfg->precision == NEO_PREC_TC) {
            // NeRFMLP (models/vanilla_nerf/model.py:44-125) layer by layer on tcgen05 (csrc/gemm_tc.cu), fp16 activations; the skip concatenation
            // by keeping h4 and the encoding in ONE buffer (layer 5 is a single K = 320 GEMM)
            const NeoVanilla::Mlp& m = v->mlp[lvl];
            __half* buf[2] = {(__half*)w.A16[0], (__half*)w.A16[1]};
            __half* B = (__half*)w.B16;
            van::enc16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(rays->rays_o, rays->viewdirs, t, total, N, buf[0] + 256, kLdA, B + 256, kLdB);
            NEO_LAUNCH_CHECK("vanilla enc16_kernel");
            if ((rc = gemm_f16(buf[0] + 256, kLdA, m.w16[0], 64, m.b[0], buf[0], kLdA, total, 256, 64, 1, s))) return rc;
            for (int l = 1; l < 8; ++l) {
                const int K = l == 5 ? 320 : 256;
                if ((rc = gemm_f16(buf[(l - 1) & 1], kLdA, m.w16[l], K, m.b[l], buf[l & 1], kLdA, total, 256, K, 1, s))) return rc;
            }
            const __half* h = buf[1];
            if ((rc = launch_rowdot_f16(h, kLdA, 256, m.wsig, m.bsig, 1, total, w.rawd, s))) return rc;
            if ((rc = gemm_f16(h, kLdA, m.wb16, 256, m.bb, B, kLdB, total, 256, 256, 0, s))) return rc;
            if ((rc = gemm_f16(B, kLdB, m.wv016, 320, m.bv0, w.V16, 128, total, 128, 320, 1, s))) return rc;
            if ((rc = launch_rowdot_f16(w.V16, 128, 128, m.wrgb, m.brgb, 3, total, w.rawc, s))) return rc;
            van::head_act_kernel<<<(unsigned)((total * 4 + 255) / 256), 256, 0, s>>>(w.rawd, w.rawc, total, w.sig, w.rgb);
            NEO_LAUNCH_CHECK("vanilla head_act_kernel");
        } else {
            van::field_kernel<<<(unsigned)((total + van::kP - 1) / van::kP), van::kThreads, 0, s>>>(v->mlp[lvl], rays->rays_o, rays->viewdirs, t, n, N,
                                                                                                      w.rgb, w.sig);
            NEO_LAUNCH_CHECK("vanilla field_kernel");
        }
        // mode 2: ascending t, last interval 1e10, scaled by |rays_d|, depth nan_to_num(inf)
        if ((rc = launch_composite(w.rgb, w.sig, t, rays->rays_d, nullptr, n, N, cfg->white_bkgd, 2, out->comp_rgb[lvl], out->acc[lvl], wt, nullptr,
                                   out->depth[lvl], s))) return rc;
        auto cp = [&](float* dst, const float* src, size_t cnt) -> int {
            if (!dst) return NEO_OK;
            NEO_CUDA(cudaMemcpyAsync(dst, src, cnt * sizeof(float), cudaMemcpyDeviceToDevice, s));
            return NEO_OK;
        };
        if ((rc = cp(out->t[lvl], t, (size_t)n * N))) return rc;
        if ((rc = cp(out->sigma[lvl], w.sig, (size_t)n * N))) return rc;
        if ((rc = cp(out->rgb_s[lvl], w.rgb, (size_t)n * N * 3))) return rc;
        if ((rc = cp(out->weights[lvl], wt, (size_t)n * N))) return rc;
    }
    return NEO_OK;
}
