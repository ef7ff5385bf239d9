// Exact-arithmetic (CUDA-core fp32) evaluation of the NeO-360 radiance field in the REFERENCE formulation:
// world->camera, tri-plane + pixel-aligned bilinear lookups, positional encoding, NeRFPPMLP with the
// cross-view means, activations.  This is the tight-parity path (NEO_PREC_FP32) and the on-GPU check for the
// tensor-core path in field_tc.cu.  Reference: models/neo360/model.py:110-158, 239-264, 339-472;
// encoder_tp_fusion_conv.py:122-209; encoder_pn.py:101-152; util.py:45-111; helper.py:121-125.
#include "common.cuh"

namespace neo {

constexpr int kP = 8;          // points per CTA
constexpr int kThreads = 128;  // one thread per hidden unit

struct RowGeo {
    float enc_in[4];    // camera-frame position fed to pos_enc (+ 1/r for bg)
    float dir[3];       // camera-frame view direction of the (quirk-Q1) conditioning ray
    Taps local;
    Taps plane[3];      // xz, xy, yz
    int valid;
};

template <int ROWS>
__device__ __forceinline__ void dense_acc(const float* __restrict__ Wt, int K, const float* __restrict__ A, int lda,
                                          float* acc, int j) {
    int k = 0;
    for (; k + 4 <= K; k += 4) {
        float w0 = __ldg(Wt + (size_t)(k + 0) * kHidden + j), w1 = __ldg(Wt + (size_t)(k + 1) * kHidden + j);
        float w2 = __ldg(Wt + (size_t)(k + 2) * kHidden + j), w3 = __ldg(Wt + (size_t)(k + 3) * kHidden + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            float4 a = *reinterpret_cast<const float4*>(A + r * lda + k);
            acc[r] = fmaf(a.w, w3, fmaf(a.z, w2, fmaf(a.y, w1, fmaf(a.x, w0, acc[r]))));
        }
    }
    for (; k < K; ++k) {
        float w0 = __ldg(Wt + (size_t)k * kHidden + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(A[r * lda + k], w0, acc[r]);
    }
}

__device__ __forceinline__ float softplus_(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_(float x) { return 1.f / (1.f + expf(-x)); }

template <int NV>
__global__ void __launch_bounds__(kThreads)
field_fp32_kernel(SceneDev sc, MLPFp32 mlp, const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                  const float* __restrict__ viewdirs, const float* __restrict__ far, const float* __restrict__ tvals,
                  int n_rays, int N, int chunk, int is_bg, float far_unc, float* __restrict__ rgb_out,
                  float* __restrict__ sigma_out) {
    constexpr int ROWS = NV * kP;
    extern __shared__ __align__(16) float smem[];
    const int in_dim = mlp.in_dim;                 // enc + 512 + 128
    const int ldx = (in_dim + 3) / 4 * 4 + 4;
    float* X = smem;                               // [ROWS][ldx]   inputs  [enc | local | world]
    float* Ha = X + ROWS * ldx;                    // [ROWS][128+4]
    float* Hb = Ha + ROWS * (kHidden + 4);         // [ROWS][128+4]
    float* Dn = Hb + ROWS * (kHidden + 4);         // [ROWS][28]  direction encodings
    float* Q = Dn + ROWS * 28;                     // [kP][64+4] x2
    RowGeo* geo = reinterpret_cast<RowGeo*>(Q + 2 * kP * 68);
    const int ldh = kHidden + 4;
    const int j = threadIdx.x;
    const long long total = (long long)n_rays * N;
    const long long tile0 = (long long)blockIdx.x * kP;

    // ---- phase A: geometry per (view, point) row ----
    if (j < ROWS) {
        int v = j / kP, p = j % kP;
        long long gp = tile0 + p;
        RowGeo& rg = geo[j];
        rg.valid = gp < total;
        if (rg.valid) {
            int b = (int)(gp / N), s = (int)(gp % N);
            RayGeom g;
            ray_geom(rays_o + 3 * b, rays_d + 3 * b, g, is_bg);
            g.far = far[b];
            float tv = tvals[gp];
            float xe[3], xl[3];
            if (is_bg) bg_point(g, tv, far_unc, xe, xl);
            else { fg_point(g, tv, xe); xl[0] = xe[0]; xl[1] = xe[1]; xl[2] = xe[2]; }
            const ViewXform& vx = sc.views[v];
            float ce[3], cl[3];
            to_camera(vx, xe, ce);
            to_camera(vx, xl, cl);
            rg.enc_in[0] = ce[0]; rg.enc_in[1] = ce[1]; rg.enc_in[2] = ce[2]; rg.enc_in[3] = tv;
            // quirk Q1: row (b_local*N + s) of the chunk is conditioned on ray ((b_local*N + s) mod B_chunk)
            int ch = chunk > 0 ? chunk : n_rays;
            int c0 = (b / ch) * ch;
            int Bc = min(ch, n_rays - c0);
            long long jl = (long long)(b - c0) * N + s;
            int src = c0 + (int)(jl % Bc);
            rotate_to_camera(vx, viewdirs + 3 * src, rg.dir);
            float gx, gy;
            local_grid_coords(sc, cl, gx, gy);
            bilinear_taps(gx, gy, sc.lat_w, sc.lat_h, rg.local);
            bilinear_taps(cl[0], cl[2], sc.plane_w, sc.plane_h, rg.plane[0]);   // xz
            bilinear_taps(cl[0], cl[1], sc.plane_w, sc.plane_h, rg.plane[1]);   // xy
            bilinear_taps(cl[1], cl[2], sc.plane_w, sc.plane_h, rg.plane[2]);   // yz
        }
    }
    __syncthreads();

    // ---- phase B: inputs  X = [pos_enc | local latent | world latent],  Dn = dir pos_enc ----
    const int ich = mlp.in_ch, enc = mlp.enc_dim;
    for (int e = j; e < ROWS * enc; e += kThreads) {
        int r = e / enc, c = e % enc;
        const RowGeo& rg = geo[r];
        float val = 0.f;
        if (rg.valid) {
            if (c < ich) val = rg.enc_in[c];
            else {
                int q = c - ich;
                int half = ich * kPosDeg;
                bool shifted = q >= half;
                if (shifted) q -= half;
                int k = q / ich, cc = q % ich;
                float xb = mul_(rg.enc_in[cc], (float)(1 << k));
                val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
            }
        }
        X[r * ldx + c] = val;
    }
    for (int e = j; e < ROWS * 28; e += kThreads) {
        int r = e / 28, c = e % 28;
        const RowGeo& rg = geo[r];
        float val = 0.f;
        if (rg.valid && c < kDirEnc) {
            if (c < 3) val = rg.dir[c];
            else {
                int q = c - 3;
                bool shifted = q >= 12;
                if (shifted) q -= 12;
                int k = q / 3, cc = q % 3;
                float xb = mul_(rg.dir[cc], (float)(1 << k));
                val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
            }
        }
        Dn[e] = val;
    }
    for (int r = 0; r < ROWS; ++r) {
        const RowGeo& rg = geo[r];
        int v = r / kP;
        float4 accl = make_float4(0.f, 0.f, 0.f, 0.f);
        float accw = 0.f;
        if (rg.valid) {
            const float* lat = sc.latent_cl + (size_t)v * sc.lat_h * sc.lat_w * kLocalCh;
#pragma unroll
            for (int tp = 0; tp < 4; ++tp) {
                float w = rg.local.w[tp];
                float4 f = __ldg(reinterpret_cast<const float4*>(lat + (size_t)rg.local.idx[tp] * kLocalCh) + j);
                accl.x += f.x * w; accl.y += f.y * w; accl.z += f.z * w; accl.w += f.w * w;
            }
            float pl[3];
#pragma unroll
            for (int pi = 0; pi < 3; ++pi) {
                const float* pp = sc.planes_cl[pi] + (size_t)v * sc.plane_h * sc.plane_w * kWorldCh;
                float a = 0.f;
#pragma unroll
                for (int tp = 0; tp < 4; ++tp)
                    a += __ldg(pp + (size_t)rg.plane[pi].idx[tp] * kWorldCh + j) * rg.plane[pi].w[tp];
                pl[pi] = a;
            }
            accw = (pl[0] + pl[1]) + pl[2];        // torch.sum(stack([xz, xy, yz]), 0)
        }
        float* xr = X + r * ldx + enc;
        xr[4 * j + 0] = accl.x; xr[4 * j + 1] = accl.y; xr[4 * j + 2] = accl.z; xr[4 * j + 3] = accl.w;
        xr[kLocalCh + j] = accw;
    }
    __syncthreads();

    // ---- phase C: MLP ----
    float acc[ROWS];
    auto bias_init = [&](const float* bptr) {
        float bv = __ldg(bptr + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = bv;
    };
    auto store_relu = [&](float* H) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) H[r * ldh + j] = fmaxf(acc[r], 0.f);
    };
    bias_init(mlp.b0); dense_acc<ROWS>(mlp.w0t, in_dim, X, ldx, acc, j); store_relu(Ha); __syncthreads();
    bias_init(mlp.b1); dense_acc<ROWS>(mlp.w1t, kHidden, Ha, ldh, acc, j); store_relu(Hb); __syncthreads();
    bias_init(mlp.b2); dense_acc<ROWS>(mlp.w2t, kHidden, Hb, ldh, acc, j); store_relu(Ha); __syncthreads();
    bias_init(mlp.b3);
    dense_acc<ROWS>(mlp.w3t, kHidden, Ha, ldh, acc, j);                              // [h2 | inputs]
    dense_acc<ROWS>(mlp.w3t + (size_t)kHidden * kHidden, in_dim, X, ldx, acc, j);
    store_relu(Hb); __syncthreads();                                                // Hb = h3 (per view)
    // bottleneck (per view) -> Ha ; hbar = mean_v h3 -> density
    bias_init(mlp.bb); dense_acc<ROWS>(mlp.wbt, kHidden, Hb, ldh, acc, j);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) Ha[r * ldh + j] = acc[r];
    {
        float ws = __ldg(mlp.wsig + j);
        float* red = Q;                     // [4 warps][kP]
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            float m = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) m += Hb[(v * kP + p) * ldh + j];
            m = m / (float)NV;
            float part = warp_sum(m * ws);
            if ((j & 31) == 0) red[(j >> 5) * kP + p] = part;
        }
    }
    __syncthreads();
    if (j < kP) {
        long long gp = tile0 + j;
        if (gp < total) {
            float raw = ((Q[0 * kP + j] + Q[1 * kP + j]) + (Q[2 * kP + j] + Q[3 * kP + j])) + __ldg(mlp.bsig);
            sigma_out[gp] = softplus_(raw - 1.0f);                                  // model.py:392-393
        }
    }
    __syncthreads();
    // view branch: [bottleneck | dir_enc] -> 64, mean over views, relu, 64->64 relu, 64->3
    float* q0 = Q;
    float* q1 = Q + kP * 68;
    if (j < 64) {
        float bv = __ldg(mlp.bv0 + j);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) acc[r] = bv;
        for (int k = 0; k < kHidden; ++k) {
            float w = __ldg(mlp.wv0t + (size_t)k * 64 + j);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(Ha[r * ldh + k], w, acc[r]);
        }
        for (int k = 0; k < kDirEnc; ++k) {
            float w = __ldg(mlp.wv0t + (size_t)(kHidden + k) * 64 + j);
#pragma unroll
            for (int r = 0; r < ROWS; ++r) acc[r] = fmaf(Dn[r * 28 + k], w, acc[r]);
        }
#pragma unroll
        for (int p = 0; p < kP; ++p) {
            float m = 0.f;
#pragma unroll
            for (int v = 0; v < NV; ++v) m += acc[v * kP + p];
            q0[p * 68 + j] = fmaxf(m / (float)NV, 0.f);
        }
    }
    __syncthreads();
    if (j < 64) {
        float a2[kP];
        float bv = __ldg(mlp.bv1 + j);
#pragma unroll
        for (int p = 0; p < kP; ++p) a2[p] = bv;
        for (int k = 0; k < 64; ++k) {
            float w = __ldg(mlp.wv1t + (size_t)k * 64 + j);
#pragma unroll
            for (int p = 0; p < kP; ++p) a2[p] = fmaf(q0[p * 68 + k], w, a2[p]);
        }
#pragma unroll
        for (int p = 0; p < kP; ++p) q1[p * 68 + j] = fmaxf(a2[p], 0.f);
    }
    __syncthreads();
    if (j < kP * 3) {
        int p = j / 3, c = j % 3;
        long long gp = tile0 + p;
        if (gp < total) {
            float a = __ldg(mlp.brgb + c);
            for (int k = 0; k < 64; ++k) a = fmaf(q1[p * 68 + k], __ldg(mlp.wrgb + c * 64 + k), a);
            rgb_out[gp * 3 + c] = sigmoid_(a) * 1.002f - 0.001f;                     // model.py:395-397
        }
    }
}

static size_t field_fp32_smem(int nv, int in_dim) {
    int rows = nv * kP;
    int ldx = (in_dim + 3) / 4 * 4 + 4;
    size_t fl = (size_t)rows * ldx + 2 * (size_t)rows * (kHidden + 4) + (size_t)rows * 28 + 2 * kP * 68;
    return fl * sizeof(float) + (size_t)rows * sizeof(RowGeo);
}

template <int NV>
static int launch_nv(const NeoScene* sc, const NeoRays* rays, const float* far, const float* t, int N, int mi,
                     float* rgb, float* sigma, cudaStream_t s) {
    const MLPFp32& m = sc->mlp32[mi];
    size_t smem = field_fp32_smem(NV, m.in_dim);
    NEO_CUDA(cudaFuncSetAttribute(field_fp32_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    long long total = (long long)rays->n_rays * N;
    unsigned grid = (unsigned)((total + kP - 1) / kP);
    field_fp32_kernel<NV><<<grid, kThreads, smem, s>>>(sc->dev, m, rays->rays_o, rays->rays_d, rays->viewdirs, far, t,
                                                       rays->n_rays, N, rays->chunk, mi & 1, 3.0f, rgb, sigma);
    NEO_LAUNCH_CHECK("field_fp32_kernel");
    return NEO_OK;
}

int launch_field_fp32(const NeoScene* sc, const NeoRays* rays, const float* far, const float* t, int N, int mlp_index,
                      float* rgb, float* sigma, cudaStream_t s) {
    if (!(sc->precision_mask & (1 << NEO_PREC_FP32))) { set_error("scene was not prepared for NEO_PREC_FP32"); return NEO_ERR_INVALID; }
    switch (sc->dev.nv) {
        case 1: return launch_nv<1>(sc, rays, far, t, N, mlp_index, rgb, sigma, s);
        case 2: return launch_nv<2>(sc, rays, far, t, N, mlp_index, rgb, sigma, s);
        case 3: return launch_nv<3>(sc, rays, far, t, N, mlp_index, rgb, sigma, s);
        case 4: return launch_nv<4>(sc, rays, far, t, N, mlp_index, rgb, sigma, s);
    }
    set_error("NEO_PREC_FP32 supports 1..4 source views (got %d)", sc->dev.nv);
    return NEO_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// stage-level lookups (index_grid / get_local_feats): rows ordered (view, point), channels contiguous
// ------------------------------------------------------------------------------------------------
// Maps are passed explicitly (channel-last (nv, H, W, C), spatial sizes = the scene's): the scene's own raw maps for index_grid /
// get_local_feats, or caller-owned maps of any channel count C % 4 == 0 for the training path's projected maps (neo_index_maps*).
struct MapSet { const float* lat; const float* pl[3]; };
struct MapSetW { float* lat; float* pl[3]; };

__global__ void index_kernel(SceneDev sc, const float* __restrict__ pts, int M, int local, int C, MapSet maps, float* __restrict__ out) {
    long long row = blockIdx.x;           // v*M + m
    int v = (int)(row / M), m = (int)(row % M);
    float c[3];
    to_camera(sc.views[v], pts + 3 * (size_t)m, c);
    if (local) {
        float gx, gy;
        Taps t;
        local_grid_coords(sc, c, gx, gy);
        bilinear_taps(gx, gy, sc.lat_w, sc.lat_h, t);
        const float* lat = maps.lat + (size_t)v * sc.lat_h * sc.lat_w * C;
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            float val = 0.f;
            for (int tp = 0; tp < 4; ++tp) val += __ldg(lat + (size_t)t.idx[tp] * C + ch) * t.w[tp];
            out[row * C + ch] = val;
        }
    } else {
        const float ga[3] = {c[0], c[0], c[1]}, gb[3] = {c[2], c[1], c[2]};
        Taps t[3];
        for (int pi = 0; pi < 3; ++pi) bilinear_taps(ga[pi], gb[pi], sc.plane_w, sc.plane_h, t[pi]);
        for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
            float pl[3];
            for (int pi = 0; pi < 3; ++pi) {
                const float* pp = maps.pl[pi] + (size_t)v * sc.plane_h * sc.plane_w * C;
                float a = 0.f;
                for (int tp = 0; tp < 4; ++tp) a += __ldg(pp + (size_t)t[pi].idx[tp] * C + ch) * t[pi].w[tp];
                pl[pi] = a;
            }
            out[row * C + ch] = (pl[0] + pl[1]) + pl[2];
        }
    }
}


// ------------------------------------------------------------------------------------------------
// backward of the stage-level lookups: scatter-add of the row gradients into channel-last gradient maps
// (nv, H, W, C): d map[v][tap texel][:] += w_tap * d out[row][:].  A row's channels are contiguous in both tensors, so the
// atomics of one warp are 16-byte vector reductions on consecutive addresses (red.global.add.v4.f32).
// ------------------------------------------------------------------------------------------------
__global__ void index_bwd_kernel(SceneDev sc, const float* __restrict__ pts, int M, int local, int C, const float* __restrict__ g_out, MapSetW maps) {
    const long long row = blockIdx.x;           // v*M + m
    const int v = (int)(row / M), m = (int)(row % M);
    float c[3];
    to_camera(sc.views[v], pts + 3 * (size_t)m, c);
    const float4* g = reinterpret_cast<const float4*>(g_out + row * C);
    if (local) {
        float gx, gy;
        Taps t;
        local_grid_coords(sc, c, gx, gy);
        bilinear_taps(gx, gy, sc.lat_w, sc.lat_h, t);
        float* base = maps.lat + (size_t)v * sc.lat_h * sc.lat_w * C;
        for (int q = threadIdx.x; q < C / 4; q += blockDim.x) {
            const float4 gv = g[q];
            for (int tp = 0; tp < 4; ++tp) {
                const float w = t.w[tp];
                if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(base + (size_t)t.idx[tp] * C) + q, make_float4(gv.x * w, gv.y * w, gv.z * w, gv.w * w));
            }
        }
    } else {
        const float ga[3] = {c[0], c[0], c[1]}, gb[3] = {c[2], c[1], c[2]};
        for (int pi = 0; pi < 3; ++pi) {
            Taps t;
            bilinear_taps(ga[pi], gb[pi], sc.plane_w, sc.plane_h, t);
            float* base = maps.pl[pi] + (size_t)v * sc.plane_h * sc.plane_w * C;
            for (int q = threadIdx.x; q < C / 4; q += blockDim.x) {
                const float4 gv = g[q];
                for (int tp = 0; tp < 4; ++tp) {
                    const float w = t.w[tp];
                    if (w != 0.f) atomicAdd(reinterpret_cast<float4*>(base + (size_t)t.idx[tp] * C) + q, make_float4(gv.x * w, gv.y * w, gv.z * w, gv.w * w));
                }
            }
        }
    }
}

int launch_index_bwd(const NeoScene* sc, const float* pts, int M, int local, const float* g_out, float* g_lat, float* g_xz, float* g_xy,
                     float* g_yz, cudaStream_t s) {
    const int C = local ? kLocalCh : kWorldCh;
    index_bwd_kernel<<<(unsigned)((long long)sc->dev.nv * M), local ? 128 : 32, 0, s>>>(sc->dev, pts, M, local, C, g_out, MapSetW{g_lat, {g_xz, g_xy, g_yz}});
    NEO_LAUNCH_CHECK("index_bwd_kernel");
    return NEO_OK;
}

int launch_index_grid(const NeoScene* sc, const float* pts, int M, float* out, cudaStream_t s) {
    index_kernel<<<(unsigned)((long long)sc->dev.nv * M), 128, 0, s>>>(sc->dev, pts, M, 0, kWorldCh,
                                                                        MapSet{nullptr, {sc->dev.planes_cl[0], sc->dev.planes_cl[1], sc->dev.planes_cl[2]}}, out);
    NEO_LAUNCH_CHECK("index_kernel(grid)");
    return NEO_OK;
}
int launch_index_local(const NeoScene* sc, const float* pts, int M, float* out, cudaStream_t s) {
    index_kernel<<<(unsigned)((long long)sc->dev.nv * M), 128, 0, s>>>(sc->dev, pts, M, 1, kLocalCh, MapSet{sc->dev.latent_cl, {nullptr, nullptr, nullptr}}, out);
    NEO_LAUNCH_CHECK("index_kernel(local)");
    return NEO_OK;
}
// caller-owned maps (training on projected maps): lookups and their scatter-add backward with the scene's cameras / grid geometry
int launch_index_maps(const NeoScene* sc, const float* pts, int M, int C, const float* lat, const float* xz, const float* xy, const float* yz,
                      float* out_local, float* out_world, cudaStream_t s) {
    const unsigned grid = (unsigned)((long long)sc->dev.nv * M);
    const int threads = C >= 128 ? 128 : 64;
    if (lat) { index_kernel<<<grid, threads, 0, s>>>(sc->dev, pts, M, 1, C, MapSet{lat, {nullptr, nullptr, nullptr}}, out_local); NEO_LAUNCH_CHECK("index_kernel(maps, local)"); }
    if (xz) { index_kernel<<<grid, threads, 0, s>>>(sc->dev, pts, M, 0, C, MapSet{nullptr, {xz, xy, yz}}, out_world); NEO_LAUNCH_CHECK("index_kernel(maps, world)"); }
    return NEO_OK;
}
int launch_index_maps_bwd(const NeoScene* sc, const float* pts, int M, int C, const float* g_local, const float* g_world, float* g_lat, float* g_xz,
                          float* g_xy, float* g_yz, cudaStream_t s) {
    const unsigned grid = (unsigned)((long long)sc->dev.nv * M);
    const int threads = C / 4 >= 64 ? 64 : 32;
    if (g_local) { index_bwd_kernel<<<grid, threads, 0, s>>>(sc->dev, pts, M, 1, C, g_local, MapSetW{g_lat, {nullptr, nullptr, nullptr}}); NEO_LAUNCH_CHECK("index_bwd_kernel(maps, local)"); }
    if (g_world) { index_bwd_kernel<<<grid, threads, 0, s>>>(sc->dev, pts, M, 0, C, g_world, MapSetW{nullptr, {g_xz, g_xy, g_yz}}); NEO_LAUNCH_CHECK("index_bwd_kernel(maps, world)"); }
    return NEO_OK;
}

}  // namespace neo
