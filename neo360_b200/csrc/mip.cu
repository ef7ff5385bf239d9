// Mip-NeRF 360 renderer of the reference, fp32 CUDA cores, reference formulation.
// Reference: models/mipnerf360/model.py:30-365; models/mipnerf360/helper.py: max_dilate_weights 152-192, sample_intervals 343-396
// (sorted_interp 207-222), construct_ray_warps 168-172, cast_rays / conical_frustum_to_gaussian / lift_gaussian 278-339,
// contract 33-66 (closed-form Jacobian instead of functorch.jacrev), lift_and_diagonalize 70-73, integrated_pos_enc 77-88,
// compute_alpha_weights 234-260, volumetric_rendering 264-274.
// Structure per level: resample (warp per ray) -> IPE features (warp per sample) -> MLP as a chain of tiled SGEMMs with fused
// bias/ReLU/concat -> compositing (warp per ray).  This is the validation-grade path for this row (no tensor cores yet).
#include "common.cuh"
#include <cuda_fp16.h>

namespace neo {
namespace mip {

constexpr float kEps = 1.1920929e-07f;
constexpr int kFeat = 504, kBasis = 21, kDeg = 12;

// ------------------------------------------------------------------------------------------------
// proposal resampling: dilation + annealed softmax + inverse CDF at interval centres  (one warp per ray)
// shared per warp: t[3n+1 -> P2] | lo[n] | hi[n] | p[n] | wd[P2] | cw[P2]
// ------------------------------------------------------------------------------------------------
__global__ void resample_kernel(const float* __restrict__ s_prev, const float* __restrict__ w_prev, int n_rays, int n_prev, int level,
                                float dilation, float anneal, int n_new, float near, float far, const float* __restrict__ jitter,
                                float* __restrict__ s_out, float* __restrict__ t_out, int p2) {
    extern __shared__ float sm[];
    const int warps = blockDim.x / 32, wid = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int per_warp = 3 * p2 + 3 * n_prev;
    float* td = sm + wid * per_warp;       // sorted dilated positions (p2)
    float* wd = td + p2;                   // dilated weights / pdf (p2)
    float* cw = wd + p2;                   // cdf (p2)
    float* lo = cw + p2;                   // t0_j (n_prev)
    float* hi = lo + n_prev;               // t1_j
    float* pp = hi + n_prev;               // p_j
    const int b = blockIdx.x * warps + wid;
    if (b >= n_rays) return;
    int ns, nw;                            // entries of the (dilated) step function: ns positions, nw = ns-1 weights
    if (level == 0) {
        if (lane == 0) { td[0] = 0.f; td[1] = 1.f; wd[0] = 1.f; }
        ns = 2; nw = 1;
        __syncwarp();
    } else {
        const float* t = s_prev + (size_t)b * (n_prev + 1);
        const float* w = w_prev + (size_t)b * n_prev;
        for (int j = lane; j < n_prev; j += 32) {
            const float a = t[j], c = t[j + 1];
            pp[j] = __fdiv_rn(w[j], fmaxf(sub_(c, a), kEps));          // weight_to_pdf
            lo[j] = sub_(a, dilation);
            hi[j] = add_(c, dilation);
        }
        const int total = 3 * n_prev + 1;
        for (int i = lane; i < p2; i += 32) {
            float v = INFINITY;
            if (i <= n_prev) v = t[i];
            else if (i < 2 * n_prev + 1) v = sub_(t[i - n_prev - 1], dilation);
            else if (i < total) v = add_(t[i - 2 * n_prev], dilation);
            td[i] = v;
        }
        __syncwarp();
        for (int k2 = 2; k2 <= p2; k2 <<= 1)
            for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                for (int i = lane; i < p2; i += 32) {
                    int l = i ^ j2;
                    if (l > i) {
                        float a = td[i], c = td[l];
                        bool up = ((i & k2) == 0);
                        if ((a > c) == up) { td[i] = c; td[l] = a; }
                    }
                }
                __syncwarp();
            }
        for (int i = lane; i < total; i += 32) td[i] = fminf(fmaxf(td[i], 0.f), 1.f);       // clip to the domain (0,1)
        __syncwarp();
        // p_dilate_i = max_j { p_j : t0_j <= td_i < t1_j }, weights = p * dt, renormalise
        float part = 0.f;
        for (int i = lane; i < total - 1; i += 32) {
            const float x = td[i];
            float m = 0.f;
            for (int j = 0; j < n_prev; ++j)
                if (lo[j] <= x && hi[j] > x) m = fmaxf(m, pp[j]);
            const float wv = mul_(m, sub_(td[i + 1], x));
            wd[i] = wv;
            part += wv;
        }
        const float tot = fmaxf(warp_sum(part), kEps);
        __syncwarp();
        // drop first/last: positions td[1..total-2], weights wd[1..total-3]
        ns = total - 2; nw = total - 3;
        for (int i = lane; i < nw; i += 32) cw[i] = __fdiv_rn(wd[i + 1], tot);
        __syncwarp();
        for (int i = lane; i < nw; i += 32) wd[i] = cw[i];
        for (int i = lane; i < ns; i += 32) cw[i] = td[i + 1];
        __syncwarp();
        for (int i = lane; i < ns; i += 32) td[i] = cw[i];
        __syncwarp();
    }
    // logits = anneal * log(w) where the interval is non-empty, else -inf ; softmax
    float mx = -INFINITY;
    for (int i = lane; i < nw; i += 32) {
        const float lg = (td[i + 1] > td[i]) ? mul_(anneal, logf(wd[i])) : -INFINITY;
        wd[i] = lg;
        mx = fmaxf(mx, lg);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int i = lane; i < nw; i += 32) { const float e = expf(wd[i] - mx); wd[i] = e; se += e; }
    se = warp_sum(se);
    __syncwarp();
    // cw = [0, min(cumsum(w[:-1]), 1), 1]  (ns entries)
    float carry = 0.f;
    for (int base = 0; base < nw - 1; base += 32) {
        const int i = base + lane;
        float v = (i < nw - 1) ? __fdiv_rn(wd[i], se) : 0.f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float nb = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = add_(v, nb); }
        v = add_(v, carry);
        if (i < nw - 1) cw[i + 1] = fminf(v, 1.0f);
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
    if (lane == 0) { cw[0] = 0.f; cw[ns - 1] = 1.0f; }
    __syncwarp();
    // centres = sorted_interp(u, cw, td) ; reuse wd for the centres
    for (int k = lane; k < n_new; k += 32) {
        float u;
        if (jitter) {
            const float u_max = kEps + (1.f - kEps) / (float)n_new;
            const float max_jitter = (1.f - u_max) / (float)(n_new - 1) - kEps;
            // torch.linspace(0, 1-u_max, n)[k] + rand * max_jitter
            const float end = 1.f - u_max, step = end / (float)(n_new - 1);
            const float base = (k < n_new / 2) ? step * (float)k : end - step * (float)(n_new - 1 - k);
            u = base + jitter[b] * max_jitter;
        } else {
            const float pad = 1.f / (2.f * (float)n_new);
            const float start = pad, end = 1.f - pad - kEps, step = (end - start) / (float)(n_new - 1);
            u = (n_new == 1) ? start : ((k < n_new / 2) ? start + step * (float)k : end - step * (float)(n_new - 1 - k));
        }
        int a = 0, c = ns;                     // last j with cw[j] <= u
        while (c - a > 1) { int mid = (a + c) >> 1; if (cw[mid] <= u) a = mid; else c = mid; }
        const float x0 = cw[a], x1 = (a + 1 < ns) ? cw[a + 1] : cw[ns - 1];
        const float f0 = td[a], f1 = (a + 1 < ns) ? td[a + 1] : td[ns - 1];
        float off = __fdiv_rn(sub_(u, x0), sub_(x1, x0));
        if (off != off) off = 0.f;
        off = fminf(fmaxf(off, 0.f), 1.f);
        wd[k] = add_(f0, mul_(off, sub_(f1, f0)));
    }
    __syncwarp();
    float* so = s_out + (size_t)b * (n_new + 1);
    float* to = t_out + (size_t)b * (n_new + 1);
    const float sn = 1.f / near, sf = 1.f / far;
    for (int k = lane; k <= n_new; k += 32) {
        float s;
        if (k == 0) s = fmaxf(2.f * wd[0] - 0.5f * (wd[1] + wd[0]), 0.f);
        else if (k == n_new) s = fminf(2.f * wd[n_new - 1] - 0.5f * (wd[n_new - 1] + wd[n_new - 2]), 1.f);
        else s = 0.5f * (wd[k] + wd[k - 1]);
        so[k] = s;
        to[k] = 1.f / (s * sf + (1.f - s) * sn);                                 // s_to_t
    }
}

// ------------------------------------------------------------------------------------------------
// conical frustum -> Gaussian -> contraction -> 21-direction lift -> integrated positional encoding (one warp per sample)
// ------------------------------------------------------------------------------------------------
// conical frustum [t0, t1] of ray b -> Gaussian (helper.py:293-339) -> contraction with its Jacobian (helper.py:33-66): z[3], zc[3][3]
__device__ __forceinline__ void frustum_gaussian(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                 const float* __restrict__ radii, const float* __restrict__ tdist, int b, int k, int n,
                                                 float (&z)[3], float (&zc)[3][3]) {
    const float t0 = tdist[(size_t)b * (n + 1) + k], t1 = tdist[(size_t)b * (n + 1) + k + 1];
    const float d[3] = {rays_d[3 * b], rays_d[3 * b + 1], rays_d[3 * b + 2]};
    const float o[3] = {rays_o[3 * b], rays_o[3 * b + 1], rays_o[3 * b + 2]};
    const float rad = radii[b];
    const float mu = (t0 + t1) / 2.f, hw = (t1 - t0) / 2.f;
    const float denom = fmaxf(3.f * mu * mu + hw * hw, kEps);
    const float t_mean = mu + (2.f * mu * hw * hw) / denom;
    const float hw4 = hw * hw * hw * hw;
    const float t_var = (hw * hw) / 3.f - (4.f / 15.f) * hw4 * (12.f * mu * mu - hw * hw) / (denom * denom);
    const float r_var = ((mu * mu) / 4.f + (5.f / 12.f) * hw * hw - (4.f / 15.f) * hw4 / denom) * rad * rad;
    const float dmag = fmaxf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2], 1e-10f);
    float mean[3], cov[3][3];
    for (int i = 0; i < 3; ++i) {
        mean[i] = d[i] * t_mean + o[i];
        for (int j = 0; j < 3; ++j) cov[i][j] = t_var * d[i] * d[j] + r_var * ((i == j ? 1.f : 0.f) - d[i] * (d[j] / dmag));
    }
    // contraction z = x (r<=1) | ((2r-1)/r^2) x ; J = f I + ((2-2r)/r^4) x x^T
    const float r2 = fmaxf(mean[0] * mean[0] + mean[1] * mean[1] + mean[2] * mean[2], 1e-32f);
    if (r2 <= 1.f) {
        for (int i = 0; i < 3; ++i) { z[i] = mean[i]; for (int j = 0; j < 3; ++j) zc[i][j] = cov[i][j]; }
    } else {
        const float r = sqrtf(r2), f = (2.f * r - 1.f) / r2, g = (2.f - 2.f * r) / (r2 * r2);
        float J[3][3], Tm[3][3];
        for (int i = 0; i < 3; ++i) { z[i] = f * mean[i]; for (int j = 0; j < 3; ++j) J[i][j] = (i == j ? f : 0.f) + g * mean[i] * mean[j]; }
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Tm[i][j] = J[i][0] * cov[0][j] + J[i][1] * cov[1][j] + J[i][2] * cov[2][j];
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) zc[i][j] = Tm[i][0] * J[j][0] + Tm[i][1] * J[j][1] + Tm[i][2] * J[j][2];
    }
}

// fp32 path (tight parity): one warp per sample, the reference's own sin(x), sin(x + pi/2) formulation
__global__ void features_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, const float* __restrict__ radii,
                                const float* __restrict__ tdist, const float* __restrict__ basis, long long M, int n,
                                float* __restrict__ X) {
    const long long m = (long long)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
    const int lane = threadIdx.x % 32;
    if (m >= M) return;
    float z[3], zc[3][3];
    frustum_gaussian(rays_o, rays_d, radii, tdist, (int)(m / n), (int)(m % n), n, z, zc);
    if (lane < kBasis) {
        const float p[3] = {basis[lane], basis[kBasis + lane], basis[2 * kBasis + lane]};
        const float lm = z[0] * p[0] + z[1] * p[1] + z[2] * p[2];
        float cp[3];
        for (int i = 0; i < 3; ++i) cp[i] = zc[i][0] * p[0] + zc[i][1] * p[1] + zc[i][2] * p[2];
        const float lv = p[0] * cp[0] + p[1] * cp[1] + p[2] * cp[2];
        float sc = 1.f;
        float* x = X + (size_t)m * kFeat;
        for (int kk = 0; kk < kDeg; ++kk) {
            const float sm_ = lm * sc, e = expf(-0.5f * (lv * sc * sc));
            x[kk * kBasis + lane] = e * sinf(sm_);
            x[kDeg * kBasis + kk * kBasis + lane] = e * sinf(sm_ + 1.57079637f);
            sc *= 2.f;
        }
    }
}

// tensor-core path: fp16 rows of the activation buffer (row stride ld16 halfs, 504 features + 8 zero columns = 1 KB per sample).
// Thread = (sample, basis direction): 12 samples x 21 directions per block, no idle lanes.  The per-sample Gaussian is computed once
// into shared memory; sin / cos of the 12 octaves come from three accurate sincosf calls (octaves 0, 4, 8) and exact angle doubling in
// between (error <= 8 ulp, below the fp32 rounding of the reference's own `x + pi/2` argument at those magnitudes and far below fp16);
// the row is assembled in shared memory and leaves as 16-byte coalesced stores (the 2-byte scattered stores of the warp-per-sample
// kernel were the bottleneck: r2 launch list, 15.9 % of the Mip-NeRF 360 frame).
constexpr int kFeatSamples = 12, kFeatThreads = kFeatSamples * kBasis;     // 252
__global__ void __launch_bounds__(kFeatThreads) features16_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                                  const float* __restrict__ radii, const float* __restrict__ tdist,
                                                                  const float* __restrict__ basis, long long M, int n,
                                                                  __half* __restrict__ X16, long long ld16) {
    __shared__ float zs[kFeatSamples][12];
    __shared__ __align__(16) __half row[kFeatSamples][kFeat + 8];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.x * kFeatSamples;
    const int nrows = (int)((M - m0) < kFeatSamples ? (M - m0) : kFeatSamples);
    if (tid < nrows) {
        float z[3], zc[3][3];
        const long long m = m0 + tid;
        frustum_gaussian(rays_o, rays_d, radii, tdist, (int)(m / n), (int)(m % n), n, z, zc);
        for (int i = 0; i < 3; ++i) { zs[tid][i] = z[i]; for (int j = 0; j < 3; ++j) zs[tid][3 + 3 * i + j] = zc[i][j]; }
    }
    if (tid < kFeatSamples * 8) row[tid / 8][kFeat + (tid % 8)] = __float2half_rn(0.f);
    __syncthreads();
    const int sidx = tid / kBasis, dir = tid % kBasis;
    if (sidx < nrows) {
        const float p[3] = {basis[dir], basis[kBasis + dir], basis[2 * kBasis + dir]};
        const float* zz = zs[sidx];
        const float lm = zz[0] * p[0] + zz[1] * p[1] + zz[2] * p[2];
        float lv = 0.f;
        for (int i = 0; i < 3; ++i) lv += p[i] * (zz[3 + 3 * i] * p[0] + zz[4 + 3 * i] * p[1] + zz[5 + 3 * i] * p[2]);
        __half* x = row[sidx];
        float sc = 1.f;
#pragma unroll
        for (int g = 0; g < kDeg / 4; ++g) {
            float sn, cs;
            sincosf(lm * sc, &sn, &cs);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int kk = 4 * g + j;
                const float e = __expf(-0.5f * (lv * sc * sc));
                x[kk * kBasis + dir] = __float2half_rn(e * sn);
                x[kDeg * kBasis + kk * kBasis + dir] = __float2half_rn(e * cs);
                const float s2 = 2.f * sn * cs, c2 = (cs - sn) * (cs + sn);
                sn = s2; cs = c2;
                sc *= 2.f;
            }
        }
    }
    __syncthreads();
    constexpr int kVec = (kFeat + 8) / 8;                       // 16-byte pieces per row
    for (int i = tid; i < nrows * kVec; i += kFeatThreads) {
        const int r = i / kVec, c = i % kVec;
        reinterpret_cast<uint4*>(X16 + (size_t)(m0 + r) * ld16)[c] = reinterpret_cast<const uint4*>(row[r])[c];
    }
}

// direction encoding broadcast to samples: DE[m][27]
__global__ void dir16_kernel(const float* __restrict__ viewdirs, long long M, int n, __half* __restrict__ out, long long ld, int pad) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= M * pad) return;
    const long long m = gid / pad;
    const int c = (int)(gid % pad);
    float val = 0.f;
    if (c < 27) {
        const float* d = viewdirs + 3 * (m / n);
        if (c < 3) val = d[c];
        else {
            int q = c - 3;
            const bool shifted = q >= 12;
            if (shifted) q -= 12;
            const float xb = mul_(d[q % 3], (float)(1 << (q / 3)));
            val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
        }
    }
    out[m * ld + c] = __float2half_rn(val);
}
__global__ void dir_kernel(const float* __restrict__ viewdirs, long long M, int n, float* __restrict__ DE) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= M * 27) return;
    const long long m = gid / 27;
    const int c = (int)(gid % 27);
    const float* d = viewdirs + 3 * (m / n);
    float val;
    if (c < 3) val = d[c];
    else {
        int q = c - 3;
        const bool shifted = q >= 12;
        if (shifted) q -= 12;
        const float xb = mul_(d[q % 3], (float)(1 << (q / 3)));
        val = sinf(shifted ? add_(xb, 1.57079637f) : xb);
    }
    DE[gid] = val;
}

// ------------------------------------------------------------------------------------------------
// out[M][N] = act( A1[M][K1] . W[:, 0:K1]^T + A2[M][K2] . W[:, K1:K1+K2]^T + bias )      W is (N, K1+K2) row-major (nn.Linear)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A1, int K1, const float* __restrict__ A2, int K2,
                                                    const float* __restrict__ W, const float* __restrict__ bias, long long M, int N,
                                                    int relu, float* __restrict__ out) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const long long m0 = (long long)blockIdx.x * 64;
    const int n0 = blockIdx.y * 64;
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int ldw = K1 + K2;
    float acc[4][4] = {};
    for (int seg = 0; seg < 2; ++seg) {
        const float* A = seg ? A2 : A1;
        const int K = seg ? K2 : K1, wo = seg ? K1 : 0;
        if (!A || K == 0) continue;
        for (int k0 = 0; k0 < K; k0 += 16) {
            for (int e = threadIdx.x; e < 16 * 64; e += 256) {
                const int mm = e / 16, kk = e % 16;
                const long long mrow = m0 + mm;
                As[kk][mm] = (mrow < M && k0 + kk < K) ? A[(size_t)mrow * K + k0 + kk] : 0.f;
                const int nn = mm;
                Bs[kk][nn] = (n0 + nn < N && k0 + kk < K) ? W[(size_t)(n0 + nn) * ldw + wo + k0 + kk] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) {
                float a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    for (int i = 0; i < 4; ++i) {
        const long long mrow = m0 + ty * 4 + i;
        if (mrow >= M) continue;
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= N) continue;
            float v = acc[i][j] + (bias ? bias[nn] : 0.f);
            if (relu) v = fmaxf(v, 0.f);
            out[(size_t)mrow * N + nn] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// activations + compute_alpha_weights(opaque_background) + volumetric_rendering (white background weight)   (one warp per ray)
// ------------------------------------------------------------------------------------------------
__global__ void composite_kernel(const float* __restrict__ raw_density, const float* __restrict__ raw_rgb, const float* __restrict__ tdist,
                                 const float* __restrict__ rays_d, int n_rays, int n, float* __restrict__ density_out,
                                 float* __restrict__ rgb_s_out, float* __restrict__ w_out, float* __restrict__ rgb_out) {
    const int warps = blockDim.x / 32, wid = threadIdx.x / 32, lane = threadIdx.x % 32;
    const int b = blockIdx.x * warps + wid;
    if (b >= n_rays) return;
    const float* dd3 = rays_d + 3 * b;
    const float dn = __fsqrt_rn(dot3_(dd3, dd3));
    const float* t = tdist + (size_t)b * (n + 1);
    float carry = 0.f, acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool ok = k < n;
        float dens = 0.f, dd = 0.f, col[3] = {0.f, 0.f, 0.f};
        if (ok) {
            const float xs = raw_density[(size_t)b * n + k] - 1.0f;
            dens = xs > 20.f ? xs : log1pf(expf(xs));
            dd = (k == n - 1) ? INFINITY : mul_(dens, mul_(sub_(t[k + 1], t[k]), dn));
            for (int c = 0; c < 3; ++c)
                col[c] = raw_rgb ? (1.f / (1.f + expf(-raw_rgb[((size_t)b * n + k) * 3 + c]))) * 1.002f - 0.001f : 0.f;
        }
        float sc = ok ? ((k == n - 1) ? 0.f : dd) : 0.f;       // cumsum runs over dd[:-1]
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { float nb = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc = add_(sc, nb); }
        const float incl = add_(sc, carry);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = carry;
        if (ok) {
            const float alpha = 1.f - expf(-dd);
            const float wv = alpha * expf(-excl);
            if (density_out) density_out[(size_t)b * n + k] = dens;
            if (w_out) w_out[(size_t)b * n + k] = wv;
            if (rgb_s_out) for (int c = 0; c < 3; ++c) rgb_s_out[((size_t)b * n + k) * 3 + c] = col[c];
            acc += wv; cr += wv * col[0]; cg += wv * col[1]; cb += wv * col[2];
        }
        carry = __shfl_sync(0xffffffffu, incl, 31);
    }
    acc = warp_sum(acc); cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
    if (lane == 0 && rgb_out) {
        const float bg = fmaxf(1.f - acc, 0.f);                 // bg_intensity_range = (1, 1)
        rgb_out[3 * b] = cr + bg; rgb_out[3 * b + 1] = cg + bg; rgb_out[3 * b + 2] = cb + bg;
    }
}

// ---- tensor-core path helpers: out[m][c] = sum_k fp16 h[m][k] * w[c][k] + b[c]  for tiny N (density: 1, rgb: 3) ----
// HBM-bound (one pass over the activation rows).  8 lanes per row, 4 rows per warp: every load instruction fetches four whole
// 128-byte lines; the weights sit in shared memory (fp32, read as broadcast float4); 3 shuffles finish a row.
constexpr int kRowdotIters = 4;          // row groups per warp: 8 warps x 4 rows x 4 = 128 rows per block
template <int N>
__global__ void __launch_bounds__(256) rowdot_f16_kernel(const __half* __restrict__ H, long long ld, int K, const float* __restrict__ Wt,
                                                         const float* __restrict__ b, long long M, float* __restrict__ out) {
    extern __shared__ __align__(16) float wsm[];          // [N][K]
    for (int i = threadIdx.x; i < N * K; i += blockDim.x) wsm[i] = Wt[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & 7, rsel = lane >> 3, warp = threadIdx.x >> 5;
    float bias[N];
#pragma unroll
    for (int c = 0; c < N; ++c) bias[c] = b[c];
#pragma unroll 1
    for (int it = 0; it < kRowdotIters; ++it) {
        const long long m = (((long long)blockIdx.x * kRowdotIters + it) * 8 + warp) * 4 + rsel;
        float acc[N];
#pragma unroll
        for (int c = 0; c < N; ++c) acc[c] = 0.f;
        if (m < M) {
            const __half* h = H + m * ld;
            for (int k = sub * 8; k < K; k += 64) {
                const uint4 v = *reinterpret_cast<const uint4*>(h + k);
                const __half2* hv = reinterpret_cast<const __half2*>(&v);
                const float2 f0 = __half22float2(hv[0]), f1 = __half22float2(hv[1]), f2 = __half22float2(hv[2]), f3 = __half22float2(hv[3]);
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wsm + c * K + k), w1 = *reinterpret_cast<const float4*>(wsm + c * K + k + 4);
                    acc[c] = fmaf(f0.x, w0.x, fmaf(f0.y, w0.y, fmaf(f1.x, w0.z, fmaf(f1.y, w0.w,
                             fmaf(f2.x, w1.x, fmaf(f2.y, w1.y, fmaf(f3.x, w1.z, fmaf(f3.y, w1.w, acc[c]))))))));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c) {
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
            acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
        }
        if (m < M && sub == 0) {
#pragma unroll
            for (int c = 0; c < N; ++c) out[m * N + c] = acc[c] + bias[c];
        }
    }
}

}  // namespace mip
int launch_rowdot_f16(const void* H, long long ld, int K, const float* W, const float* b, int N, long long M, float* out, cudaStream_t s) {
    if (M <= 0) return NEO_OK;
    if ((K % 8) || (ld % 8) || (N != 1 && N != 3)) { set_error("rowdot_f16: K %% 8, ld %% 8 and N in {1, 3} required (K=%d ld=%lld N=%d)", K, ld, N); return NEO_ERR_INVALID; }
    const unsigned grid = (unsigned)((M + 8 * 4 * mip::kRowdotIters - 1) / (8 * 4 * mip::kRowdotIters));
    const size_t smem = (size_t)N * K * sizeof(float);
    if (N == 1) mip::rowdot_f16_kernel<1><<<grid, 256, smem, s>>>((const __half*)H, ld, K, W, b, M, out);
    else mip::rowdot_f16_kernel<3><<<grid, 256, smem, s>>>((const __half*)H, ld, K, W, b, M, out);
    NEO_LAUNCH_CHECK("rowdot_f16_kernel");
    return NEO_OK;
}
// csrc/gemm_tc.cu
int gemm_f16(const void* A, long long lda, const void* W, long long ldw, const float* bias, void* C, long long ldc, long long M, int N, int K,
             int relu, cudaStream_t s);
int f32_to_f16_pad(const float* in, long long rows, int cols_in, long long ld_in, void* out, int cols_out, long long ld_out, cudaStream_t s);
}  // namespace neo

using namespace neo;

namespace {
struct Cv { float* base; size_t used; float* take(size_t n) { n = (n + 63) & ~size_t(63); float* p = base ? base + used : nullptr; used += n; return p; } };
struct WSM {
    float *s[3], *t, *w[3], *X, *Ha, *Hb, *beta, *DE, *V, *rawd, *rawc;
    // tensor-core path (fp16): two activation buffers [M][width + 512] whose tail columns hold the padded IPE features of buffer 0,
    // [M][256 + 64] = bottleneck | padded direction encoding, [M][128], and the packed weights
    void *A16[2], *B16, *V16, *W16;
};
constexpr int kFeatPad = 512, kDirPad = 64;
size_t tc_weight_halves(int width) {      // fp16 elements of one MLP's packed weights (upper bound: the 8-layer NeRF MLP)
    return (size_t)width * kFeatPad + 6 * (size_t)width * width + (size_t)width * (width + kFeatPad) + 256 * (size_t)width + 128 * (256 + kDirPad);
}
size_t carve(Cv& c, int n, const NeoMipCfg* cfg, int width, WSM& w) {
    const int ns[3] = {cfg->n_prop, cfg->n_prop, cfg->n_nerf};
    int nmax = cfg->n_prop > cfg->n_nerf ? cfg->n_prop : cfg->n_nerf;
    for (int l = 0; l < 3; ++l) { w.s[l] = c.take((size_t)n * (ns[l] + 1)); w.w[l] = c.take((size_t)n * ns[l]); }
    w.t = c.take((size_t)n * (nmax + 1));
    const size_t M = (size_t)n * nmax;
    const bool tcp = cfg->precision == NEO_PREC_TC;          // the fp32 activation buffers are not needed on the tensor-core path
    w.X = c.take(tcp ? 0 : M * mip::kFeat);
    w.Ha = c.take(tcp ? 0 : M * width); w.Hb = c.take(tcp ? 0 : M * width);
    w.beta = c.take(tcp ? 0 : M * 256); w.DE = c.take(tcp ? 0 : M * 27); w.V = c.take(tcp ? 0 : M * 128);
    w.rawd = c.take(M); w.rawc = c.take(M * 3);
    if (cfg->precision == NEO_PREC_TC) {
        for (int i = 0; i < 2; ++i) w.A16[i] = c.take((M * (size_t)(width + kFeatPad) + 1) / 2);
        w.B16 = c.take((M * (256 + kDirPad) + 1) / 2);
        w.V16 = c.take((M * 128 + 1) / 2);
        w.W16 = c.take((tc_weight_halves(width) + 1) / 2);
    }
    return c.used * sizeof(float);
}
int check(const NeoMipCfg* c) {
    if (!c || c->n_prop < 2 || c->n_nerf < 2 || c->n_prop > 160 || c->n_nerf > 160) { set_error("mip: sample counts must be in [2,160]"); return NEO_ERR_INVALID; }
    if (!(c->near_plane > 0.f) || !(c->far_plane > c->near_plane)) { set_error("mip: need 0 < near < far"); return NEO_ERR_INVALID; }
    if (c->precision != NEO_PREC_FP32 && c->precision != NEO_PREC_TC) { set_error("mip: bad precision %d", c->precision); return NEO_ERR_INVALID; }
    return NEO_OK;
}
int gemm(const float* A1, int K1, const float* A2, int K2, const float* W, const float* b, long long M, int N, int relu, float* out, cudaStream_t s) {
    dim3 grid((unsigned)((M + 63) / 64), (unsigned)((N + 63) / 64));
    mip::sgemm_kernel<<<grid, 256, 0, s>>>(A1, K1, A2, K2, W, b, M, N, relu, out);
    NEO_LAUNCH_CHECK("mip sgemm_kernel");
    return NEO_OK;
}

// One MLP of Mip-NeRF 360 (models/mipnerf360/model.py:30-195) on tensor cores: fp16 activations, every dense layer a gemm_f16 launch
// (csrc/gemm_tc.cu), the skip concatenation by laying h4 and the features out in ONE buffer (layer 5 is a single K = width + 512 GEMM).
int mlp_tc(const NeoMipMLPParams& p, const WSM& w, long long M, int n, const float* viewdirs, float* rawd, float* rawc, cudaStream_t s) {
    const int W = p.width, ld = W + kFeatPad;
    if (W % 64) { set_error("mip tc: width must be a multiple of 64 (got %d)", W); return NEO_ERR_UNSUPPORTED; }
    __half* buf[2] = {(__half*)w.A16[0], (__half*)w.A16[1]};
    __half* wp = (__half*)w.W16;
    int rc;
    // the IPE features were written by features_kernel as fp16 into the tail columns of buffer 0 (zero padded to 512)
    auto pack = [&](const float* src, int rows, int cols_in, int cols_out) -> __half* {
        __half* dst = wp;
        wp += (size_t)rows * cols_out;
        return f32_to_f16_pad(src, rows, cols_in, cols_in, dst, cols_out, cols_out, s) ? nullptr : dst;
    };
    __half* w0 = pack(p.w[0], W, mip::kFeat, kFeatPad);
    if (!w0) return NEO_ERR_CUDA;
    if ((rc = gemm_f16(buf[0] + W, ld, w0, kFeatPad, p.b[0], buf[0], ld, M, W, kFeatPad, 1, s))) return rc;
    for (int l = 1; l < p.depth; ++l) {
        const bool skip_in = (l == 5);                      // cat([h, inputs]) after layer 4 feeds layer 5 (model.py:122-128)
        __half* wl;
        int K;
        if (skip_in) {
            // W5 is (width, width + 504): columns [0, width) stay, the 504 feature columns are padded to 512
            wl = wp;
            wp += (size_t)W * (W + kFeatPad);
            if ((rc = f32_to_f16_pad(p.w[l], W, W, W + mip::kFeat, wl, W, W + kFeatPad, s))) return rc;
            if ((rc = f32_to_f16_pad(p.w[l] + W, W, mip::kFeat, W + mip::kFeat, wl + W, kFeatPad, W + kFeatPad, s))) return rc;
            K = W + kFeatPad;
        } else {
            wl = pack(p.w[l], W, W, W);
            if (!wl) return NEO_ERR_CUDA;
            K = W;
        }
        if ((rc = gemm_f16(buf[(l - 1) & 1], ld, wl, K, p.b[l], buf[l & 1], ld, M, W, K, 1, s))) return rc;
    }
    const __half* h = buf[(p.depth - 1) & 1];
    if ((rc = launch_rowdot_f16(h, ld, W, p.wsig, p.bsig, 1, M, rawd, s))) return rc;
    if (p.wrgb) {
        __half* B = (__half*)w.B16;
        __half* V = (__half*)w.V16;
        const int ldb = 256 + kDirPad;
        __half* wb = pack(p.wb, 256, W, W);
        if (!wb) return NEO_ERR_CUDA;
        if ((rc = gemm_f16(h, ld, wb, W, p.bb, B, ldb, M, 256, W, 0, s))) return rc;
        mip::dir16_kernel<<<(unsigned)((M * kDirPad + 255) / 256), 256, 0, s>>>(viewdirs, M, n, B + 256, ldb, kDirPad);
        NEO_LAUNCH_CHECK("mip dir16_kernel");
        __half* wv = wp;
        wp += (size_t)128 * ldb;
        if ((rc = f32_to_f16_pad(p.wv0, 128, 256, 256 + 27, wv, 256, ldb, s))) return rc;
        if ((rc = f32_to_f16_pad(p.wv0 + 256, 128, 27, 256 + 27, wv + 256, kDirPad, ldb, s))) return rc;
        if ((rc = gemm_f16(B, ldb, wv, ldb, p.bv0, V, 128, M, 128, ldb, 1, s))) return rc;
        if ((rc = launch_rowdot_f16(V, 128, 128, p.wrgb, p.brgb, 3, M, rawc, s))) return rc;
    }
    return NEO_OK;
}
}  // namespace

extern "C" size_t neo_mip_workspace_bytes(int n_rays, const NeoMipCfg* cfg, int nerf_width) {
    if (n_rays <= 0 || check(cfg) || nerf_width < 64) return 0;
    Cv c{nullptr, 0};
    WSM w;
    return carve(c, n_rays, cfg, nerf_width > 256 ? nerf_width : 256, w);
}

extern "C" int neo_mip_render_fwd(const NeoMipMLPParams mlps[3], const float* rays_o, const float* rays_d, const float* viewdirs,
                                  const float* radii, int n_rays, const NeoMipCfg* cfg, NeoMipOut* out, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    if (!mlps || !rays_o || !rays_d || !viewdirs || !radii || !out || n_rays <= 0) { set_error("neo_mip_render_fwd: bad arguments"); return NEO_ERR_INVALID; }
    int rc = check(cfg);
    if (rc) return rc;
    for (int l = 0; l < 3; ++l) {
        const NeoMipMLPParams& p = mlps[l];
        if (p.depth < 1 || p.depth > 8 || p.width < 64 || p.width % 4 || !p.basis) { set_error("mip mlp %d: bad depth/width", l); return NEO_ERR_INVALID; }
        if ((l == 2) != (p.wrgb != nullptr)) { set_error("mip: mlps must be {prop, prop, nerf}"); return NEO_ERR_INVALID; }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int width = mlps[2].width > 256 ? mlps[2].width : 256;
    Cv c{reinterpret_cast<float*>(workspace), 0};
    WSM w;
    size_t need = carve(c, n_rays, cfg, width, w);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes, got %zu", need, workspace_bytes); return NEO_ERR_WORKSPACE; }
    const int ns[3] = {cfg->n_prop, cfg->n_prop, cfg->n_nerf};
    const float anneal = (10.f * cfg->train_frac) / (9.f * cfg->train_frac + 1.f);      // bias(train_frac, anneal_slope=10)
    long long prod = 1;
    for (int lvl = 0; lvl < 3; ++lvl) {
        const int n = ns[lvl], n_prev = lvl ? ns[lvl - 1] : 1;
        const float dilation = 0.0025f + 0.5f / (float)prod;
        prod *= n;
        int p2 = 2;
        while (p2 < 3 * n_prev + 1 || p2 < n + 1) p2 <<= 1;
        const int warps = 4;
        const size_t smem = (size_t)warps * (3 * p2 + 3 * n_prev) * sizeof(float);
        if (smem > 48 * 1024) NEO_CUDA(cudaFuncSetAttribute(mip::resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mip::resample_kernel<<<(n_rays + warps - 1) / warps, warps * 32, smem, s>>>(lvl ? w.s[lvl - 1] : nullptr, lvl ? w.w[lvl - 1] : nullptr, n_rays,
                                                                                   n_prev, lvl, dilation, anneal, n, cfg->near_plane, cfg->far_plane,
                                                                                   cfg->jitter[lvl], w.s[lvl], w.t, p2);
        NEO_LAUNCH_CHECK("mip resample_kernel");
        const long long M = (long long)n_rays * n;
        const bool tcp = cfg->precision == NEO_PREC_TC;
        if (tcp) mip::features16_kernel<<<(unsigned)((M + mip::kFeatSamples - 1) / mip::kFeatSamples), mip::kFeatThreads, 0, s>>>(
                     rays_o, rays_d, radii, w.t, mlps[lvl].basis, M, n, (__half*)w.A16[0] + mlps[lvl].width, mlps[lvl].width + kFeatPad);
        else mip::features_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(rays_o, rays_d, radii, w.t, mlps[lvl].basis, M, n, w.X);
        NEO_LAUNCH_CHECK("mip features_kernel");
        const NeoMipMLPParams& p = mlps[lvl];
        const float* rawc = nullptr;
        if (cfg->precision == NEO_PREC_TC) {
            if ((rc = mlp_tc(p, w, M, n, viewdirs, w.rawd, w.rawc, s))) return rc;
            if (p.wrgb) rawc = w.rawc;
            else if (out->rgb_s[lvl]) NEO_CUDA(cudaMemsetAsync(out->rgb_s[lvl], 0, (size_t)M * 3 * sizeof(float), s));
        } else {
            float* src = w.Ha;
            float* dst = w.Hb;
            if ((rc = gemm(w.X, mip::kFeat, nullptr, 0, p.w[0], p.b[0], M, p.width, 1, src, s))) return rc;
            for (int l = 1; l < p.depth; ++l) {
                const bool skip_in = (l == 5);                      // cat([h, inputs]) after layer 4 feeds layer 5
                if ((rc = gemm(src, p.width, skip_in ? w.X : nullptr, skip_in ? mip::kFeat : 0, p.w[l], p.b[l], M, p.width, 1, dst, s))) return rc;
                float* tmp = src; src = dst; dst = tmp;
            }
            if ((rc = gemm(src, p.width, nullptr, 0, p.wsig, p.bsig, M, 1, 0, w.rawd, s))) return rc;
                    if (p.wrgb) {
                if ((rc = gemm(src, p.width, nullptr, 0, p.wb, p.bb, M, 256, 0, w.beta, s))) return rc;
                mip::dir_kernel<<<(unsigned)((M * 27 + 255) / 256), 256, 0, s>>>(viewdirs, M, n, w.DE);
                NEO_LAUNCH_CHECK("mip dir_kernel");
                if ((rc = gemm(w.beta, 256, w.DE, 27, p.wv0, p.bv0, M, 128, 1, w.V, s))) return rc;
                if ((rc = gemm(w.V, 128, nullptr, 0, p.wrgb, p.brgb, M, 3, 0, w.rawc, s))) return rc;
                rawc = w.rawc;
            } else {
                if (out->rgb_s[lvl]) NEO_CUDA(cudaMemsetAsync(out->rgb_s[lvl], 0, (size_t)M * 3 * sizeof(float), s));     // disable_rgb: zeros
            }
        }
        const int cw = 8;
        mip::composite_kernel<<<(n_rays + cw - 1) / cw, cw * 32, 0, s>>>(w.rawd, rawc, w.t, rays_d, n_rays, n, out->density[lvl],
                                                                        rawc ? out->rgb_s[lvl] : nullptr, w.w[lvl], out->rgb[lvl]);
        NEO_LAUNCH_CHECK("mip composite_kernel");
        if (out->sdist[lvl]) NEO_CUDA(cudaMemcpyAsync(out->sdist[lvl], w.s[lvl], (size_t)n_rays * (n + 1) * sizeof(float), cudaMemcpyDeviceToDevice, s));
        if (out->weights[lvl]) NEO_CUDA(cudaMemcpyAsync(out->weights[lvl], w.w[lvl], (size_t)M * sizeof(float), cudaMemcpyDeviceToDevice, s));
    }
    return NEO_OK;
}
