// Training-mode BatchNorm + activation (forward apply, backward reduce / apply) on NHWC fp16 tensors with fp32
// statistics, and the fused SGD-Nesterov step over a flat fp32 parameter buffer.
// Reference semantics: nn.BatchNorm2d(momentum=0.1, eps=1e-5) + activation under autograd (models.py:100-113),
// optimizer train.py:135-144.  The batch sums themselves come out of the conv epilogue (conv_tc.cuh).
#include "b200yolo.h"
#include "common.cuh"

using namespace b2y;

static inline int grid_for(long long n, int block, int cap = 148 * 16) {
    long long g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ s1, const float* __restrict__ s2, float count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd,
                                   float* __restrict__ scale, float* __restrict__ shift, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float mean = s1[c] / count;
    float var = s2[c] / count - mean * mean;  // biased variance used for normalisation
    var = fmaxf(var, 0.f);
    const float invstd = 1.f / sqrtf(var + eps);
    if (rmean != nullptr) {
        const float unbiased = count > 1.f ? var * (count / (count - 1.f)) : var;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * unbiased;
    }
    save_mean[c] = mean;
    save_invstd[c] = invstd;
    const float g = gamma != nullptr ? gamma[c] : 1.f;
    const float sc = g * invstd;
    scale[c] = sc;
    shift[c] = (beta != nullptr ? beta[c] : 0.f) - mean * sc;
}

extern "C" int b2y_bn_finalize(const float* stat_sum, const float* stat_sqsum, long long count, const float* gamma,
                               const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                               float* save_mean, float* save_invstd, float* scale, float* shift, int c,
                               void* stream) {
    if (!stat_sum || !stat_sqsum || !save_mean || !save_invstd || !scale || !shift || c <= 0 || count <= 0)
        return B2Y_ERR_INVALID;
    bn_finalize_kernel<<<(c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        stat_sum, stat_sqsum, (float)count, gamma, beta, eps, momentum, running_mean, running_var, save_mean,
        save_invstd, scale, shift, c);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void bn_act_fwd_kernel(const __half* __restrict__ x, long long xp, const float* __restrict__ scale,
                                  const float* __restrict__ shift, const __half* __restrict__ res, long long rp,
                                  __half* __restrict__ y, long long yp, long long pixels, int C, int act,
                                  float slope) {
    const int CV = C / 8;
    const long long total = pixels * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        const long long pix = idx / CV;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + pix * xp) + cv);
        const __half* h = reinterpret_cast<const __half*>(&v);
        const float4 sa = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2);
        const float4 sb = __ldg(reinterpret_cast<const float4*>(scale) + cv * 2 + 1);
        const float4 ta = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2);
        const float4 tb = __ldg(reinterpret_cast<const float4*>(shift) + cv * 2 + 1);
        const float sc[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float sh[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = apply_act(fmaf(__half2float(h[j]), sc[j], sh[j]), act, slope);
        if (res != nullptr) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res + pix * rp) + cv);
            const __half* rh = reinterpret_cast<const __half*>(&rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += __half2float(rh[j]);
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(r[2 * j], r[2 * j + 1]);
        reinterpret_cast<uint4*>(y + pix * yp)[cv] = o;
    }
}

// v2 layout for the three BatchNorm passes: a thread owns ONE 8-channel vector (its per-channel coefficients live in
// registers for the whole kernel) and walks the pixels with a fixed stride, so the loop body is loads + math + store:
// no index divisions, no per-element coefficient loads, the activation resolved at compile time.  Needs 256 % (C/8) == 0.
template <int ACT>
__global__ void __launch_bounds__(256)
bn_act_fwd_v2_kernel(const __half* __restrict__ x, long long xp, const float* __restrict__ scale,
                     const float* __restrict__ shift, const __half* __restrict__ res, long long rp,
                     __half* __restrict__ y, long long yp, long long pixels, int CV, float slope) {
    const int cv = threadIdx.x % CV;
    const int ppb = 256 / CV;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = __ldg(scale + cv * 8 + j);
        sh[j] = __ldg(shift + cv * 8 + j);
    }
    for (long long pix = (long long)blockIdx.x * ppb + threadIdx.x / CV; pix < pixels; pix += (long long)gridDim.x * ppb) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + pix * xp) + cv);
        const __half* h = reinterpret_cast<const __half*>(&v);
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = apply_act(fmaf(__half2float(h[j]), sc[j], sh[j]), ACT, slope);
        if (res != nullptr) {
            const uint4 rv = __ldg(reinterpret_cast<const uint4*>(res + pix * rp) + cv);
            const __half* rh = reinterpret_cast<const __half*>(&rv);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += __half2float(rh[j]);
        }
        uint4 o;
        __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(r[2 * j], r[2 * j + 1]);
        reinterpret_cast<uint4*>(y + pix * yp)[cv] = o;
    }
}

static inline int v2_grid(long long pixels, int CV) {
    const int ppb = 256 / CV;
    long long g = (pixels + ppb - 1) / ppb;
    const long long cap = 148 * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static inline bool v2_ok(int c) { return c % 8 == 0 && c / 8 <= 256 && 256 % (c / 8) == 0; }

extern "C" int b2y_bn_act_fwd(const void* x, long long x_pitch, const float* scale, const float* shift,
                              const void* residual, long long res_pitch, void* y, long long y_pitch,
                              long long pixels, int c, int act, float slope, void* stream) {
    if (!x || !y || !scale || !shift || c % 8 != 0 || x_pitch % 8 != 0 || y_pitch % 8 != 0) return B2Y_ERR_INVALID;
    if (residual != nullptr && res_pitch % 8 != 0) return B2Y_ERR_INVALID;
    if (v2_ok(c) && (act == B2Y_ACT_LEAKY || act == B2Y_ACT_MISH || act == B2Y_ACT_LINEAR)) {
        const int CV = c / 8;
        const int grid = v2_grid(pixels, CV);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2Y_FWD_V2(A)                                                                                                  \
    bn_act_fwd_v2_kernel<A><<<grid, 256, 0, st>>>(reinterpret_cast<const __half*>(x), x_pitch, scale, shift,           \
                                                  reinterpret_cast<const __half*>(residual), res_pitch,               \
                                                  reinterpret_cast<__half*>(y), y_pitch, pixels, CV, slope)
        if (act == B2Y_ACT_LEAKY) B2Y_FWD_V2(B2Y_ACT_LEAKY);
        else if (act == B2Y_ACT_MISH) B2Y_FWD_V2(B2Y_ACT_MISH);
        else B2Y_FWD_V2(B2Y_ACT_LINEAR);
#undef B2Y_FWD_V2
        B2Y_CUDA_CHECK(cudaGetLastError());
        return B2Y_OK;
    }
    bn_act_fwd_kernel<<<grid_for(pixels * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half*>(x), x_pitch, scale, shift, reinterpret_cast<const __half*>(residual), res_pitch,
        reinterpret_cast<__half*>(y), y_pitch, pixels, c, act, slope);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------
// backward pass 1: dbeta[c] = sum du, dgamma[c] = sum du * xhat   with  u = x*scale+shift, du = dy*act'(u)
// grid.y tiles the channels in slabs of 256 (32 lanes x 8 channels); each warp walks a strip of pixels.
// ------------------------------------------------------------------------------------------------
template <typename GT>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy, long long dp,
                         const float* __restrict__ scale, const float* __restrict__ shift,
                         const float* __restrict__ mean, const float* __restrict__ invstd,
                         float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ du_absmax,
                         long long pixels, int C, int act, float slope) {
    __shared__ float red[8][32][17];
    float amax = 0.f;
    const int CV = C / 8;
    const int cv0 = blockIdx.y * 32;
    const int cvn = min(32, CV - cv0);                 // channel vectors in this slab
    // lanes -> (pixel sub-index, channel vector); pack several pixels per warp when the slab is narrow
    int lanes_c = 1;
    while (lanes_c < cvn) lanes_c <<= 1;               // next pow2 >= cvn (<= 32)
    const int ppw = 32 / lanes_c;                      // pixels per warp iteration
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int lc = lane % lanes_c, lp = lane / lanes_c;
    const bool active = lc < cvn;
    const int cv = cv0 + lc;
    float sc[8], sh[8], mu[8], is[8], gb[8], gg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gb[j] = gg[j] = 0.f;
        const int c = cv * 8 + j;
        sc[j] = active ? scale[c] : 0.f;
        sh[j] = active ? shift[c] : 0.f;
        mu[j] = (active && mean) ? mean[c] : 0.f;
        is[j] = (active && invstd) ? invstd[c] : 1.f;
    }
    const long long warps_total = (long long)gridDim.x * 8;
    const long long wid = (long long)blockIdx.x * 8 + warp;
    for (long long p0 = wid * ppw; p0 < pixels; p0 += warps_total * ppw) {
        const long long pix = p0 + lp;
        if (active && pix < pixels) {
            float xf8[8], g8[8];
            Half8<__half>::load(x + pix * xp + cv * 8, xf8);
            Half8<GT>::load(dy + pix * dp + cv * 8, g8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float xf = xf8[j];
                const float u = fmaf(xf, sc[j], sh[j]);
                const float du = g8[j] * act_grad(u, act, slope);
                gb[j] += du;
                gg[j] += du * ((xf - mu[j]) * is[j]);
                amax = fmaxf(amax, fabsf(du));
            }
        }
    }
    if (du_absmax != nullptr) {   // max |du| over the tensor (non-negative floats order like their bit patterns)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if (lane == 0 && amax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(du_absmax), __float_as_uint(amax));
    }
    // combine the pixel sub-lanes of a warp, then the 8 warps of the CTA, then one atomic per channel
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        for (int o = lanes_c; o < 32; o <<= 1) {
            gb[j] += __shfl_xor_sync(0xffffffffu, gb[j], o);
            gg[j] += __shfl_xor_sync(0xffffffffu, gg[j], o);
        }
    }
    if (lp == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            red[warp][lc][j] = gb[j];
            red[warp][lc][8 + j] = gg[j];
        }
    }
    __syncthreads();
    if (warp == 0 && lp == 0 && active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float b = 0.f, g = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                b += red[w][lc][j];
                g += red[w][lc][8 + j];
            }
            atomicAdd(dbeta + cv * 8 + j, b);
            if (dgamma != nullptr) atomicAdd(dgamma + cv * 8 + j, g);
        }
    }
}

template <typename GT, int ACT>
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_v2_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy, long long dp,
                            const float* __restrict__ scale, const float* __restrict__ shift,
                            const float* __restrict__ mean, const float* __restrict__ invstd,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ du_absmax,
                            long long pixels, int CV, float slope) {
    __shared__ float red[256][17];
    const int cv = threadIdx.x % CV;
    const int ppb = 256 / CV;
    float sc[8], sh[8], mu[8], is[8], gb[8], gg[8];
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        gb[j] = gg[j] = 0.f;
        const int c = cv * 8 + j;
        sc[j] = __ldg(scale + c);
        sh[j] = __ldg(shift + c);
        mu[j] = mean ? __ldg(mean + c) : 0.f;
        is[j] = invstd ? __ldg(invstd + c) : 1.f;
    }
    for (long long pix = (long long)blockIdx.x * ppb + threadIdx.x / CV; pix < pixels; pix += (long long)gridDim.x * ppb) {
        float xf8[8], g8[8];
        Half8<__half>::load(x + pix * xp + cv * 8, xf8);
        Half8<GT>::load(dy + pix * dp + cv * 8, g8);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xf = xf8[j];
            const float u = fmaf(xf, sc[j], sh[j]);
            const float du = g8[j] * act_grad(u, ACT, slope);
            gb[j] += du;
            gg[j] += du * ((xf - mu[j]) * is[j]);
            amax = fmaxf(amax, fabsf(du));
        }
    }
    if (du_absmax != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        if ((threadIdx.x & 31) == 0 && amax > 0.f) atomicMax(reinterpret_cast<unsigned int*>(du_absmax), __float_as_uint(amax));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[threadIdx.x][j] = gb[j];
        red[threadIdx.x][8 + j] = gg[j];
    }
    __syncthreads();
    // threads 0 .. CV*16-1: one (channel vector, component) each, summed over the 256/CV pixel sub-rows of the CTA
    for (int i = threadIdx.x; i < CV * 16; i += 256) {
        const int v = i / 16, comp = i % 16;
        float sum = 0.f;
        for (int k = 0; k < ppb; ++k) sum += red[v + k * CV][comp];
        if (comp < 8)
            atomicAdd(dbeta + v * 8 + comp, sum);
        else if (dgamma != nullptr)
            atomicAdd(dgamma + v * 8 + comp - 8, sum);
    }
}

extern "C" int b2y_bn_act_bwd_reduce(const void* x, long long x_pitch, const void* dy, long long dy_pitch,
                                     const float* scale, const float* shift, const float* save_mean,
                                     const float* save_invstd, float* dgamma, float* dbeta, float* du_absmax,
                                     long long pixels, int c, int act, float slope, int grad_dtype, void* stream) {
    if (!x || !dy || !scale || !shift || !dbeta || c % 8 != 0 || x_pitch % 8 != 0 || dy_pitch % 8 != 0)
        return B2Y_ERR_INVALID;
    const int CV = c / 8;
    if (v2_ok(c) && (act == B2Y_ACT_LEAKY || act == B2Y_ACT_MISH || act == B2Y_ACT_LINEAR)) {
        const int g2 = v2_grid(pixels, CV);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2Y_RED_V2(T, A)                                                                                               \
    bn_act_bwd_reduce_v2_kernel<T, A><<<g2, 256, 0, st>>>(reinterpret_cast<const __half*>(x), x_pitch,                  \
                                                          reinterpret_cast<const T*>(dy), dy_pitch, scale, shift,      \
                                                          save_mean, save_invstd, dgamma, dbeta, du_absmax, pixels,   \
                                                          CV, slope)
        if (grad_dtype == B2Y_DT_BF16) {
            if (act == B2Y_ACT_LEAKY) B2Y_RED_V2(__nv_bfloat16, B2Y_ACT_LEAKY);
            else if (act == B2Y_ACT_MISH) B2Y_RED_V2(__nv_bfloat16, B2Y_ACT_MISH);
            else B2Y_RED_V2(__nv_bfloat16, B2Y_ACT_LINEAR);
        } else {
            if (act == B2Y_ACT_LEAKY) B2Y_RED_V2(__half, B2Y_ACT_LEAKY);
            else if (act == B2Y_ACT_MISH) B2Y_RED_V2(__half, B2Y_ACT_MISH);
            else B2Y_RED_V2(__half, B2Y_ACT_LINEAR);
        }
#undef B2Y_RED_V2
        B2Y_CUDA_CHECK(cudaGetLastError());
        return B2Y_OK;
    }
    dim3 grid(1, (CV + 31) / 32);
    long long want = (pixels + 63) / 64;
    grid.x = (unsigned)(want < 1 ? 1 : (want > 148 * 4 ? 148 * 4 : want));
    if (grad_dtype == B2Y_DT_BF16)
        bn_act_bwd_reduce_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, scale,
            shift, save_mean, save_invstd, dgamma, dbeta, du_absmax, pixels, c, act, slope);
    else
        bn_act_bwd_reduce_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __half*>(dy), dy_pitch, scale, shift,
            save_mean, save_invstd, dgamma, dbeta, du_absmax, pixels, c, act, slope);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// backward pass 2: dx = gamma*invstd * (du - dbeta/N - xhat*dgamma/N)
// The data gradient dz feeds two tensor-core GEMMs whose operands must share one 16-bit format with the fp16
// activations / weights, so dz is written in fp16 multiplied by a per-layer power of two s chosen on the device from
// a bound of max|dz| (no host sync); the GEMM epilogues multiply by 1/s (scale_out[1]) read from device memory.
template <typename GT>
__global__ void bn_act_bwd_apply_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy,
                                        long long dp, const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ dgamma,
                                        const float* __restrict__ dbeta, __half* __restrict__ dx, long long dxp,
                                        long long pixels, int C, int act, float slope,
                                        const float* __restrict__ du_absmax, float* __restrict__ scale_out) {
    const int CV = C / 8;
    const long long total = pixels * CV;
    const float inv_n = 1.f / (float)pixels;
    // every CTA derives the same scale: bound = max_c |gamma_c*invstd_c| * (max|du| + |dbeta_c|/N + 16*|dgamma_c|/N)
    __shared__ float s_red[32];
    __shared__ float s_scale;
    float bound = 0.f;
    {
        const float dumax = du_absmax != nullptr ? *du_absmax : 1.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float g = gamma != nullptr ? gamma[c] : 1.f;
            const float b = fabsf(g * invstd[c]) * (dumax + fabsf(dbeta[c]) * inv_n + 16.f * fabsf(dgamma[c]) * inv_n);
            bound = fmaxf(bound, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bound;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s_red[w]);
            float sc = 1.f;
            if (m > 0.f && m < 3.0e38f) sc = exp2f(floorf(log2f(4096.f / m)));   // target max |dz|*s <= 2^12
            sc = fminf(fmaxf(sc, 1.0e-30f), 1.0e30f);
            s_scale = sc;
            if (blockIdx.x == 0 && scale_out != nullptr) {
                scale_out[0] = sc;
                scale_out[1] = 1.f / sc;
            }
        }
        __syncthreads();
    }
    const float sc = s_scale;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        const long long pix = idx / CV;
        float xf8[8], g8[8];
        Half8<__half>::load(x + pix * xp + cv * 8, xf8);
        Half8<GT>::load(dy + pix * dp + cv * 8, g8);
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            const float xf = xf8[j];
            const float u = fmaf(xf, scale[c], shift[c]);
            const float du = g8[j] * act_grad(u, act, slope);
            const float xhat = (xf - mean[c]) * invstd[c];
            const float g = gamma != nullptr ? gamma[c] : 1.f;
            r[j] = sc * (g * invstd[c] * (du - dbeta[c] * inv_n - xhat * dgamma[c] * inv_n));
        }
        Half8<__half>::store(dx + pix * dxp + cv * 8, r);
    }
}

template <typename GT, int ACT>
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_v2_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy, long long dp,
                           const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ gamma, const float* __restrict__ mean,
                           const float* __restrict__ invstd, const float* __restrict__ dgamma,
                           const float* __restrict__ dbeta, __half* __restrict__ dx, long long dxp, long long pixels,
                           int CV, float slope, const float* __restrict__ du_absmax, float* __restrict__ scale_out) {
    const int C = CV * 8;
    const float inv_n = 1.f / (float)pixels;
    __shared__ float s_red[8];
    __shared__ float s_scale;
    float bound = 0.f;
    {   // same bound / scale as bn_act_bwd_apply_kernel (every CTA derives it identically)
        const float dumax = du_absmax != nullptr ? *du_absmax : 1.f;
        for (int c = threadIdx.x; c < C; c += 256) {
            const float g = gamma != nullptr ? gamma[c] : 1.f;
            const float b = fabsf(g * invstd[c]) * (dumax + fabsf(dbeta[c]) * inv_n + 16.f * fabsf(dgamma[c]) * inv_n);
            bound = fmaxf(bound, b);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = bound;
        __syncthreads();
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int w = 0; w < 8; ++w) m = fmaxf(m, s_red[w]);
            float sc = 1.f;
            if (m > 0.f && m < 3.0e38f) sc = exp2f(floorf(log2f(4096.f / m)));
            sc = fminf(fmaxf(sc, 1.0e-30f), 1.0e30f);
            s_scale = sc;
            if (blockIdx.x == 0 && scale_out != nullptr) {
                scale_out[0] = sc;
                scale_out[1] = 1.f / sc;
            }
        }
        __syncthreads();
    }
    const float scl = s_scale;
    const int cv = threadIdx.x % CV;
    const int ppb = 256 / CV;
    float sc[8], sh[8], mu[8], is[8], gi[8], db[8], dg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cv * 8 + j;
        sc[j] = __ldg(scale + c);
        sh[j] = __ldg(shift + c);
        mu[j] = __ldg(mean + c);
        is[j] = __ldg(invstd + c);
        gi[j] = (gamma != nullptr ? __ldg(gamma + c) : 1.f) * is[j];
        db[j] = __ldg(dbeta + c) * inv_n;
        dg[j] = __ldg(dgamma + c) * inv_n;
    }
    for (long long pix = (long long)blockIdx.x * ppb + threadIdx.x / CV; pix < pixels; pix += (long long)gridDim.x * ppb) {
        float xf8[8], g8[8];
        Half8<__half>::load(x + pix * xp + cv * 8, xf8);
        Half8<GT>::load(dy + pix * dp + cv * 8, g8);
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float xf = xf8[j];
            const float u = fmaf(xf, sc[j], sh[j]);
            const float du = g8[j] * act_grad(u, ACT, slope);
            const float xhat = (xf - mu[j]) * is[j];
            r[j] = scl * (gi[j] * (du - db[j] - xhat * dg[j]));
        }
        Half8<__half>::store(dx + pix * dxp + cv * 8, r);
    }
}

extern "C" int b2y_bn_act_bwd_apply(const void* x, long long x_pitch, const void* dy, long long dy_pitch,
                                    const float* scale, const float* shift, const float* gamma,
                                    const float* save_mean, const float* save_invstd, const float* dgamma,
                                    const float* dbeta, void* dx, long long dx_pitch, long long pixels, int c, int act,
                                    float slope, int grad_dtype, const float* du_absmax, float* scale_out,
                                    void* stream) {
    if (!x || !dy || !scale || !shift || !save_mean || !save_invstd || !dgamma || !dbeta || !dx || c % 8 != 0)
        return B2Y_ERR_INVALID;
    if (v2_ok(c) && (act == B2Y_ACT_LEAKY || act == B2Y_ACT_MISH || act == B2Y_ACT_LINEAR)) {
        const int CV = c / 8;
        const int g2 = v2_grid(pixels, CV);
        cudaStream_t st = static_cast<cudaStream_t>(stream);
#define B2Y_APP_V2(T, A)                                                                                               \
    bn_act_bwd_apply_v2_kernel<T, A><<<g2, 256, 0, st>>>(                                                               \
        reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const T*>(dy), dy_pitch, scale, shift, gamma,    \
        save_mean, save_invstd, dgamma, dbeta, reinterpret_cast<__half*>(dx), dx_pitch, pixels, CV, slope, du_absmax,  \
        scale_out)
        if (grad_dtype == B2Y_DT_BF16) {
            if (act == B2Y_ACT_LEAKY) B2Y_APP_V2(__nv_bfloat16, B2Y_ACT_LEAKY);
            else if (act == B2Y_ACT_MISH) B2Y_APP_V2(__nv_bfloat16, B2Y_ACT_MISH);
            else B2Y_APP_V2(__nv_bfloat16, B2Y_ACT_LINEAR);
        } else {
            if (act == B2Y_ACT_LEAKY) B2Y_APP_V2(__half, B2Y_ACT_LEAKY);
            else if (act == B2Y_ACT_MISH) B2Y_APP_V2(__half, B2Y_ACT_MISH);
            else B2Y_APP_V2(__half, B2Y_ACT_LINEAR);
        }
#undef B2Y_APP_V2
        B2Y_CUDA_CHECK(cudaGetLastError());
        return B2Y_OK;
    }
    if (grad_dtype == B2Y_DT_BF16)
        bn_act_bwd_apply_kernel<__nv_bfloat16><<<grid_for(pixels * (c / 8), 256), 256, 0,
                                                 static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, scale,
            shift, gamma, save_mean, save_invstd, dgamma, dbeta, reinterpret_cast<__half*>(dx), dx_pitch, pixels, c,
            act, slope, du_absmax, scale_out);
    else
        bn_act_bwd_apply_kernel<__half><<<grid_for(pixels * (c / 8), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __half*>(dy), dy_pitch, scale, shift,
            gamma, save_mean, save_invstd, dgamma, dbeta, reinterpret_cast<__half*>(dx), dx_pitch, pixels, c, act,
            slope, du_absmax, scale_out);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------
// SGD with Nesterov momentum (torch.optim.SGD semantics, train.py:135-144):
//   g = grad*grad_scale + wd*p ; buf = first ? g : mu*buf + g ; p -= lr * (g + mu*buf)
// ------------------------------------------------------------------------------------------------
__global__ void sgd_nesterov_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                    long long n, float lr, float mu, float wd, float gs, int first,
                                    float* __restrict__ ema, float ema_decay) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float w = p[i];
        const float grad = fmaf(wd, w, g[i] * gs);
        const float b = first ? grad : fmaf(mu, buf[i], grad);
        buf[i] = b;
        const float wn = w - lr * fmaf(mu, b, grad);
        p[i] = wn;
        // ModelEMA.update (utils/torch_utils.py:171-183) fused into the optimiser pass: ema = d*ema + (1-d)*w_new
        if (ema != nullptr) ema[i] = fmaf(ema_decay, ema[i], (1.f - ema_decay) * wn);
    }
}

extern "C" int b2y_sgd_nesterov(float* param, const float* grad, float* momentum_buf, long long n, float lr,
                                float momentum, float weight_decay, float grad_scale, int first_step, void* stream) {
    if (!param || !grad || !momentum_buf || n < 0) return B2Y_ERR_INVALID;
    sgd_nesterov_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first_step, nullptr, 0.f);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_sgd_nesterov_ema(float* param, const float* grad, float* momentum_buf, float* ema, long long n,
                                    float lr, float momentum, float weight_decay, float grad_scale, int first_step,
                                    float ema_decay, void* stream) {
    if (!param || !grad || !momentum_buf || !ema || n < 0) return B2Y_ERR_INVALID;
    sgd_nesterov_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        param, grad, momentum_buf, n, lr, momentum, weight_decay, grad_scale, first_step, ema, ema_decay);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// Network-slimming sparsity term of the BatchNorm scales (prune_utils.py:133-138, BNOptimizer.updateBN, called between
// backward and optimizer.step at train.py:444-445): grad += coeff * sign(w) on a table of [offset, length] ranges of
// the flat buffers.  One CTA per range (a BatchNorm layer: 32 .. 1024 scales).
__global__ void l1_subgrad_ranges_kernel(float* __restrict__ g, const float* __restrict__ p,
                                         const long long* __restrict__ ranges, float coeff) {
    const long long off = ranges[2 * blockIdx.x], len = ranges[2 * blockIdx.x + 1];
    for (long long i = threadIdx.x; i < len; i += blockDim.x) {
        const float w = p[off + i];
        const float sgn = (float)((w > 0.f) - (w < 0.f));           // torch.sign: 0 at 0
        g[off + i] = fmaf(coeff, sgn, g[off + i]);
    }
}

extern "C" int b2y_l1_subgrad_ranges(float* grad, const float* param, const long long* ranges_dev, int n_ranges,
                                     float coeff, void* stream) {
    if (!grad || !param || n_ranges < 0 || (n_ranges > 0 && !ranges_dev)) return B2Y_ERR_INVALID;
    if (n_ranges == 0) return B2Y_OK;
    l1_subgrad_ranges_kernel<<<n_ranges, 128, 0, static_cast<cudaStream_t>(stream)>>>(grad, param, ranges_dev, coeff);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------
// backward of the data-movement layers (gradients are NHWC fp16, accumulated in place)
// ------------------------------------------------------------------------------------------------
// YOLO head: dp fp32 [B][na][ny][nx][no] -> d(raw) fp16 [B][ny][nx][pitch] (channel a*no+o), times `scale`
template <typename GT>
__global__ void yolo_grad_to_raw_kernel(const float* __restrict__ dp, GT* __restrict__ draw, long long pitch,
                                        int B, int na, int no, int ny, int nx, float scale,
                                        const float* __restrict__ scale_ptr) {
    if (scale_ptr != nullptr) scale *= __ldg(scale_ptr);
    const long long total = (long long)B * ny * nx * pitch;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(idx % pitch);
        long long t = idx / pitch;
        const int x = (int)(t % nx);
        t /= nx;
        const int y = (int)(t % ny);
        const int b = (int)(t / ny);
        float v = 0.f;
        if (c < na * no) {
            const int a = c / no, o = c - a * no;
            v = dp[((((long long)b * na + a) * ny + y) * nx + x) * no + o] * scale;
        }
        draw[idx] = Half8<GT>::from_f(v);
    }
}
// The same permutation, one warp per PIXEL (pitch <= 256): the (anchor, output) split of every element is computed once
// per thread, a lane issues its 8 loads (three contiguous runs of `no` floats in dp) back to back and the warp writes one
// contiguous 512-byte row -- no per-element index divisions.
template <typename GT>
__global__ void __launch_bounds__(256) yolo_grad_to_raw_pixel_kernel(const float* __restrict__ dp, GT* __restrict__ draw,
                                                                     long long pitch, int B, int na, int no, int ny,
                                                                     int nx, float scale,
                                                                     const float* __restrict__ scale_ptr) {
    if (scale_ptr != nullptr) scale *= __ldg(scale_ptr);
    const int lane = threadIdx.x & 31;
    const unsigned plane = (unsigned)(ny * nx);
    const unsigned pixels = (unsigned)B * plane;
    const unsigned warps = (gridDim.x * blockDim.x) >> 5;
    const int nel = na * no;
    long long eoff[8];       // offset of element e inside dp relative to the pixel's (b, a = 0, y, x, o = 0) entry, or -1
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int e = lane + 32 * i;
        const int a = e / no;
        eoff[i] = e < nel ? (long long)a * plane * no + (e - a * no) : -1;
    }
    for (unsigned pix = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < pixels; pix += warps) {
        const unsigned b = pix / plane, yx = pix - b * plane;
        const float* src = dp + ((long long)b * na * plane + yx) * no;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = eoff[i] >= 0 ? __ldg(src + eoff[i]) * scale : 0.f;
        GT* dst = draw + (long long)pix * pitch;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (lane + 32 * i < pitch) dst[lane + 32 * i] = Half8<GT>::from_f(v[i]);
    }
}

extern "C" int b2y_yolo_grad_to_raw(const float* dp, void* draw, long long raw_pitch, int batch, int na, int no,
                                    int ny, int nx, float scale, const float* scale_ptr, int grad_dtype,
                                    void* stream) {
    if (!dp || !draw || raw_pitch < (long long)na * no) return B2Y_ERR_INVALID;
    const long long total = (long long)batch * ny * nx * raw_pitch;
    if (raw_pitch <= 256 && (long long)batch * ny * nx < 0x7fffffffLL) {
        long long pb = ((long long)batch * ny * nx + 7) / 8;
        if (pb > 148 * 8) pb = 148 * 8;
        if (grad_dtype == B2Y_DT_BF16)
            yolo_grad_to_raw_pixel_kernel<__nv_bfloat16><<<(int)pb, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                dp, reinterpret_cast<__nv_bfloat16*>(draw), raw_pitch, batch, na, no, ny, nx, scale, scale_ptr);
        else
            yolo_grad_to_raw_pixel_kernel<__half><<<(int)pb, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                dp, reinterpret_cast<__half*>(draw), raw_pitch, batch, na, no, ny, nx, scale, scale_ptr);
        B2Y_CUDA_CHECK(cudaGetLastError());
        return B2Y_OK;
    }
    if (grad_dtype == B2Y_DT_BF16)
        yolo_grad_to_raw_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            dp, reinterpret_cast<__nv_bfloat16*>(draw), raw_pitch, batch, na, no, ny, nx, scale, scale_ptr);
    else
        yolo_grad_to_raw_kernel<__half><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            dp, reinterpret_cast<__half*>(draw), raw_pitch, batch, na, no, ny, nx, scale, scale_ptr);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// nearest upsample backward: dx[n,y,x,:] += sum_{dy,dx<s} dup[n, y*s+dy, x*s+dx, :]
template <typename GT>
__global__ void upsample_bwd_kernel(const GT* __restrict__ dy, long long dyp, GT* __restrict__ dx,
                                    long long dxp, int B, int H, int W, int C, int s) {
    const int CV = C / 8;
    const long long total = (long long)B * H * W * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long pix = idx / CV;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const int n = (int)(pix / ((long long)W * H));
        float acc[8], t8[8];
        Half8<GT>::load(dx + pix * dxp + cv * 8, acc);
        for (int a = 0; a < s; ++a)
            for (int b = 0; b < s; ++b) {
                const long long op = ((long long)n * H * s + (y * s + a)) * (W * s) + (x * s + b);
                Half8<GT>::load(dy + op * dyp + cv * 8, t8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] += t8[j];
            }
        Half8<GT>::store(dx + pix * dxp + cv * 8, acc);
    }
}
extern "C" int b2y_upsample_nearest_bwd(const void* dy, long long dy_pitch, void* dx, long long dx_pitch, int batch,
                                        int in_h, int in_w, int c, int scale, int grad_dtype, void* stream) {
    if (!dy || !dx || c % 8 != 0 || dy_pitch % 8 != 0 || dx_pitch % 8 != 0 || scale < 1) return B2Y_ERR_INVALID;
    const long long total = (long long)batch * in_h * in_w * (c / 8);
    if (grad_dtype == B2Y_DT_BF16)
        upsample_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch, batch,
            in_h, in_w, c, scale);
    else
        upsample_bwd_kernel<__half><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(dy), dy_pitch, reinterpret_cast<__half*>(dx), dx_pitch, batch, in_h, in_w,
            c, scale);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// maxpool backward: every output pixel re-finds its arg-max (first maximum in row-major window order, as
// torch's max_pool2d does) and adds its gradient there (windows overlap for stride 1 -> fp16x2 atomics).
template <typename GT, typename GT2>
__global__ void maxpool_bwd_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy,
                                   long long dyp, GT* __restrict__ dx, long long dxp, int B, int H, int W, int C,
                                   int k, int stride, int pad, int Ho, int Wo, int zero_pad) {
    const int CV = C / 2;
    const long long total = (long long)B * Ho * Wo * CV;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        long long pix = idx / CV;
        const int xo = (int)(pix % Wo);
        const int yo = (int)((pix / Wo) % Ho);
        const int n = (int)(pix / ((long long)Wo * Ho));
        float best0 = -INFINITY, best1 = -INFINITY;
        long long arg0 = -1, arg1 = -1;
        for (int kh = 0; kh < k; ++kh) {
            const int yi = yo * stride - pad + kh;
            for (int kw = 0; kw < k; ++kw) {
                const int xi = xo * stride - pad + kw;
                if (yi < 0 || yi >= H || xi < 0 || xi >= W) {
                    if (zero_pad) {  // the zero padding competes (and swallows the gradient when it wins)
                        if (0.f > best0) { best0 = 0.f; arg0 = -1; }
                        if (0.f > best1) { best1 = 0.f; arg1 = -1; }
                    }
                    continue;
                }
                const long long ip = ((long long)n * H + yi) * W + xi;
                const __half2 v = *reinterpret_cast<const __half2*>(x + ip * xp + cv * 2);
                const float a = __low2float(v), b = __high2float(v);
                if (a > best0) { best0 = a; arg0 = ip; }
                if (b > best1) { best1 = b; arg1 = ip; }
            }
        }
        const GT g0 = dy[pix * dyp + cv * 2], g1 = dy[pix * dyp + cv * 2 + 1];
        const GT zero = Half8<GT>::from_f(0.f);
        GT2 v;
        if (arg0 >= 0 && arg0 == arg1) {
            v.x = g0;
            v.y = g1;
            atomicAdd(reinterpret_cast<GT2*>(dx + arg0 * dxp + cv * 2), v);
        } else {
            if (arg0 >= 0) {
                v.x = g0;
                v.y = zero;
                atomicAdd(reinterpret_cast<GT2*>(dx + arg0 * dxp + cv * 2), v);
            }
            if (arg1 >= 0) {
                v.x = zero;
                v.y = g1;
                atomicAdd(reinterpret_cast<GT2*>(dx + arg1 * dxp + cv * 2), v);
            }
        }
    }
}
// Stride-1 "same" pooling over a small map (SPP: 5/9/13 over 20x20): one CTA owns the plane of one image x 8 channels.
// Pass 1 finds, per input row and window position, the row-window maximum and its FIRST column; pass 2 walks the window
// rows with a strict '>' -- together the first maximum in row-major window order, as the k*k scan of the general kernel
// (and torch's max_pool2d) finds it, with k + k loads.  Gradients are summed in shared memory (fp32) and added to dx by
// the owner, so there are no global atomics and the result does not depend on scheduling.
constexpr int POOL_PLANE_MAX = 576;
template <typename GT>
__global__ void __launch_bounds__(256) maxpool_plane_bwd_kernel(const __half* __restrict__ x, long long xp,
                                                                const GT* __restrict__ dy, long long dyp,
                                                                GT* __restrict__ dx, long long dxp, int H, int W,
                                                                int C, int k) {
    __shared__ __align__(16) __half sx[POOL_PLANE_MAX * 8];
    __shared__ __align__(16) __half sr[POOL_PLANE_MAX * 8];
    __shared__ unsigned char sa[POOL_PLANE_MAX * 8];
    __shared__ float acc[POOL_PLANE_MAX * 8];
    const int CV = C / 8;
    const int n = blockIdx.x / CV, cv = blockIdx.x - n * CV;
    const int HW = H * W, pad = (k - 1) / 2;
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
        reinterpret_cast<uint4*>(sx)[i] = __ldg(reinterpret_cast<const uint4*>(x + ((long long)n * HW + i) * xp) + cv);
    for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) acc[i] = 0.f;
    __syncthreads();
    for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {
        const int c = i & 7, pix = i >> 3;
        const int yi = pix / W, xo = pix - yi * W;
        const int x0 = max(xo - pad, 0), x1 = min(xo - pad + k, W);
        float m = __half2float(sx[(yi * W + x0) * 8 + c]);
        int a = x0;
        for (int xi = x0 + 1; xi < x1; ++xi) {
            const float v = __half2float(sx[(yi * W + xi) * 8 + c]);
            if (v > m) { m = v; a = xi; }
        }
        sr[i] = __float2half(m);        // exact: m is one of the fp16 inputs
        sa[i] = (unsigned char)a;
    }
    __syncthreads();
    // sx is dead now: stage the plane of dy in it with 16-byte loads (the scalar 2-byte global loads of the first version
    // put one DRAM round trip into every iteration of the pass below)
    GT* sdy = reinterpret_cast<GT*>(sx);
    for (int i = threadIdx.x; i < HW; i += blockDim.x)
        reinterpret_cast<uint4*>(sdy)[i] = __ldg(reinterpret_cast<const uint4*>(dy + ((long long)n * HW + i) * dyp) + cv);
    __syncthreads();
    for (int i = threadIdx.x; i < HW * 8; i += blockDim.x) {
        const int c = i & 7, pix = i >> 3;
        const int yo = pix / W, xo = pix - yo * W;
        const int y0 = max(yo - pad, 0), y1 = min(yo - pad + k, H);
        float m = __half2float(sr[(y0 * W + xo) * 8 + c]);
        int ay = y0;
        for (int yi = y0 + 1; yi < y1; ++yi) {
            const float v = __half2float(sr[(yi * W + xo) * 8 + c]);
            if (v > m) { m = v; ay = yi; }
        }
        // NaN rows never win a strict '>' (same as the general kernel, whose -inf start is replaced by the first value)
        const int ax = sa[(ay * W + xo) * 8 + c];
        const float g = Half8<GT>::to_f(sdy[i]);
        atomicAdd(&acc[(ay * W + ax) * 8 + c], g);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < HW; i += blockDim.x) {
        GT* d = dx + ((long long)n * HW + i) * dxp + cv * 8;
        float f[8];
        Half8<GT>::load(d, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] += acc[i * 8 + j];
        Half8<GT>::store(d, f);
    }
}

extern "C" int b2y_maxpool_bwd(const void* x, long long x_pitch, const void* dy, long long dy_pitch, void* dx,
                               long long dx_pitch, int batch, int in_h, int in_w, int c, int ksize, int stride,
                               int pad_mode, int grad_dtype, void* stream) {
    if (!x || !dy || !dx || c % 2 != 0 || x_pitch % 2 != 0 || dy_pitch % 2 != 0 || dx_pitch % 2 != 0)
        return B2Y_ERR_INVALID;
    int pad, Ho, Wo;
    if (pad_mode == 1) {
        pad = 0;
        Ho = (in_h + 1 - ksize) / stride + 1;
        Wo = (in_w + 1 - ksize) / stride + 1;
    } else {
        pad = (ksize - 1) / 2;
        Ho = (in_h + 2 * pad - ksize) / stride + 1;
        Wo = (in_w + 2 * pad - ksize) / stride + 1;
    }
    const long long total = (long long)batch * Ho * Wo * (c / 2);
    if (pad_mode != 1 && stride == 1 && (ksize & 1) && c % 8 == 0 && x_pitch % 8 == 0 && in_w <= 255 &&
        in_h * in_w <= POOL_PLANE_MAX && dy_pitch % 8 == 0 && dx_pitch % 8 == 0 &&
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0) {
        const int grid = batch * (c / 8);
        if (grad_dtype == B2Y_DT_BF16)
            maxpool_plane_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch,
                reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch, in_h, in_w, c, ksize);
        else
            maxpool_plane_bwd_kernel<__half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
                reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __half*>(dy), dy_pitch,
                reinterpret_cast<__half*>(dx), dx_pitch, in_h, in_w, c, ksize);
        B2Y_CUDA_CHECK(cudaGetLastError());
        return B2Y_OK;
    }
    if (grad_dtype == B2Y_DT_BF16)
        maxpool_bwd_kernel<__nv_bfloat16, __nv_bfloat162><<<grid_for(total, 256), 256, 0,
                                                            static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch,
            reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch, batch, in_h, in_w, c, ksize, stride, pad, Ho, Wo,
            pad_mode == 1);
    else
        maxpool_bwd_kernel<__half, __half2><<<grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __half*>(dy), dy_pitch,
            reinterpret_cast<__half*>(dx), dx_pitch, batch, in_h, in_w, c, ksize, stride, pad, Ho, Wo, pad_mode == 1);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

// ------------------------------------------------------------------------------------------------
// stem weight gradient (Cin <= 4, NCHW fp32 image): dW[co][ci][kh][kw] += scale * sum_pix dz[pix][co] * x[...]
// CTA = 32 output channels x 8 tap groups; each thread keeps <= TPG taps of one output channel in registers.
// ------------------------------------------------------------------------------------------------
template <int TPG, typename GT>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const float* __restrict__ x, const GT* __restrict__ dz, long long dzp, float* __restrict__ dw,
                  int B, int Cin, int H, int W, int Cout, int k, int stride, int pad, int Ho, int Wo, float scale) {
    const int taps = Cin * k * k;
    const int co_l = threadIdx.x & 31;
    const int grp = threadIdx.x >> 5;  // 0..7
    const long long M = (long long)B * Ho * Wo;
    for (int co0 = 0; co0 < Cout; co0 += 32) {
        const int co = co0 + co_l;
        float acc[TPG];
        int t_ci[TPG], t_kh[TPG], t_kw[TPG];
#pragma unroll
        for (int j = 0; j < TPG; ++j) {
            acc[j] = 0.f;
            const int t = grp * TPG + j;
            t_ci[j] = t / (k * k);
            t_kh[j] = (t / k) % k;
            t_kw[j] = t % k;
        }
        const long long per = (M + gridDim.x - 1) / gridDim.x;
        const long long m0 = (long long)blockIdx.x * per;
        const long long m1 = m0 + per < M ? m0 + per : M;
        for (long long m = m0; m < m1; ++m) {
            const int xo = (int)(m % Wo);
            const int yo = (int)((m / Wo) % Ho);
            const int n = (int)(m / ((long long)Wo * Ho));
            const float g = co < Cout ? Half8<GT>::to_f(dz[m * dzp + co]) : 0.f;
#pragma unroll
            for (int j = 0; j < TPG; ++j) {
                if (grp * TPG + j < taps) {
                    const int yi = yo * stride - pad + t_kh[j], xi = xo * stride - pad + t_kw[j];
                    float v = 0.f;
                    if (yi >= 0 && yi < H && xi >= 0 && xi < W)
                        v = __ldg(x + (((long long)n * Cin + t_ci[j]) * H + yi) * W + xi);
                    acc[j] = fmaf(g, v, acc[j]);
                }
            }
        }
        if (co < Cout) {
#pragma unroll
            for (int j = 0; j < TPG; ++j)
                if (grp * TPG + j < taps) atomicAdd(dw + (long long)co * taps + grp * TPG + j, acc[j] * scale);
        }
    }
}
extern "C" int b2y_stem_conv_bwd_weight(const b2y_conv_desc* d, const float* x_nchw, const void* dz, float* dw_oihw,
                                        float scale, int grad_dtype, void* stream) {
    if (!d || !x_nchw || !dz || !dw_oihw) return B2Y_ERR_INVALID;
    const int taps = d->in_c * d->ksize * d->ksize;
    if (d->in_c > 4 || taps > 8 * 5) return B2Y_ERR_UNSUPPORTED;
    const long long M = (long long)d->batch * d->out_h * d->out_w;
    int grid = (int)((M + 1023) / 1024);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    if (grad_dtype == B2Y_DT_BF16)
        stem_wgrad_kernel<5, __nv_bfloat16><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            x_nchw, reinterpret_cast<const __nv_bfloat16*>(dz), d->out_pitch, dw_oihw, d->batch, d->in_c, d->in_h,
            d->in_w, d->out_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w, scale);
    else
        stem_wgrad_kernel<5, __half><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
            x_nchw, reinterpret_cast<const __half*>(dz), d->out_pitch, dw_oihw, d->batch, d->in_c, d->in_h, d->in_w,
            d->out_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w, scale);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}
