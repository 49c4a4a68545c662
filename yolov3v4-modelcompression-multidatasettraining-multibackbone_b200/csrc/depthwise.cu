// Depthwise convolution (groups == channels) and squeeze-excite, forward and backward, NHWC fp16 (C % 8 == 0).
// Reference: the `depthwise` cfg block models.py:115-197 (nn.Conv2d(groups=Cin) + BatchNorm2d + activation, k = 3 / 5,
// stride 1 / 2) and SE utils/layers.py:176-192 (avgpool -> Linear(C, C/4) -> ReLU -> Linear(C/4, C) -> HardSigmoid ->
// x * y).  0.7 % of yolov3-mobilenet's MACs but a third of its layers: these kernels are HBM / launch bound, so they run
// on the CUDA cores with 16-byte vector accesses; a thread owns one 8-channel vector (its per-channel state lives in
// registers) and walks pixels, like the BatchNorm passes in bn_train.cu.
#include "b200yolo.h"
#include "common.cuh"

using namespace b2y;

namespace {

struct Geo {
    int CV, PPB;
};
inline Geo geo_for(int c) {
    Geo g;
    g.CV = c / 8;
    g.PPB = 256 / g.CV;
    if (g.PPB < 1) g.PPB = 1;
    return g;
}
inline int sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& s = sms[dev & 63];
    if (s <= 0 && cudaDeviceGetAttribute(&s, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) s = 148;
    if (s <= 0) s = 148;
    return s;
}
inline int grid_for_rows(long long rows, int ppb, int per_sm) {
    long long g = (rows + ppb - 1) / ppb;
    const long long cap = (long long)sm_count() * per_sm;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// block-level reduction of per-thread 8-vectors over the PPB pixel rows of a CTA, then one atomic per channel
__device__ __forceinline__ void block_reduce_atomic(float (&a)[8], float* red /*[256][9]*/, float* dst, int CV, int PPB,
                                                    float mul) {
    const int tid = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 9 + j] = a[j];
    __syncthreads();
    for (int i = tid; i < CV * 8; i += blockDim.x) {
        const int v = i / 8, comp = i % 8;
        float s = 0.f;
        for (int k = 0; k < PPB; ++k) s += red[(v + k * CV) * 9 + comp];
        atomicAdd(dst + v * 8 + comp, s * mul);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------
// forward: y = act(dwconv(x, w) * scale + bias); optional channel sums of the raw output (training BatchNorm)
// weights arrive as fp32 [C][k][k] (the nn.Conv2d(groups=C) parameter) and are staged as fp16 [tap][C] in smem.
__global__ void __launch_bounds__(256)
dwconv_fwd_kernel(const __half* __restrict__ x, long long xp, const float* __restrict__ w,
                  const float* __restrict__ scale, const float* __restrict__ bias, __half* __restrict__ y,
                  long long yp, float* __restrict__ s1, float* __restrict__ s2, int B, int H, int W, int C, int k,
                  int stride, int pad, int Ho, int Wo, int act, float slope, int CV, int PPB) {
    extern __shared__ uint8_t smem_dw[];
    __half* ws = reinterpret_cast<__half*>(smem_dw);                                  // [k*k][C]
    float* red = reinterpret_cast<float*>(smem_dw + (((size_t)k * k * C * 2 + 15) & ~(size_t)15));   // [256][9]
    const int k2 = k * k;
    for (int i = threadIdx.x; i < C * k2; i += blockDim.x) {
        const int c = i / k2, t = i - c * k2;
        ws[t * C + c] = __float2half_rn(__ldg(w + i));      // fp16 weights: the engine's precision policy
    }
    __syncthreads();
    const int tid = threadIdx.x;
    const bool active = tid < CV * PPB;
    const int cv = active ? tid % CV : 0, prow = tid / CV;
    float sc[8], bi[8], a1[8], a2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = scale != nullptr ? __ldg(scale + cv * 8 + j) : 1.f;
        bi[j] = bias != nullptr ? __ldg(bias + cv * 8 + j) : 0.f;
        a1[j] = a2[j] = 0.f;
    }
    const long long pixels = (long long)B * Ho * Wo;
    const long long stride_p = (long long)gridDim.x * PPB;
    if (active) {
        for (long long pix = (long long)blockIdx.x * PPB + prow; pix < pixels; pix += stride_p) {
            const int xo = (int)(pix % Wo);
            const int yo = (int)((pix / Wo) % Ho);
            const int n = (int)(pix / ((long long)Wo * Ho));
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            for (int r = 0; r < k; ++r) {
                const int yi = yo * stride - pad + r;
                if (yi < 0 || yi >= H) continue;
                for (int s = 0; s < k; ++s) {
                    const int xi = xo * stride - pad + s;
                    if (xi < 0 || xi >= W) continue;
                    const uint4 xv = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + yi) * W + xi) * xp) + cv);
                    const uint4 wv = *reinterpret_cast<const uint4*>(ws + (r * k + s) * C + cv * 8);
                    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
                    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 xf = __half22float2(xh[j]), wf = __half22float2(wh[j]);
                        acc[2 * j] = fmaf(xf.x, wf.x, acc[2 * j]);
                        acc[2 * j + 1] = fmaf(xf.y, wf.y, acc[2 * j + 1]);
                    }
                }
            }
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // the statistics are taken on the fp16-rounded value that is stored (what the BatchNorm pass will read)
                const float raw = fmaf(acc[j], sc[j], bi[j]);
                o[j] = apply_act(raw, act, slope);
            }
            Half8<__half>::store(y + pix * yp + cv * 8, o);
            if (s1 != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float q = __half2float(__float2half_rn(o[j]));
                    a1[j] += q;
                    a2[j] = fmaf(q, q, a2[j]);
                }
            }
        }
    }
    if (s1 != nullptr) {
        block_reduce_atomic(a1, red, s1, CV, PPB, 1.f);
        block_reduce_atomic(a2, red, s2, CV, PPB, 1.f);
    }
}

// data gradient: dx[n,yi,xi,c] (+)= inv_s * sum_{r,s} w[c][r][s] * dz[n,(yi+pad-r)/st,(xi+pad-s)/st,c]
template <typename GT>
__global__ void __launch_bounds__(256)
dwconv_bwd_data_kernel(const __half* __restrict__ dz, long long dzp, const float* __restrict__ w, GT* __restrict__ dx,
                       long long dxp, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                       int accumulate, const float* __restrict__ inv_scale_ptr, int CV, int PPB) {
    extern __shared__ uint8_t smem_dw[];
    __half* ws = reinterpret_cast<__half*>(smem_dw);
    const int k2 = k * k;
    for (int i = threadIdx.x; i < C * k2; i += blockDim.x) {
        const int c = i / k2, t = i - c * k2;
        ws[t * C + c] = __float2half_rn(__ldg(w + i));
    }
    __syncthreads();
    const int tid = threadIdx.x;
    if (tid >= CV * PPB) return;
    const float inv_s = inv_scale_ptr != nullptr ? __ldg(inv_scale_ptr) : 1.f;
    const int cv = tid % CV, prow = tid / CV;
    const long long pixels = (long long)B * H * W;
    const long long stride_p = (long long)gridDim.x * PPB;
    for (long long pix = (long long)blockIdx.x * PPB + prow; pix < pixels; pix += stride_p) {
        const int xi = (int)(pix % W);
        const int yi = (int)((pix / W) % H);
        const int n = (int)(pix / ((long long)W * H));
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int r = 0; r < k; ++r) {
            const int ty = yi + pad - r;
            if (ty < 0 || ty % stride != 0) continue;
            const int yo = ty / stride;
            if (yo >= Ho) continue;
            for (int s = 0; s < k; ++s) {
                const int tx = xi + pad - s;
                if (tx < 0 || tx % stride != 0) continue;
                const int xo = tx / stride;
                if (xo >= Wo) continue;
                const uint4 gv = __ldg(reinterpret_cast<const uint4*>(dz + (((long long)n * Ho + yo) * Wo + xo) * dzp) + cv);
                const uint4 wv = *reinterpret_cast<const uint4*>(ws + (r * k + s) * C + cv * 8);
                const __half2* gh = reinterpret_cast<const __half2*>(&gv);
                const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 gf = __half22float2(gh[j]), wf = __half22float2(wh[j]);
                    acc[2 * j] = fmaf(gf.x, wf.x, acc[2 * j]);
                    acc[2 * j + 1] = fmaf(gf.y, wf.y, acc[2 * j + 1]);
                }
            }
        }
        GT* dp = dx + pix * dxp + cv * 8;
        float o[8];
        if (accumulate) {
            Half8<GT>::load(dp, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(acc[j], inv_s, o[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = acc[j] * inv_s;
        }
        Half8<GT>::store(dp, o);
    }
}

// weight gradient: dw[c][r][s] += alpha * inv_s * sum_pixels dz[pix][c] * x[pix @ (r,s)][c];  grid.y = tap
__global__ void __launch_bounds__(256)
dwconv_bwd_weight_kernel(const __half* __restrict__ x, long long xp, const __half* __restrict__ dz, long long dzp,
                         float* __restrict__ dw, int B, int H, int W, int C, int k, int stride, int pad, int Ho, int Wo,
                         float alpha, const float* __restrict__ inv_scale_ptr, int CV, int PPB) {
    __shared__ float red[256 * 9];
    const int tap = blockIdx.y;
    const int r = tap / k, s = tap % k;
    const int tid = threadIdx.x;
    const bool active = tid < CV * PPB;
    const int cv = active ? tid % CV : 0, prow = tid / CV;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const long long pixels = (long long)B * Ho * Wo;
    const long long stride_p = (long long)gridDim.x * PPB;
    if (active) {
        for (long long pix = (long long)blockIdx.x * PPB + prow; pix < pixels; pix += stride_p) {
            const int xo = (int)(pix % Wo);
            const int yo = (int)((pix / Wo) % Ho);
            const int n = (int)(pix / ((long long)Wo * Ho));
            const int yi = yo * stride - pad + r, xi = xo * stride - pad + s;
            if (yi < 0 || yi >= H || xi < 0 || xi >= W) continue;
            float gf[8], xf[8];
            Half8<__half>::load(dz + pix * dzp + cv * 8, gf);
            Half8<__half>::load(x + (((long long)n * H + yi) * W + xi) * xp + cv * 8, xf);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(gf[j], xf[j], acc[j]);
        }
    }
    const float mul = alpha * (inv_scale_ptr != nullptr ? __ldg(inv_scale_ptr) : 1.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) red[tid * 9 + j] = acc[j];
    __syncthreads();
    const int k2 = k * k;
    for (int i = tid; i < CV * 8; i += blockDim.x) {
        const int v = i / 8, comp = i % 8;
        float sum = 0.f;
        for (int q = 0; q < PPB; ++q) sum += red[(v + q * CV) * 9 + comp];
        atomicAdd(dw + (long long)(v * 8 + comp) * k2 + tap, sum * mul);
    }
}

// ------------------------------------------------------------------------------------------------------------
// squeeze-excite
// pooled[n][c] += sum over the image's pixels of x (caller zeroes; the mean is taken in the fc kernel)
__global__ void __launch_bounds__(256)
se_pool_kernel(const __half* __restrict__ x, long long xp, float* __restrict__ pooled, int HW, int C, int CV, int PPB) {
    __shared__ float red[256 * 9];
    const int n = blockIdx.y;
    const int tid = threadIdx.x;
    const bool active = tid < CV * PPB;
    const int cv = active ? tid % CV : 0, prow = tid / CV;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (active) {
        for (int p = blockIdx.x * PPB + prow; p < HW; p += gridDim.x * PPB) {
            float xf[8];
            Half8<__half>::load(x + ((long long)n * HW + p) * xp + cv * 8, xf);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += xf[j];
        }
    }
    block_reduce_atomic(acc, red, pooled + (long long)n * C, CV, PPB, 1.f);
}

// one block per image: mean -> h = relu(W1 mean) -> s = hsigmoid(W2 h).  W1 [Cr][C], W2 [C][Cr] fp32 (nn.Linear)
__global__ void se_fc_kernel(const float* __restrict__ pooled_sum, const float* __restrict__ w1,
                             const float* __restrict__ w2, float* __restrict__ mean_out, float* __restrict__ h_out,
                             float* __restrict__ v_out, float* __restrict__ s_out, int C, int Cr, float inv_hw) {
    extern __shared__ float sh[];       // [C] mean, [Cr] h
    float* m = sh;
    float* h = sh + C;
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        m[c] = pooled_sum[(long long)n * C + c] * inv_hw;
        if (mean_out != nullptr) mean_out[(long long)n * C + c] = m[c];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < Cr; j += nw) {
        float a = 0.f;
        for (int c = lane; c < C; c += 32) a = fmaf(__ldg(w1 + (long long)j * C + c), m[c], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            h[j] = fmaxf(a, 0.f);
            if (h_out != nullptr) h_out[(long long)n * Cr + j] = h[j];
        }
    }
    __syncthreads();
    for (int c = warp; c < C; c += nw) {
        float a = 0.f;
        for (int j = lane; j < Cr; j += 32) a = fmaf(__ldg(w2 + (long long)c * Cr + j), h[j], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            if (v_out != nullptr) v_out[(long long)n * C + c] = a;
            s_out[(long long)n * C + c] = fminf(fmaxf(a + 3.f, 0.f), 6.f) / 6.f;     // HardSigmoid utils/layers.py:167-173
        }
    }
}

// y = x * s[n][c]
__global__ void __launch_bounds__(256)
se_scale_kernel(const __half* __restrict__ x, long long xp, const float* __restrict__ s, __half* __restrict__ y,
                long long yp, int HW, int C, long long total_vec) {
    const int CV = C / 8;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        const long long pix = idx / CV;
        const int n = (int)(pix / HW);
        float xf[8];
        Half8<__half>::load(x + pix * xp + cv * 8, xf);
        const float4 sa = __ldg(reinterpret_cast<const float4*>(s + (long long)n * C + cv * 8));
        const float4 sb = __ldg(reinterpret_cast<const float4*>(s + (long long)n * C + cv * 8) + 1);
        const float sv[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) xf[j] *= sv[j];
        Half8<__half>::store(y + pix * yp + cv * 8, xf);
    }
}

// backward, pass 1: ds_raw[n][c] += sum_pixels dy * x
template <typename GT>
__global__ void __launch_bounds__(256)
se_bwd_reduce_kernel(const __half* __restrict__ x, long long xp, const GT* __restrict__ dy, long long dyp,
                     float* __restrict__ ds, int HW, int C, int CV, int PPB) {
    __shared__ float red[256 * 9];
    const int n = blockIdx.y;
    const int tid = threadIdx.x;
    const bool active = tid < CV * PPB;
    const int cv = active ? tid % CV : 0, prow = tid / CV;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (active) {
        for (int p = blockIdx.x * PPB + prow; p < HW; p += gridDim.x * PPB) {
            float xf[8], gf[8];
            Half8<__half>::load(x + ((long long)n * HW + p) * xp + cv * 8, xf);
            Half8<GT>::load(dy + ((long long)n * HW + p) * dyp + cv * 8, gf);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(xf[j], gf[j], acc[j]);
        }
    }
    block_reduce_atomic(acc, red, ds + (long long)n * C, CV, PPB, 1.f);
}

// backward of the two Linear layers, one block per image: dv = ds * hsig'(v); dh = (W2^T dv) * [h > 0];
// dmean = W1^T dh.  Outputs dv [N][C], dh [N][Cr], dmean [N][C] (the weight gradients are outer-product sums over
// the images, se_fc_wgrad_kernel).
__global__ void se_fc_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ v, const float* __restrict__ h,
                                 const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ dv_out,
                                 float* __restrict__ dh_out, float* __restrict__ dmean_out, int C, int Cr) {
    extern __shared__ float sh[];       // [C] dv, [Cr] dh
    float* dv = sh;
    float* dh = sh + C;
    const int n = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float vv = v[(long long)n * C + c];
        const float g = (vv > -3.f && vv < 3.f) ? (1.f / 6.f) : 0.f;
        dv[c] = ds[(long long)n * C + c] * g;
        dv_out[(long long)n * C + c] = dv[c];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
        float a = 0.f;
        for (int c = 0; c < C; ++c) a = fmaf(__ldg(w2 + (long long)c * Cr + j), dv[c], a);
        dh[j] = h[(long long)n * Cr + j] > 0.f ? a : 0.f;
        dh_out[(long long)n * Cr + j] = dh[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f;
        for (int j = 0; j < Cr; ++j) a = fmaf(__ldg(w1 + (long long)j * C + c), dh[j], a);
        dmean_out[(long long)n * C + c] = a;
    }
}

// dW[i][j] = alpha * sum_n a[n][i] * b[n][j]   (dW2 = dv (x) h, dW1 = dh (x) mean)
__global__ void se_fc_wgrad_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dw,
                                   int N, int I, int J, float alpha) {
    const long long total = (long long)I * J;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / J), j = (int)(idx % J);
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(a[(long long)n * I + i], b[(long long)n * J + j], s);
        dw[idx] = s * alpha;
    }
}

// backward, pass 2: dx (+)= dy * s[n][c] + dmean[n][c] / HW
template <typename GT>
__global__ void __launch_bounds__(256)
se_bwd_apply_kernel(const GT* __restrict__ dy, long long dyp, const float* __restrict__ s,
                    const float* __restrict__ dmean, GT* __restrict__ dx, long long dxp, int HW, int C, float inv_hw,
                    int accumulate, long long total_vec) {
    const int CV = C / 8;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total_vec;
         idx += (long long)gridDim.x * blockDim.x) {
        const int cv = (int)(idx % CV);
        const long long pix = idx / CV;
        const int n = (int)(pix / HW);
        float gf[8], o[8];
        Half8<GT>::load(dy + pix * dyp + cv * 8, gf);
        if (accumulate) Half8<GT>::load(dx + pix * dxp + cv * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            const float t = fmaf(gf[j], __ldg(s + (long long)n * C + c), __ldg(dmean + (long long)n * C + c) * inv_hw);
            o[j] = accumulate ? o[j] + t : t;
        }
        Half8<GT>::store(dx + pix * dxp + cv * 8, o);
    }
}

inline bool vec_ok(int c, long long p0, long long p1) { return c > 0 && c % 8 == 0 && c / 8 <= 256 && p0 % 8 == 0 && p1 % 8 == 0; }
inline size_t dw_smem(int c, int k) { return (((size_t)k * k * c * 2 + 15) & ~(size_t)15) + 256 * 9 * sizeof(float); }

template <typename K>
inline int set_smem(K kern, size_t bytes, unsigned long long& mask) {
    if (bytes > 48 * 1024 && b2y_first_use_on_device(mask))
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)) != cudaSuccess)
            return -1;
    return 0;
}

}  // namespace

extern "C" int b2y_dwconv_fwd(const b2y_conv_desc* d, const void* x, const float* w, const float* scale,
                              const float* bias, void* y, float* stat_sum, float* stat_sqsum, void* stream) {
    if (!d || !x || !w || !y) return B2Y_ERR_INVALID;
    if (d->in_c != d->out_c || !vec_ok(d->in_c, d->in_pitch, d->out_pitch) || d->ksize < 1 || d->ksize > 7)
        return B2Y_ERR_UNSUPPORTED;
    const Geo g = geo_for(d->in_c);
    const long long pixels = (long long)d->batch * d->out_h * d->out_w;
    const size_t smem = dw_smem(d->in_c, d->ksize);
    if (smem > 200 * 1024) return B2Y_ERR_UNSUPPORTED;
    static unsigned long long mask = 0;
    if (set_smem(dwconv_fwd_kernel, smem, mask) != 0) return B2Y_ERR_CUDA;
    dwconv_fwd_kernel<<<grid_for_rows(pixels, g.PPB, 4), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half*>(x), d->in_pitch, w, scale, bias, reinterpret_cast<__half*>(y), d->out_pitch,
        stat_sum, stat_sqsum, d->batch, d->in_h, d->in_w, d->in_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w,
        d->act, d->slope, g.CV, g.PPB);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_dwconv_bwd_data(const b2y_conv_desc* d, const void* dz, const float* w, void* dx, int accumulate,
                                   int out_dtype, const float* inv_scale_ptr, void* stream) {
    if (!d || !dz || !w || !dx) return B2Y_ERR_INVALID;
    if (d->in_c != d->out_c || !vec_ok(d->in_c, d->in_pitch, d->out_pitch) || d->ksize < 1 || d->ksize > 7)
        return B2Y_ERR_UNSUPPORTED;
    const Geo g = geo_for(d->in_c);
    const long long pixels = (long long)d->batch * d->in_h * d->in_w;
    const size_t smem = dw_smem(d->in_c, d->ksize);
    if (smem > 200 * 1024) return B2Y_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = grid_for_rows(pixels, g.PPB, 4);
    static unsigned long long m16 = 0, mbf = 0;
    if (out_dtype == B2Y_DT_BF16) {
        if (set_smem(dwconv_bwd_data_kernel<__nv_bfloat16>, smem, mbf) != 0) return B2Y_ERR_CUDA;
        dwconv_bwd_data_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(
            reinterpret_cast<const __half*>(dz), d->out_pitch, w, reinterpret_cast<__nv_bfloat16*>(dx), d->in_pitch,
            d->batch, d->in_h, d->in_w, d->in_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w, accumulate,
            inv_scale_ptr, g.CV, g.PPB);
    } else {
        if (set_smem(dwconv_bwd_data_kernel<__half>, smem, m16) != 0) return B2Y_ERR_CUDA;
        dwconv_bwd_data_kernel<__half><<<grid, 256, smem, st>>>(
            reinterpret_cast<const __half*>(dz), d->out_pitch, w, reinterpret_cast<__half*>(dx), d->in_pitch, d->batch,
            d->in_h, d->in_w, d->in_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w, accumulate, inv_scale_ptr,
            g.CV, g.PPB);
    }
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_dwconv_bwd_weight(const b2y_conv_desc* d, const void* x, const void* dz, float* dw, float alpha,
                                     const float* inv_scale_ptr, void* stream) {
    if (!d || !x || !dz || !dw) return B2Y_ERR_INVALID;
    if (d->in_c != d->out_c || !vec_ok(d->in_c, d->in_pitch, d->out_pitch) || d->ksize < 1 || d->ksize > 7)
        return B2Y_ERR_UNSUPPORTED;
    const Geo g = geo_for(d->in_c);
    const long long pixels = (long long)d->batch * d->out_h * d->out_w;
    dim3 grid(grid_for_rows(pixels, g.PPB, 1), d->ksize * d->ksize);
    dwconv_bwd_weight_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __half*>(x), d->in_pitch, reinterpret_cast<const __half*>(dz), d->out_pitch, dw,
        d->batch, d->in_h, d->in_w, d->in_c, d->ksize, d->stride, d->pad, d->out_h, d->out_w, alpha, inv_scale_ptr,
        g.CV, g.PPB);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_se_fwd(const void* x, long long x_pitch, const float* w1, const float* w2, void* y,
                          long long y_pitch, int batch, int hw, int c, int cr, float* ws /* fp32 [batch][3c+cr] */,
                          void* stream) {
    if (!x || !w1 || !w2 || !y || !ws || batch <= 0 || hw <= 0 || cr <= 0) return B2Y_ERR_INVALID;
    if (!vec_ok(c, x_pitch, y_pitch)) return B2Y_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // workspace layout: pooled sum / mean [N][C] | h [N][Cr] | v [N][C] | s [N][C]
    float* pooled = ws;
    float* h = pooled + (long long)batch * c;
    float* v = h + (long long)batch * cr;
    float* s = v + (long long)batch * c;
    B2Y_CUDA_CHECK(cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)batch * c, st));
    const Geo g = geo_for(c);
    int gx = (hw + g.PPB - 1) / g.PPB;
    const int cap = (sm_count() * 4 + batch - 1) / batch;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    se_pool_kernel<<<dim3(gx, batch), 256, 0, st>>>(reinterpret_cast<const __half*>(x), x_pitch, pooled, hw, c, g.CV,
                                                    g.PPB);
    se_fc_kernel<<<batch, 256, sizeof(float) * (c + cr), st>>>(pooled, w1, w2, pooled, h, v, s, c, cr, 1.f / (float)hw);
    const long long total = (long long)batch * hw * (c / 8);
    long long gs = (total + 255) / 256;
    if (gs > (long long)sm_count() * 8) gs = (long long)sm_count() * 8;
    se_scale_kernel<<<(int)gs, 256, 0, st>>>(reinterpret_cast<const __half*>(x), x_pitch, s, reinterpret_cast<__half*>(y),
                                             y_pitch, hw, c, total);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_se_bwd(const void* x, long long x_pitch, const void* dy, long long dy_pitch, const float* w1,
                          const float* w2, const float* ws /* from b2y_se_fwd */, float* ws_bwd /* fp32 [batch][3c+cr] */,
                          void* dx, long long dx_pitch, int accumulate, float* dw1, float* dw2, float grad_scale,
                          int batch, int hw, int c, int cr, int grad_dtype, void* stream) {
    if (!x || !dy || !w1 || !w2 || !ws || !ws_bwd || !dx || !dw1 || !dw2) return B2Y_ERR_INVALID;
    if (!vec_ok(c, x_pitch, dy_pitch) || dx_pitch % 8 != 0) return B2Y_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float* mean = ws;
    const float* h = mean + (long long)batch * c;
    const float* v = h + (long long)batch * cr;
    const float* s = v + (long long)batch * c;
    float* ds = ws_bwd;                                   // [N][C]  sum dy*x, then reused as dv
    float* dh = ds + (long long)batch * c;                // [N][Cr]
    float* dmean = dh + (long long)batch * cr;            // [N][C]
    float* dv = dmean + (long long)batch * c;             // [N][C]
    B2Y_CUDA_CHECK(cudaMemsetAsync(ds, 0, sizeof(float) * (size_t)batch * c, st));
    const Geo g = geo_for(c);
    int gx = (hw + g.PPB - 1) / g.PPB;
    const int cap = (sm_count() * 4 + batch - 1) / batch;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    const long long total = (long long)batch * hw * (c / 8);
    long long gs = (total + 255) / 256;
    if (gs > (long long)sm_count() * 8) gs = (long long)sm_count() * 8;
    if (grad_dtype == B2Y_DT_BF16)
        se_bwd_reduce_kernel<__nv_bfloat16><<<dim3(gx, batch), 256, 0, st>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, ds, hw, c,
            g.CV, g.PPB);
    else
        se_bwd_reduce_kernel<__half><<<dim3(gx, batch), 256, 0, st>>>(
            reinterpret_cast<const __half*>(x), x_pitch, reinterpret_cast<const __half*>(dy), dy_pitch, ds, hw, c, g.CV,
            g.PPB);
    se_fc_bwd_kernel<<<batch, 256, sizeof(float) * (c + cr), st>>>(ds, v, h, w1, w2, dv, dh, dmean, c, cr);
    {
        const long long t2 = (long long)c * cr;
        int gw = (int)((t2 + 255) / 256);
        se_fc_wgrad_kernel<<<gw, 256, 0, st>>>(dv, h, dw2, batch, c, cr, grad_scale);        // dW2 [C][Cr]
        se_fc_wgrad_kernel<<<gw, 256, 0, st>>>(dh, mean, dw1, batch, cr, c, grad_scale);     // dW1 [Cr][C]
    }
    if (grad_dtype == B2Y_DT_BF16)
        se_bwd_apply_kernel<__nv_bfloat16><<<(int)gs, 256, 0, st>>>(
            reinterpret_cast<const __nv_bfloat16*>(dy), dy_pitch, s, dmean, reinterpret_cast<__nv_bfloat16*>(dx), dx_pitch,
            hw, c, 1.f / (float)hw, accumulate, total);
    else
        se_bwd_apply_kernel<__half><<<(int)gs, 256, 0, st>>>(
            reinterpret_cast<const __half*>(dy), dy_pitch, s, dmean, reinterpret_cast<__half*>(dx), dx_pitch, hw, c,
            1.f / (float)hw, accumulate, total);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}
