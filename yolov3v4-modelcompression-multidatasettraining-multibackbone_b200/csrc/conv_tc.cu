// Host side of the tcgen05 implicit-GEMM convolution: TMA tensor-map encoding, tile-shape dispatch,
// C-ABI entry points (include/b200yolo.h).
#include "conv_tc.cuh"

#include <mutex>

#include "b200yolo.h"

namespace b2y {

// ---- driver entry points (resolved lazily; the library does not link libcuda) -------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_encodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encodeTiled = nullptr;
static PFN_encodeIm2col g_encodeIm2col = nullptr;
static int g_driver_version = 0;
static int g_num_sms = 0;
static std::once_flag g_once;

static void resolve_driver() {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        g_encodeTiled = reinterpret_cast<PFN_encodeTiled>(fn);
    fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        g_encodeIm2col = reinterpret_cast<PFN_encodeIm2col>(fn);
    cudaDriverGetVersion(&g_driver_version);
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
}

static CUtensorMapSwizzle swizzle_enum(int kbytes) {
    return kbytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                         : (kbytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

// 2-D tiled map over a row-major [rows][cols] matrix with row pitch `pitch_elems`.
static int make_map_2d(CUtensorMap* m, const void* base, int esize, long long rows, long long cols,
                       long long pitch_elems, int box_cols, int box_rows, int kbytes) {
    if (!g_encodeTiled) return B2Y_ERR_DRIVER;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)(pitch_elems * esize)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encodeTiled(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                               const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle_enum(kbytes), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? B2Y_OK : B2Y_ERR_DRIVER;
}

// im2col map over an NHWC activation tensor (dims C,W,H,N) for an RxS / stride / pad convolution.
static int make_map_im2col(CUtensorMap* m, const void* base, int esize, int N, int H, int W, int C,
                           long long pitch_elems, int R, int S, int stride, int pad, int block_k, int kbytes) {
    if (!g_encodeIm2col) return B2Y_ERR_DRIVER;
    cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstride[3] = {(cuuint64_t)(pitch_elems * esize), (cuuint64_t)(pitch_elems * esize * W),
                             (cuuint64_t)(pitch_elems * esize * W * (long long)H)};
    // bounding box of the *base pixel* (top-left tap): lower = -pad, upper = pad - (filter-1)
    int lower[2] = {-pad, -pad};
    int upper[2] = {pad - (S - 1), pad - (R - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = g_encodeIm2col(m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 4,
                                const_cast<void*>(base), gdim, gstride, lower, upper, (cuuint32_t)block_k,
                                /*pixelsPerColumn*/ 128, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(kbytes),
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return B2Y_ERR_DRIVER;
    // Drivers <= 13.1 mis-encode im2col maps of tensors smaller than 128 KiB (same workaround as CUTLASS).
    if (g_driver_version <= 13010) {
        long long bytes = (long long)N * H * W * pitch_elems * esize;
        if (bytes < 131072) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
    }
    return B2Y_OK;
}

template <int BLOCK_N, int KBYTES, int KIND>
static int launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcParams& p, cudaStream_t st) {
    using Cfg = ConvTcCfg<BLOCK_N, KBYTES>;
    auto kern = conv_tc_kernel<BLOCK_N, KBYTES, KIND>;
    static bool attr_set = false;
    if (!attr_set) {
        B2Y_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int tiles = p.num_m_tiles * p.num_n_tiles;
    int grid = tiles < g_num_sms ? tiles : g_num_sms;
    kern<<<grid, 256, Cfg::SMEM_BYTES, st>>>(tmA, tmB, p);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

template <int KIND>
static int dispatch(int block_n, int kbytes, const CUtensorMap& a, const CUtensorMap& b, const ConvTcParams& p,
                    cudaStream_t st) {
#define B2Y_CASE(BN, KB) \
    if (block_n == BN && kbytes == KB) return launch_cfg<BN, KB, KIND>(a, b, p, st);
    B2Y_CASE(32, 32) B2Y_CASE(32, 64) B2Y_CASE(32, 128)
    B2Y_CASE(64, 32) B2Y_CASE(64, 64) B2Y_CASE(64, 128)
    B2Y_CASE(128, 32) B2Y_CASE(128, 64) B2Y_CASE(128, 128)
    B2Y_CASE(256, 32) B2Y_CASE(256, 64) B2Y_CASE(256, 128)
#undef B2Y_CASE
    return B2Y_ERR_UNSUPPORTED;
}

struct EpilogueArgs {
    const float* bias = nullptr;
    int act = 0;
    float slope = 0.1f;
    float acc_scale = 1.f;
    const void* res = nullptr;
    long long res_pitch = 0;
    void* out = nullptr;
    long long out_pitch = 0;
    int out_dtype = OUT_F16;
    float out_scale = 1.f;
    int out_fakequant = 0;
    float q_lo = -128.f, q_hi = 127.f;
    float* stat_sum = nullptr;
    float* stat_sqsum = nullptr;
};

// Shared launcher: x NHWC (esize 2 = fp16, 1 = int8), w [Cout][R][S][Cin].
int conv_tc_launch(int kind, const b2y_conv_desc* d, const void* x, const void* w, const EpilogueArgs& e,
                   cudaStream_t st) {
    std::call_once(g_once, resolve_driver);
    const int esize = kind == CONV_KIND_F16 ? 2 : 1;
    if (!d || !x || !w || !e.out) return B2Y_ERR_INVALID;
    if (d->batch <= 0 || d->in_c <= 0 || d->out_c <= 0 || d->ksize <= 0 || d->stride <= 0) return B2Y_ERR_INVALID;
    const long long kb_total = (long long)d->in_c * esize;
    if (kb_total % 32 != 0) return B2Y_ERR_UNSUPPORTED;
    int kbytes = kb_total % 128 == 0 ? 128 : (kb_total % 64 == 0 ? 64 : 32);
    if ((d->in_pitch * esize) % 16 != 0) return B2Y_ERR_INVALID;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(w) & 15)) return B2Y_ERR_INVALID;
    const int Ho = (d->in_h + 2 * d->pad - d->ksize) / d->stride + 1;
    const int Wo = (d->in_w + 2 * d->pad - d->ksize) / d->stride + 1;
    if (Ho != d->out_h || Wo != d->out_w) return B2Y_ERR_INVALID;
    const long long M = (long long)d->batch * Ho * Wo;
    if (M > 0x7fffff00LL) return B2Y_ERR_UNSUPPORTED;

    int block_n = d->out_c <= 32 ? 32 : (d->out_c <= 64 ? 64 : (d->out_c <= 128 ? 128 : 256));
    const int block_k = kbytes / esize;

    ConvTcParams p{};
    p.M_total = (int)M;
    p.Cout = d->out_c;
    p.num_m_tiles = (int)((M + 127) / 128);
    p.num_n_tiles = (d->out_c + block_n - 1) / block_n;
    p.k_chunks = (int)(kb_total / kbytes);
    p.Cin = d->in_c;
    p.R = d->ksize;
    p.S = d->ksize;
    p.Ho = Ho;
    p.Wo = Wo;
    p.stride = d->stride;
    p.pad = d->pad;
    p.bias = e.bias;
    p.act = e.act;
    p.slope = e.slope;
    p.acc_scale = e.acc_scale;
    p.res = reinterpret_cast<const __half*>(e.res);
    p.res_pitch = e.res_pitch;
    p.out = e.out;
    p.out_pitch = e.out_pitch;
    p.out_dtype = e.out_dtype;
    p.out_scale = e.out_scale;
    p.out_inv_scale = 1.f / e.out_scale;
    p.out_fakequant = e.out_fakequant;
    p.q_lo = e.q_lo;
    p.q_hi = e.q_hi;
    p.stat_sum = e.stat_sum;
    p.stat_sqsum = e.stat_sqsum;

    CUtensorMap tmA, tmB;
    int rc;
    const bool pointwise = (d->ksize == 1 && d->stride == 1 && d->pad == 0);
    if (pointwise) {
        p.a_mode = A_MODE_TILED2D;
        rc = make_map_2d(&tmA, x, esize, M, d->in_c, d->in_pitch, block_k, 128, kbytes);
    } else {
        p.a_mode = A_MODE_IM2COL;
        rc = make_map_im2col(&tmA, x, esize, d->batch, d->in_h, d->in_w, d->in_c, d->in_pitch, d->ksize, d->ksize,
                             d->stride, d->pad, block_k, kbytes);
    }
    if (rc != B2Y_OK) return rc;
    const long long Ktot = (long long)d->ksize * d->ksize * d->in_c;
    rc = make_map_2d(&tmB, w, esize, d->out_c, Ktot, Ktot, block_k, block_n, kbytes);
    if (rc != B2Y_OK) return rc;

    if (kind == CONV_KIND_F16) return dispatch<CONV_KIND_F16>(block_n, kbytes, tmA, tmB, p, st);
    return dispatch<CONV_KIND_I8>(block_n, kbytes, tmA, tmB, p, st);
}

}  // namespace b2y

using namespace b2y;

extern "C" int b2y_conv2d_fwd(const b2y_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                              const void* residual, void* y, void* stream) {
    if (!d) return B2Y_ERR_INVALID;
    EpilogueArgs e;
    e.bias = bias;
    e.act = d->act;
    e.slope = d->slope;
    e.res = residual;
    e.res_pitch = d->res_pitch;
    e.out = y;
    e.out_pitch = d->out_pitch;
    e.out_dtype = d->out_dtype == B2Y_OUT_F32 ? OUT_F32 : OUT_F16;
    return conv_tc_launch(CONV_KIND_F16, d, x, w_packed, e, static_cast<cudaStream_t>(stream));
}

extern "C" int b2y_conv2d_fwd_stats(const b2y_conv_desc* d, const void* x, const void* w_packed, const float* bias,
                                    void* y, float* stat_sum, float* stat_sqsum, void* stream) {
    if (!d || !stat_sum || !stat_sqsum) return B2Y_ERR_INVALID;
    EpilogueArgs e;
    e.bias = bias;
    e.act = B2Y_ACT_LINEAR;
    e.out = y;
    e.out_pitch = d->out_pitch;
    e.out_dtype = d->out_dtype == B2Y_OUT_F32 ? OUT_F32 : OUT_F16;
    e.stat_sum = stat_sum;
    e.stat_sqsum = stat_sqsum;
    return conv_tc_launch(CONV_KIND_F16, d, x, w_packed, e, static_cast<cudaStream_t>(stream));
}

extern "C" int b2y_qconv2d_fwd(const b2y_qconv_desc* d, const void* x_i8, const void* w_i8, const float* bias,
                               void* y, void* stream) {
    if (!d) return B2Y_ERR_INVALID;
    EpilogueArgs e;
    e.bias = bias;
    e.act = d->conv.act;
    e.slope = d->conv.slope;
    e.acc_scale = d->acc_scale;
    e.out = y;
    e.out_pitch = d->conv.out_pitch;
    e.out_scale = d->out_scale;
    e.q_lo = d->q_lo;
    e.q_hi = d->q_hi;
    if (d->out_kind == B2Y_OUT_I8) {
        e.out_dtype = OUT_I8;
    } else {
        e.out_dtype = d->out_kind == B2Y_OUT_F32 ? OUT_F32 : OUT_F16;
        e.out_fakequant = d->requant ? 1 : 0;
    }
    return conv_tc_launch(CONV_KIND_I8, &d->conv, x_i8, w_i8, e, static_cast<cudaStream_t>(stream));
}
