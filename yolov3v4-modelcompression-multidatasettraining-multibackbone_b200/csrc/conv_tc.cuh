// tcgen05 implicit-GEMM convolution, sm_100a.
//
//   D[M = B*Ho*Wo pixels][N = Cout] = sum_{tap (r,s)} sum_{c} A_tap[M][c] * W[N][(r,s,c)]
//
// * A (activations, NHWC) is never im2col'ed in memory: the TMA unit gathers each
//   (tap, 64-channel) slab straight from the NHWC tensor in im2col mode (zero fill for the
//   padding halo, traversal stride = conv stride) into a 128B-swizzled K-major smem tile.
// * B (weights, [Cout][R][S][Cin]) comes in through a tiled 2-D TMA map.
// * One elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=32 bytes) with the fp32/int32
//   accumulator in TMEM, double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Persistent grid (<= #SMs CTAs), warp-specialised: warp0 = TMA producer, warp1 = MMA issuer,
//   warp2 = TMEM allocator, warps 4-11 = epilogue (TMEM -> regs -> bias/act/residual -> global).
#pragma once
#include "common.cuh"

namespace b2y {

enum { CONV_KIND_F16 = 0, CONV_KIND_I8 = 1 };
enum { A_MODE_IM2COL = 0, A_MODE_TILED2D = 1 };
enum { OUT_F16 = 0, OUT_F32 = 1, OUT_I8 = 2, OUT_BF16 = 3 };

struct ConvTcParams {
    int M_total;      // B*Ho*Wo
    int Cout;         // valid output channels
    int num_m_tiles;  // ceil(M_total/128)
    int num_n_tiles;  // ceil(Cout/BLOCK_N)
    int k_chunks;     // Cin*esize / KBYTES
    int Cin;          // elements per tap (GEMM K per tap)
    int ntaps;        // filter taps visited (9 for 3x3; a subset for the stride-2 data-gradient phases)
    int MH, MW;       // GEMM row space = batch x MH x MW "base pixels"
    int stride;       // TMA traversal stride along W (conv stride for fprop, 1 for dgrad)
    int stride_h;     // ... along H (differs from `stride` only for the pixel-pair view of narrow stride-2 layers)
    int lower_w, lower_h;     // coordinate of base pixel (0,0): base = q*stride + lower
    unsigned char tap_ow[16]; // per-tap im2col offsets (>= 0)
    unsigned char tap_oh[16];
    unsigned long long tap_w_packed, tap_h_packed;   // the same offsets, 4 bits per tap (what the producer reads)
    int a_mode;
    // row -> output pixel mapping: pixel(n,p,q) = ((n*out_OH + p*out_ys + out_y0)*out_OW + q*out_xs + out_x0)
    int out_identity;         // 1: output pixel index == GEMM row index
    int out_OH, out_OW, out_ys, out_xs, out_y0, out_x0;
    // epilogue
    const float* bias;  // [Cout] or null
    int act;
    float slope;
    float acc_scale;        // multiplies the accumulator (1 for fp16; s_a*s_w for int8)
    const float* acc_scale_ptr;  // optional device scalar multiplied in as well (per-layer gradient un-scaling)
    unsigned idesc_ab;      // a_format<<7 | b_format<<10 for kind::f16 (0 = f16, 1 = bf16 per operand)
    int res_bf16;           // residual / accumulate tensor is bf16 instead of fp16
    const __half* res;      // optional residual (16-bit, NHWC) added after the activation
    long long res_pitch;    // elements per pixel row
    void* out;
    long long out_pitch;    // elements per pixel row
    int out_dtype;          // OUT_*
    float out_inv_scale;    // int8 output: q = clamp(round_half_away(v * out_inv_scale))
    float out_scale;        //              (and v_dequant = q * out_scale when out is fp16/fp32 fake-quant)
    int out_fakequant;      // 1: write the dequantised value q*out_scale in out_dtype (fp16/fp32)
    float q_lo, q_hi;       // clamp range, e.g. -128, 127
    // training extras: per-channel sum / sum of squares of the raw (pre-bias) output, fp32 atomics
    float* stat_sum;
    float* stat_sqsum;
    int b_resident;         // 1: the whole weight panel of the (single) N tile stays in smem, loaded once per CTA
    int num_stages;         // smem ring depth for this launch (<= 32)
    int pair;               // host: launch the cta_group::2 variant (cluster of two)
    int epi_tma;            // the short epilogue stores through smem + TMA (tmC valid); rows / columns clipped by the map
    int epi_fast;           // host-checked: 16-bit out, whole channel tiles, 16-byte aligned bias/residual/output
    // int8 graph: quantised shortcut folded into the epilogue (COSPTQuantizedShortcut eval, ptq_cos.py:876-884 / 931-933):
    //   code = clamp(rha((rha(q*out_scale*qs_rx)*qs_x + rha(a*qs_a_in*qs_ra)*qs_a) * qs_rsum)),  q = this conv's int8 code,
    // all scales powers of two (host-checked), so every product is exact and equals the stand-alone kernel bit for bit
    const int8_t* qres;     // the other addend's int8 codes (NHWC) or null
    long long qres_pitch;
    float qs_rx, qs_x, qs_a_in, qs_ra, qs_a, qs_rsum, qs_lo, qs_hi;
    int qs_simple;          // both addend roundings are identities: code = clamp(rha(q*qs_cx + a*qs_ca))
    float qs_cx, qs_ca;     // out_scale / scale_sum, a_in / scale_sum
};

template <int BLOCK_N, int KBYTES>
struct ConvTcCfg {
    static constexpr int BLOCK_M = 128;
    static constexpr int A_BYTES = BLOCK_M * KBYTES;
    static constexpr int B_BYTES = BLOCK_N * KBYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // always a multiple of 1024
    static constexpr int MAX_STAGE_SMEM = 192 * 1024;
    static constexpr int NUM_STAGES_RAW = MAX_STAGE_SMEM / STAGE_BYTES;
    // small-K layers have small stages: keep up to 32 of them in flight so that enough bytes are outstanding per SM
    // to cover HBM latency (Little's law: 148 SMs x ~190 KB / ~2 us)
    static constexpr int NUM_STAGES = NUM_STAGES_RAW > 32 ? 32 : NUM_STAGES_RAW;
    // accumulator ring in TMEM: narrow tiles get a deeper ring so the MMA warp can run several tiles ahead of
    // the (latency-bound) epilogue; 256-wide tiles use the whole 512-column TMEM with 2 stages
    static constexpr int ACC_STAGES = BLOCK_N >= 256 ? 2 : (BLOCK_N == 128 ? 4 : 8);
    static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;   // 512, 512, 512, 256 columns
    static constexpr int AUX_BYTES = 1024;  // barriers + tmem ptr, at a fixed offset behind the tile area
    static constexpr int EPI_STAGE_BYTES = 8 * 4096;   // per epilogue warp: 2 x 2 KB output staging for TMA stores
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + MAX_STAGE_SMEM + AUX_BYTES + EPI_STAGE_BYTES;
};

__device__ __noinline__ float mish_noinline(float x) { return mish_f(x); }
__device__ __noinline__ float swish_noinline(float x) { return x * sigmoid_f(x); }

__device__ __forceinline__ float round_half_away(float x) {
    // reference utils/quantized/quantized_ptq_cos.py:14-20  sign(x)*floor(|x|+0.5)
    return copysignf(floorf(fabsf(x) + 0.5f), x);
}

// The same rounding for |x| < 2^22 without the conversion pipe (FRND / F2I issue at a quarter of the FP32 rate and bound
// the int8 epilogues): floor(t) = (t + 2^23 rounded DOWN) - 2^23 for 0 <= t < 2^23.
__device__ __forceinline__ float round_half_away_small(float x) {
    const float t = fabsf(x) + 0.5f;
    return copysignf(__fadd_rd(t, 8388608.f) - 8388608.f, x);
}
// four integer-valued floats in [-128, 127] -> packed int8: the low byte of (v + 1.5 * 2^23) is v's two's complement
__device__ __forceinline__ uint32_t pack_i8x4(float a, float b, float c, float d) {
    const uint32_t ua = __float_as_uint(a + 12582912.f), ub = __float_as_uint(b + 12582912.f);
    const uint32_t uc = __float_as_uint(c + 12582912.f), ud = __float_as_uint(d + 12582912.f);
    return __byte_perm(__byte_perm(ua, ub, 0x0040), __byte_perm(uc, ud, 0x0040), 0x5410);
}

template <int BLOCK_N>
struct ConvTcEpi {
    // 8 epilogue warps (two groups of 4, one warp per TMEM lane quarter each):
    //   wide tiles  (BLOCK_N >= 128): both groups work on the same tile, half of the columns each;
    //   narrow tiles (BLOCK_N < 128): the groups take alternate tiles, so two accumulators drain concurrently.
    static constexpr int WARPS = 8;
    static constexpr bool SPLIT_TILES = BLOCK_N < 128;
    static constexpr int ARRIVALS = SPLIT_TILES ? 4 : 8;    // epilogue warps that release one accumulator stage
    static constexpr int THREADS = 128 + 32 * WARPS;
};

// General (slow-path) epilogue for one 32-column chunk: every output type, partial channel tiles, batch statistics,
// int8 requantisation.  `v` holds the scaled accumulators on entry.
template <int KIND>
__device__ __forceinline__ void epi_chunk_general(float (&v)[32], const ConvTcParams& p, int n0c0, long long row,
                                               bool row_ok, int lane, uint8_t* stage, bool defer_stats, float& ds1,
                                               float& ds2) {
    ds1 = ds2 = 0.f;
    if (p.stat_sum != nullptr) {
        // per-channel batch statistics of the raw conv output (training BN): the warp's 32x32 chunk goes through its smem
        // staging tile (rotated columns: conflict-free both ways), then lane j sums column j -- 32 STS + 32 LDS instead of
        // 640 shuffles per chunk
        float* tile = reinterpret_cast<float*>(stage);
#pragma unroll
        for (int j = 0; j < 32; ++j) tile[lane * 32 + ((j + lane) & 31)] = row_ok ? v[j] : 0.f;
        __syncwarp();
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float t = tile[i * 32 + ((lane + i) & 31)];
            s1 += t;
            s2 = fmaf(t, t, s2);
        }
        __syncwarp();
        if (defer_stats) {          // single N tile: the warp keeps per-column partial sums across its tiles
            ds1 = s1;
            ds2 = s2;
        } else if (n0c0 + lane < p.Cout) {
            atomicAdd(p.stat_sum + n0c0 + lane, s1);
            atomicAdd(p.stat_sqsum + n0c0 + lane, s2);
        }
    }
    const int nvalid = min(32, p.Cout - n0c0);
    if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid) v[j] += __ldg(p.bias + n0c0 + j);
    }
    switch (p.act) {
        case B2Y_ACT_LEAKY:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
            break;
        case B2Y_ACT_MISH:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = mish_noinline(v[j]);
            break;
        case B2Y_ACT_RELU:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            break;
        case B2Y_ACT_RELU6:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], 0.f), 6.f);
            break;
        case B2Y_ACT_HSWISH:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = v[j] * (fminf(fmaxf(v[j] + 3.f, 0.f), 6.f) / 6.f);
            break;
        case B2Y_ACT_SWISH:
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = swish_noinline(v[j]);
            break;
        default:
            break;
    }
    if (!row_ok) return;
    if (p.res != nullptr) {
        const __half* rp = p.res + row * p.res_pitch + n0c0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nvalid)
                v[j] += p.res_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rp)[j]) : __half2float(rp[j]);
    }
    if (p.out_fakequant || p.out_dtype == OUT_I8) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float q = round_half_away(v[j] * p.out_inv_scale);
            q = fminf(fmaxf(q, p.q_lo), p.q_hi);
            v[j] = (p.out_dtype == OUT_I8) ? q : q * p.out_scale;
        }
    }
    if (p.out_dtype == OUT_BF16) {
        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.out_pitch + n0c0;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                reinterpret_cast<uint4*>(op)[q] = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = __float2bfloat16_rn(v[j]);
        }
    } else if (p.out_dtype == OUT_F16) {
        __half* op = reinterpret_cast<__half*>(p.out) + row * p.out_pitch + n0c0;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 u;
                __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int t = 0; t < 4; ++t) h2[t] = __floats2half2_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                reinterpret_cast<uint4*>(op)[q] = u;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = __float2half_rn(v[j]);
        }
    } else if (p.out_dtype == OUT_F32) {
        float* op = reinterpret_cast<float*>(p.out) + row * p.out_pitch + n0c0;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                reinterpret_cast<float4*>(op)[q] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = v[j];
        }
    } else {  // OUT_I8
        int8_t* op = reinterpret_cast<int8_t*>(p.out) + row * p.out_pitch + n0c0;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint32_t w[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int b = q * 16 + t * 4;
                    w[t] = ((uint32_t)(uint8_t)(int8_t)(int)v[b]) | ((uint32_t)(uint8_t)(int8_t)(int)v[b + 1] << 8) |
                           ((uint32_t)(uint8_t)(int8_t)(int)v[b + 2] << 16) |
                           ((uint32_t)(uint8_t)(int8_t)(int)v[b + 3] << 24);
                }
                reinterpret_cast<uint4*>(op)[q] = make_uint4(w[0], w[1], w[2], w[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nvalid) op[j] = (int8_t)(int)v[j];
        }
    }
}

// CLUSTER > 1: the CTAs of a cluster work on CLUSTER consecutive M tiles of the same N tile; each loads 1/CLUSTER of
// the weight tile and TMA-multicasts it to all of them, so the weight re-reads from L2 drop by CLUSTER x.
//
// PAIR (CLUSTER == 2): the two CTAs of the cluster are one cta_group::2 MMA pair.  Each CTA loads its own 128-row A tile
// and HALF of the weight tile; the leader CTA issues one M=256 tcgen05.mma per k-slice that reads both halves, each
// CTA's TMEM receives its own 128 accumulator rows.  Per CTA the smem fill per k-step drops from A + B to A + B/2,
// which is what bounds the wide layers (L2 -> SM bandwidth), with no extra instructions on the critical loops.
template <int BLOCK_N, int KBYTES, int KIND, int CLUSTER, int PAIR = 0>
__global__ void __launch_bounds__(ConvTcEpi<BLOCK_N>::THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ ConvTcParams p) {
    using Cfg = ConvTcCfg<BLOCK_N, KBYTES>;
    using Epi = ConvTcEpi<BLOCK_N>;
    constexpr int MAX_NS = 32;
    constexpr int ESIZE = (KIND == CONV_KIND_F16) ? 2 : 1;
    constexpr int BLOCK_K = KBYTES / ESIZE;  // elements per k-chunk
    constexpr uint32_t LAYOUT = swizzle_layout_type(KBYTES);
    constexpr uint32_t IDESC = (KIND == CONV_KIND_F16)
                                   ? make_idesc(/*c=F32*/ 1, /*a=F16*/ 0, /*b=F16*/ 0, 0, 0, PAIR ? 256 : 128, BLOCK_N)
                                   : make_idesc(/*c=S32*/ 2, /*a=S8*/ 1, /*b=S8*/ 1, 0, 0, PAIR ? 256 : 128, BLOCK_N);
    static_assert(!PAIR || CLUSTER == 2, "a CTA pair is a cluster of two");

    // smem: [ns stages of A (+B)] [resident weight panel, if any] ... [aux at a fixed offset]
    const bool b_res = CLUSTER == 1 && p.b_resident != 0;
    const int ns = p.num_stages;
    const uint32_t stage_bytes = b_res ? (uint32_t)Cfg::A_BYTES
                                       : (PAIR ? (uint32_t)(Cfg::A_BYTES + Cfg::B_BYTES / 2) : (uint32_t)Cfg::STAGE_BYTES);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* aux = smem + Cfg::MAX_STAGE_SMEM;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);           // [MAX_NS]
    uint64_t* empty_bar = full_bar + MAX_NS;                         // [MAX_NS]
    constexpr int AS = Cfg::ACC_STAGES;
    uint64_t* tmem_full_bar = empty_bar + MAX_NS;                    // [AS]
    uint64_t* tmem_empty_bar = tmem_full_bar + AS;                   // [AS]
    uint64_t* bres_bar = tmem_empty_bar + AS;                        // [1] resident weight panel landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bres_bar + 1);
    uint8_t* epi_stage = aux + Cfg::AUX_BYTES;                       // [8 warps][4096] output staging for TMA stores

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
        prefetch_tmap(&tmC);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < ns; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], PAIR ? 1 : CLUSTER);   // multicast: every CTA that writes this stage releases it
        }
        mbar_init(bres_bar, 1);
        for (int i = 0; i < AS; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], PAIR ? 2 * Epi::ARRIVALS : Epi::ARRIVALS);   // pair: both CTAs' epilogues
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if (PAIR) {
            tmem_alloc2(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish2();
        } else {
            tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();   // barriers of every CTA are initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // work items: (group of CLUSTER consecutive M tiles, N tile); this CTA takes M tile group*CLUSTER + rank.
    // A "ghost" M tile past the end still runs the pipeline (TMA zero fill) but stores nothing.
    const int cta_rank = CLUSTER > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = (int)blockIdx.x / CLUSTER;
    const int num_clusters = (int)gridDim.x / CLUSTER;
    const int num_n_tiles = p.num_n_tiles;
    const int num_mgroups = (p.num_m_tiles + CLUSTER - 1) / CLUSTER;
    const int num_items = num_mgroups * num_n_tiles;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CLUSTER) - 1u);
    const int taps = p.ntaps;
    const int k_chunks = p.k_chunks;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full_base = smem_u32(full_bar);
    const uint32_t empty_base = smem_u32(empty_bar);
    const uint32_t bres_base = smem_base + (uint32_t)ns * stage_bytes;     // resident weight panel (k_steps x B tile)

    // The two issue loops below are executed by the whole (converged) warp with one elected lane doing the issue:
    // all loop state is then warp-uniform and lives in the uniform datapath.  (Run by a single divergent lane, each
    // TMA / MMA issue cost ~10 extra R2UR/ELECT instructions and the loops, at ~8 cycles per dependent instruction,
    // became the bottleneck of every layer whose k-step is shorter than ~800 cycles.)
    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t stage = 0, phase = 0;
        const int HoWo = p.MH * p.MW;
        const int MW = p.MW;
        const int cstride = p.stride, lower_w = p.lower_w, lower_h = p.lower_h, Cin = p.Cin;
        const bool im2col = p.a_mode == A_MODE_IM2COL;
        const unsigned long long tw = p.tap_w_packed, th = p.tap_h_packed;
        if (b_res && elect_one()) {
            // weight-stationary: all k-steps of the (only) N tile are fetched once; the ring then carries A alone
            const int k_steps = taps * k_chunks;
            mbar_expect_tx_s(smem_u32(bres_bar), (uint32_t)k_steps * Cfg::B_BYTES);
            for (int ks = 0; ks < k_steps; ++ks)
                tma_load_2d_s(bres_base + ks * Cfg::B_BYTES, &tmB, smem_u32(bres_bar), ks * BLOCK_K, 0);
        }
        __syncwarp();
        // everything above (barriers, TMEM, the static weights) may overlap the tail of the previous layer's kernel;
        // activations must not be touched before it has completed (programmatic dependent launch)
        pdl_wait();
        pdl_launch_dependents();
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const int mgroup = item / num_n_tiles;
            const int n_tile = item - mgroup * num_n_tiles;
            const int m0 = (mgroup * CLUSTER + cta_rank) * 128;
            int img = 0, base_w = 0, base_h = 0;
            if (im2col) {
                img = m0 / HoWo;
                const int rem = m0 - img * HoWo;
                const int po = rem / MW;
                base_w = (rem - po * MW) * cstride + lower_w;
                base_h = po * p.stride_h + lower_h;
            }
            const int b_row = n_tile * BLOCK_N + (CLUSTER > 1 ? cta_rank * (BLOCK_N / CLUSTER) : 0);
            int b_k = 0;
            for (int tap = 0; tap < taps; ++tap) {
                const uint16_t s = (uint16_t)((tw >> (4 * tap)) & 15);
                const uint16_t r = (uint16_t)((th >> (4 * tap)) & 15);
                int a_c = 0;
                for (int kc = 0; kc < k_chunks; ++kc) {
                    mbar_wait_s(empty_base + stage * 8, phase ^ 1);
                    if (elect_one()) {
                        const uint32_t a_dst = smem_base + stage * stage_bytes;
                        const uint32_t b_dst = a_dst + Cfg::A_BYTES;
                        const uint32_t fb = full_base + stage * 8;
                        if (PAIR) {
                            // all four loads of the pair complete on the leader's barrier, which expects both CTAs' bytes
                            if (cta_rank == 0) mbar_expect_tx_s(fb, 2 * stage_bytes);
                            const uint32_t lb = mapa_u32(fb, 0);
                            if (im2col)
                                tma2_load_im2col_4d(a_dst, &tmA, lb, a_c, base_w, base_h, img, s, r);
                            else
                                tma2_load_2d(a_dst, &tmA, lb, a_c, m0);
                            tma2_load_2d(b_dst, &tmB, lb, b_k, b_row);
                        } else {
                        mbar_expect_tx_s(fb, stage_bytes);
                        if (im2col)
                            tma_load_im2col_4d_s(a_dst, &tmA, fb, a_c, base_w, base_h, img, s, r);
                        else
                            tma_load_2d_s(a_dst, &tmA, fb, a_c, m0);
                        if (CLUSTER > 1)
                            tma_load_2d_multicast_s(b_dst + cta_rank * (BLOCK_N / CLUSTER) * KBYTES, &tmB, fb, b_k, b_row,
                                                    MC_MASK);
                        else if (!b_res)
                            tma_load_2d_s(b_dst, &tmB, fb, b_k, b_row);
                        }
                    }
                    a_c += BLOCK_K;
                    b_k += BLOCK_K;
                    if (++stage == (uint32_t)ns) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                b_k += Cin - k_chunks * BLOCK_K;   // == 0 (Cin is a whole number of chunks); keeps the intent explicit
            }
        }
    } else if (warp == 1 && (!PAIR || cta_rank == 0)) {
        // ===================== MMA issuer (pair: the leader CTA issues for both) =====================
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        const int k_steps = taps * k_chunks;
        const uint32_t idesc = IDESC | (KIND == CONV_KIND_F16 ? p.idesc_ab : 0u);
        const uint64_t desc_base = smem_desc_base(16, 8 * KBYTES, LAYOUT);
        const uint32_t tfull_base = smem_u32(tmem_full_bar), tempty_base = smem_u32(tmem_empty_bar);
        if (b_res) {
            mbar_wait_s(smem_u32(bres_bar), 0);
            tc_fence_after();
        }
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            mbar_wait_s(tempty_base + acc * 8, acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
            uint32_t accum = 0;
            for (int ks = 0; ks < k_steps; ++ks) {
                mbar_wait_s(full_base + stage * 8, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_base + stage * stage_bytes;
                    const uint32_t b_addr = b_res ? bres_base + ks * Cfg::B_BYTES : a_addr + Cfg::A_BYTES;
                    const uint64_t adesc = desc_base | (uint64_t)((a_addr >> 4) & 0x3FFF);
                    const uint64_t bdesc = desc_base | (uint64_t)((b_addr >> 4) & 0x3FFF);
#pragma unroll
                    for (int k = 0; k < KBYTES / 32; ++k) {
                        if (PAIR) {
                            if (KIND == CONV_KIND_F16)
                                mma2_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                            k > 0 ? 1u : accum);
                            else
                                mma2_i8_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                           k > 0 ? 1u : accum);
                        } else if (KIND == CONV_KIND_F16)
                            mma_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                       k > 0 ? 1u : accum);
                        else
                            mma_i8_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc,
                                      k > 0 ? 1u : accum);
                    }
                    if (PAIR)
                        tc_commit2_multicast_s(empty_base + stage * 8, MC_MASK);   // frees the slot in both CTAs
                    else if (CLUSTER > 1)
                        tc_commit_multicast_s(empty_base + stage * 8, MC_MASK);  // release the slot in every CTA
                    else
                        tc_commit_s(empty_base + stage * 8);  // frees the smem slot when these MMAs retire
                }
                accum = 1;
                if (++stage == (uint32_t)ns) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            if (elect_one()) {   // accumulator complete -> epilogue (of both CTAs for a pair)
                if (PAIR)
                    tc_commit2_multicast_s(tfull_base + acc * 8, MC_MASK);
                else
                    tc_commit_s(tfull_base + acc * 8);
            }
            if (++acc == AS) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        constexpr int COLS_PER_GROUP = Epi::SPLIT_TILES ? BLOCK_N : BLOCK_N / 2;
        pdl_wait();
        const float acc_mul = p.acc_scale * (p.acc_scale_ptr != nullptr ? __ldg(p.acc_scale_ptr) : 1.f);
        const int ew = (warp - 4) & 3;          // TMEM lane quarter == warp_id % 4
        const int cg = (warp - 4) >> 2;         // group
        const bool fast = p.epi_fast != 0;
        const int act = p.act;
        const float slope = p.slope;
        const int c_begin = Epi::SPLIT_TILES ? 0 : cg * COLS_PER_GROUP;
        const int hw = p.MH * p.MW;
        const bool epi_tma = p.epi_tma != 0;
        uint8_t* my_stage = epi_stage + (warp - 4) * 4096;
        int sbuf = 0;
        int it = 0;
        // training BN statistics: with a single N tile every tile of this warp covers the same columns, so the channel
        // sums are kept in registers (lane j = column j of chunk q) and flushed with ONE atomic per column per warp at the
        // end -- instead of one per column per 32-row chunk (millions of same-address L2 atomics on the large early layers)
        const bool defer_stats = p.stat_sum != nullptr && num_n_tiles == 1;
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
        for (int item = cluster_id; item < num_items; item += num_clusters, ++it) {
            if (Epi::SPLIT_TILES && (it & 1) != cg) continue;
            const int acc = it % AS;
            const uint32_t acc_phase = (uint32_t)(it / AS) & 1u;
            const int mgroup = item / num_n_tiles;
            const int n_tile = item - mgroup * num_n_tiles;
            const int m_tile = mgroup * CLUSTER + cta_rank;
            const int n0 = n_tile * BLOCK_N;
            const long long grow = (long long)m_tile * 128 + ew * 32 + lane;   // GEMM row
            const bool row_ok = grow < p.M_total;
            long long row = grow;                                              // output pixel index
            if (!p.out_identity && row_ok) {
                const int n_ = (int)(grow / hw);
                const int r_ = (int)(grow - (long long)n_ * hw);
                const int p_ = r_ / p.MW, q_ = r_ - p_ * p.MW;
                row = ((long long)n_ * p.out_OH + (long long)p_ * p.out_ys + p.out_y0) * p.out_OW +
                      (long long)q_ * p.out_xs + p.out_x0;
            }
            const uint32_t taddr_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BLOCK_N);

            if (fast) {
                // short path: 16-bit output (or fp32 through TMA), vector-aligned bias / residual / output (checked on the host);
                // a clipped last N tile is allowed on TMA-store launches
                const float* bias_p = p.bias != nullptr ? p.bias + n0 + c_begin : nullptr;
                const uint4* res_p = (p.res != nullptr && row_ok)
                                         ? reinterpret_cast<const uint4*>(p.res + row * p.res_pitch + n0 + c_begin)
                                         : nullptr;
                uint4* out_p = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + row * p.out_pitch + n0 + c_begin);
                uint4 rnext[4];
                if (res_p != nullptr) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) rnext[q] = __ldg(res_p + q);
                }
                if (p.res != nullptr) {
                    // pull the residual rows of this warp's NEXT tile into L2 now: by the time its accumulator is ready
                    // the loads above hit L2 instead of paying HBM latency inside the epilogue
                    const int nitem = item + (Epi::SPLIT_TILES ? 2 : 1) * num_clusters;
                    if (nitem < num_items) {
                        const int nmg = nitem / num_n_tiles;
                        const int nnt = nitem - nmg * num_n_tiles;
                        const long long ngrow = (long long)(nmg * CLUSTER + cta_rank) * 128 + ew * 32 + lane;
                        if (ngrow < p.M_total && p.out_identity) {
                            const char* np_ = reinterpret_cast<const char*>(p.res + ngrow * p.res_pitch + nnt * BLOCK_N + c_begin);
#pragma unroll
                            for (int b = 0; b < COLS_PER_GROUP * 2; b += 128)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(np_ + b));
                        }
                    }
                }
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < COLS_PER_GROUP; c += 32) {
                    const int cvalid = p.Cout - (n0 + c_begin + c);     // < 32 only in the last chunk of a clipped N tile
                    if (cvalid <= 0) break;                              // (TMA-store launches only; warp-uniform)
                    uint32_t raw[32];
                    tmem_ld_32x32(taddr_row + (uint32_t)(c_begin + c), raw);
                    float4 bv[8];
                    if (bias_p != nullptr && cvalid >= 32) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) bv[q] = __ldg(reinterpret_cast<const float4*>(bias_p + c) + q);
                    } else if (bias_p != nullptr) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            bv[q].x = 4 * q < cvalid ? __ldg(bias_p + c + 4 * q) : 0.f;
                            bv[q].y = 4 * q + 1 < cvalid ? __ldg(bias_p + c + 4 * q + 1) : 0.f;
                            bv[q].z = 4 * q + 2 < cvalid ? __ldg(bias_p + c + 4 * q + 2) : 0.f;
                            bv[q].w = 4 * q + 3 < cvalid ? __ldg(bias_p + c + 4 * q + 3) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q) bv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    uint4 qa[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                    if (KIND == CONV_KIND_I8 && p.qres != nullptr && row_ok) {
                        const uint4* ap = reinterpret_cast<const uint4*>(p.qres + row * p.qres_pitch + n0 + c_begin + c);
                        qa[0] = __ldg(ap);
                        qa[1] = __ldg(ap + 1);
                    }
                    uint4 rcur[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
                    if (res_p != nullptr && c + 32 < COLS_PER_GROUP) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) rnext[q] = __ldg(res_p + (c + 32) / 8 + q);
                    }
                    tc_wait_ld();
                    float v[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float a0 = (KIND == CONV_KIND_F16) ? __uint_as_float(raw[4 * q]) : (float)(int)raw[4 * q];
                        const float a1 = (KIND == CONV_KIND_F16) ? __uint_as_float(raw[4 * q + 1]) : (float)(int)raw[4 * q + 1];
                        const float a2 = (KIND == CONV_KIND_F16) ? __uint_as_float(raw[4 * q + 2]) : (float)(int)raw[4 * q + 2];
                        const float a3 = (KIND == CONV_KIND_F16) ? __uint_as_float(raw[4 * q + 3]) : (float)(int)raw[4 * q + 3];
                        v[4 * q] = fmaf(a0, acc_mul, bv[q].x);
                        v[4 * q + 1] = fmaf(a1, acc_mul, bv[q].y);
                        v[4 * q + 2] = fmaf(a2, acc_mul, bv[q].z);
                        v[4 * q + 3] = fmaf(a3, acc_mul, bv[q].w);
                    }
                    if (KIND == CONV_KIND_F16 && p.stat_sum != nullptr) {
                        // training BN statistics on the short path (host: TMA store, no bias, no residual): column sums of
                        // the warp's 32 x 32 chunk through the 2 KB staging buffer the NEXT TMA store will use, 16 columns
                        // at a time (rotated columns; lane L sums column L & 15 over rows 16 (L >> 4) .. + 15)
                        float* tile = reinterpret_cast<float*>(my_stage + sbuf * 2048);
                        if (lane == 0) bulk_wait_read<1>();      // the store that last read this buffer is done with it
                        __syncwarp();
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) tile[lane * 16 + ((j + lane) & 15)] = row_ok ? v[h * 16 + j] : 0.f;
                            __syncwarp();
                            const int col = lane & 15, r0 = (lane >> 4) * 16;
                            float t1 = 0.f, t2 = 0.f;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float t = tile[(r0 + i) * 16 + ((col + r0 + i) & 15)];
                                t1 += t;
                                t2 = fmaf(t, t, t2);
                            }
                            t1 += __shfl_xor_sync(0xffffffffu, t1, 16);
                            t2 += __shfl_xor_sync(0xffffffffu, t2, 16);
                            if ((lane >> 4) == h) {
                                s1 = t1;
                                s2 = t2;
                            }
                            __syncwarp();
                        }
                        if (defer_stats) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                if (c == q * 32) {
                                    st1[q] += s1;
                                    st2[q] += s2;
                                }
                        } else if (n0 + c_begin + c + lane < p.Cout) {
                            atomicAdd(p.stat_sum + n0 + c_begin + c + lane, s1);
                            atomicAdd(p.stat_sqsum + n0 + c_begin + c + lane, s2);
                        }
                    }
                    if (act == B2Y_ACT_LEAKY) {       // 0 <= slope <= 1 on this path: max(v, slope*v)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], v[j] * slope);
                    } else if (act == B2Y_ACT_MISH) {
                        // inlined closed form (one ex2, one rcp per element): 32 independent dependency chains the
                        // scheduler can interleave -- as a call per element the epilogue was bound by the serial
                        // MUFU latency of each chain (yolov4's Mish layers ran 2-6x slower than their leaky twins)
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = mish_f(v[j]);
                    }
                    if (row_ok && res_p != nullptr) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint4 u = rcur[q];
                            if (p.res_bf16) {
                                const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 f = __bfloat1622float2(b2[t]);
                                    v[q * 8 + t * 2] += f.x;
                                    v[q * 8 + t * 2 + 1] += f.y;
                                }
                            } else {
                                const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const float2 f = __half22float2(h2[t]);
                                    v[q * 8 + t * 2] += f.x;
                                    v[q * 8 + t * 2 + 1] += f.y;
                                }
                            }
                        }
                    }
                    if (KIND == CONV_KIND_I8 && (p.out_dtype == OUT_I8 || p.out_fakequant)) {
                        // int8 graph: requantise exactly like the general path (round half away, clamp); values are
                        // identical because acc_mul is a power of two there
                        // clamp first: the bounds are integers and the rounding is monotone, so the result is the same
                        // and the rounding operand is small
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            v[j] = round_half_away_small(fminf(fmaxf(v[j] * p.out_inv_scale, p.q_lo), p.q_hi));
                        if (p.qres != nullptr && p.qs_simple) {
                            // both addends already sit on the common grid (scale_x <= out_scale, scale_a <= sa_in: what the
                            // _min / _max shortcuts vote): their roundings are identities, the sum is exact in fp32
                            const uint32_t* aw = reinterpret_cast<const uint32_t*>(qa);
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                const uint32_t ax = aw[t] ^ 0x80808080u;
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float af = __uint_as_float(__byte_perm(ax, 0x4B400000u, 0x7650u | e)) - 12583040.f;
                                    const float tt = fmaf(af, p.qs_ca, v[t * 4 + e] * p.qs_cx);
                                    v[t * 4 + e] = round_half_away_small(fminf(fmaxf(tt, p.qs_lo), p.qs_hi));
                                }
                            }
                        } else if (p.qres != nullptr) {
                            const int8_t* ab = reinterpret_cast<const int8_t*>(qa);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float xf = round_half_away(v[j] * p.out_scale * p.qs_rx) * p.qs_x;   // rounded, NOT clamped
                                const float af = round_half_away((float)ab[j] * p.qs_a_in * p.qs_ra) * p.qs_a;
                                const float q = round_half_away((xf + af) * p.qs_rsum);
                                v[j] = fminf(fmaxf(q, p.qs_lo), p.qs_hi);
                            }
                        } else if (p.out_fakequant) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= p.out_scale;
                        }
                    }
                    if (p.out_dtype == OUT_I8) {
                        uint32_t w[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) w[t] = pack_i8x4(v[t * 4], v[t * 4 + 1], v[t * 4 + 2], v[t * 4 + 3]);
                        if (epi_tma) {
                            // 32 rows x 32 B through smem (SWIZZLE_32B image) and one TMA store: whole 32-byte sectors per
                            // row reach L2 in one request instead of 2 x 32 scattered 16-byte pieces
                            uint8_t* buf = my_stage + sbuf * 2048;
                            if (lane == 0) bulk_wait_read<1>();
                            __syncwarp();
                            const int sw = (lane >> 2) & 1;
                            *reinterpret_cast<uint4*>(buf + lane * 32 + ((0 ^ sw) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
                            *reinterpret_cast<uint4*>(buf + lane * 32 + ((1 ^ sw) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
                            fence_proxy_async_cta();
                            __syncwarp();
                            if (lane == 0) {
                                tma_store_2d(&tmC, smem_u32(buf), n0 + c_begin + c, m_tile * 128 + ew * 32);
                                bulk_commit();
                            }
                            sbuf ^= 1;
                        } else if (row_ok) {
                            uint4* op8 = reinterpret_cast<uint4*>(reinterpret_cast<int8_t*>(p.out) + row * p.out_pitch + n0 +
                                                                  c_begin + c);
                            op8[0] = make_uint4(w[0], w[1], w[2], w[3]);
                            op8[1] = make_uint4(w[4], w[5], w[6], w[7]);
                        }
                        continue;
                    }
                    if (p.out_dtype == OUT_F32) {
                        // fp32 rows (YOLO head): 32 rows x 128 B, SWIZZLE_128B image, one TMA store; the map clips column 255
                        uint8_t* buf = my_stage;
                        if (lane == 0) bulk_wait_read<0>();
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            *reinterpret_cast<float4*>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        fence_proxy_async_cta();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC, smem_u32(buf), n0 + c_begin + c, m_tile * 128 + ew * 32);
                            bulk_commit();
                        }
                        continue;
                    }
                    uint4 ou[4];
                    if (p.out_dtype == OUT_BF16) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&ou[q]);
#pragma unroll
                            for (int t = 0; t < 4; ++t) h2[t] = __floats2bfloat162_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            __half2* h2 = reinterpret_cast<__half2*>(&ou[q]);
#pragma unroll
                            for (int t = 0; t < 4; ++t) h2[t] = __floats2half2_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                        }
                    }
                    if (epi_tma) {
                        // 32 rows x 64 B through smem (SWIZZLE_64B image) and one TMA store: full 64-byte row segments reach
                        // L2 instead of 32 scattered 16-byte pieces per store instruction; the map clips the M / Cout tails
                        uint8_t* buf = my_stage + sbuf * 2048;
                        if (lane == 0) bulk_wait_read<1>();      // the store that last read this buffer is done with it
                        __syncwarp();
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<uint4*>(buf + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = ou[q];
                        fence_proxy_async_cta();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&tmC, smem_u32(buf), n0 + c_begin + c, m_tile * 128 + ew * 32);
                            bulk_commit();
                        }
                        sbuf ^= 1;
                    } else if (row_ok) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) out_p[c / 8 + q] = ou[q];
                    }
                }
            } else {
                mbar_wait(&tmem_full_bar[acc], acc_phase);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < COLS_PER_GROUP; c += 32) {
                    const int c0 = c_begin + c;
                    uint32_t raw[32];
                    tmem_ld_32x32(taddr_row + (uint32_t)c0, raw);
                    tc_wait_ld();
                    if (n0 + c0 >= p.Cout) continue;  // warp-uniform
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        v[j] = ((KIND == CONV_KIND_F16) ? __uint_as_float(raw[j]) : (float)(int)raw[j]) * acc_mul;
                    float ds1, ds2;
                    epi_chunk_general<KIND>(v, p, n0 + c0, row, row_ok, lane, my_stage, defer_stats, ds1, ds2);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (c == q * 32) {
                            st1[q] += ds1;
                            st2[q] += ds2;
                        }
                }
            }
            // release this accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR)
                    mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty_bar[acc]), 0));   // the leader's MMA warp waits on it
                else
                    mbar_arrive(&tmem_empty_bar[acc]);
            }
        }
        if (defer_stats) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int col = c_begin + q * 32 + lane;
                if (q * 32 < COLS_PER_GROUP && col < p.Cout && (st1[q] != 0.f || st2[q] != 0.f)) {
                    atomicAdd(p.stat_sum + col, st1[q]);
                    atomicAdd(p.stat_sqsum + col, st2[q]);
                }
            }
        }
        if (epi_tma && lane == 0) bulk_wait_all();   // outstanding TMA stores complete before the CTA exits
    }

    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast / arrive into its smem
    if (warp == 2) {
        tc_fence_after();
        if (PAIR)
            tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
        else
            tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace b2y
