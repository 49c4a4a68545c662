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
//   warp2 = TMEM allocator, warps 4-7 = epilogue (TMEM -> regs -> bias/act/residual -> global).
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
    int stride;       // TMA traversal stride (conv stride for fprop, 1 for dgrad)
    int lower_w, lower_h;     // coordinate of base pixel (0,0): base = q*stride + lower
    unsigned char tap_ow[16]; // per-tap im2col offsets (>= 0)
    unsigned char tap_oh[16];
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
};

template <int BLOCK_N, int KBYTES>
struct ConvTcCfg {
    static constexpr int BLOCK_M = 128;
    static constexpr int A_BYTES = BLOCK_M * KBYTES;
    static constexpr int B_BYTES = BLOCK_N * KBYTES;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // always a multiple of 1024
    static constexpr int MAX_STAGE_SMEM = 196 * 1024;
    static constexpr int NUM_STAGES_RAW = MAX_STAGE_SMEM / STAGE_BYTES;
    // small-K layers have small stages: keep up to 32 of them in flight so that enough bytes are outstanding per SM
    // to cover HBM latency (Little's law: 148 SMs x ~190 KB / ~2 us)
    static constexpr int NUM_STAGES = NUM_STAGES_RAW > 32 ? 32 : NUM_STAGES_RAW;
    // accumulator ring in TMEM: narrow tiles get a deeper ring so the MMA warp can run several tiles ahead of
    // the (latency-bound) epilogue; 256-wide tiles use the whole 512-column TMEM with 2 stages
    static constexpr int ACC_STAGES = BLOCK_N >= 256 ? 2 : (BLOCK_N == 128 ? 4 : 8);
    static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;   // 512, 512, 512, 256 columns
    static constexpr int AUX_BYTES = 1024;  // barriers + tmem ptr + bias staging
    static constexpr int BIAS_BYTES = BLOCK_N * 4;
    static constexpr int SMEM_BYTES = 1024 /*align slack*/ + NUM_STAGES * STAGE_BYTES + AUX_BYTES + BIAS_BYTES;
};

__device__ __noinline__ float mish_noinline(float x) { return mish_f(x); }
__device__ __noinline__ float swish_noinline(float x) { return x * sigmoid_f(x); }

__device__ __forceinline__ float round_half_away(float x) {
    // reference utils/quantized/quantized_ptq_cos.py:14-20  sign(x)*floor(|x|+0.5)
    return copysignf(floorf(fabsf(x) + 0.5f), x);
}

template <int BLOCK_N>
struct ConvTcEpi {
    // two column groups of epilogue warps for wide tiles: more loads/stores in flight per SM
    static constexpr int WARPS = BLOCK_N >= 128 ? 8 : 4;
    static constexpr int THREADS = 128 + 32 * WARPS;
};

// CLUSTER > 1: the CTAs of a cluster work on CLUSTER consecutive M tiles of the same N tile; each loads 1/CLUSTER of
// the weight tile and TMA-multicasts it to all of them, so the (dominant) weight re-reads from L2 drop by CLUSTER x.
template <int BLOCK_N, int KBYTES, int KIND, int CLUSTER>
__global__ void __launch_bounds__(ConvTcEpi<BLOCK_N>::THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const ConvTcParams p) {
    using Cfg = ConvTcCfg<BLOCK_N, KBYTES>;
    constexpr int NS = Cfg::NUM_STAGES;
    constexpr int ESIZE = (KIND == CONV_KIND_F16) ? 2 : 1;
    constexpr int BLOCK_K = KBYTES / ESIZE;  // elements per k-chunk
    constexpr uint32_t LAYOUT = swizzle_layout_type(KBYTES);
    constexpr uint32_t IDESC = (KIND == CONV_KIND_F16)
                                   ? make_idesc(/*c=F32*/ 1, /*a=F16*/ 0, /*b=F16*/ 0, 0, 0, 128, BLOCK_N)
                                   : make_idesc(/*c=S32*/ 2, /*a=S8*/ 1, /*b=S8*/ 1, 0, 0, 128, BLOCK_N);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* aux = smem + NS * Cfg::STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);           // [NS]
    uint64_t* empty_bar = full_bar + NS;                             // [NS]
    constexpr int AS = Cfg::ACC_STAGES;
    uint64_t* tmem_full_bar = empty_bar + NS;                        // [AS]
    uint64_t* tmem_empty_bar = tmem_full_bar + AS;                   // [AS]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + AS);
    float* sbias = reinterpret_cast<float*>(aux + Cfg::AUX_BYTES);   // [BLOCK_N]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA);
        prefetch_tmap(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], CLUSTER);   // every CTA that multicasts into this stage must see it released
        }
        for (int i = 0; i < AS; ++i) {
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], ConvTcEpi<BLOCK_N>::WARPS);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
        tmem_relinquish();
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
    const int num_mgroups = (p.num_m_tiles + CLUSTER - 1) / CLUSTER;
    const int num_items = num_mgroups * p.num_n_tiles;
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CLUSTER) - 1u);
    const int taps = p.ntaps;
    const int k_steps = taps * p.k_chunks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int HoWo = p.MH * p.MW;
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                const int mgroup = item / p.num_n_tiles;
                const int n_tile = item - mgroup * p.num_n_tiles;
                const int m_tile = mgroup * CLUSTER + cta_rank;
                const int m0 = m_tile * 128;
                const int img = m0 / HoWo;
                const int rem = m0 - img * HoWo;
                const int po = rem / p.MW;
                const int qo = rem - po * p.MW;
                const int base_w = qo * p.stride + p.lower_w;
                const int base_h = po * p.stride + p.lower_h;
                for (int tap = 0; tap < taps; ++tap) {
                    const int r = p.tap_oh[tap];
                    const int s = p.tap_ow[tap];
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
                        uint8_t* b_dst = a_dst + Cfg::A_BYTES;
                        mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                        if (p.a_mode == A_MODE_IM2COL) {
                            tma_load_im2col_4d(a_dst, &tmA, &full_bar[stage], kc * BLOCK_K, base_w, base_h, img,
                                               (uint16_t)s, (uint16_t)r);
                        } else {
                            tma_load_2d(a_dst, &tmA, &full_bar[stage], kc * BLOCK_K, m0);
                        }
                        if (CLUSTER > 1) {
                            constexpr int SLICE_ROWS = BLOCK_N / CLUSTER;
                            tma_load_2d_multicast(b_dst + cta_rank * SLICE_ROWS * KBYTES, &tmB, &full_bar[stage],
                                                  tap * p.Cin + kc * BLOCK_K, n_tile * BLOCK_N + cta_rank * SLICE_ROWS,
                                                  MC_MASK);
                        } else {
                            tma_load_2d(b_dst, &tmB, &full_bar[stage], tap * p.Cin + kc * BLOCK_K, n_tile * BLOCK_N);
                        }
                        if (++stage == NS) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint64_t desc_base = smem_desc_base(16, 8 * KBYTES, LAYOUT);
            for (int item = cluster_id; item < num_items; item += num_clusters) {
                mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_N);
                for (int ks = 0; ks < k_steps; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + Cfg::A_BYTES;
                    const uint64_t adesc = smem_desc_at(desc_base, a_addr);
                    const uint64_t bdesc = smem_desc_at(desc_base, b_addr);
#pragma unroll
                    for (int k = 0; k < KBYTES / 32; ++k) {
                        const uint32_t accum = (ks > 0 || k > 0) ? 1u : 0u;
                        if (KIND == CONV_KIND_F16)
                            mma_f16_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC | p.idesc_ab,
                                       accum);
                        else
                            mma_i8_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, accum);
                    }
                    if (CLUSTER > 1)
                        tc_commit_multicast(&empty_bar[stage], MC_MASK);  // release the slot in every CTA of the cluster
                    else
                        tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == NS) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
                if (++acc == AS) {
                    acc = 0;
                    acc_phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        constexpr int EPI_WARPS = ConvTcEpi<BLOCK_N>::WARPS;
        constexpr int EPI_THREADS = 32 * EPI_WARPS;
        constexpr int COLS_PER_GROUP = BLOCK_N / (EPI_WARPS / 4);
        const float acc_mul = p.acc_scale * (p.acc_scale_ptr != nullptr ? __ldg(p.acc_scale_ptr) : 1.f);
        const int ew = (warp - 4) & 3;          // TMEM lane quarter == warp_id % 4
        const int cg = (warp - 4) >> 2;         // column group
        const int et = threadIdx.x - 128;       // 0..EPI_THREADS-1
        int acc = 0;
        uint32_t acc_phase = 0;
        bool first_item = true;
        for (int item = cluster_id; item < num_items; item += num_clusters) {
            const int mgroup = item / p.num_n_tiles;
            const int n_tile = item - mgroup * p.num_n_tiles;
            const int m_tile = mgroup * CLUSTER + cta_rank;
            const int n0 = n_tile * BLOCK_N;
            const long long grow = (long long)m_tile * 128 + ew * 32 + lane;   // GEMM row
            const bool row_ok = grow < p.M_total;
            long long row = grow;                                              // output pixel index
            if (!p.out_identity && row_ok) {
                const int hw = p.MH * p.MW;
                const int n_ = (int)(grow / hw);
                const int r_ = (int)(grow - (long long)n_ * hw);
                const int p_ = r_ / p.MW, q_ = r_ - p_ * p.MW;
                row = ((long long)n_ * p.out_OH + (long long)p_ * p.out_ys + p.out_y0) * p.out_OW +
                      (long long)q_ * p.out_xs + p.out_x0;
            }
            const int c_begin = cg * COLS_PER_GROUP;
            const int c_end = c_begin + COLS_PER_GROUP;
            // residual prefetch (one 32-column chunk ahead; the first chunk is issued before the accumulator wait)
            const __half* res_row = (p.res != nullptr && row_ok) ? p.res + row * p.res_pitch + n0 : nullptr;
            const bool res_vec = res_row != nullptr && ((reinterpret_cast<uintptr_t>(res_row) & 15) == 0);
            uint4 rnext[4];
            if (res_vec && n0 + c_begin + 32 <= p.Cout) {
#pragma unroll
                for (int q = 0; q < 4; ++q) rnext[q] = __ldg(reinterpret_cast<const uint4*>(res_row + c_begin) + q);
            }

            // stage the bias slice for this tile
            if (p.num_n_tiles > 1 || first_item) {   // a single N tile: the bias slice never changes
                first_item = false;
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
                for (int i = et; i < BLOCK_N; i += EPI_THREADS) {
                    const int n = n0 + i;
                    sbias[i] = (p.bias != nullptr && n < p.Cout) ? __ldg(p.bias + n) : 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
            }

            mbar_wait(&tmem_full_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr_row = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BLOCK_N);

#pragma unroll 1
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(taddr_row + (uint32_t)c0, raw);
                uint4 rcur[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) rcur[q] = rnext[q];
                if (res_vec && c0 + 32 < c_end && n0 + c0 + 64 <= p.Cout) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        rnext[q] = __ldg(reinterpret_cast<const uint4*>(res_row + c0 + 32) + q);
                }
                tc_wait_ld();
                if (n0 + c0 >= p.Cout) continue;  // warp-uniform

                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float a = (KIND == CONV_KIND_F16) ? __uint_as_float(raw[j]) : (float)(int)raw[j];
                    v[j] = a * acc_mul;
                }

                if (p.stat_sum != nullptr) {
                    // per-channel batch statistics of the raw conv output (training BN):
                    // butterfly-reduce each column over the 32 rows of this warp.
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float s1 = row_ok ? v[j] : 0.f;
                        float s2 = s1 * s1;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                        }
                        if (lane == j && n0 + c0 + j < p.Cout) {
                            atomicAdd(p.stat_sum + n0 + c0 + j, s1);
                            atomicAdd(p.stat_sqsum + n0 + c0 + j, s2);
                        }
                    }
                }

                // bias + activation: the switch is hoisted out of the element loop so that the executed path is one
                // short straight-line block (a per-element switch made the unrolled body ~64 KB and I-cache bound)
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += sbias[c0 + j];
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

                const int nvalid = min(32, p.Cout - (n0 + c0));
                if (row_ok) {
                    if (p.res != nullptr) {
                        const __half* rp = p.res + row * p.res_pitch + n0 + c0;
                        if (nvalid == 32 && res_vec) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint4 u = rcur[q];
                                if (p.res_bf16) {
                                    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        float2 f = __bfloat1622float2(b2[t]);
                                        v[q * 8 + t * 2] += f.x;
                                        v[q * 8 + t * 2 + 1] += f.y;
                                    }
                                } else {
                                    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                                    for (int t = 0; t < 4; ++t) {
                                        float2 f = __half22float2(h2[t]);
                                        v[q * 8 + t * 2] += f.x;
                                        v[q * 8 + t * 2 + 1] += f.y;
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid)
                                    v[j] += p.res_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(rp)[j])
                                                       : __half2float(rp[j]);
                        }
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
                        __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.out_pitch + n0 + c0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u;
                                __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    h2[t] = __floats2bfloat162_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                                reinterpret_cast<uint4*>(op)[q] = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid) op[j] = __float2bfloat16_rn(v[j]);
                        }
                    } else if (p.out_dtype == OUT_F16) {
                        __half* op = reinterpret_cast<__half*>(p.out) + row * p.out_pitch + n0 + c0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                uint4 u;
                                __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    h2[t] = __floats2half2_rn(v[q * 8 + t * 2], v[q * 8 + t * 2 + 1]);
                                reinterpret_cast<uint4*>(op)[q] = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid) op[j] = __float2half_rn(v[j]);
                        }
                    } else if (p.out_dtype == OUT_F32) {
                        float* op = reinterpret_cast<float*>(p.out) + row * p.out_pitch + n0 + c0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                reinterpret_cast<float4*>(op)[q] =
                                    make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid) op[j] = v[j];
                        }
                    } else {  // OUT_I8
                        int8_t* op = reinterpret_cast<int8_t*>(p.out) + row * p.out_pitch + n0 + c0;
                        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                uint32_t w[4];
#pragma unroll
                                for (int t = 0; t < 4; ++t) {
                                    const int b = q * 16 + t * 4;
                                    w[t] = ((uint32_t)(uint8_t)(int8_t)(int)v[b]) |
                                           ((uint32_t)(uint8_t)(int8_t)(int)v[b + 1] << 8) |
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
                __syncwarp();
            }
            // release this accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
            if (++acc == AS) {
                acc = 0;
                acc_phase ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CLUSTER > 1) cluster_sync_all();   // nobody leaves while a peer may still multicast / arrive into its smem
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace b2y
