// Shared device helpers for the b200yolo kernels (sm_100a only).
// Thin inline-PTX wrappers around mbarrier / TMA / tcgen05 / TMEM.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "b200yolo.h"  // status codes + activation ids (single source of truth)

namespace b2y {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug traps (kernel error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s at 2 GHz
            __trap();
        }
    }
}

// shared-window (u32) address variants for the single-thread issue loops: no generic->shared conversion per call
__device__ __forceinline__ bool mbar_try_wait_s(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_s(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_s(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void mbar_expect_tx_s(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// programmatic dependent launch: wait for the upstream grid's memory / let the downstream grid start its prologue
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Host side: launch `kern` as a programmatic dependent of the previous kernel in the stream (B2Y_PDL=0: plain stream
// order).  The grid may become resident while its predecessor drains, so the kernel MUST execute pdl_wait() before it
// touches anything the predecessor wrote (and should call pdl_launch_dependents() right after, so that at most one
// grid is ever parked behind a running one).  In a captured graph the edge becomes a programmatic dependency.
inline bool b2y_pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B2Y_PDL");
        v = (e && atoi(e) == 0) ? 0 : 1;
    }
    return v != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    int na = 0;
    if (b2y_pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        na = 1;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// multicast variant: the box lands at the same smem offset in every CTA of `mask` and signals each one's mbarrier
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                                      int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// im2col-mode load of an NHWC tensor: coords (c, w, h, n) are the *base pixel*
// (already offset by the lower corner), (off_w, off_h) the filter tap.
__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h),
        "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

__device__ __forceinline__ void tma_load_2d_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1,
                                                        uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d_s(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c, int w, int h,
                                                     int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
        : "memory");
}

// TMA store of a smem box (written by generic stores + fence.proxy.async) to global; bulk-group completion
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_cta() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) variants: both CTAs of the pair issue their own loads, all of them complete on the
// mbarrier of the leader CTA (rank 0), addressed through the shared::cluster window (mapa).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_im2col_4d(uint32_t dst, const CUtensorMap* m, uint32_t leader_bar, int c, int w,
                                                    int h, int n, uint16_t off_w, uint16_t off_h) {
    asm volatile(
        "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w),
        "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives on the barrier at this smem offset in every CTA of `mask` once all prior cta_group::2 MMAs retire
__device__ __forceinline__ void tc_commit2_multicast_s(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}
// D[tmem of both CTAs, 256 rows] (+)= A[smem of both CTAs] * B[N/2 rows in each CTA's smem]
__device__ __forceinline__ void mma2_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma2_i8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier when all prior MMAs issued by this thread retire.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void tc_commit_s(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit_multicast_s(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
}

// commit that arrives on the barrier at the same smem offset in every CTA of `mask` (cluster multicast)
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t mask) {
    asm volatile(
        "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc];  kind::f16 (fp16/bf16 in, fp32 accumulate)
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::i8 (int8 in, int32 accumulate)
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns -> 32 registers per thread (thread = TMEM lane).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// Shared-memory matrix descriptor (tcgen05), swizzled canonical layouts.
//   K-major : rows of `row_bytes` (32/64/128 = swizzle span), 8-row groups SBO apart.
//   bits: [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type
__host__ __device__ __forceinline__ uint64_t smem_desc_base(uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}
__device__ __forceinline__ uint64_t smem_desc_at(uint64_t base, uint32_t smem_addr) {
    return base | (uint64_t)((smem_addr >> 4) & 0x3FFF);
}
__host__ __device__ constexpr uint32_t swizzle_layout_type(int row_bytes) {
    return row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : (row_bytes == 32 ? 6u : 0u));
}

// Instruction descriptor (upper 32 bits of the idesc operand).
//   c_format [4,6): 1=F32 2=S32 | a_format [7,10) | b_format [10,13) | a_major 15 | b_major 16
//   n_dim [17,23) = N>>3 | m_dim [24,29) = M>>4
__host__ __device__ constexpr uint32_t make_idesc(uint32_t c_fmt, uint32_t a_fmt, uint32_t b_fmt, uint32_t a_mn_major,
                                                  uint32_t b_mn_major, uint32_t M, uint32_t N) {
    return (c_fmt << 4) | (a_fmt << 7) | (b_fmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
           ((M >> 4) << 24);
}

// ---------------------------------------------------------------- activations (fp32)
__device__ __forceinline__ float softplus_f(float x) {
    // torch F.softplus(beta=1, threshold=20)  (reference utils/layers.py:148)
    return x > 20.f ? x : log1pf(expf(x));
}
// Mish (reference utils/layers.py:118-128, 146-148): tanh(softplus(x)) in closed form.  With e = exp(x):
//   tanh(log(1+e)) = ((1+e)^2 - 1) / ((1+e)^2 + 1) = n / (n + 2),  n = e (e + 2)
// one exponential and one division instead of exp + log1p + tanh; no cancellation for either sign of x (for x << 0 it
// tends to e like the original).  x is clamped at 20, where torch's softplus switches to the identity and tanh is 1
// in fp32 anyway.  Agreement with the literal formula: a few fp32 ulp.
__device__ __forceinline__ float mish_tanh_softplus(float x, float& e) {
    e = __expf(fminf(x, 20.f));
    const float n = e * (e + 2.f);
    return __fdividef(n, n + 2.f);
}
__device__ __forceinline__ float mish_f(float x) {
    float e;
    return x * mish_tanh_softplus(x, e);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    switch (act) {
        case B2Y_ACT_LEAKY: return v > 0.f ? v : v * slope;
        case B2Y_ACT_MISH: return mish_f(v);
        case B2Y_ACT_RELU: return fmaxf(v, 0.f);
        case B2Y_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
        case B2Y_ACT_HSWISH: return v * (fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f);
        case B2Y_ACT_SWISH: return v * sigmoid_f(v);
        default: return v;
    }
}
// d act(v) / dv
__device__ __forceinline__ float act_grad(float v, int act, float slope) {
    switch (act) {
        case B2Y_ACT_LEAKY: return v > 0.f ? 1.f : slope;
        case B2Y_ACT_MISH: {
            // reference utils/layers.py:123-128: fx + x * sigmoid(x) * (1 - fx^2), fx = tanh(softplus(x))
            // fx = n / (n + 2) and sigmoid(x) = e / (1 + e) share ONE reciprocal, 1 / ((n + 2)(1 + e)) (<= e^60, no
            // overflow): the backward BN passes are bound by the MUFU pipe (exp + reciprocals), this drops a third of it
            const float e = __expf(fminf(v, 20.f));
            const float n = e * (e + 2.f), d1 = n + 2.f, d2 = 1.f + e;
            const float r = __fdividef(1.f, d1 * d2);
            const float fx = n * d2 * r;
            const float sx = e * d1 * r;
            return fx + v * sx * (1.f - fx * fx);
        }
        case B2Y_ACT_RELU: return v > 0.f ? 1.f : 0.f;
        case B2Y_ACT_RELU6: return (v > 0.f && v < 6.f) ? 1.f : 0.f;
        case B2Y_ACT_HSWISH: {
            float r = fminf(fmaxf(v + 3.f, 0.f), 6.f) / 6.f;
            float dr = (v > -3.f && v < 3.f) ? (1.f / 6.f) : 0.f;
            return r + v * dr;
        }
        case B2Y_ACT_SWISH: {
            float sx = sigmoid_f(v);
            return sx * (1.f + v * (1.f - sx));
        }
        default: return 1.f;
    }
}

}  // namespace b2y

namespace b2y {
// 16-bit storage helpers so that gradient kernels can be instantiated for fp16 or bf16 tensors
template <typename T> struct Half8;
template <> struct Half8<__half> {
    __device__ static void unpack(const uint4& u, float (&f)[8]) {
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    __device__ static void load(const __half* p, float (&f)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __half22float2(h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    __device__ static void store(__half* p, const float (&f)[8]) {
        uint4 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    __device__ static float to_f(__half v) { return __half2float(v); }
    __device__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <> struct Half8<__nv_bfloat16> {
    __device__ static void unpack(const uint4& u, float (&f)[8]) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __bfloat1622float2(h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __bfloat1622float2(h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&f)[8]) {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = u;
    }
    __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
}  // namespace b2y

// host-side helpers ---------------------------------------------------------
// true the first time it is called for the current device with this mask (launch attributes such as
// MaxDynamicSharedMemorySize are per device; a process may drive several GPUs)
static inline bool b2y_first_use_on_device(unsigned long long& mask) {
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (mask & bit) return false;
    mask |= bit;
    return true;
}

#define B2Y_CUDA_CHECK(expr)                       \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) {                   \
            b2y_set_last_cuda_error((int)_e);      \
            return B2Y_ERR_CUDA;                   \
        }                                          \
    } while (0)

extern "C" void b2y_set_last_cuda_error(int e);
