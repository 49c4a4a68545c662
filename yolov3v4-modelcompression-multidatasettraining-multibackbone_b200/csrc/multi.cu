// Table-driven ("multi-tensor") layout kernels of the training step: ONE launch re-packs the fp32 master weights of
// every convolution into the two fp16 operand layouts of the tensor-core kernels, ONE launch turns all packed fp32
// weight gradients back into the OIHW parameter layout.  Round 1 issued 4 small launches per layer and step for this
// (pack_weights / pack_dgrad / unpack_wgrad / axpby: ~10 % of the step, launch bound).
//
//   master   W  fp32 [O][I][k][k]                       (nn.Conv2d.weight, checkpoint layout, models.py:92-99)
//   forward  Wf fp16 [Opad][k][k][Ipad]                 (K-major B operand of the implicit GEMM, conv_tc.cuh; columns
//                                                        >= I are never written: the caller zero-fills them once)
//   dgrad    Wd fp16 [phase][I][tap][Opad]              (per output-phase slabs, conv_tc.cu enumerate_dgrad_phases)
//   wgrad    dW fp32 [Opad][k][k][Ipad] -> OIHW fp32    (wgrad_tc.cu writes the packed form with red.global.add)
//
// Work unit = a tile of 32 output channels x TI input channels x all taps, staged through shared memory so that both
// the global reads and the global writes are contiguous runs (>= 64 bytes).  b2y_layout_tile_i(k) gives TI.
#include "b200yolo.h"
#include "common.cuh"

using namespace b2y;

namespace {

constexpr int TO = 32;          // output channels per tile
constexpr int ROW = 289;        // floats per smem row (k*k*TI <= 288, +1 -> conflict-free column reads)

__host__ __device__ inline int tile_i(int k) {
    int t = 288 / (k * k);
    t -= t % 8;
    return t > 256 ? 256 : (t < 8 ? 8 : t);
}

template <typename Item>
__device__ __forceinline__ int find_item(const Item* __restrict__ items, int n, int tile) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {            // last item with tile_begin <= tile
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].tile_begin <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
pack_multi_kernel(const b2y_pack_item* __restrict__ items, int n_items) {
    __shared__ float tile[TO][ROW];
    __shared__ long long tap_base[16];     // dgrad: element offset of (phase slab + tap index * Opad) per source tap
    __shared__ int tap_ntaps[16];          // dgrad: taps of that phase
    const b2y_pack_item it = items[find_item(items, n_items, blockIdx.x)];
    const int k2 = it.k * it.k, TI = tile_i(it.k);
    const int t = blockIdx.x - it.tile_begin;
    const int tiles_i = (it.I + TI - 1) / TI;
    const int o0 = (t / tiles_i) * TO, i0 = (t % tiles_i) * TI;
    const int ni = min(TI, it.I - i0);
    if (threadIdx.x < k2) {
        // phase of tap (r, s): the output phase py with (py + pad - r) % stride == 0; inside the phase taps are
        // ordered by ascending (r, s) -- same enumeration as enumerate_dgrad_phases (conv_tc.cu)
        const int st = it.stride, pad = it.pad, k = it.k;
        const int r = threadIdx.x / k, s = threadIdx.x % k;
        const int py = ((r - pad) % st + st) % st, px = ((s - pad) % st + st) % st;
        long long off = 0;
        int my_nh = 0, my_nw = 0, r_idx = 0, s_idx = 0;
        for (int qy = 0; qy < st; ++qy)
            for (int qx = 0; qx < st; ++qx) {
                const int rmin = (qy + pad) % st, smin = (qx + pad) % st;
                const int nh = rmin < k ? (k - 1 - rmin) / st + 1 : 0;
                const int nw = smin < k ? (k - 1 - smin) / st + 1 : 0;
                if (qy == py && qx == px) {
                    my_nh = nh;
                    my_nw = nw;
                    r_idx = (r - rmin) / st;
                    s_idx = (s - smin) / st;
                } else if (qy < py || (qy == py && qx < px)) {
                    off += (long long)it.I * nh * nw * it.Opad;
                }
            }
        tap_base[threadIdx.x] = off + (long long)(r_idx * my_nw + s_idx) * it.Opad;
        tap_ntaps[threadIdx.x] = my_nh * my_nw;
    }
    // ---- load: per output channel one contiguous run of ni*k2 floats ----
    const int run = ni * k2;
    for (int idx = threadIdx.x; idx < TO * run; idx += 256) {
        const int ol = idx / run, rem = idx - ol * run;
        const int o = o0 + ol;
        tile[ol][rem] = o < it.O ? __ldg(it.w + ((long long)o * it.I + i0) * k2 + rem) : 0.f;
    }
    __syncthreads();
    // ---- forward layout [o][tap][i]: runs of ni halves ----
    if (it.w_fwd != nullptr) {
        __half* wf = reinterpret_cast<__half*>(it.w_fwd);
        for (int idx = threadIdx.x; idx < TO * k2 * ni; idx += 256) {
            const int il = idx % ni;
            const int tmp = idx / ni;
            const int tap = tmp % k2, ol = tmp / k2;
            const int o = o0 + ol;
            if (o < it.Opad)
                wf[((long long)o * k2 + tap) * it.Ipad + i0 + il] = __float2half_rn(tile[ol][il * k2 + tap]);
        }
    }
    // ---- data-gradient layout [phase][i][tap][o]: runs of 32 output channels ----
    if (it.w_dgrad != nullptr) {
        __half* wd = reinterpret_cast<__half*>(it.w_dgrad);
        for (int idx = threadIdx.x; idx < ni * k2 * TO; idx += 256) {
            const int ol = idx % TO;
            const int tmp = idx / TO;
            const int tap = tmp % k2, il = tmp / k2;
            const int o = o0 + ol;
            if (o < it.Opad)
                wd[tap_base[tap] + (long long)(i0 + il) * tap_ntaps[tap] * it.Opad + o] =
                    __float2half_rn(tile[ol][il * k2 + tap]);
        }
    }
}

__global__ void __launch_bounds__(256)
unpack_multi_kernel(const b2y_unpack_item* __restrict__ items, int n_items) {
    __shared__ float tile[TO][ROW];
    const b2y_unpack_item it = items[find_item(items, n_items, blockIdx.x)];
    const int k2 = it.k * it.k, TI = tile_i(it.k);
    const int t = blockIdx.x - it.tile_begin;
    const int tiles_i = (it.I + TI - 1) / TI;
    const int o0 = (t / tiles_i) * TO, i0 = (t % tiles_i) * TI;
    const int ni = min(TI, it.I - i0);
    const int no = min(TO, it.O - o0);
    if (no <= 0) return;
    // packed [o][tap][i]: runs of ni floats
    for (int idx = threadIdx.x; idx < no * k2 * ni; idx += 256) {
        const int il = idx % ni;
        const int tmp = idx / ni;
        const int tap = tmp % k2, ol = tmp / k2;
        tile[ol][il * k2 + tap] = __ldg(it.src + ((long long)(o0 + ol) * k2 + tap) * it.Ipad + i0 + il);
    }
    __syncthreads();
    const int run = ni * k2;
    for (int idx = threadIdx.x; idx < no * run; idx += 256) {
        const int ol = idx / run, rem = idx - ol * run;
        float* d = it.dst + ((long long)(o0 + ol) * it.I + i0) * k2 + rem;
        const float v = tile[ol][rem];
        *d = it.accumulate ? *d + v : v;
    }
}

}  // namespace

extern "C" int b2y_layout_tile_i(int ksize) { return ksize >= 1 && ksize <= 4 ? tile_i(ksize) : 0; }

extern "C" int b2y_pack_conv_weights_multi(const b2y_pack_item* items_dev, int n_items, int total_tiles, void* stream) {
    if (!items_dev || n_items <= 0 || total_tiles <= 0) return B2Y_ERR_INVALID;
    pack_multi_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" int b2y_unpack_wgrad_multi(const b2y_unpack_item* items_dev, int n_items, int total_tiles, void* stream) {
    if (!items_dev || n_items <= 0 || total_tiles <= 0) return B2Y_ERR_INVALID;
    unpack_multi_kernel<<<total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(items_dev, n_items);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}
