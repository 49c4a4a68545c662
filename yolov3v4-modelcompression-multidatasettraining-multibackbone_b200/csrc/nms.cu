// Detection post-processing on the device: the reference's non_max_suppression
// (utils/utils.py:782-860: confidence / size filter, multi-label candidates in nonzero() order, per-class offset boxes,
// torchvision NMS with its stable score sort and double-precision threshold compare, 'merge' box refinement) and the
// per-image true-positive matching loop of test.py:150-170.  HBM / latency bound integer + fp32 work: no tensor cores.
//
//   rows pass (count)  : one warp per 32 prediction rows (a lane per row header) -> number of candidates of every row
//   cub exclusive scan : candidate offsets = the order nonzero() enumerates them (row-major, class ascending)
//   rows pass (emit)   : candidates (xyxy, conf, class) + 64-bit sort keys (image | descending score)
//   cub radix sort     : stable, so equal scores keep ascending candidate order = torch's stable descending sort
//   greedy kernel      : one CTA per image walks the sorted list in chunks of 512: every candidate is tested against
//                        the boxes kept so far (tiles through shared memory), survivors of the chunk are resolved among
//                        themselves with a 512 x 512 bit matrix and a one-warp serial scan; no n^2 matrix in HBM
//   finish kernel      : one warp per kept box: score-weighted mean of the overlapping candidates (1 < n < 3000) and
//                        the output row (x1, y1, x2, y2, conf, cls)
#include <cstdint>

#include <cub/cub.cuh>

#include "b200yolo.h"
#include "common.cuh"

namespace {

constexpr float kMinWH = 2.f, kMaxWH = 4096.f;      // utils.py:790
constexpr int kMergeMax = 3000;                     // utils.py:844: merge only for 1 < n < 3000
constexpr int NT = 512;                             // candidates per chunk of the greedy kernel = threads per CTA
constexpr int NW = NT / 32;

__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) <= 3.402823466e38f; }   // false for NaN / inf

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// descending-score key: larger score -> smaller key (scores are finite)
__device__ __forceinline__ unsigned desc_key(float s) {
    unsigned b = __float_as_uint(s);
    unsigned ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~ord;
}

// torchvision nms_kernel / utils.box_iou arithmetic, operation by operation in fp32 (no contraction)
__device__ __forceinline__ float box_area(float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }
__device__ __forceinline__ float box_iou(float4 a, float aarea, float4 b, float barea) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter));
}
// `IoU > threshold` without the division when the boxes do not intersect: w == 0 or h == 0 gives inter = +0, so the
// quotient is +-0 or NaN and the comparison is false for every threshold >= 0 -- the result is bit-identical to evaluating
// the quotient (callers pass nonneg = threshold >= 0; otherwise the quotient is always formed).
__device__ __forceinline__ bool iou_gt_d(float4 a, float aarea, float4 b, float barea, double thr, bool nonneg) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    if (nonneg && (w == 0.f || h == 0.f)) return false;
    const float inter = __fmul_rn(w, h);
    return (double)__fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter)) > thr;     // torchvision: double threshold
}
__device__ __forceinline__ bool iou_gt_f(float4 a, float aarea, float4 b, float barea, float thr, bool nonneg) {
    const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y), xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
    if (nonneg && (w == 0.f || h == 0.f)) return false;
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(aarea, barea), inter)) > thr;             // torch: scalar cast to fp32
}
__device__ __forceinline__ float4 shift_box(float4 b, float cls, int agnostic) {
    const float o = agnostic ? 0.f : __fmul_rn(cls, kMaxWH);        // boxes + c * max_wh (utils.py:840-841)
    return make_float4(__fadd_rn(b.x, o), __fadd_rn(b.y, o), __fadd_rn(b.z, o), __fadd_rn(b.w, o));
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per 32 consecutive prediction rows: every lane loads the 5-float header of its own row (32 rows in flight per
// warp instead of one), the rows that pass the confidence / size filters -- a few per cent for a trained detector -- are
// then expanded one after the other by the whole warp.  EMIT = false: counts[row] = number of candidates of the row.
// EMIT = true: the candidates are written at row_off[row] in class order.
template <bool EMIT>
__global__ void nms_rows_kernel(const float* __restrict__ pred, long long rows, int R, int nc, float conf_thres,
                                int multi_label, const unsigned char* __restrict__ allow, int* __restrict__ counts,
                                const int* __restrict__ row_off, float4* __restrict__ cand_box,
                                float* __restrict__ cand_conf, float* __restrict__ cand_cls,
                                unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const long long row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
    if (row0 >= rows) return;
    const long long my_row = row0 + lane;
    const int no = 5 + nc;
    float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f, h4 = 0.f;
    bool my_ok = false;
    if (my_row < rows) {
        const float* hx = pred + my_row * (long long)no;
        h0 = hx[0]; h1 = hx[1]; h2 = hx[2]; h3 = hx[3]; h4 = hx[4];
        // x[:, 4] > conf_thres, then ((x[:, 2:4] > min_wh) & (x[:, 2:4] < max_wh)).all(1)     (utils.py:799-802)
        my_ok = h4 > conf_thres && h2 > kMinWH && h2 < kMaxWH && h3 > kMinWH && h3 < kMaxWH;
        if (!EMIT && !my_ok) counts[my_row] = 0;
    }
    unsigned todo = __ballot_sync(~0u, my_ok);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const long long row = row0 + src;
        const float* x = pred + row * (long long)no;
        const float cx = __shfl_sync(~0u, h0, src), cy = __shfl_sync(~0u, h1, src), w = __shfl_sync(~0u, h2, src),
                    h = __shfl_sync(~0u, h3, src), obj = __shfl_sync(~0u, h4, src);
        // xywh2xyxy (utils.py:138-146)
        const float hw = __fmul_rn(w, 0.5f), hh = __fmul_rn(h, 0.5f);
        const float4 box = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
        const bool box_ok = is_finite(box.x) && is_finite(box.y) && is_finite(box.z) && is_finite(box.w);
        const unsigned long long img = (unsigned long long)(row / R);
        const int base = EMIT ? row_off[row] : 0;
        int n = 0;
        if (multi_label) {
            // (x[:, 5:] * obj > conf_thres).nonzero(): classes in ascending order within the row (utils.py:816-818)
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                float conf = 0.f;
                bool take = false;
                if (c < nc) {
                    conf = __fmul_rn(x[5 + c], obj);
                    take = conf > conf_thres && is_finite(conf) && box_ok && (allow == nullptr || allow[c] != 0);
                }
                const unsigned bal = __ballot_sync(~0u, take);
                if (EMIT && take) {
                    const int pos = base + n + __popc(bal & lanemask_lt());
                    cand_box[pos] = box;
                    cand_conf[pos] = conf;
                    cand_cls[pos] = (float)c;
                    keys[pos] = (img << 32) | desc_key(conf);
                    vals[pos] = (unsigned)pos;
                }
                n += __popc(bal);
            }
        } else {
            // conf, j = x[:, 5:].max(1): first index of the maximum, NaN propagates (and is dropped as non-finite)
            float best = -INFINITY;
            int arg = 0x7fffffff;
            bool nan = false;
            for (int c = lane; c < nc; c += 32) {
                const float conf = __fmul_rn(x[5 + c], obj);
                nan |= (conf != conf);
                if (conf > best) { best = conf; arg = c; }
            }
            nan = __any_sync(~0u, nan);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(~0u, best, o);
                const int oa = __shfl_xor_sync(~0u, arg, o);
                if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
            }
            const bool take = !nan && arg < nc && is_finite(best) && box_ok && (allow == nullptr || allow[arg] != 0);
            if (EMIT && take && lane == 0) {
                cand_box[base] = box;
                cand_conf[base] = best;
                cand_cls[base] = (float)arg;
                keys[base] = (img << 32) | desc_key(best);
                vals[base] = (unsigned)base;
            }
            n = take ? 1 : 0;
        }
        if (!EMIT && lane == 0) counts[row] = n;
    }
}

__global__ void nms_image_offsets_kernel(const int* __restrict__ row_off, int B, int R, int* __restrict__ img_off) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b <= B) img_off[b] = row_off[(long long)b * R];
}

// ---------------------------------------------------------------------------------------------------------------
// Greedy suppression, one CTA per image, identical to torchvision's loop over the score-sorted list:
// a candidate survives iff no EARLIER surviving candidate overlaps it by more than the threshold.
__global__ void __launch_bounds__(NT) nms_greedy_kernel(const int* __restrict__ img_off,
                                                        const unsigned* __restrict__ order,
                                                        const float4* __restrict__ cand_box,
                                                        const float* __restrict__ cand_cls, double iou_thres,
                                                        int agnostic, float4* __restrict__ kept_box,
                                                        unsigned* __restrict__ kept_idx, int* __restrict__ det_count) {
    __shared__ float4 s_box[NT];
    __shared__ float s_area[NT];
    __shared__ unsigned s_mask[NT][NW + 1];          // row i: chunk members j > i that i suppresses (+1: no conflicts)
    __shared__ unsigned char s_alive[NT], s_keep[NT];
    __shared__ int s_wsum[NW];
    __shared__ int s_kept;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int off = img_off[b], n = img_off[b + 1] - off;
    const bool nonneg = iou_thres >= 0.0;
    if (tid == 0) s_kept = 0;
    __syncthreads();

    for (int base = 0; base < n; base += NT) {
        const int cnt = min(NT, n - base);
        const bool valid = tid < cnt;
        unsigned idx = 0;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        float area = 0.f;
        if (valid) {
            idx = order[off + base + tid];
            box = shift_box(cand_box[idx], cand_cls[idx], agnostic);
            area = box_area(box);
        }
        bool alive = valid;
        // ---- phase 1: against everything kept in earlier chunks ----
        const int kept = s_kept;
        for (int kt = 0; kt < kept; kt += NT) {
            const int kn = min(NT, kept - kt);
            __syncthreads();                                  // previous tile fully consumed
            if (tid < kn) {
                const float4 kb = kept_box[off + kt + tid];
                s_box[tid] = kb;
                s_area[tid] = box_area(kb);
            }
            __syncthreads();
            if (alive) {
                for (int q = 0; q < kn; ++q) {
                    // torchvision: iarea (the kept box) + areas[j] - inter; `ovr > iou_threshold` with a double threshold
                    if (iou_gt_d(s_box[q], s_area[q], box, area, iou_thres, nonneg)) { alive = false; break; }
                }
            }
        }
        __syncthreads();
        // ---- phase 2: among the members of this chunk ----
        s_box[tid] = box;
        s_area[tid] = area;
        s_alive[tid] = alive ? 1 : 0;
        __syncthreads();
        for (int wd = 0; wd < NW; ++wd) {
            unsigned word = 0;
            if (alive && wd >= warp) {
                const int j0 = wd * 32;
#pragma unroll 4
                for (int bit = 0; bit < 32; ++bit) {
                    const int j = j0 + bit;
                    if (j > tid && j < cnt && s_alive[j] && iou_gt_d(box, area, s_box[j], s_area[j], iou_thres, nonneg))
                        word |= 1u << bit;
                }
            }
            s_mask[tid][wd] = word;
        }
        __syncthreads();
        if (warp == 0) {
            unsigned removed = 0;                             // lane l < NW owns bits [32 l, 32 l + 32)
            for (int i = 0; i < cnt; ++i) {
                const unsigned word = __shfl_sync(~0u, removed, i >> 5);
                const bool keep = s_alive[i] && !((word >> (i & 31)) & 1u);
                if (keep && lane < NW) removed |= s_mask[i][lane];
                if (lane == 0) s_keep[i] = keep ? 1 : 0;
            }
        }
        __syncthreads();
        // ---- append the survivors to the kept list in score order ----
        const bool k = valid && s_keep[tid];
        const unsigned bal = __ballot_sync(~0u, k);
        if (lane == 0) s_wsum[warp] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int wq = 0; wq < NW; ++wq) {
            const int c = s_wsum[wq];
            before += wq < warp ? c : 0;
            total += c;
        }
        if (k) {
            const int pos = off + kept + before + __popc(bal & lanemask_lt());
            kept_box[pos] = box;
            kept_idx[pos] = idx;
        }
        __syncthreads();                                      // kept_box visible to the CTA, s_kept read by everyone
        if (tid == 0) s_kept = kept + total;
        __syncthreads();
    }
    if (tid == 0) det_count[b] = s_kept;
}

// ---------------------------------------------------------------------------------------------------------------
// Output rows; merge-NMS (utils.py:843-850): x[i, :4] = (w @ x[:, :4]) / w.sum(1), w = (iou(kept, all) > thres) * scores
__global__ void nms_finish_kernel(const int* __restrict__ img_off, const int* __restrict__ det_count,
                                  const float4* __restrict__ kept_box, const unsigned* __restrict__ kept_idx,
                                  const float4* __restrict__ cand_box, const float* __restrict__ cand_conf,
                                  const float* __restrict__ cand_cls, float iou_thres, int agnostic,
                                  float* __restrict__ det) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int off = img_off[b], n = img_off[b + 1] - off, kept = det_count[b];
    const int warps = gridDim.x * (blockDim.x >> 5);
    for (int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); k < kept; k += warps) {
        const unsigned idx = kept_idx[off + k];
        float4 out = cand_box[idx];
        if (n > 1 && n < kMergeMax) {
            const float4 kb = kept_box[off + k];
            const float karea = box_area(kb);
            double sx1 = 0, sy1 = 0, sx2 = 0, sy2 = 0, sw = 0;
            for (int j = lane; j < n; j += 32) {
                const float4 raw = cand_box[off + j];
                const float4 sb = shift_box(raw, cand_cls[off + j], agnostic);
                if (iou_gt_f(kb, karea, sb, box_area(sb), iou_thres, iou_thres >= 0.f)) {
                    const double wgt = (double)cand_conf[off + j];
                    sx1 += wgt * raw.x; sy1 += wgt * raw.y; sx2 += wgt * raw.z; sy2 += wgt * raw.w;
                    sw += wgt;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                sx1 += __shfl_xor_sync(~0u, sx1, o); sy1 += __shfl_xor_sync(~0u, sy1, o);
                sx2 += __shfl_xor_sync(~0u, sx2, o); sy2 += __shfl_xor_sync(~0u, sy2, o);
                sw += __shfl_xor_sync(~0u, sw, o);
            }
            const float fw = (float)sw;
            out = make_float4(__fdiv_rn((float)sx1, fw), __fdiv_rn((float)sy1, fw), __fdiv_rn((float)sx2, fw),
                              __fdiv_rn((float)sy2, fw));
        }
        if (lane == 0) {
            float* r = det + (long long)(off + k) * 6;
            r[0] = out.x; r[1] = out.y; r[2] = out.z; r[3] = out.w;
            r[4] = cand_conf[idx];
            r[5] = cand_cls[idx];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// test.py:137 (clip_coords) + :150-170: one CTA per image.
__global__ void tp_match_kernel(float* __restrict__ det, const int* __restrict__ det_off,
                                const int* __restrict__ det_count, const float* __restrict__ tcls,
                                const float* __restrict__ tbox, const int* __restrict__ lab_off,
                                const float* __restrict__ iouv, int niou, float clip_w, float clip_h,
                                int* __restrict__ claim, int* __restrict__ best_t, float* __restrict__ best_iou,
                                unsigned char* __restrict__ correct) {
    const int b = blockIdx.x;
    const int d0 = det_off[b], nd = det_count[b], l0 = lab_off[b], nl = lab_off[b + 1] - l0;
    const float thr0 = iouv[0];
    for (int j = threadIdx.x; j < nd; j += blockDim.x) {
        float* r = det + (long long)(d0 + j) * 6;
        float4 pb = make_float4(r[0], r[1], r[2], r[3]);
        if (clip_w > 0.f) {                                   // boxes[:, 0].clamp_(0, w) ...
            pb.x = fminf(fmaxf(pb.x, 0.f), clip_w); pb.y = fminf(fmaxf(pb.y, 0.f), clip_h);
            pb.z = fminf(fmaxf(pb.z, 0.f), clip_w); pb.w = fminf(fmaxf(pb.w, 0.f), clip_h);
            r[0] = pb.x; r[1] = pb.y; r[2] = pb.z; r[3] = pb.w;
        }
        const float pc = r[5], parea = box_area(pb);
        float best = -INFINITY;
        int arg = -1;
        for (int t = 0; t < nl; ++t) {                        // box_iou(pred[pi], tbox[ti]).max(1): first maximum
            if (tcls[l0 + t] != pc) continue;
            const float4 tb = *reinterpret_cast<const float4*>(tbox + (long long)(l0 + t) * 4);
            const float iou = box_iou(pb, parea, tb, box_area(tb));
            if (iou > best || arg < 0) { best = iou; arg = t; }
        }
        best_t[d0 + j] = arg;
        best_iou[d0 + j] = best;
        if (arg >= 0 && best > thr0) atomicMin(&claim[l0 + arg], j);   // first prediction (score order) wins the target
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nd; j += blockDim.x) {
        const int arg = best_t[d0 + j];
        const float iou = best_iou[d0 + j];
        const bool won = arg >= 0 && iou > thr0 && claim[l0 + arg] == j;
        unsigned char* c = correct + (long long)(d0 + j) * niou;
        for (int q = 0; q < niou; ++q) c[q] = (won && iou > iouv[q]) ? 1 : 0;
    }
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct RunLayout {
    size_t cand_box, cand_conf, cand_cls, keys_in, keys_out, vals_in, vals_out, kept_box, kept_idx, cub, total;
};

size_t sort_temp_bytes(long long n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                    (const unsigned*)nullptr, (unsigned*)nullptr, (int)n, 0, 64);
    return bytes;
}

RunLayout run_layout(long long n) {
    RunLayout L;
    size_t o = 0;
    const size_t N = (size_t)(n > 0 ? n : 1);
    L.cand_box = o; o = align256(o + N * sizeof(float4));
    L.cand_conf = o; o = align256(o + N * sizeof(float));
    L.cand_cls = o; o = align256(o + N * sizeof(float));
    L.keys_in = o; o = align256(o + N * sizeof(unsigned long long));
    L.keys_out = o; o = align256(o + N * sizeof(unsigned long long));
    L.vals_in = o; o = align256(o + N * sizeof(unsigned));
    L.vals_out = o; o = align256(o + N * sizeof(unsigned));
    L.kept_box = o; o = align256(o + N * sizeof(float4));
    L.kept_idx = o; o = align256(o + N * sizeof(unsigned));
    L.cub = o; o = align256(o + sort_temp_bytes((long long)N));
    L.total = o;
    return L;
}

}  // namespace

extern "C" size_t b2y_nms_count_workspace_bytes(int batch, int rows) {
    size_t bytes = 0;
    const long long n = (long long)batch * rows + 1;
    if (batch <= 0 || rows <= 0 || n > 0x7fffffffLL) return 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (int)n);
    return align256(bytes);
}

extern "C" int b2y_nms_count(const float* pred, int batch, int rows, int nc, float conf_thres, int multi_label,
                             const unsigned char* class_allow, int* row_off, int* img_off, void* workspace,
                             size_t workspace_bytes, void* stream) {
    if (!pred || !row_off || !img_off || !workspace || batch <= 0 || rows <= 0 || nc < 1) return B2Y_ERR_INVALID;
    const long long total_rows = (long long)batch * rows;
    if (total_rows + 1 > 0x7fffffffLL) return B2Y_ERR_UNSUPPORTED;
    if (workspace_bytes < b2y_nms_count_workspace_bytes(batch, rows)) return B2Y_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int multi = (multi_label && nc > 1) ? 1 : 0;              // multi_label &= nc > 1 (utils.py:794)
    const int wpb = 4;                                              // 4 warps x 32 rows per CTA
    const unsigned grid = (unsigned)((total_rows + wpb * 32 - 1) / (wpb * 32));
    nms_rows_kernel<false><<<grid, wpb * 32, 0, st>>>(pred, total_rows, rows, nc, conf_thres, multi, class_allow,
                                                       row_off, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    B2Y_CUDA_CHECK(cudaGetLastError());
    B2Y_CUDA_CHECK(cudaMemsetAsync(row_off + total_rows, 0, sizeof(int), st));
    size_t bytes = workspace_bytes;
    B2Y_CUDA_CHECK(cub::DeviceScan::ExclusiveSum(workspace, bytes, row_off, row_off, (int)(total_rows + 1), st));
    nms_image_offsets_kernel<<<(batch + 1 + 127) / 128, 128, 0, st>>>(row_off, batch, rows, img_off);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" size_t b2y_nms_run_workspace_bytes(long long total) {
    if (total < 0 || total > 0x7fffffffLL) return 0;
    return run_layout(total).total;
}

extern "C" int b2y_nms_run(const float* pred, int batch, int rows, int nc, float conf_thres, double iou_thres,
                           int multi_label, int agnostic, const unsigned char* class_allow, const int* row_off,
                           const int* img_off, long long total, void* workspace, size_t workspace_bytes, float* det,
                           int* det_count, void* stream) {
    if (!pred || !row_off || !img_off || !det_count || batch <= 0 || rows <= 0 || nc < 1 || total < 0)
        return B2Y_ERR_INVALID;
    if (total > 0x7fffffffLL || batch > 0xffff) return B2Y_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (total == 0) {
        B2Y_CUDA_CHECK(cudaMemsetAsync(det_count, 0, sizeof(int) * batch, st));
        return B2Y_OK;
    }
    if (!workspace || !det) return B2Y_ERR_INVALID;
    const RunLayout L = run_layout(total);
    if (workspace_bytes < L.total) return B2Y_ERR_INVALID;
    char* ws = static_cast<char*>(workspace);
    float4* cand_box = reinterpret_cast<float4*>(ws + L.cand_box);
    float* cand_conf = reinterpret_cast<float*>(ws + L.cand_conf);
    float* cand_cls = reinterpret_cast<float*>(ws + L.cand_cls);
    unsigned long long* keys_in = reinterpret_cast<unsigned long long*>(ws + L.keys_in);
    unsigned long long* keys_out = reinterpret_cast<unsigned long long*>(ws + L.keys_out);
    unsigned* vals_in = reinterpret_cast<unsigned*>(ws + L.vals_in);
    unsigned* vals_out = reinterpret_cast<unsigned*>(ws + L.vals_out);
    float4* kept_box = reinterpret_cast<float4*>(ws + L.kept_box);
    unsigned* kept_idx = reinterpret_cast<unsigned*>(ws + L.kept_idx);

    const long long total_rows = (long long)batch * rows;
    const int multi = (multi_label && nc > 1) ? 1 : 0;
    const int wpb = 4;
    const unsigned grid = (unsigned)((total_rows + wpb * 32 - 1) / (wpb * 32));
    nms_rows_kernel<true><<<grid, wpb * 32, 0, st>>>(pred, total_rows, rows, nc, conf_thres, multi, class_allow, nullptr,
                                                      row_off, cand_box, cand_conf, cand_cls, keys_in, vals_in);
    B2Y_CUDA_CHECK(cudaGetLastError());
    int img_bits = 1;
    while ((1 << img_bits) < batch) ++img_bits;
    size_t bytes = L.total - L.cub;
    B2Y_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.cub, bytes, keys_in, keys_out, vals_in, vals_out, (int)total, 0,
                                                   32 + img_bits, st));
    nms_greedy_kernel<<<batch, NT, 0, st>>>(img_off, vals_out, cand_box, cand_cls, iou_thres, agnostic, kept_box, kept_idx,
                                            det_count);
    B2Y_CUDA_CHECK(cudaGetLastError());
    const dim3 fgrid(64, batch);
    nms_finish_kernel<<<fgrid, 256, 0, st>>>(img_off, det_count, kept_box, kept_idx, cand_box, cand_conf, cand_cls,
                                             (float)iou_thres, agnostic, det);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}

extern "C" size_t b2y_tp_match_workspace_bytes(long long n_det, long long n_lab) {
    if (n_det < 0 || n_lab < 0) return 0;
    return align256((size_t)(n_lab + 1) * sizeof(int)) + align256((size_t)(n_det + 1) * sizeof(int)) +
           align256((size_t)(n_det + 1) * sizeof(float));
}

extern "C" int b2y_tp_match(float* det, const int* det_off, const int* det_count, long long n_det, const float* tcls,
                            const float* tbox, const int* lab_off, long long n_lab, const float* iouv, int niou,
                            int batch, float clip_w, float clip_h, void* workspace, size_t workspace_bytes,
                            unsigned char* correct, void* stream) {
    if (!det_off || !det_count || !lab_off || !iouv || !workspace || batch <= 0 || niou <= 0 || n_det < 0 || n_lab < 0)
        return B2Y_ERR_INVALID;
    if (n_det > 0 && (!det || !correct)) return B2Y_ERR_INVALID;
    if (n_lab > 0 && (!tcls || !tbox)) return B2Y_ERR_INVALID;
    if (n_det > 0x7fffffffLL || n_lab > 0x7fffffffLL) return B2Y_ERR_UNSUPPORTED;
    if (workspace_bytes < b2y_tp_match_workspace_bytes(n_det, n_lab)) return B2Y_ERR_INVALID;
    if (n_lab > 0 && (reinterpret_cast<uintptr_t>(tbox) & 15) != 0) return B2Y_ERR_INVALID;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    char* ws = static_cast<char*>(workspace);
    int* claim = reinterpret_cast<int*>(ws);
    int* best_t = reinterpret_cast<int*>(ws + align256((size_t)(n_lab + 1) * sizeof(int)));
    float* best_iou = reinterpret_cast<float*>(ws + align256((size_t)(n_lab + 1) * sizeof(int)) +
                                               align256((size_t)(n_det + 1) * sizeof(int)));
    B2Y_CUDA_CHECK(cudaMemsetAsync(claim, 0x7f, (size_t)(n_lab + 1) * sizeof(int), st));
    tp_match_kernel<<<batch, 256, 0, st>>>(det, det_off, det_count, tcls, tbox, lab_off, iouv, niou, clip_w, clip_h, claim,
                                           best_t, best_iou, correct);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}
