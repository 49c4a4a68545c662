// Input pipeline on the device, inference slice: the reference's letterbox (utils/datasets.py:611-646:
// cv2.resize INTER_LINEAR + cv2.copyMakeBorder with colour 114) fused with the BGR -> RGB / HWC -> CHW shuffle of
// LoadImages.__next__ (datasets.py:108-118), writing straight into the uint8 NCHW batch slot that the stem kernel
// consumes (it applies the "/ 256").  The arithmetic is OpenCV's (third-party dependency of the reference, not vendored;
// this container: 4.13.0) -- restated in csrc/preprocess_core.h / oracle/preprocess_oracle.py and pinned against cv2's
// own outputs; bit-exact.  HBM / gather bound: 4 source pixels read per output pixel, one byte written per channel.
#include "b200yolo.h"
#include "common.cuh"
#include "preprocess_core.h"

namespace {

__global__ void __launch_bounds__(256) letterbox_u8_kernel(const unsigned char* __restrict__ src,
                                                           unsigned char* __restrict__ dst, b2y_lb_params p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.dst_w) return;
    b2y_lb_pixel(src, dst, p, x, (int)blockIdx.y);
}

}  // namespace

extern "C" int b2y_letterbox_u8(const unsigned char* src, int src_h, int src_w, int channels, long long src_pitch,
                                int resized_h, int resized_w, int top, int left, unsigned char* dst, int dst_h, int dst_w,
                                int swap_rb, int color, void* stream) {
    if (!src || !dst || src_h <= 0 || src_w <= 0 || channels < 1 || channels > 4 || resized_h <= 0 || resized_w <= 0 ||
        dst_h <= 0 || dst_w <= 0 || top < 0 || left < 0 || src_pitch < (long long)src_w * channels || color < 0 ||
        color > 255)
        return B2Y_ERR_INVALID;
    if (top + resized_h > dst_h || left + resized_w > dst_w) return B2Y_ERR_INVALID;
    if (dst_h > 65535) return B2Y_ERR_UNSUPPORTED;
    b2y_lb_params p;
    p.src_h = src_h; p.src_w = src_w; p.channels = channels; p.src_pitch = src_pitch;
    p.rs_h = resized_h; p.rs_w = resized_w; p.top = top; p.left = left; p.dst_h = dst_h; p.dst_w = dst_w;
    p.swap_rb = swap_rb ? 1 : 0; p.color = (unsigned char)color;
    b2y_lb_set_scales(p);
    const dim3 grid((dst_w + 255) / 256, dst_h);
    letterbox_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, dst, p);
    B2Y_CUDA_CHECK(cudaGetLastError());
    return B2Y_OK;
}
