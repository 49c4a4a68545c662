// Host harness of the CPU tests: runs csrc/preprocess_core.h (the device letterbox's per-pixel code) over a whole image.
#include "preprocess_core.h"

extern "C" int b2y_letterbox_u8_host(const unsigned char* src, int src_h, int src_w, int channels, long long src_pitch,
                                     int resized_h, int resized_w, int top, int left, unsigned char* dst, int dst_h,
                                     int dst_w, int swap_rb, int color) {
    b2y_lb_params p;
    p.src_h = src_h; p.src_w = src_w; p.channels = channels; p.src_pitch = src_pitch;
    p.rs_h = resized_h; p.rs_w = resized_w; p.top = top; p.left = left; p.dst_h = dst_h; p.dst_w = dst_w;
    p.swap_rb = swap_rb ? 1 : 0; p.color = (unsigned char)color;
    b2y_lb_set_scales(p);
    for (int y = 0; y < dst_h; ++y)
        for (int x = 0; x < dst_w; ++x) b2y_lb_pixel(src, dst, p, x, y);
    return 0;
}
