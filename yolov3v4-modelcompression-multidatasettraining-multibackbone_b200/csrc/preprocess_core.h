// Per-pixel arithmetic of the device letterbox (csrc/preprocess.cu), shared with the host harness of the CPU tests
// (tests/host/letterbox_host.cpp) so that the very same code is checked against cv2's outputs without a GPU.
//
// OpenCV 4.x modules/imgproc/src/resize.cpp, 8-bit INTER_LINEAR (fixed point, INTER_RESIZE_COEF_BITS = 11):
//   fx = (float)((dx + 0.5) * scale_x - 0.5), sx = floor(fx), fx -= sx; columns clamp (sx < 0 -> fx = 0, sx = 0;
//   sx >= w - 1 -> fx = 0, sx = w - 1); rows clamp the ROW INDICES only, not the weights;
//   coefficients = cvRound(c * 2048) (int16); horizontal pass S[sx] * a0 + S[sx + 1] * a1 (int32);
//   vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
#ifndef B2Y_PREPROCESS_CORE_H_
#define B2Y_PREPROCESS_CORE_H_

#if defined(__CUDACC__)
#define B2Y_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define B2Y_HD inline
#endif

struct b2y_lb_params {
    int src_h, src_w, channels;       // source image, HWC uint8
    long long src_pitch;              // bytes per source row
    int rs_h, rs_w;                   // size after cv2.resize (== source size: plain copy)
    int top, left;                    // border
    int dst_h, dst_w;                 // output plane size
    int swap_rb;                      // 1: output channel c = source channel (channels - 1 - c)   (BGR -> RGB)
    unsigned char color;
    double scale_x, scale_y;          // resize.cpp: scale_x = 1. / ((double)dsize.width / ssize.width)
};

// OpenCV's scalar set-up code runs without FMA: (dx + 0.5) * scale - 0.5 rounds twice.  Device code must not contract it.
B2Y_HD float b2y_lb_src_coord(int d, double scale) {
#if defined(__CUDA_ARCH__)
    return (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
#else
    volatile double t = ((double)d + 0.5) * scale;      // volatile: no contraction on the host either
    return (float)(t - 0.5);
#endif
}

B2Y_HD int b2y_lb_round(float v) {                       // saturate_cast<short>(float) = cvRound: nearest, ties to even
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)std::nearbyint(v);
#endif
}

B2Y_HD void b2y_lb_coeff(int d, double scale, int src, bool clamp, int& s, int& a0, int& a1) {
    float f = b2y_lb_src_coord(d, scale);
    int si = (int)floorf(f);
#if defined(__CUDA_ARCH__)
    f = __fsub_rn(f, (float)si);
    const float g = __fsub_rn(1.f, f);
#else
    f = f - (float)si;
    const float g = 1.f - f;
#endif
    float f1 = f, f0 = g;
    if (clamp) {
        if (si < 0) { f1 = 0.f; f0 = 1.f; si = 0; }
        if (si >= src - 1) { f1 = 0.f; f0 = 1.f; si = src - 1; }
    }
#if defined(__CUDA_ARCH__)
    a0 = b2y_lb_round(__fmul_rn(f0, 2048.f));
    a1 = b2y_lb_round(__fmul_rn(f1, 2048.f));
#else
    a0 = b2y_lb_round(f0 * 2048.f);
    a1 = b2y_lb_round(f1 * 2048.f);
#endif
    s = si;
}

// One output pixel (all channels): dst is planar [channels][dst_h][dst_w].
B2Y_HD void b2y_lb_pixel(const unsigned char* src, unsigned char* dst, const b2y_lb_params& p, int x, int y) {
    const int C = p.channels;
    const int rx = x - p.left, ry = y - p.top;
    const long long plane = (long long)p.dst_h * p.dst_w;
    unsigned char* out = dst + (long long)y * p.dst_w + x;
    if (rx < 0 || ry < 0 || rx >= p.rs_w || ry >= p.rs_h) {                 // copyMakeBorder(..., value = color)
        for (int c = 0; c < C; ++c) out[c * plane] = p.color;
        return;
    }
    if (p.rs_w == p.src_w && p.rs_h == p.src_h) {                            // shape[::-1] == new_unpad: no resize
        const unsigned char* s = src + (long long)ry * p.src_pitch + (long long)rx * C;
        for (int c = 0; c < C; ++c) out[c * plane] = s[p.swap_rb ? C - 1 - c : c];
        return;
    }
    int sx, a0, a1, sy, b0, b1;
    b2y_lb_coeff(rx, p.scale_x, p.src_w, true, sx, a0, a1);
    b2y_lb_coeff(ry, p.scale_y, p.src_h, false, sy, b0, b1);
    const int x1 = sx + 1 < p.src_w ? sx + 1 : p.src_w - 1;
    int y0 = sy < 0 ? 0 : sy, y1 = sy + 1 < 0 ? 0 : sy + 1;
    y0 = y0 < p.src_h ? y0 : p.src_h - 1;
    y1 = y1 < p.src_h ? y1 : p.src_h - 1;
    const unsigned char* r0 = src + (long long)y0 * p.src_pitch;
    const unsigned char* r1 = src + (long long)y1 * p.src_pitch;
    for (int c = 0; c < C; ++c) {
        const int sc = p.swap_rb ? C - 1 - c : c;
        const int h0 = (int)r0[sx * C + sc] * a0 + (int)r0[x1 * C + sc] * a1;
        const int h1 = (int)r1[sx * C + sc] * a0 + (int)r1[x1 * C + sc] * a1;
        int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        out[c * plane] = (unsigned char)v;
    }
}

B2Y_HD void b2y_lb_set_scales(b2y_lb_params& p) {
    p.scale_x = 1.0 / ((double)p.rs_w / (double)p.src_w);
    p.scale_y = 1.0 / ((double)p.rs_h / (double)p.src_h);
}

#endif  // B2Y_PREPROCESS_CORE_H_
