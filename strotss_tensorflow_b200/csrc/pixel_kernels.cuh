// Pixel-side step of the STROTSS iteration (SURVEY 8f "next #3"): the Laplacian-pyramid fold that turns the six
// optimisation variables into the image (nn/strotss_utils.py:159-163), its backward, the pyramid construction
// (:139-156) and the RMSprop update of all variables in one launch (run_strotss.py:63,148).
// Images are NHWC fp32 with batch 1, exactly like the reference's tensors: element (y, x, ch) at (y*w + x)*c + ch.
//
// tf.image.resize(..., method='bilinear') in TF2 = half-pixel centres, no antialiasing:
//   in = (out + 0.5) * (in_size / out_size) - 0.5 ; lo = max(floor(in), 0) ; hi = min(ceil(in), in_size - 1) ;
//   lerp = in - floor(in) ; value = top + (bottom - top) * y_lerp, top = tl + (tr - tl) * x_lerp
#pragma once
#include "common.cuh"

namespace sb {

struct Lerp { int lo, hi; float w; };

__device__ __forceinline__ Lerp lerp_of(int o, float scale, int in_size) {
    const float in = (static_cast<float>(o) + 0.5f) * scale - 0.5f;
    const float f = floorf(in);
    Lerp l;
    l.lo = max(static_cast<int>(f), 0);
    l.hi = min(static_cast<int>(ceilf(in)), in_size - 1);
    l.w = in - f;
    return l;
}

// out[y][x][:] = (add ? add[y][x][:] : 0) + sign * resize(src)[y][x][:]
// (fold level: add = pyramid variable, sign = +1;  make_laplacian: add = image, sign = -1;  plain resize: add = null)
__global__ void __launch_bounds__(256) resize_add_kernel(const float* __restrict__ src, int sh, int sw, int c,
                                                         const float* __restrict__ add, float sign,
                                                         float* __restrict__ out, int oh, int ow) {
    const long long total = static_cast<long long>(oh) * ow * c;
    const float ys = static_cast<float>(sh) / static_cast<float>(oh), xs = static_cast<float>(sw) / static_cast<float>(ow);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(e % c);
        const long long p = e / c;
        const int x = static_cast<int>(p % ow), y = static_cast<int>(p / ow);
        const Lerp ly = lerp_of(y, ys, sh), lx = lerp_of(x, xs, sw);
        const float tl = src[(static_cast<long long>(ly.lo) * sw + lx.lo) * c + ch];
        const float tr = src[(static_cast<long long>(ly.lo) * sw + lx.hi) * c + ch];
        const float bl = src[(static_cast<long long>(ly.hi) * sw + lx.lo) * c + ch];
        const float br = src[(static_cast<long long>(ly.hi) * sw + lx.hi) * c + ch];
        const float top = tl + (tr - tl) * lx.w;
        const float bot = bl + (br - bl) * lx.w;
        const float v = top + (bot - top) * ly.w;
        out[e] = (add ? add[e] : 0.f) + sign * v;
    }
}

// Transpose of the resize: gsrc[j][i][:] = sum over output pixels (y, x) of weight(y -> j) * weight(x -> i) * gout[y][x][:]
// as a GATHER (deterministic, no atomics): every source pixel scans the few output rows / columns that can reference it.
// `accumulate` adds to gsrc instead of overwriting it.
__device__ __forceinline__ void tap_range(int j, float scale, int out_size, int* a, int* b) {
    // outputs o with floor(in_o) in {j-1, j} (or clamped onto j at the borders): in_o in (j - 1, j + 1)
    const float inv = 1.f / scale;
    int lo = static_cast<int>(floorf((static_cast<float>(j) - 1.f + 0.5f) * inv - 0.5f)) - 1;
    int hi = static_cast<int>(ceilf((static_cast<float>(j) + 1.f + 0.5f) * inv - 0.5f)) + 1;
    *a = max(lo, 0);
    *b = min(hi, out_size - 1);
}

__global__ void __launch_bounds__(256) resize_transpose_kernel(const float* __restrict__ gout, int oh, int ow, int c,
                                                               float* __restrict__ gsrc, int sh, int sw, int accumulate) {
    const long long total = static_cast<long long>(sh) * sw * c;
    const float ys = static_cast<float>(sh) / static_cast<float>(oh), xs = static_cast<float>(sw) / static_cast<float>(ow);
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ch = static_cast<int>(e % c);
        const long long p = e / c;
        const int i = static_cast<int>(p % sw), j = static_cast<int>(p / sw);
        int y0, y1, x0, x1;
        // border source pixels also collect the clamped taps (lo = max(.,0), hi = min(., size-1))
        tap_range(j, ys, oh, &y0, &y1);
        tap_range(i, xs, ow, &x0, &x1);
        if (j == 0) y0 = 0;
        if (j == sh - 1) y1 = oh - 1;
        if (i == 0) x0 = 0;
        if (i == sw - 1) x1 = ow - 1;
        float acc = 0.f;
        for (int y = y0; y <= y1; ++y) {
            const Lerp ly = lerp_of(y, ys, sh);
            const float wy = (ly.lo == j ? 1.f - ly.w : 0.f) + (ly.hi == j ? ly.w : 0.f);
            if (wy == 0.f) continue;
            for (int x = x0; x <= x1; ++x) {
                const Lerp lx = lerp_of(x, xs, sw);
                const float wx = (lx.lo == i ? 1.f - lx.w : 0.f) + (lx.hi == i ? lx.w : 0.f);
                if (wx != 0.f) acc = fmaf(wy * wx, gout[(static_cast<long long>(y) * ow + x) * c + ch], acc);
            }
        }
        gsrc[e] = accumulate ? gsrc[e] + acc : acc;
    }
}

// RMSprop (Keras, momentum 0, not centred) over up to kMaxVars tensors in one launch:
//   rms = rho * rms + (1 - rho) * g^2 ;  var -= lr * g / (sqrt(rms) + eps)
constexpr int kMaxVars = 8;
struct RmspropArgs {
    float* var[kMaxVars]; float* rms[kMaxVars]; const float* grad[kMaxVars];
    long long end[kMaxVars];          // exclusive prefix of element counts
    int nvars; float lr, rho, eps;
};

__global__ void __launch_bounds__(256) rmsprop_kernel(const RmspropArgs a) {
    const long long total = a.end[a.nvars - 1];
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        int k = 0;
        while (e >= a.end[k]) ++k;
        const long long off = e - (k ? a.end[k - 1] : 0);
        const float g = a.grad[k][off];
        const float r = a.rho * a.rms[k][off] + (1.f - a.rho) * (g * g);
        a.rms[k][off] = r;
        a.var[k][off] = a.var[k][off] - a.lr * g / (sqrtf(r) + a.eps);
    }
}

}  // namespace sb
