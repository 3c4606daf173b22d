// vis_heatmap.cu — device half of create_heatmap_overlay (utils/image_utils.py:320-604; SURVEY.md 8f "next" row 2).
//
// Per frame (one stream-ordered sequence, no synchronisation, caller-owned scratch of 3 float planes):
//   k_heat_clear      heat = 0
//   per defect        k_heat_local: analytic heat of the defect on its region (float64 exp, boosts, 4-sigma cut-off)
//                     -> tmp_a; k_heat_blur_h / k_heat_blur_v: separable Gaussian (float32, BORDER_REFLECT_101 inside
//                     the REGION, as cv2.GaussianBlur on the sliced array does); the vertical pass max-combines into
//                     heat.  Widespread defects max-combine their Gaussian directly.
//   final blur        k_heat_blur_h / k_heat_blur_v over the whole mask (reflect at the image border)
//   k_heat_max        global maximum (non-negative floats: atomicMax on the bit pattern)
//   k_heat_colorize   idx = uint8(heat / max * 255) (truncation), JET colour, saturate(round(0.6*img + 0.4*colour))
// Floating point: cv2's separable filter accumulates in float32 in a SIMD-dependent order, so this path is specified
// with a tolerance (tests: <= 2 levels on isolated pixels where the 8-bit heat index flips), not bit-exactness.
// Bound: HBM (a few passes over H*W floats + the frame).
#include "vis_internal.h"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__global__ void k_heat_clear(float* __restrict__ p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// analytic heat of one defect on its region (row-major region buffer `out`, or max into `heat` for kind 1 / ksize 1)
__global__ void __launch_bounds__(kT)
k_heat_local(VisHeatDefect d, int img_w, float* __restrict__ out, float* __restrict__ heat, int direct) {
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rw * rh) return;
    const int ly = i / rw, lx = i - ly * rw;
    const int gx = d.x1 + lx, gy = d.y1 + ly;
    const double ddx = (double)gx - d.cx, ddy = (double)gy - d.cy;
    const double dist_sq = ddx * ddx + ddy * ddy;
    float v;
    if (d.kind == 1) {
        v = (float)(d.intensity * exp(-dist_sq / (2.0 * (d.sigma * d.sigma))));
    } else {
        double g = d.intensity * exp(-dist_sq / (2.0 * (d.sigma * d.sigma)));
        const bool in_box = gx >= d.x && gx < d.x + d.w && gy >= d.y && gy < d.y + d.h;
        const double ex = ddx / fmax(d.w / 2.0, 1.0), ey = ddy / fmax(d.h / 2.0, 1.0);
        const double boost = (ex * ex + ey * ey < 1.2 * 1.2) ? 1.8 : (in_box ? 1.4 : 1.0);
        g = fmin(1.0, g * boost);
        const double lim = 4.0 * d.sigma;
        v = dist_sq < lim * lim ? (float)g : 0.f;
    }
    if (direct) {
        float* h = heat + (size_t)gy * img_w + gx;
        *h = fmaxf(*h, v);
    } else {
        out[i] = v;
    }
}

// separable Gaussian on a region of `rw` x `rh` floats with pitch `pitch` (elements); BORDER_REFLECT_101 inside the
// region.  Both passes stage their inputs in shared memory (each input element is read from global memory ~1.2x
// instead of ksize times) and accumulate taps in kernel order with fmaf.
constexpr int kBlurR = 25;                       // largest radius (ksize <= 51)
constexpr int kHSeg = 256;                       // outputs per block of the horizontal pass
constexpr int kVCols = 32, kVRows = 64;          // tile of the vertical pass

__global__ void __launch_bounds__(kT)
k_heat_blur_h(const float* __restrict__ src, int src_pitch, float* __restrict__ dst, int dst_pitch, int rw, int rh,
              const float* __restrict__ kern, int ksize) {
    __shared__ float row[kHSeg + 2 * kBlurR];
    __shared__ float kk[2 * kBlurR + 1];
    const int y = blockIdx.y, x0 = blockIdx.x * kHSeg, r = ksize >> 1;
    const float* in = src + (size_t)y * src_pitch;
    for (int i = threadIdx.x; i < kHSeg + 2 * r; i += kT) row[i] = in[reflect101(x0 + i - r, rw)];
    if (threadIdx.x < ksize) kk[threadIdx.x] = __ldg(kern + threadIdx.x);
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= rw) return;
    float s = 0.f;
    for (int k = 0; k < ksize; ++k) s = fmaf(kk[k], row[threadIdx.x + k], s);
    dst[(size_t)y * dst_pitch + x] = s;
}
// vertical pass; combine: 0 = store, 1 = max into dst
__global__ void __launch_bounds__(kT)
k_heat_blur_v(const float* __restrict__ src, int src_pitch, float* __restrict__ dst, int dst_pitch, int rw, int rh,
              const float* __restrict__ kern, int ksize, int combine) {
    __shared__ float tile[kVRows + 2 * kBlurR][kVCols];
    __shared__ float kk[2 * kBlurR + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8 threads
    const int x = blockIdx.x * kVCols + tx, y0 = blockIdx.y * kVRows, r = ksize >> 1;
    for (int i = ty; i < kVRows + 2 * r; i += kT / 32)
        tile[i][tx] = x < rw ? src[(size_t)reflect101(y0 + i - r, rh) * src_pitch + x] : 0.f;
    if (threadIdx.x < ksize) kk[threadIdx.x] = __ldg(kern + threadIdx.x);
    __syncthreads();
    if (x >= rw) return;
    for (int j = ty; j < kVRows; j += kT / 32) {
        const int y = y0 + j;
        if (y >= rh) break;
        float s = 0.f;
        for (int k = 0; k < ksize; ++k) s = fmaf(kk[k], tile[j + k][tx], s);
        float* o = dst + (size_t)y * dst_pitch + x;
        *o = combine ? fmaxf(*o, s) : s;
    }
}

__global__ void __launch_bounds__(kT)
k_heat_max(const float* __restrict__ p, size_t n, unsigned int* __restrict__ out) {
    float m = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));     // non-negative: bit order = value order
}

__global__ void __launch_bounds__(kT)
k_heat_colorize(const uint8_t* __restrict__ img, int64_t img_pitch, const float* __restrict__ heat, int w, int h,
                const unsigned int* __restrict__ max_bits, const uint8_t* __restrict__ jet, uint8_t* __restrict__ dst,
                int64_t dst_pitch) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= w * h) return;
    const int y = i / w, x = i - y * w;
    const float mx = __uint_as_float(*max_bits);
    const float v = heat[i];
    // numpy: (heat / max * 255).astype(uint8) in float32, truncation toward zero
    const float t = mx > 0.f ? __fmul_rn(__fdiv_rn(v, mx), 255.f) : __fmul_rn(v, 255.f);
    const int idx = min(max((int)t, 0), 255);
    const uint8_t* s = img + (size_t)y * img_pitch + (size_t)x * 3;
    uint8_t* o = dst + (size_t)y * dst_pitch + (size_t)x * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // cv2.addWeighted(img, 0.6, colour, 0.4, 0): float32, round half to even, saturate
        const float r = __fadd_rn(__fmul_rn((float)s[c], 0.6f), __fmul_rn((float)__ldg(jet + idx * 3 + c), 0.4f));
        o[c] = (uint8_t)min(max(__float2int_rn(r), 0), 255);
    }
}

inline int blocks_for(size_t n) { return (int)((n + kT - 1) / kT); }

}  // namespace

extern "C" int vis_heatmap_overlay(const uint8_t* img, int64_t img_pitch, int h, int w,
                                   const VisHeatDefect* defects, int n_defects, const float* kernels,
                                   int final_ksize, int final_koff, const uint8_t* jet768,
                                   float* scratch, uint8_t* dst, int64_t dst_pitch, void* stream) {
    if (!img || !dst || !scratch || !jet768 || !kernels || h <= 0 || w <= 0 || n_defects < 0 || (n_defects && !defects) ||
        img_pitch < (int64_t)w * 3 || dst_pitch < (int64_t)w * 3 || final_ksize < 1 || (final_ksize & 1) == 0) {
        vis::set_error("vis_heatmap_overlay: bad arguments");
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)w * h;
    float* heat = scratch;                 // h*w
    float* ta = scratch + n;               // h*w (region buffers live at its start)
    float* tb = scratch + 2 * n;           // h*w, then one uint for the maximum
    unsigned int* mx = reinterpret_cast<unsigned int*>(scratch + 3 * n);
    k_heat_clear<<<592, kT, 0, st>>>(heat, n);
    cudaMemsetAsync(mx, 0, sizeof(unsigned int), st);
    for (int i = 0; i < n_defects; ++i) {
        const VisHeatDefect& d = defects[i];
        const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
        if (rw <= 0 || rh <= 0 || d.x1 < 0 || d.y1 < 0 || d.x2 > w || d.y2 > h || d.ksize < 1 || d.ksize > 51 || (d.ksize & 1) == 0) {
            vis::set_error("vis_heatmap_overlay: defect %d has an invalid region or kernel", i);
            return VIS_E_INVALID;
        }
        const int nb = blocks_for((size_t)rw * rh);
        const int direct = d.kind == 1 || d.ksize == 1;
        k_heat_local<<<nb, kT, 0, st>>>(d, w, ta, heat, direct);
        if (!direct) {
            k_heat_blur_h<<<dim3((rw + kHSeg - 1) / kHSeg, rh), kT, 0, st>>>(ta, rw, tb, rw, rw, rh, kernels + d.koff, d.ksize);
            k_heat_blur_v<<<dim3((rw + kVCols - 1) / kVCols, (rh + kVRows - 1) / kVRows), kT, 0, st>>>(
                tb, rw, heat + (size_t)d.y1 * w + d.x1, w, rw, rh, kernels + d.koff, d.ksize, 1);
        }
    }
    const float* fin = heat;
    if (final_ksize > 1) {
        k_heat_blur_h<<<dim3((w + kHSeg - 1) / kHSeg, h), kT, 0, st>>>(heat, w, ta, w, w, h, kernels + final_koff, final_ksize);
        k_heat_blur_v<<<dim3((w + kVCols - 1) / kVCols, (h + kVRows - 1) / kVRows), kT, 0, st>>>(
            ta, w, tb, w, w, h, kernels + final_koff, final_ksize, 0);
        fin = tb;
    }
    k_heat_max<<<592, kT, 0, st>>>(fin, n, mx);
    k_heat_colorize<<<blocks_for(n), kT, 0, st>>>(img, img_pitch, fin, w, h, mx, jet768, dst, dst_pitch);
    return vis::check_launch("vis_heatmap_overlay");
}
