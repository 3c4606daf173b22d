// vis_heatmap.cu — device half of create_heatmap_overlay (utils/image_utils.py:320-604; SURVEY.md 8f "next" row 2).
//
// BATCH form (round 2): any number of frames and defects in SIX launches, whatever the batch size —
//   k_heat_tables     per defect, the two 1-D factors of its Gaussian, exp(-(x-cx)^2 / 2 sigma^2) and the same in y, in
//                     float64 like the reference's numpy expression (one exp per region column / row instead of one per
//                     region pixel: the 2-D value is their product, within 2 ulp(double) of exp of the sum, long before
//                     the float32 rounding that follows)
//   k_heat_defect_h   analytic heat of a defect (intensity * Gaussian, boosts inside the box, min(1, .), 4-sigma cut-off,
//                     all float64, cast to float32) evaluated straight into shared memory, reflected at the REGION
//                     border as cv2.GaussianBlur on the sliced array does, and blurred horizontally; widespread defects
//                     and 1-tap kernels max-combine into the heat plane directly
//   k_heat_defect_v   vertical blur of every defect region, max-combined into its frame's heat plane (atomicMax on the
//                     bit pattern: the values are non-negative floats)
//   k_heat_final_h/v  the whole-mask blur (reflect at the image border) + the frame's maximum
//   k_heat_colorize   idx = uint8(heat / max * 255) (truncation), JET colour, saturate(round(0.6*img + 0.4*colour))
// Both blur passes are register tiled: a thread produces 8 consecutive outputs from one sliding window (8 FMA per pair of
// shared-memory reads), the horizontal pass keeps its row de-interleaved by 8 so that lanes read consecutive words.
// Floating point: cv2's separable filter accumulates in float32 in a SIMD-dependent order, so this path is specified
// with a tolerance (+-1 on the 8-bit heat index = <= 2 output levels, tests/test_oracle_heatmap.py), not bit-exactness.
// Bound: fp32 FMA issue — the reference's defects reach the sigma cap, so every defect is a ~50-tap separable blur over
// a region of up to (8 sigma + 31)^2 pixels: ~640 M FMA per 1080p frame with 3-4 defects against ~140 MB of traffic.
#include "vis_internal.h"

namespace {

constexpr int kT = 256;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

constexpr int kBlurR = 25;                       // largest radius (ksize <= 51)
constexpr int kOut = 8;                          // outputs per thread of both passes
constexpr int kHSeg = 512, kHRows = 4;           // horizontal pass: 4 rows x 512 outputs per block (64 threads per row)
constexpr int kHLen = kHSeg + 2 * kBlurR + 6;    // staged row, padded to a multiple of 8 (568)
constexpr int kHPhase = kHLen / kOut;            // 71: element i lives at (i % 8) * 71 + i / 8 -> lanes read consecutive words
constexpr int kVCols = 32, kVRows = 64;          // vertical pass: 32 columns x 64 output rows per block
constexpr int kVTile = kVRows + 56;              // rows staged: the last thread's rounded window ends at row 56 + 63
constexpr int kWPad = kOut - 1;                  // zero weights in front of / behind the kernel: no tap predicates
static_assert(kHLen % kOut == 0, "staged row must de-interleave evenly");

constexpr int kWLen = kWPad + 64 + 8;           // padded weights: kWPad zeros, the kernel, zeros up to the rounded window
__device__ __forceinline__ void load_weights(float* wp, const float* __restrict__ kern, int ksize) {
    for (int i = threadIdx.x; i < kWLen; i += kT) {
        const int t = i - kWPad;
        wp[i] = (t >= 0 && t < ksize) ? __ldg(kern + t) : 0.f;
    }
}

// analytic heat of a box defect / widespread defect at region-local (lx, ly), float64 like the reference, from the 1-D tables
__device__ __forceinline__ float heat_value(const VisHeatDefect& d, const double* __restrict__ tx, const double* __restrict__ ty,
                                            int lx, int ly) {
    const int gx = d.x1 + lx, gy = d.y1 + ly;
    const double g0 = d.intensity * (tx[lx] * ty[ly]);
    if (d.kind == 1) return (float)g0;
    const double ddx = (double)gx - d.cx, ddy = (double)gy - d.cy;
    const bool in_box = gx >= d.x && gx < d.x + d.w && gy >= d.y && gy < d.y + d.h;
    const double ex = ddx / fmax(d.w / 2.0, 1.0), ey = ddy / fmax(d.h / 2.0, 1.0);
    const double boost = (ex * ex + ey * ey < 1.2 * 1.2) ? 1.8 : (in_box ? 1.4 : 1.0);
    const double g = fmin(1.0, g0 * boost);
    const double lim = 4.0 * d.sigma;
    return (ddx * ddx + ddy * ddy) < lim * lim ? (float)g : 0.f;
}

__global__ void __launch_bounds__(kT) k_heat_tables(const VisHeatItem* __restrict__ items, double* __restrict__ tabs) {
    const VisHeatItem& it = items[blockIdx.y];
    const int rw = it.d.x2 - it.d.x1, rh = it.d.y2 - it.d.y1;
    const int t = blockIdx.x * kT + threadIdx.x;
    if (t >= rw + rh) return;
    const double c = t < rw ? (double)(it.d.x1 + t) - it.d.cx : (double)(it.d.y1 + t - rw) - it.d.cy;
    tabs[it.tab_off + t] = exp(-(c * c) / (2.0 * (it.d.sigma * it.d.sigma)));
}

// 8 outputs from one sliding window: value(k) for k = 0 .. n-1 with n = 8 + 2r rounded up to 8 (the surplus meets zero
// weights and zero-filled staging).  Weights sit in a 15-register window that advances 8 taps per block of 8 values, so
// a block is 8 value reads + 2 vector weight reads for 64 FMA; output j accumulates its taps in ascending order.
template <typename F>
__device__ __forceinline__ void window8(float (&acc)[kOut], const float* wp, int r, F value) {
#pragma unroll
    for (int j = 0; j < kOut; ++j) acc[j] = 0.f;
    float w[2 * kOut - 1];
#pragma unroll
    for (int j = 0; j < kOut - 1; ++j) w[j] = wp[j];              // taps -7 .. -1 of the padded kernel (zeros)
    const int n = (kOut + 2 * r + kOut - 1) & ~(kOut - 1);
#pragma unroll 1
    for (int kb = 0; kb < n; kb += kOut) {
#pragma unroll
        for (int q = 0; q < kOut; ++q) w[kOut - 1 + q] = wp[kWPad + kb + q];
        // w[i] = weight of tap (kb + i - 7): output j at value kb + kk uses tap kb + kk - j -> w[7 + kk - j]
#pragma unroll
        for (int kk = 0; kk < kOut; ++kk) {
            const float v = value(kb + kk);
#pragma unroll
            for (int j = 0; j < kOut; ++j) acc[j] = fmaf(w[kOut - 1 + kk - j], v, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < kOut - 1; ++j) w[j] = w[j + kOut];
    }
}

__device__ __forceinline__ void atomic_max_f(float* p, float v) {      // non-negative floats: bit order = value order
    atomicMax(reinterpret_cast<unsigned int*>(p), __float_as_uint(fmaxf(v, 0.f)));
}

__global__ void __launch_bounds__(kT)
k_heat_defect_h(const VisHeatItem* __restrict__ items, const VisHeatFrame* __restrict__ frames, const double* __restrict__ tabs,
                const float* __restrict__ kernels, float* __restrict__ tmp, float* __restrict__ heat) {
    __shared__ float rows[kHRows][kHLen];
    __shared__ float wp[kWLen];
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHRows;
    if (x0 >= rw || y0 >= rh) return;
    const double* tx = tabs + it.tab_off;
    const double* ty = tx + rw;
    const VisHeatFrame& fr = frames[it.frame];
    const bool direct = d.kind == 1 || d.ksize == 1;
    if (direct) {                                      // no blur: max-combine the analytic heat itself
        float* plane = heat + fr.plane_off;
        for (int i = threadIdx.x; i < kHRows * kHSeg; i += kT) {
            const int ly = y0 + i / kHSeg, lx = x0 + i % kHSeg;
            if (ly < rh && lx < rw) atomic_max_f(plane + (size_t)(d.y1 + ly) * fr.w + d.x1 + lx, heat_value(d, tx, ty, lx, ly));
        }
        return;
    }
    const int r = d.ksize >> 1;
    load_weights(wp, kernels + d.koff, d.ksize);
    const int span = kHSeg + 2 * r;
    for (int i = threadIdx.x; i < kHRows * kHLen; i += kT) {             // the padding behind `span` is zero-filled
        const int row = i / kHLen, idx = i - row * kHLen;
        const int ly = y0 + row;
        float v = 0.f;
        if (ly < rh && idx < span) v = heat_value(d, tx, ty, reflect101(x0 + idx - r, rw), ly);
        rows[row][(idx % kOut) * kHPhase + idx / kOut] = v;
    }
    __syncthreads();
    const int row = threadIdx.x >> 6, c = threadIdx.x & 63;
    const int ly = y0 + row, lx = x0 + c * kOut;
    if (ly >= rh || lx >= rw) return;
    float acc[kOut];
    const float* rp = rows[row];
    window8(acc, wp, r, [&](int k) { const int i = c * kOut + k; return rp[(i % kOut) * kHPhase + i / kOut]; });
    float* o = tmp + it.tmp_off + (size_t)ly * rw + lx;
#pragma unroll
    for (int j = 0; j < kOut; ++j)
        if (lx + j < rw) o[j] = acc[j];
}

// vertical pass of a defect region: tmp -> max into the frame's heat plane
__global__ void __launch_bounds__(kT)
k_heat_defect_v(const VisHeatItem* __restrict__ items, const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels,
                const float* __restrict__ tmp, float* __restrict__ heat) {
    __shared__ float tile[kVTile][kVCols];
    __shared__ float wp[kWLen];
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    if (d.kind == 1 || d.ksize == 1) return;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * kVCols + lane, y0 = blockIdx.y * kVRows;
    if (blockIdx.x * kVCols >= rw || y0 >= rh) return;
    const int r = d.ksize >> 1;
    load_weights(wp, kernels + d.koff, d.ksize);
    const float* src = tmp + it.tmp_off;
    for (int i = grp; i < kVTile; i += kT / 32)                            // rows behind the window are zero-filled
        tile[i][lane] = (x < rw && i < kVRows + 2 * r) ? src[(size_t)reflect101(y0 + i - r, rh) * rw + x] : 0.f;
    __syncthreads();
    if (x >= rw) return;
    float acc[kOut];
    window8(acc, wp, r, [&](int k) { return tile[grp * kOut + k][lane]; });
    const VisHeatFrame& fr = frames[it.frame];
    float* plane = heat + fr.plane_off;
#pragma unroll
    for (int j = 0; j < kOut; ++j) {
        const int y = y0 + grp * kOut + j;
        if (y < rh) atomic_max_f(plane + (size_t)(d.y1 + y) * fr.w + d.x1 + x, acc[j]);
    }
}

// whole-mask blur, horizontal: heat -> fa (reflect at the image border)
__global__ void __launch_bounds__(kT)
k_heat_final_h(const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels, const float* __restrict__ heat,
               float* __restrict__ fa) {
    __shared__ float rows[kHRows][kHLen];
    __shared__ float wp[kWLen];
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHRows;
    if (x0 >= fr.w || y0 >= fr.h) return;
    const int r = fr.final_ksize >> 1;
    load_weights(wp, kernels + fr.final_koff, fr.final_ksize);
    const float* plane = heat + fr.plane_off;
    const int span = kHSeg + 2 * r;
    for (int i = threadIdx.x; i < kHRows * kHLen; i += kT) {
        const int row = i / kHLen, idx = i - row * kHLen;
        const int y = y0 + row;
        rows[row][(idx % kOut) * kHPhase + idx / kOut] =
            (y < fr.h && idx < span) ? plane[(size_t)y * fr.w + reflect101(x0 + idx - r, fr.w)] : 0.f;
    }
    __syncthreads();
    const int row = threadIdx.x >> 6, c = threadIdx.x & 63;
    const int y = y0 + row, x = x0 + c * kOut;
    if (y >= fr.h || x >= fr.w) return;
    float acc[kOut];
    const float* rp = rows[row];
    window8(acc, wp, r, [&](int k) { const int i = c * kOut + k; return rp[(i % kOut) * kHPhase + i / kOut]; });
    float* o = fa + fr.plane_off + (size_t)y * fr.w + x;
#pragma unroll
    for (int j = 0; j < kOut; ++j)
        if (x + j < fr.w) o[j] = acc[j];
}

// whole-mask blur, vertical: fa -> fb, and the frame's maximum
__global__ void __launch_bounds__(kT)
k_heat_final_v(const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels, const float* __restrict__ fa,
               float* __restrict__ fb, unsigned int* __restrict__ max_bits) {
    __shared__ float tile[kVTile][kVCols];
    __shared__ float wp[kWLen];
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * kVCols + lane, y0 = blockIdx.y * kVRows;
    if (blockIdx.x * kVCols >= fr.w || y0 >= fr.h) return;
    const int r = fr.final_ksize >> 1;
    load_weights(wp, kernels + fr.final_koff, fr.final_ksize);
    const float* src = fa + fr.plane_off;
    for (int i = grp; i < kVTile; i += kT / 32)
        tile[i][lane] = (x < fr.w && i < kVRows + 2 * r) ? src[(size_t)reflect101(y0 + i - r, fr.h) * fr.w + x] : 0.f;
    __syncthreads();
    float m = 0.f;
    if (x < fr.w) {
        float acc[kOut];
        window8(acc, wp, r, [&](int k) { return tile[grp * kOut + k][lane]; });
        float* o = fb + fr.plane_off;
#pragma unroll
        for (int j = 0; j < kOut; ++j) {
            const int y = y0 + grp * kOut + j;
            if (y < fr.h) {
                o[(size_t)y * fr.w + x] = acc[j];
                m = fmaxf(m, acc[j]);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f) atomicMax(max_bits + blockIdx.z, __float_as_uint(m));
}

__global__ void __launch_bounds__(kT)
k_heat_colorize(const VisHeatFrame* __restrict__ frames, const float* __restrict__ fb, const unsigned int* __restrict__ max_bits,
                const uint8_t* __restrict__ jet) {
    const VisHeatFrame& fr = frames[blockIdx.y];
    const float mx = __uint_as_float(max_bits[blockIdx.y]);
    const float* plane = fb + fr.plane_off;
    const long long n = (long long)fr.w * fr.h;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const int y = (int)(i / fr.w), x = (int)(i - (long long)y * fr.w);
        const float v = plane[i];
        // numpy: (heat / max * 255).astype(uint8) in float32, truncation toward zero
        const float t = mx > 0.f ? __fmul_rn(__fdiv_rn(v, mx), 255.f) : __fmul_rn(v, 255.f);
        const int idx = min(max((int)t, 0), 255);
        const uint8_t* s = fr.src + (size_t)y * fr.src_pitch + (size_t)x * 3;
        uint8_t* o = fr.dst + (size_t)y * fr.dst_pitch + (size_t)x * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // cv2.addWeighted(img, 0.6, colour, 0.4, 0): float32, round half to even, saturate
            const float rr = __fadd_rn(__fmul_rn((float)s[c], 0.6f), __fmul_rn((float)__ldg(jet + idx * 3 + c), 0.4f));
            o[c] = (uint8_t)min(max(__float2int_rn(rr), 0), 255);
        }
    }
}

}  // namespace

extern "C" int vis_heatmap_batch(const VisHeatFrame* frames, int n_frames, const VisHeatItem* items, int n_items,
                                 int max_w, int max_h, int max_rw, int max_rh, int64_t plane_floats,
                                 const float* kernels, const uint8_t* jet768, float* heat, float* fa, float* fb,
                                 float* tmp, double* tabs, unsigned int* max_bits, void* stream) {
    if (!frames || n_frames <= 0 || n_frames > 65535 || n_items < 0 || n_items > 65535 || (n_items && (!items || !tmp || !tabs)) ||
        !kernels || !jet768 || !heat || !fa || !fb || !max_bits || max_w <= 0 || max_h <= 0 || plane_floats <= 0 ||
        (n_items && (max_rw <= 0 || max_rh <= 0))) {
        vis::set_error("vis_heatmap_batch: bad arguments (frames=%d items=%d)", n_frames, n_items);
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(heat, 0, sizeof(float) * (size_t)plane_floats, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(max_bits, 0, sizeof(unsigned int) * (size_t)n_frames, st);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_heatmap_batch: cudaMemsetAsync");
    if (n_items) {
        k_heat_tables<<<dim3((max_rw + max_rh + kT - 1) / kT, n_items), kT, 0, st>>>(items, tabs);
        k_heat_defect_h<<<dim3((max_rw + kHSeg - 1) / kHSeg, (max_rh + kHRows - 1) / kHRows, n_items), kT, 0, st>>>(
            items, frames, tabs, kernels, tmp, heat);
        k_heat_defect_v<<<dim3((max_rw + kVCols - 1) / kVCols, (max_rh + kVRows - 1) / kVRows, n_items), kT, 0, st>>>(
            items, frames, kernels, tmp, heat);
    }
    k_heat_final_h<<<dim3((max_w + kHSeg - 1) / kHSeg, (max_h + kHRows - 1) / kHRows, n_frames), kT, 0, st>>>(frames, kernels, heat, fa);
    k_heat_final_v<<<dim3((max_w + kVCols - 1) / kVCols, (max_h + kVRows - 1) / kVRows, n_frames), kT, 0, st>>>(frames, kernels, fa, fb, max_bits);
    k_heat_colorize<<<dim3(592, n_frames), kT, 0, st>>>(frames, fb, max_bits, jet768);
    return vis::check_launch("vis_heatmap_batch");
}
