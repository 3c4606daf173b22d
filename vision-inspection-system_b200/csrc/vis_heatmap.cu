// vis_heatmap.cu — device half of create_heatmap_overlay (utils/image_utils.py:320-604; SURVEY.md 8f "next" row 2).
//
// BATCH form (round 2): any number of frames and defects in SIX launches, whatever the batch size —
//   k_heat_tables     per defect, everything that depends on one coordinate: the two 1-D factors of its Gaussian,
//                     exp(-(x-cx)^2 / 2 sigma^2) and the same in y, in float64 like the reference's numpy expression (one
//                     exp per region column / row instead of one per region pixel: the 2-D value is their product, within
//                     2 ulp(double) of exp of the sum, long before the float32 rounding that follows), and the squared
//                     distances of the boost and cut-off tests
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
constexpr int kHSeg = 512, kHRows = 4;           // horizontal pass: 512 outputs x 4 rows per step (64 threads per row) ...
constexpr int kHBlkRows = 16;                    // ... and four steps per block: 16 rows staged at once
constexpr int kHLen = kHSeg + 2 * kBlurR + 6;    // staged row, padded to a multiple of 8 (568)
constexpr int kHPhase = kHLen / kOut;            // 71: element i lives at (i % 8) * 71 + i / 8 -> lanes read consecutive words
constexpr int kVCols = 32, kVRows = 64;          // vertical pass: 32 columns x 64 output rows per step ...
constexpr int kVBlkRows = 192;                   // ... and three steps per block: 192 + 2R rows staged at once
constexpr int kRDefect = 25, kRFinal = 15;       // window radii of the two launch classes (51 / 31 taps at most)
static_assert(kHLen % kOut == 0, "staged row must de-interleave evenly");
static_assert(kHBlkRows % kHRows == 0 && kVBlkRows % kVRows == 0, "whole steps per block");

// The kernel of a launch class sits in shared memory, centred in a window of compile-time radius R (smaller kernels are
// padded with zero weights; the reference's kernels are 49..51 taps for defects, <= 31 for the final blur), loaded ONCE
// per block; the taps unroll completely and a thread keeps the eight weights its sliding window needs in registers.
template <int R>
__device__ __forceinline__ void load_weights(float* sw, const float* __restrict__ kern, int ksize) {
    const int pad = R - (ksize >> 1);
    for (int t = threadIdx.x; t < 2 * R + 1; t += kT) sw[t] = (t >= pad && t < pad + ksize) ? __ldg(kern + t - pad) : 0.f;
}

// analytic heat of a box defect / widespread defect at region-local (lx, ly), float64 like the reference.  Everything
// that depends on one coordinate only comes from the defect's 1-D tables (k_heat_tables): the Gaussian factors, the
// squared normalised distances of the boost test ((dx / max(w/2, 1))^2: the divisions leave the per-pixel path) and the
// squared distances of the 4-sigma cut-off — same operations, same roundings, evaluated once per column / row.
struct HeatTabs { const double *tx, *ty, *bx, *by, *qx, *qy; };
__device__ __forceinline__ HeatTabs heat_tabs(const double* __restrict__ tabs, const VisHeatItem& it) {
    const int rw = it.d.x2 - it.d.x1, rh = it.d.y2 - it.d.y1;
    const double* p = tabs + it.tab_off;
    return {p, p + rw, p + rw + rh, p + 2 * rw + rh, p + 2 * (rw + rh), p + 3 * rw + 2 * rh};
}
__device__ __forceinline__ float heat_value(const VisHeatDefect& d, const HeatTabs& t, int lx, int ly) {
    const double g0 = d.intensity * (t.tx[lx] * t.ty[ly]);
    if (d.kind == 1) return (float)g0;
    const int gx = d.x1 + lx, gy = d.y1 + ly;
    const bool in_box = gx >= d.x && gx < d.x + d.w && gy >= d.y && gy < d.y + d.h;
    const double boost = (t.bx[lx] + t.by[ly] < 1.2 * 1.2) ? 1.8 : (in_box ? 1.4 : 1.0);
    const double g = fmin(1.0, g0 * boost);
    const double lim = 4.0 * d.sigma;
    return (t.qx[lx] + t.qy[ly]) < lim * lim ? (float)g : 0.f;
}

__global__ void __launch_bounds__(kT) k_heat_tables(const VisHeatItem* __restrict__ items, double* __restrict__ tabs) {
    const VisHeatItem& it = items[blockIdx.y];
    const VisHeatDefect& d = it.d;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int t = blockIdx.x * kT + threadIdx.x;
    if (t >= rw + rh) return;
    const bool is_x = t < rw;
    const double c = is_x ? (double)(d.x1 + t) - d.cx : (double)(d.y1 + t - rw) - d.cy;
    const double e = c / fmax((is_x ? d.w : d.h) / 2.0, 1.0);
    double* p = tabs + it.tab_off;
    p[t] = exp(-(c * c) / (2.0 * (d.sigma * d.sigma)));          // tx | ty
    p[rw + rh + t] = e * e;                                        // bx | by
    p[2 * (rw + rh) + t] = c * c;                                  // qx | qy
}

// 8 outputs from one sliding window of radius R: value(k), k = 0 .. 8 + 2R - 1, is staged element (first output - R + k);
// output j accumulates taps t = k - j in ascending order.
template <int R, typename F>
__device__ __forceinline__ void window8(float (&acc)[kOut], const float* sw, F value) {
#pragma unroll
    for (int j = 0; j < kOut; ++j) acc[j] = 0.f;
    float w[kOut];                                // w[i] = weight of tap (k - i) at step k: a sliding set of eight
#pragma unroll
    for (int i = 0; i < kOut; ++i) w[i] = 0.f;
#pragma unroll
    for (int k = 0; k < kOut + 2 * R; ++k) {
#pragma unroll
        for (int i = kOut - 1; i > 0; --i) w[i] = w[i - 1];
        w[0] = k <= 2 * R ? sw[k] : 0.f;          // one broadcast read per step
        const float v = value(k);
#pragma unroll
        for (int j = 0; j < kOut; ++j)
            if (k - j >= 0 && k - j <= 2 * R) acc[j] = fmaf(w[j], v, acc[j]);
    }
}

__device__ __forceinline__ void atomic_max_f(float* p, float v) {      // non-negative floats: bit order = value order
    atomicMax(reinterpret_cast<unsigned int*>(p), __float_as_uint(fmaxf(v, 0.f)));
}

// horizontal step shared by both classes: 4 staged rows (de-interleaved by 8) x 512 outputs, thread = 8 outputs
template <int R>
__device__ __forceinline__ void h_steps(const float (*rows)[kHLen], const float* sw, int n_rows, int n_cols, float* out, size_t out_pitch) {
    const int c = threadIdx.x & 63;
    if (c * kOut >= n_cols) return;
#pragma unroll 1
    for (int row = threadIdx.x >> 6; row < n_rows; row += kHRows) {
        float acc[kOut];
        const float* rp = rows[row] + c;
        window8<R>(acc, sw, [&](int k) { return rp[(k % kOut) * kHPhase + k / kOut]; });      // element c*8 + k
        float* o = out + (size_t)row * out_pitch + c * kOut;
#pragma unroll
        for (int j = 0; j < kOut; ++j)
            if (c * kOut + j < n_cols) o[j] = acc[j];
    }
}

__global__ void __launch_bounds__(kT, 4)
k_heat_defect_h(const VisHeatItem* __restrict__ items, const VisHeatFrame* __restrict__ frames, const double* __restrict__ tabs,
                const float* __restrict__ kernels, float* __restrict__ tmp, float* __restrict__ heat) {
    __shared__ float rows[kHBlkRows][kHLen];
    __shared__ float sw[2 * kRDefect + 1];
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHBlkRows;
    if (x0 >= rw || y0 >= rh) return;
    const HeatTabs tb = heat_tabs(tabs, it);
    const VisHeatFrame& fr = frames[it.frame];
    const int n_rows = min(kHBlkRows, rh - y0);
    const bool direct = d.kind == 1 || d.ksize == 1;
    if (direct) {                                      // no blur: max-combine the analytic heat itself
        float* plane = heat + fr.plane_off;
        for (int i = threadIdx.x; i < n_rows * kHSeg; i += kT) {
            const int ly = y0 + i / kHSeg, lx = x0 + i % kHSeg;
            if (lx < rw) atomic_max_f(plane + (size_t)(d.y1 + ly) * fr.w + d.x1 + lx, heat_value(d, tb, lx, ly));
        }
        return;
    }
    constexpr int R = kRDefect;
    load_weights<R>(sw, kernels + d.koff, d.ksize);
    // Staging, thread = column: everything that depends on the column (reflected at the REGION border) is fetched once,
    // then the rows of the block are walked with the row's three table values read as warp-uniform loads.  The
    // operations on a pixel are exactly heat_value()'s (same float64 roundings, same comparisons).
    constexpr int span = kHSeg + 2 * R;
    const double lim2 = (4.0 * d.sigma) * (4.0 * d.sigma);
    for (int idx = threadIdx.x; idx < span; idx += kT) {
        const int lx = reflect101(x0 + idx - R, rw), gx = d.x1 + lx;
        const double txv = tb.tx[lx], bxv = tb.bx[lx], qxv = tb.qx[lx];
        const bool in_x = gx >= d.x && gx < d.x + d.w;
        float* col = &rows[0][(idx % kOut) * kHPhase + idx / kOut];
#pragma unroll 4
        for (int r = 0; r < n_rows; ++r) {
            const int ly = y0 + r, gy = d.y1 + ly;
            const double g0 = d.intensity * (txv * tb.ty[ly]);
            const bool in_box = in_x && gy >= d.y && gy < d.y + d.h;
            const double boost = (bxv + tb.by[ly] < 1.2 * 1.2) ? 1.8 : (in_box ? 1.4 : 1.0);
            const double g = fmin(1.0, g0 * boost);
            col[r * kHLen] = (qxv + tb.qy[ly]) < lim2 ? (float)g : 0.f;
        }
    }
    __syncthreads();
    h_steps<R>(rows, sw, n_rows, min(kHSeg, rw - x0), tmp + it.tmp_off + (size_t)y0 * rw + x0, (size_t)rw);
}

// vertical step shared by both classes: the staged tile holds rows y0 - R .. y0 + kVBlkRows + R of 32 columns; a warp
// produces 8 consecutive rows of its lane's column per step, 64 rows per step and block
template <int R, typename Emit>
__device__ __forceinline__ void v_steps(const float (*tile)[kVCols], const float* sw, int n_rows, Emit emit) {
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int r0 = (threadIdx.x >> 5) * kOut; r0 < n_rows; r0 += kVRows) {
        float acc[kOut];
        window8<R>(acc, sw, [&](int k) { return tile[r0 + k][lane]; });
#pragma unroll
        for (int j = 0; j < kOut; ++j)
            if (r0 + j < n_rows) emit(r0 + j, acc[j]);
    }
}

template <int R>
__device__ __forceinline__ void v_stage(float (*tile)[kVCols], const float* __restrict__ src, int pitch, int x, bool x_ok, int y0, int h) {
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int n = min(kVBlkRows, h - y0) + 2 * R;
    if (!x_ok) {
        for (int i = grp; i < n; i += kT / 32) tile[i][lane] = 0.f;
    } else if (y0 >= R && y0 + n - R <= h) {            // interior block: no reflection, loads four at a time
        const float* p = src + (size_t)(y0 - R) * pitch + x;
#pragma unroll 4
        for (int i = grp; i < n; i += kT / 32) tile[i][lane] = p[(size_t)i * pitch];
    } else {
        for (int i = grp; i < n; i += kT / 32) tile[i][lane] = src[(size_t)reflect101(y0 + i - R, h) * pitch + x];
    }
}

// vertical pass of a defect region: tmp -> max into the frame's heat plane
__global__ void __launch_bounds__(kT)
k_heat_defect_v(const VisHeatItem* __restrict__ items, const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels,
                const float* __restrict__ tmp, float* __restrict__ heat) {
    __shared__ float tile[kVBlkRows + 2 * kRDefect][kVCols];
    __shared__ float sw[2 * kRDefect + 1];
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    if (d.kind == 1 || d.ksize == 1) return;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int x = blockIdx.x * kVCols + (threadIdx.x & 31), y0 = blockIdx.y * kVBlkRows;
    if (blockIdx.x * kVCols >= rw || y0 >= rh) return;
    constexpr int R = kRDefect;
    load_weights<R>(sw, kernels + d.koff, d.ksize);
    v_stage<R>(tile, tmp + it.tmp_off, rw, x, x < rw, y0, rh);
    __syncthreads();
    if (x >= rw) return;
    const VisHeatFrame& fr = frames[it.frame];
    float* plane = heat + fr.plane_off + (size_t)(d.y1 + y0) * fr.w + d.x1 + x;
    const int fw = fr.w;
    v_steps<R>(tile, sw, min(kVBlkRows, rh - y0), [&](int r, float v) { atomic_max_f(plane + (size_t)r * fw, v); });
}

// whole-mask blur, horizontal: heat -> fa (reflect at the image border)
__global__ void __launch_bounds__(kT)
k_heat_final_h(const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels, const float* __restrict__ heat,
               float* __restrict__ fa) {
    __shared__ float rows[kHBlkRows][kHLen];
    __shared__ float sw[2 * kRFinal + 1];
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHBlkRows;
    if (x0 >= fr.w || y0 >= fr.h) return;
    constexpr int R = kRFinal;
    load_weights<R>(sw, kernels + fr.final_koff, fr.final_ksize);
    const float* plane = heat + fr.plane_off + (size_t)y0 * fr.w;
    const int n_rows = min(kHBlkRows, fr.h - y0);
    constexpr int span = kHSeg + 2 * R;
    for (int idx = threadIdx.x; idx < span; idx += kT) {            // thread = column, reflected once
        const float* src = plane + reflect101(x0 + idx - R, fr.w);
        float* col = &rows[0][(idx % kOut) * kHPhase + idx / kOut];
#pragma unroll 4
        for (int r = 0; r < n_rows; ++r) col[r * kHLen] = src[(size_t)r * fr.w];
    }
    __syncthreads();
    h_steps<R>(rows, sw, n_rows, min(kHSeg, fr.w - x0), fa + fr.plane_off + (size_t)y0 * fr.w + x0, (size_t)fr.w);
}

// whole-mask blur, vertical: fa -> fb, and the frame's maximum
__global__ void __launch_bounds__(kT, 5)
k_heat_final_v(const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels, const float* __restrict__ fa,
               float* __restrict__ fb, unsigned int* __restrict__ max_bits) {
    __shared__ float tile[kVBlkRows + 2 * kRFinal][kVCols];
    __shared__ float sw[2 * kRFinal + 1];
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int lane = threadIdx.x & 31;
    const int x = blockIdx.x * kVCols + lane, y0 = blockIdx.y * kVBlkRows;
    if (blockIdx.x * kVCols >= fr.w || y0 >= fr.h) return;
    constexpr int R = kRFinal;
    load_weights<R>(sw, kernels + fr.final_koff, fr.final_ksize);
    v_stage<R>(tile, fa + fr.plane_off, fr.w, x, x < fr.w, y0, fr.h);
    __syncthreads();
    float m = 0.f;
    if (x < fr.w) {
        float* o = fb + fr.plane_off + (size_t)y0 * fr.w + x;
        const int fw = fr.w;
        v_steps<R>(tile, sw, min(kVBlkRows, fr.h - y0), [&](int r, float v) { o[(size_t)r * fw] = v; m = fmaxf(m, v); });
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0.f) atomicMax(max_bits + blockIdx.z, __float_as_uint(m));
}

__device__ __forceinline__ int heat_index(float v, float mx) {
    // numpy: (heat / max * 255).astype(uint8) in float32, truncation toward zero
    const float t = mx > 0.f ? __fmul_rn(__fdiv_rn(v, mx), 255.f) : __fmul_rn(v, 255.f);
    return min(max((int)t, 0), 255);
}
__device__ __forceinline__ uint32_t blend_byte(uint32_t img, uint32_t col) {
    // cv2.addWeighted(img, 0.6, colour, 0.4, 0): float32, round half to even, saturate
    const float r = __fadd_rn(__fmul_rn((float)img, 0.6f), __fmul_rn((float)col, 0.4f));
    return (uint32_t)min(max(__float2int_rn(r), 0), 255);
}

// thread = 4 pixels (one float4 of heat, three 32-bit words of BGR in, three out) when the frame allows; JET in shared memory
__global__ void __launch_bounds__(kT)
k_heat_colorize(const VisHeatFrame* __restrict__ frames, const float* __restrict__ fb, const unsigned int* __restrict__ max_bits,
                const uint8_t* __restrict__ jet) {
    __shared__ uint8_t sjet[768];
    for (int i = threadIdx.x; i < 768; i += kT) sjet[i] = __ldg(jet + i);
    __syncthreads();
    const VisHeatFrame& fr = frames[blockIdx.y];
    const float mx = __uint_as_float(max_bits[blockIdx.y]);
    const float* plane = fb + fr.plane_off;
    const bool vec = (fr.w & 3) == 0 && ((fr.src_pitch | fr.dst_pitch) & 3) == 0 && (fr.plane_off & 3) == 0 &&
                     (((uintptr_t)fr.src | (uintptr_t)fr.dst) & 3) == 0;
    if (vec) {
        const int wq = fr.w >> 2;
        const long long n = (long long)wq * fr.h;
        for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
            const int y = (int)(i / wq), q = (int)(i - (long long)y * wq);
            const float4 hv = *reinterpret_cast<const float4*>(plane + (size_t)y * fr.w + 4 * q);
            const uint32_t* s = reinterpret_cast<const uint32_t*>(fr.src + (size_t)y * fr.src_pitch) + 3 * q;
            uint32_t* o = reinterpret_cast<uint32_t*>(fr.dst + (size_t)y * fr.dst_pitch) + 3 * q;
            const uint32_t in[3] = {__ldcs(s), __ldcs(s + 1), __ldcs(s + 2)};
            const int idx[4] = {heat_index(hv.x, mx), heat_index(hv.y, mx), heat_index(hv.z, mx), heat_index(hv.w, mx)};
            uint32_t out[3] = {0u, 0u, 0u};
#pragma unroll
            for (int b = 0; b < 12; ++b) {                     // byte b = pixel b / 3, channel b % 3
                const uint32_t v = blend_byte((in[b >> 2] >> (8 * (b & 3))) & 0xffu, sjet[idx[b / 3] * 3 + b % 3]);
                out[b >> 2] |= v << (8 * (b & 3));
            }
            __stcs(o, out[0]); __stcs(o + 1, out[1]); __stcs(o + 2, out[2]);
        }
        return;
    }
    const long long n = (long long)fr.w * fr.h;
    for (long long i = (long long)blockIdx.x * kT + threadIdx.x; i < n; i += (long long)gridDim.x * kT) {
        const int y = (int)(i / fr.w), x = (int)(i - (long long)y * fr.w);
        const int idx = heat_index(plane[i], mx);
        const uint8_t* s = fr.src + (size_t)y * fr.src_pitch + (size_t)x * 3;
        uint8_t* o = fr.dst + (size_t)y * fr.dst_pitch + (size_t)x * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[c] = (uint8_t)blend_byte(s[c], sjet[idx * 3 + c]);
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
        k_heat_defect_h<<<dim3((max_rw + kHSeg - 1) / kHSeg, (max_rh + kHBlkRows - 1) / kHBlkRows, n_items), kT, 0, st>>>(
            items, frames, tabs, kernels, tmp, heat);
        k_heat_defect_v<<<dim3((max_rw + kVCols - 1) / kVCols, (max_rh + kVBlkRows - 1) / kVBlkRows, n_items), kT, 0, st>>>(
            items, frames, kernels, tmp, heat);
    }
    k_heat_final_h<<<dim3((max_w + kHSeg - 1) / kHSeg, (max_h + kHBlkRows - 1) / kHBlkRows, n_frames), kT, 0, st>>>(frames, kernels, heat, fa);
    k_heat_final_v<<<dim3((max_w + kVCols - 1) / kVCols, (max_h + kVBlkRows - 1) / kVBlkRows, n_frames), kT, 0, st>>>(frames, kernels, fa, fb, max_bits);
    k_heat_colorize<<<dim3(592, n_frames), kT, 0, st>>>(frames, fb, max_bits, jet768);
    return vis::check_launch("vis_heatmap_batch");
}
