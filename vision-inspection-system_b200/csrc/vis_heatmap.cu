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
constexpr int kHSeg = 512, kHRows = 4;           // horizontal pass: 4 rows x 512 outputs per block (64 threads per row)
constexpr int kHLen = kHSeg + 2 * kBlurR + 6;    // staged row, padded to a multiple of 8 (568)
constexpr int kHPhase = kHLen / kOut;            // 71: element i lives at (i % 8) * 71 + i / 8 -> lanes read consecutive words
constexpr int kVCols = 32, kVRows = 64;          // vertical pass: 32 columns x 64 output rows per block
constexpr int kVTile = kVRows + 2 * kBlurR;       // rows staged
constexpr int kRDefect = 25, kRFinal = 15;       // window radii of the two launch classes (51 / 31 taps at most)
static_assert(kHLen % kOut == 0, "staged row must de-interleave evenly");

// The kernel of a launch class is held in REGISTERS, centred in a window of compile-time radius R (smaller kernels are
// padded with zero weights; the reference's kernels are 49..51 taps for defects, <= 31 for the final blur), so the taps
// unroll completely: per staged value one shared-memory read and up to 8 FMA with register operands only.
template <int R>
__device__ __forceinline__ void load_weights(float (&w)[2 * R + 1], const float* __restrict__ kern, int ksize) {
    const int pad = R - (ksize >> 1);
#pragma unroll
    for (int t = 0; t < 2 * R + 1; ++t) w[t] = (t >= pad && t < pad + ksize) ? __ldg(kern + t - pad) : 0.f;
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
__device__ __forceinline__ void window8(float (&acc)[kOut], const float (&w)[2 * R + 1], F value) {
#pragma unroll
    for (int j = 0; j < kOut; ++j) acc[j] = 0.f;
#pragma unroll
    for (int k = 0; k < kOut + 2 * R; ++k) {
        const float v = value(k);
#pragma unroll
        for (int j = 0; j < kOut; ++j)
            if (k - j >= 0 && k - j <= 2 * R) acc[j] = fmaf(w[k - j], v, acc[j]);
    }
}

__device__ __forceinline__ void atomic_max_f(float* p, float v) {      // non-negative floats: bit order = value order
    atomicMax(reinterpret_cast<unsigned int*>(p), __float_as_uint(fmaxf(v, 0.f)));
}

__global__ void __launch_bounds__(kT)
k_heat_defect_h(const VisHeatItem* __restrict__ items, const VisHeatFrame* __restrict__ frames, const double* __restrict__ tabs,
                const float* __restrict__ kernels, float* __restrict__ tmp, float* __restrict__ heat) {
    __shared__ float rows[kHRows][kHLen];
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHRows;
    if (x0 >= rw || y0 >= rh) return;
    const HeatTabs tb = heat_tabs(tabs, it);
    const VisHeatFrame& fr = frames[it.frame];
    const bool direct = d.kind == 1 || d.ksize == 1;
    if (direct) {                                      // no blur: max-combine the analytic heat itself
        float* plane = heat + fr.plane_off;
        for (int i = threadIdx.x; i < kHRows * kHSeg; i += kT) {
            const int ly = y0 + i / kHSeg, lx = x0 + i % kHSeg;
            if (ly < rh && lx < rw) atomic_max_f(plane + (size_t)(d.y1 + ly) * fr.w + d.x1 + lx, heat_value(d, tb, lx, ly));
        }
        return;
    }
    constexpr int R = kRDefect;
    float w[2 * R + 1];
    load_weights<R>(w, kernels + d.koff, d.ksize);
    constexpr int span = kHSeg + 2 * R;
    for (int i = threadIdx.x; i < kHRows * span; i += kT) {
        const int row = i / span, idx = i - row * span;
        const int ly = y0 + row;
        float v = 0.f;
        if (ly < rh) v = heat_value(d, tb, reflect101(x0 + idx - R, rw), ly);
        rows[row][(idx % kOut) * kHPhase + idx / kOut] = v;
    }
    __syncthreads();
    const int row = threadIdx.x >> 6, c = threadIdx.x & 63;
    const int ly = y0 + row, lx = x0 + c * kOut;
    if (ly >= rh || lx >= rw) return;
    float acc[kOut];
    const float* rp = rows[row] + c;
    window8<R>(acc, w, [&](int k) { return rp[(k % kOut) * kHPhase + k / kOut]; });      // element c*8 + k
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
    const VisHeatItem& it = items[blockIdx.z];
    const VisHeatDefect& d = it.d;
    if (d.kind == 1 || d.ksize == 1) return;
    const int rw = d.x2 - d.x1, rh = d.y2 - d.y1;
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * kVCols + lane, y0 = blockIdx.y * kVRows;
    if (blockIdx.x * kVCols >= rw || y0 >= rh) return;
    constexpr int R = kRDefect;
    float w[2 * R + 1];
    load_weights<R>(w, kernels + d.koff, d.ksize);
    const float* src = tmp + it.tmp_off;
    for (int i = grp; i < kVRows + 2 * R; i += kT / 32)
        tile[i][lane] = x < rw ? src[(size_t)reflect101(y0 + i - R, rh) * rw + x] : 0.f;
    __syncthreads();
    if (x >= rw) return;
    float acc[kOut];
    window8<R>(acc, w, [&](int k) { return tile[grp * kOut + k][lane]; });
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
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int x0 = blockIdx.x * kHSeg, y0 = blockIdx.y * kHRows;
    if (x0 >= fr.w || y0 >= fr.h) return;
    constexpr int R = kRFinal;
    float w[2 * R + 1];
    load_weights<R>(w, kernels + fr.final_koff, fr.final_ksize);
    const float* plane = heat + fr.plane_off;
    constexpr int span = kHSeg + 2 * R;
    for (int i = threadIdx.x; i < kHRows * span; i += kT) {
        const int row = i / span, idx = i - row * span;
        const int y = y0 + row;
        rows[row][(idx % kOut) * kHPhase + idx / kOut] = y < fr.h ? plane[(size_t)y * fr.w + reflect101(x0 + idx - R, fr.w)] : 0.f;
    }
    __syncthreads();
    const int row = threadIdx.x >> 6, c = threadIdx.x & 63;
    const int y = y0 + row, x = x0 + c * kOut;
    if (y >= fr.h || x >= fr.w) return;
    float acc[kOut];
    const float* rp = rows[row] + c;
    window8<R>(acc, w, [&](int k) { return rp[(k % kOut) * kHPhase + k / kOut]; });
    float* o = fa + fr.plane_off + (size_t)y * fr.w + x;
#pragma unroll
    for (int j = 0; j < kOut; ++j)
        if (x + j < fr.w) o[j] = acc[j];
}

// whole-mask blur, vertical: fa -> fb, and the frame's maximum
__global__ void __launch_bounds__(kT)
k_heat_final_v(const VisHeatFrame* __restrict__ frames, const float* __restrict__ kernels, const float* __restrict__ fa,
               float* __restrict__ fb, unsigned int* __restrict__ max_bits) {
    __shared__ float tile[kVRows + 2 * kRFinal][kVCols];
    const VisHeatFrame& fr = frames[blockIdx.z];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int x = blockIdx.x * kVCols + lane, y0 = blockIdx.y * kVRows;
    if (blockIdx.x * kVCols >= fr.w || y0 >= fr.h) return;
    constexpr int R = kRFinal;
    float w[2 * R + 1];
    load_weights<R>(w, kernels + fr.final_koff, fr.final_ksize);
    const float* src = fa + fr.plane_off;
    for (int i = grp; i < kVRows + 2 * R; i += kT / 32)
        tile[i][lane] = x < fr.w ? src[(size_t)reflect101(y0 + i - R, fr.h) * fr.w + x] : 0.f;
    __syncthreads();
    float m = 0.f;
    if (x < fr.w) {
        float acc[kOut];
        window8<R>(acc, w, [&](int k) { return tile[grp * kOut + k][lane]; });
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
