// vis_host.cpp — host-only entry points of libvis_b200.so: error string, resampling tables, normalisation
// table, push-order records and strip planning for the fused kernel.
//
// The coefficient arithmetic follows Pillow's precompute_coeffs / normalize_coeffs_8bpc
// (libImaging/Resample.c, Pillow 12.2.0), reached by the reference through Image.resize at
// utils/image_utils.py:75 and, on the VLM-input side, tf:image_transforms.py:367.  Everything is IEEE double
// with contraction disabled (-ffp-contract=off in the build), because the results are truncated to
// 22-bit fixed point and one ulp would change a coefficient.
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

#include "vis_internal.h"

namespace vis {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
    set_error("%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
    return VIS_E_CUDA;
}

// ---- filter kernels ------------------------------------------------------------------------------
static double cubic_kernel(double x) {          // Keys cubic, a = -0.5, support 2
    const double a = -0.5;
    x = std::fabs(x);
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static double sinc_pi(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return std::sin(x) / x;
}
static double lanczos3_kernel(double x) {       // truncated sinc, support 3
    if (-3.0 <= x && x < 3.0) return sinc_pi(x) * sinc_pi(x / 3);
    return 0.0;
}

struct FilterDef {
    double (*fn)(double);
    double support;
};
static bool filter_def(int filter, FilterDef* out) {
    switch (filter) {
        case VIS_FILTER_BICUBIC: *out = {cubic_kernel, 2.0}; return true;
        case VIS_FILTER_LANCZOS: *out = {lanczos3_kernel, 3.0}; return true;
        default: return false;
    }
}

struct Window {
    double scale, filterscale, support;
    int ksize;
};
static Window window_for(float in0, float in1, int out_size, const FilterDef& f) {
    Window w;
    // Pillow keeps the source box as float (ImagingResample's float box[4]) and divides in double
    w.scale = (double)(in1 - in0) / out_size;
    w.filterscale = w.scale < 1.0 ? 1.0 : w.scale;
    w.support = f.support * w.filterscale;
    w.ksize = (int)std::ceil(w.support) * 2 + 1;
    return w;
}

}  // namespace vis

using namespace vis;

extern "C" {

int vis_abi_version(void) { return VIS_B200_ABI_VERSION; }

const char* vis_last_error(void) { return g_err; }

#ifndef VIS_SOURCE_HASH_VALUE
#define VIS_SOURCE_HASH_VALUE "unstamped-build!"
#endif
// the marker is what build.py greps in the binary to decide whether it is up to date (content, not mtimes)
static const char g_source_hash[] = "VIS_SOURCE_HASH=" VIS_SOURCE_HASH_VALUE;
const char* vis_source_hash(void) { return g_source_hash + 16; }

int vis_coeff_ksize(int in_size, int out_size, int filter) {
    FilterDef f;
    if (in_size <= 0 || out_size <= 0 || !filter_def(filter, &f)) {
        set_error("vis_coeff_ksize: bad arguments (in=%d out=%d filter=%d)", in_size, out_size, filter);
        return VIS_E_INVALID;
    }
    return window_for(0.f, (float)in_size, out_size, f).ksize;
}

int vis_coeff_ksize_box(float in0, float in1, int out_size, int filter) {
    FilterDef f;
    if (out_size <= 0 || !(in1 > in0) || !filter_def(filter, &f)) {
        set_error("vis_coeff_ksize_box: bad arguments (box=[%g,%g) out=%d filter=%d)", (double)in0, (double)in1, out_size, filter);
        return VIS_E_INVALID;
    }
    return window_for(in0, in1, out_size, f).ksize;
}

int vis_build_coeffs(int in_size, int out_size, int filter, int32_t* k, int32_t* bounds, int* ksize_out) {
    if (in_size <= 0) {
        set_error("vis_build_coeffs: bad arguments (in=%d out=%d filter=%d)", in_size, out_size, filter);
        return VIS_E_INVALID;
    }
    return vis_build_coeffs_box(in_size, 0.f, (float)in_size, out_size, filter, k, bounds, ksize_out);
}

int vis_build_coeffs_box(int in_size, float in0, float in1, int out_size, int filter, int32_t* k, int32_t* bounds,
                         int* ksize_out) {
    FilterDef f;
    if (in_size <= 0 || out_size <= 0 || !k || !bounds || !(in1 > in0) || in0 < 0.f || in1 > (float)in_size ||
        !filter_def(filter, &f)) {
        set_error("vis_build_coeffs: bad arguments (in=%d box=[%g,%g) out=%d filter=%d)", in_size, (double)in0,
                  (double)in1, out_size, filter);
        return VIS_E_INVALID;
    }
    const Window win = window_for(in0, in1, out_size, f);
    const double inv_fs = 1.0 / win.filterscale;
    const double fixed_one = (double)(1 << VIS_PRECISION_BITS);
    std::vector<double> w((size_t)win.ksize);
    for (int o = 0; o < out_size; ++o) {
        const double center = in0 + (o + 0.5) * win.scale;
        int first = (int)(center - win.support + 0.5);      // C truncation, as Pillow
        if (first < 0) first = 0;
        int last = (int)(center + win.support + 0.5);
        if (last > in_size) last = in_size;
        const int taps = last - first;
        double total = 0.0;
        for (int t = 0; t < taps; ++t) {
            w[t] = f.fn((t + first - center + 0.5) * inv_fs);
            total += w[t];                                   // sequential, in tap order
        }
        int32_t* row = k + (size_t)o * win.ksize;
        for (int t = 0; t < win.ksize; ++t) {
            double v = 0.0;
            if (t < taps) v = (total != 0.0) ? w[t] / total : w[t];
            row[t] = v < 0 ? (int32_t)(-0.5 + v * fixed_one) : (int32_t)(0.5 + v * fixed_one);
        }
        bounds[2 * o] = first;
        bounds[2 * o + 1] = taps;
    }
    if (ksize_out) *ksize_out = win.ksize;
    return VIS_OK;
}

// Pillow's precompute_coeffs for the modes it resamples in double precision (I;16, I, F — libImaging/Resample.c
// ImagingResampleHorizontal_16bpc / _32bpc): the normalised weights stay doubles, nothing is converted to fixed point.
int vis_build_coeffs_f64(int in_size, int out_size, int filter, double* k, int32_t* bounds, int* ksize_out) {
    FilterDef f;
    if (in_size <= 0 || out_size <= 0 || !k || !bounds || !filter_def(filter, &f)) {
        set_error("vis_build_coeffs_f64: bad arguments (in=%d out=%d filter=%d)", in_size, out_size, filter);
        return VIS_E_INVALID;
    }
    const Window win = window_for(0.f, (float)in_size, out_size, f);
    const double inv_fs = 1.0 / win.filterscale;
    for (int o = 0; o < out_size; ++o) {
        const double center = (o + 0.5) * win.scale;
        int first = (int)(center - win.support + 0.5);
        if (first < 0) first = 0;
        int last = (int)(center + win.support + 0.5);
        if (last > in_size) last = in_size;
        const int taps = last - first;
        double* row = k + (size_t)o * win.ksize;
        double total = 0.0;
        for (int t = 0; t < taps; ++t) {
            row[t] = f.fn((t + first - center + 0.5) * inv_fs);
            total += row[t];
        }
        for (int t = 0; t < taps; ++t)
            if (total != 0.0) row[t] /= total;
        for (int t = taps; t < win.ksize; ++t) row[t] = 0.0;
        bounds[2 * o] = first;
        bounds[2 * o + 1] = taps;
    }
    if (ksize_out) *ksize_out = win.ksize;
    return VIS_OK;
}

int vis_build_lut(const float mean[3], const float stdv[3], double rescale, float* lut768) {
    if (!mean || !stdv || !lut768) {
        set_error("vis_build_lut: null argument");
        return VIS_E_INVALID;
    }
    for (int v = 0; v < 256; ++v) {
        // tf:image_transforms.py:118-122: float64 multiply, then cast to float32
        volatile float scaled = (float)((double)v * rescale);
        for (int c = 0; c < 3; ++c) {
            // tf:image_transforms.py:439: float32 subtract and divide
            volatile float centred = scaled - mean[c];
            lut768[v * 3 + c] = centred / stdv[c];
        }
    }
    return VIS_OK;
}

// ---- push-order records for the fused kernel -------------------------------------------------------
// record layout (int32 slots): [0 .. kt)   coefficients, newest tap first: slot t multiplies input (last - t)
//                              [stride-2]  first input index of the window
//                              [stride-1]  last input index of the window
int vis_record_stride(int kt) {
    if (kt <= 0) return VIS_E_INVALID;
    return (kt + 2 + 3) & ~3;
}

int vis_max_taps(const int32_t* bounds, int out_size) {
    if (!bounds || out_size <= 0) return VIS_E_INVALID;
    int kt = 1;
    for (int o = 0; o < out_size; ++o) kt = bounds[2 * o + 1] > kt ? bounds[2 * o + 1] : kt;
    return kt;
}

int vis_pack_records(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int kt,
                     int32_t* rec, int64_t rec_capacity) {
    if (out_size <= 0 || !k || !bounds || !rec || ksize <= 0 || kt <= 0) {
        set_error("vis_pack_records: bad arguments");
        return VIS_E_INVALID;
    }
    if (vis_max_taps(bounds, out_size) > kt) {
        set_error("vis_pack_records: table has %d taps, record holds %d", vis_max_taps(bounds, out_size), kt);
        return VIS_E_INVALID;
    }
    const int stride = vis_record_stride(kt);
    if (rec_capacity < (int64_t)(out_size + 1) * stride) {
        set_error("vis_pack_records: capacity %lld < %lld", (long long)rec_capacity,
                  (long long)(out_size + 1) * stride);
        return VIS_E_CAPACITY;
    }
    std::memset(rec, 0, sizeof(int32_t) * (size_t)(out_size + 1) * stride);
    for (int o = 0; o < out_size; ++o) {
        const int first = bounds[2 * o], taps = bounds[2 * o + 1];
        int32_t* r = rec + (size_t)o * stride;
        for (int t = 0; t < taps; ++t) r[t] = k[(size_t)o * ksize + (taps - 1 - t)];
        r[stride - 2] = first;
        r[stride - 1] = first + taps - 1;
    }
    int32_t* s = rec + (size_t)out_size * stride;   // sentinel: never matches an input index
    s[stride - 2] = INT_MAX;
    s[stride - 1] = INT_MAX;
    return VIS_OK;
}

}  // extern "C"
