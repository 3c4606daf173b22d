// vis_generic.cu — generic (any geometry, any filter) device passes of libvis_b200.so.
//
// These are the always-correct building blocks: one horizontal pass, one vertical pass, one
// normalize+patchify pass, each a separate launch on one image.  They serve
//   * resize_image / agent thumbnails (LANCZOS, uint8 out) — utils/image_utils.py:75,
//     src/agents/vlm_inspector.py:64, src/agents/vlm_auditor.py:91;
//   * geometries the fused kernel (vis_fused_sched*.cu, vis_fused_ws.cu) declines: > 16 taps, the tall-image vertical-first
//     branch (PIL:Image.py:2431-2435), unaligned row pitches.
// Arithmetic: Pillow ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc — int32 accumulate of
// uint8 x 22-bit coefficients, +2^21, arithmetic >>22, clamp to 0..255, uint8 between the passes.
#include <algorithm>
#include <climits>

#include "vis_internal.h"

namespace {

__device__ __forceinline__ uint8_t clip8(int acc) {
    return (uint8_t)min(max(acc >> VIS_PRECISION_BITS, 0), 255);
}

// one thread = one output pixel (all channels); taps are contiguous pixels of the same row
template <int CH>
__global__ void __launch_bounds__(256)
k_resample_h(const uint8_t* __restrict__ src, int64_t src_pitch, int rows,
             uint8_t* __restrict__ dst, int64_t dst_pitch, int out_w,
             const int32_t* __restrict__ k, const int32_t* __restrict__ bounds, int ksize) {
    const int xo = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (xo >= out_w || y >= rows) return;
    const int first = __ldg(bounds + 2 * xo), taps = __ldg(bounds + 2 * xo + 1);
    const int32_t* kk = k + (size_t)xo * ksize;
    const uint8_t* p = src + (size_t)y * src_pitch + (size_t)first * CH;
    int acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 1 << (VIS_PRECISION_BITS - 1);
    for (int t = 0; t < taps; ++t) {
        const int w = __ldg(kk + t);
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] += (int)p[t * CH + c] * w;
    }
    uint8_t* d = dst + (size_t)y * dst_pitch + (size_t)xo * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) d[c] = clip8(acc[c]);
}

// one thread = VEC consecutive bytes of one output row; taps are the same bytes of consecutive source rows
template <int VEC>
__global__ void __launch_bounds__(256)
k_resample_v(const uint8_t* __restrict__ src, int64_t src_pitch, int row_bytes,
             uint8_t* __restrict__ dst, int64_t dst_pitch, int out_h,
             const int32_t* __restrict__ k, const int32_t* __restrict__ bounds, int ksize) {
    const int xb = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const int yo = blockIdx.y * blockDim.y + threadIdx.y;
    if (xb >= row_bytes || yo >= out_h) return;
    const int first = __ldg(bounds + 2 * yo), taps = __ldg(bounds + 2 * yo + 1);
    const int32_t* kk = k + (size_t)yo * ksize;
    int acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 1 << (VIS_PRECISION_BITS - 1);
    const uint8_t* p = src + (size_t)first * src_pitch + xb;
    if (VEC == 4) {
        for (int t = 0; t < taps; ++t) {
            const int w = __ldg(kk + t);
            const uint32_t q = __ldg(reinterpret_cast<const uint32_t*>(p + (size_t)t * src_pitch));
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] += (int)((q >> (8 * i)) & 0xff) * w;
        }
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) o |= (uint32_t)clip8(acc[i]) << (8 * i);
        *reinterpret_cast<uint32_t*>(dst + (size_t)yo * dst_pitch + xb) = o;
    } else {
        for (int t = 0; t < taps; ++t) acc[0] += (int)p[(size_t)t * src_pitch] * __ldg(kk + t);
        dst[(size_t)yo * dst_pitch + xb] = clip8(acc[0]);
    }
}

// one thread = 4 consecutive floats (16 B) of one pixel_values row
__global__ void __launch_bounds__(128)
k_normalize_patchify(const uint8_t* __restrict__ src, int64_t src_pitch, int gw,
                     const float* __restrict__ lut, float* __restrict__ out, int64_t row0, int n_rows) {
    const int q = blockIdx.y * blockDim.x + threadIdx.x;      // 16-byte chunk inside the row
    const int r = blockIdx.x;                                 // patch row of this frame
    if (q >= VIS_ROW_FLOATS / 4 || r >= n_rows) return;
    // invert row = ((bh*(gw/2) + bw)*2 + mh)*2 + mw   (tf:...pil_qwen2_vl.py:198-214)
    const int mw = r & 1, mh = (r >> 1) & 1, blk = r >> 2;
    const int bw = blk % (gw / 2), bh = blk / (gw / 2);
    const int gy = bh * 2 + mh, gx = bw * 2 + mw;
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int col = q * 4 + e;                            // col = ((c*2 + t)*14 + py)*14 + px
        const int c = col / 392, rem = col % 392, in_plane = rem % 196;
        const int py = in_plane / 14, px = in_plane % 14;
        const uint8_t s = __ldg(src + (size_t)(gy * 14 + py) * src_pitch + (size_t)(gx * 14 + px) * 3 + c);
        v[e] = __ldg(lut + (int)s * 3 + c);
    }
    *reinterpret_cast<float4*>(out + (size_t)(row0 + r) * VIS_ROW_FLOATS + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
}


// Image.reduce((fx, fy), box) — libImaging/Reduce.c.  One thread per output pixel (all channels): fy rows of fx * CH
// contiguous bytes; a warp covers 32 * fx * CH contiguous bytes of every source row it touches, so the source is read
// once with full sectors.  The four cell kinds (interior, clipped at the right edge, at the bottom edge, at both) have
// their own sample count n, multiplier M(n) and rounding term n / 2, computed on the host exactly as Reduce.c does
// (float division, truncated).
struct ReduceParams {
    int x0, y0, x1, y1, fx, fy, ow, oh;
    unsigned mult[4], amend[4];           // index = (clipped column) | (clipped row) << 1
};

template <int CH>
__global__ void __launch_bounds__(256)
k_reduce(const uint8_t* __restrict__ src, int64_t src_pitch, uint8_t* __restrict__ dst, int64_t dst_pitch,
         const ReduceParams p) {
    const int ox = blockIdx.x * 64 + (threadIdx.x & 63), oy = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (ox >= p.ow || oy >= p.oh) return;
    const int xs = p.x0 + ox * p.fx, xe = min(xs + p.fx, p.x1);
    const int ys = p.y0 + oy * p.fy, ye = min(ys + p.fy, p.y1);
    const int kind = (xe - xs < p.fx ? 1 : 0) | (ye - ys < p.fy ? 2 : 0);
    unsigned ss[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) ss[c] = p.amend[kind];
    for (int y = ys; y < ye; ++y) {
        const uint8_t* row = src + (int64_t)y * src_pitch + (int64_t)xs * CH;
        for (int i = 0; i < (xe - xs) * CH; i += CH)
#pragma unroll
            for (int c = 0; c < CH; ++c) ss[c] += __ldg(row + i + c);
    }
    uint8_t* d = dst + (int64_t)oy * dst_pitch + (int64_t)ox * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) d[c] = (uint8_t)((ss[c] * p.mult[kind]) >> 24);
}

inline unsigned reduce_multiplier(int n) {        // Reduce.c division_UINT32(n, 8)
    const uint32_t max_dividend = (uint32_t)(1 << 8) * (uint32_t)n;
    const float max_int = (1 << 30) * 4.0f;
    return (unsigned)(max_int / (float)max_dividend);
}


// Image.resize(NEAREST): a gather through host-built index tables (-1 = outside the image: 0)
__global__ void __launch_bounds__(256)
k_gather(const uint8_t* __restrict__ src, int64_t src_pitch, int ch, uint8_t* __restrict__ dst, int64_t dst_pitch,
         int out_h, int out_w, const int32_t* __restrict__ xtab, const int32_t* __restrict__ ytab) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= out_w || y >= out_h) return;
    const int sx = __ldg(xtab + x), sy = __ldg(ytab + y);
    uint8_t* d = dst + (int64_t)y * dst_pitch + (int64_t)x * ch;
    const uint8_t* s = src + (int64_t)max(sy, 0) * src_pitch + (int64_t)max(sx, 0) * ch;
    const bool inside = sx >= 0 && sy >= 0;
    for (int c = 0; c < ch; ++c) d[c] = inside ? __ldg(s + c) : (uint8_t)0;
}


// Convert.c rgbA2rgba / rgba2rgbA (and the LA pair): premultiply / un-premultiply by the last channel, in place
template <int CH, bool FORWARD>
__global__ void __launch_bounds__(256)
k_alpha(uint8_t* __restrict__ img, int64_t pitch, int h, int w) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= w || y >= h) return;
    uint8_t* p = img + (int64_t)y * pitch + (int64_t)x * CH;
    const unsigned a = p[CH - 1];
    if (!FORWARD && (a == 0 || a == 255)) return;
#pragma unroll
    for (int c = 0; c < CH - 1; ++c) {
        const unsigned v = p[c];
        if (FORWARD) {
            const unsigned t = v * a + 128;
            p[c] = (uint8_t)(((t >> 8) + t) >> 8);
        } else {
            p[c] = (uint8_t)min(255u, 255u * v / a);
        }
    }
}

// Row re-pitch: frames whose base or row pitch is not a multiple of 16 bytes (502-pixel rows: 1506 bytes) cannot be
// staged by the fused kernels' bulk copies; this copies them, a batch per launch, into a staging buffer with an aligned
// pitch.  thread = 16 destination bytes: the covering aligned source words are funnel-shifted into place, one 16-byte store.
__global__ void __launch_bounds__(256) k_repitch(const VisRepitch* __restrict__ descs) {
    const VisRepitch r = descs[blockIdx.y];
    const int chunks = (r.row_bytes + 15) >> 4;
    const long long total = (long long)r.rows * chunks;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int row = (int)(i / chunks), c = (int)(i - (long long)row * chunks);
        const int n = min(16, r.row_bytes - 16 * c);
        const uintptr_t a = (uintptr_t)(r.src + (size_t)row * r.src_pitch + 16 * c);
        const int sh = (int)(a & 3);
        const uint32_t* w = reinterpret_cast<const uint32_t*>(a - sh);
        const int nwords = (sh + n + 3) >> 2;                   // aligned words holding the n bytes
        uint32_t v[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) v[q] = q < nwords ? __ldg(w + q) : 0u;
        uint4 o;
        o.x = __funnelshift_r(v[0], v[1], 8 * sh); o.y = __funnelshift_r(v[1], v[2], 8 * sh);
        o.z = __funnelshift_r(v[2], v[3], 8 * sh); o.w = __funnelshift_r(v[3], v[4], 8 * sh);
        uint8_t* d = r.dst + (size_t)row * r.dst_pitch + 16 * c;
        if (n == 16) {
            *reinterpret_cast<uint4*>(d) = o;
        } else {                                                  // row tail: only the row's own bytes are written
            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
            for (int b = 0; b < n; ++b) d[b] = (uint8_t)(ow[b >> 2] >> (8 * (b & 3)));
        }
    }
}

// Pillow's double-precision resample passes for the modes that are not 8 bits per channel (libImaging/Resample.c
// ImagingResampleHorizontal/Vertical_16bpc and _32bpc — what img.resize(..., LANCZOS) at utils/image_utils.py:75 runs for
// "I;16", "I" and "F" frames): ss = sum of pixel * k[x] in tap order, one rounding per multiply and per add (no FMA: the
// reference's wheels target baseline x86-64), then the mode's conversion.  One thread per output sample; correctness path.
__device__ __forceinline__ int round_up_x86(double f) {                 // ROUND_UP + x86 cvttsd2si (out of range: INT_MIN)
    const double v = f >= 0.0 ? __dadd_rn(f, 0.5) : __dadd_rn(f, -0.5);
    if (!(v > -2147483649.0 && v < 2147483648.0)) return INT_MIN;
    return (int)v;
}
__device__ __forceinline__ int clip8_int(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

template <int KIND, bool VERT>
__global__ void __launch_bounds__(256)
k_resample_hp(const uint8_t* __restrict__ src, int64_t src_pitch, uint8_t* __restrict__ dst, int64_t dst_pitch, int out_w,
              int out_h, const double* __restrict__ k, const int* __restrict__ bounds, int ksize) {
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= out_w || y >= out_h) return;
    const int o = VERT ? y : x;
    const int first = bounds[2 * o], taps = bounds[2 * o + 1];
    const double* kk = k + (size_t)o * ksize;
    constexpr int BPP = KIND <= VIS_HP_U16BE ? 2 : 4;
    double ss = 0.0;
    for (int t = 0; t < taps; ++t) {
        const uint8_t* p = VERT ? src + (size_t)(first + t) * src_pitch + (size_t)x * BPP
                                : src + (size_t)y * src_pitch + (size_t)(first + t) * BPP;
        double v;
        if (KIND == VIS_HP_U16LE) v = (double)(p[0] + (p[1] << 8));
        else if (KIND == VIS_HP_U16BE) v = (double)(p[1] + (p[0] << 8));
        else if (KIND == VIS_HP_I32) v = (double)*reinterpret_cast<const int*>(p);
        else v = (double)*reinterpret_cast<const float*>(p);
        ss = __dadd_rn(ss, __dmul_rn(v, kk[t]));
    }
    uint8_t* q = dst + (size_t)y * dst_pitch + (size_t)x * BPP;
    if (KIND <= VIS_HP_U16BE) {
        const int r = round_up_x86(ss);
        const uint8_t lo = (uint8_t)clip8_int(r % 256), hi = (uint8_t)clip8_int(r >> 8);     // Pillow's two CLIP8s, as is
        q[KIND == VIS_HP_U16LE ? 0 : 1] = lo;
        q[KIND == VIS_HP_U16LE ? 1 : 0] = hi;
    } else if (KIND == VIS_HP_I32) {
        *reinterpret_cast<int*>(q) = round_up_x86(ss);
    } else {
        *reinterpret_cast<float*>(q) = (float)ss;
    }
}

}  // namespace

extern "C" {

int vis_resample_h_u8(const uint8_t* src, int64_t src_pitch, int rows, int in_w, int channels,
                      uint8_t* dst, int64_t dst_pitch, int out_w,
                      const int32_t* k, const int32_t* bounds, int ksize, void* stream) {
    if (!src || !dst || !k || !bounds || rows <= 0 || in_w <= 0 || out_w <= 0 || ksize <= 0 ||
        src_pitch < (int64_t)in_w * channels || dst_pitch < (int64_t)out_w * channels) {
        vis::set_error("vis_resample_h_u8: bad arguments");
        return VIS_E_INVALID;
    }
    dim3 block(64, 4), grid((out_w + 63) / 64, (rows + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    switch (channels) {
        case 1: k_resample_h<1><<<grid, block, 0, st>>>(src, src_pitch, rows, dst, dst_pitch, out_w, k, bounds, ksize); break;
        case 2: k_resample_h<2><<<grid, block, 0, st>>>(src, src_pitch, rows, dst, dst_pitch, out_w, k, bounds, ksize); break;
        case 3: k_resample_h<3><<<grid, block, 0, st>>>(src, src_pitch, rows, dst, dst_pitch, out_w, k, bounds, ksize); break;
        case 4: k_resample_h<4><<<grid, block, 0, st>>>(src, src_pitch, rows, dst, dst_pitch, out_w, k, bounds, ksize); break;
        default:
            vis::set_error("vis_resample_h_u8: %d channels unsupported (1..4)", channels);
            return VIS_E_INVALID;
    }
    return vis::check_launch("vis_resample_h_u8");
}

int vis_resample_v_u8(const uint8_t* src, int64_t src_pitch, int in_h, int row_bytes,
                      uint8_t* dst, int64_t dst_pitch, int out_h,
                      const int32_t* k, const int32_t* bounds, int ksize, void* stream) {
    if (!src || !dst || !k || !bounds || in_h <= 0 || row_bytes <= 0 || out_h <= 0 || ksize <= 0 ||
        src_pitch < row_bytes || dst_pitch < row_bytes) {
        vis::set_error("vis_resample_v_u8: bad arguments");
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = (row_bytes % 4 == 0) && (src_pitch % 4 == 0) && (dst_pitch % 4 == 0) &&
                     (((uintptr_t)src | (uintptr_t)dst) % 4 == 0);
    dim3 block(64, 4);
    if (vec) {
        dim3 grid((row_bytes / 4 + 63) / 64, (out_h + 3) / 4);
        k_resample_v<4><<<grid, block, 0, st>>>(src, src_pitch, row_bytes, dst, dst_pitch, out_h, k, bounds, ksize);
    } else {
        dim3 grid((row_bytes + 63) / 64, (out_h + 3) / 4);
        k_resample_v<1><<<grid, block, 0, st>>>(src, src_pitch, row_bytes, dst, dst_pitch, out_h, k, bounds, ksize);
    }
    return vis::check_launch("vis_resample_v_u8");
}

int vis_reduce_u8(const uint8_t* src, int64_t src_pitch, int h, int w, int channels, int fx, int fy,
                  int x0, int y0, int x1, int y1, uint8_t* dst, int64_t dst_pitch, void* stream) {
    if (!src || !dst || h <= 0 || w <= 0 || channels < 1 || channels > 4 || fx < 1 || fy < 1 || x0 < 0 || y0 < 0 ||
        x1 > w || y1 > h || x1 <= x0 || y1 <= y0 || src_pitch < (int64_t)w * channels || (int64_t)fx * fy > (1 << 16)) {
        vis::set_error("vis_reduce_u8: bad arguments (%dx%d, factor %dx%d, box %d,%d,%d,%d)", w, h, fx, fy, x0, y0, x1, y1);
        return VIS_E_INVALID;
    }
    ReduceParams p;
    p.x0 = x0; p.y0 = y0; p.x1 = x1; p.y1 = y1; p.fx = fx; p.fy = fy;
    p.ow = (x1 - x0 + fx - 1) / fx;
    p.oh = (y1 - y0 + fy - 1) / fy;
    if (dst_pitch < (int64_t)p.ow * channels) {
        vis::set_error("vis_reduce_u8: destination pitch %lld < %d", (long long)dst_pitch, p.ow * channels);
        return VIS_E_INVALID;
    }
    const int rx = (x1 - x0) % fx ? (x1 - x0) % fx : fx, ry = (y1 - y0) % fy ? (y1 - y0) % fy : fy;
    const int n[4] = {fx * fy, rx * fy, fx * ry, rx * ry};
    for (int i = 0; i < 4; ++i) { p.mult[i] = reduce_multiplier(n[i]); p.amend[i] = (unsigned)n[i] / 2; }
    const dim3 grid((p.ow + 63) / 64, (p.oh + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    switch (channels) {
        case 1: k_reduce<1><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, p); break;
        case 2: k_reduce<2><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, p); break;
        case 3: k_reduce<3><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, p); break;
        default: k_reduce<4><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, p); break;
    }
    return vis::check_launch("vis_reduce_u8");
}

int vis_nearest_table(int in_size, float in0, float in1, int out_size, int32_t* tab) {
    if (in_size <= 0 || out_size <= 0 || !tab) {
        vis::set_error("vis_nearest_table: bad arguments (in=%d out=%d)", in_size, out_size);
        return VIS_E_INVALID;
    }
    const double a = (double)(in1 - in0) / out_size;       // float subtraction, as _imaging.c _resize
    double xo = (double)in0 + a * 0.5;
    for (int x = 0; x < out_size; ++x) {
        const int xin = xo < 0.0 ? -1 : (int)xo;
        tab[x] = (xin >= 0 && xin < in_size) ? xin : -1;
        xo += a;                                            // accumulated, as ImagingScaleAffine
    }
    return VIS_OK;
}

int vis_gather_u8(const uint8_t* src, int64_t src_pitch, int h, int w, int channels, uint8_t* dst, int64_t dst_pitch,
                  int out_h, int out_w, const int32_t* xtab, const int32_t* ytab, void* stream) {
    if (!src || !dst || !xtab || !ytab || h <= 0 || w <= 0 || out_h <= 0 || out_w <= 0 || channels < 1 || channels > 4 ||
        src_pitch < (int64_t)w * channels || dst_pitch < (int64_t)out_w * channels) {
        vis::set_error("vis_gather_u8: bad arguments (%dx%d -> %dx%d, %d channels)", w, h, out_w, out_h, channels);
        return VIS_E_INVALID;
    }
    const dim3 grid((out_w + 63) / 64, (out_h + 3) / 4);
    k_gather<<<grid, 256, 0, (cudaStream_t)stream>>>(src, src_pitch, channels, dst, dst_pitch, out_h, out_w, xtab, ytab);
    return vis::check_launch("vis_gather_u8");
}

int vis_alpha_premultiply_u8(uint8_t* img, int64_t pitch, int h, int w, int channels, int forward, void* stream) {
    if (!img || h <= 0 || w <= 0 || (channels != 2 && channels != 4) || pitch < (int64_t)w * channels) {
        vis::set_error("vis_alpha_premultiply_u8: bad arguments (%dx%d, %d channels)", w, h, channels);
        return VIS_E_INVALID;
    }
    const dim3 grid((w + 63) / 64, (h + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 4) {
        if (forward) k_alpha<4, true><<<grid, 256, 0, st>>>(img, pitch, h, w);
        else         k_alpha<4, false><<<grid, 256, 0, st>>>(img, pitch, h, w);
    } else {
        if (forward) k_alpha<2, true><<<grid, 256, 0, st>>>(img, pitch, h, w);
        else         k_alpha<2, false><<<grid, 256, 0, st>>>(img, pitch, h, w);
    }
    return vis::check_launch("vis_alpha_premultiply_u8");
}

int vis_normalize_patchify(const uint8_t* src, int64_t src_pitch, int h, int w,
                           const float* lut768, float* pixel_values, int64_t row0, void* stream) {
    if (!src || !lut768 || !pixel_values || h <= 0 || w <= 0 || h % 28 || w % 28 || src_pitch < (int64_t)w * 3 ||
        row0 < 0 || ((uintptr_t)pixel_values % 16)) {
        vis::set_error("vis_normalize_patchify: bad arguments (h=%d w=%d must be multiples of 28)", h, w);
        return VIS_E_INVALID;
    }
    const int gh = h / VIS_PATCH, gw = w / VIS_PATCH, n_rows = gh * gw;
    dim3 block(128), grid(n_rows, (VIS_ROW_FLOATS / 4 + 127) / 128);
    k_normalize_patchify<<<grid, block, 0, (cudaStream_t)stream>>>(src, src_pitch, gw, lut768, pixel_values, row0, n_rows);
    return vis::check_launch("vis_normalize_patchify");
}

int vis_repitch_u8(const VisRepitch* descs, int n, int64_t max_frame_bytes, void* stream) {
    if (!descs || n <= 0 || n > 65535 || max_frame_bytes <= 0) {
        vis::set_error("vis_repitch_u8: bad arguments (n=%d)", n);
        return VIS_E_INVALID;
    }
    const int blocks = (int)std::min<int64_t>(1024, (max_frame_bytes / 16 + 255) / 256 + 1);
    k_repitch<<<dim3(blocks, n), 256, 0, (cudaStream_t)stream>>>(descs);
    return vis::check_launch("vis_repitch_u8");
}

int vis_resample_hp(const uint8_t* src, int64_t src_pitch, int h, int w, int kind, int vertical,
                    uint8_t* dst, int64_t dst_pitch, int out_size, const double* k, const int32_t* bounds, int ksize,
                    void* stream) {
    const int bpp = (kind == VIS_HP_U16LE || kind == VIS_HP_U16BE) ? 2 : 4;
    if (!src || !dst || !k || !bounds || h <= 0 || w <= 0 || out_size <= 0 || ksize <= 0 || kind < VIS_HP_U16LE || kind > VIS_HP_F32 ||
        src_pitch < (int64_t)w * bpp || dst_pitch < (int64_t)(vertical ? w : out_size) * bpp ||
        ((kind >= VIS_HP_I32) && (((uintptr_t)src | (uintptr_t)dst | (uintptr_t)src_pitch | (uintptr_t)dst_pitch) & 3))) {
        vis::set_error("vis_resample_hp: bad arguments (%dx%d -> %d, kind %d)", w, h, out_size, kind);
        return VIS_E_INVALID;
    }
    const int out_w = vertical ? w : out_size, out_h = vertical ? out_size : h;
    const dim3 grid((out_w + 63) / 64, (out_h + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
#define VIS_HP(K) (vertical ? k_resample_hp<K, true><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, out_w, out_h, k, bounds, ksize) \
                            : k_resample_hp<K, false><<<grid, 256, 0, st>>>(src, src_pitch, dst, dst_pitch, out_w, out_h, k, bounds, ksize))
    switch (kind) {
        case VIS_HP_U16LE: VIS_HP(VIS_HP_U16LE); break;
        case VIS_HP_U16BE: VIS_HP(VIS_HP_U16BE); break;
        case VIS_HP_I32:   VIS_HP(VIS_HP_I32); break;
        default:           VIS_HP(VIS_HP_F32); break;
    }
#undef VIS_HP
    return vis::check_launch("vis_resample_hp");
}

}  // extern "C"
