// vis_compose.cu — comparison panel of create_side_by_side_comparison (utils/image_utils.py:608-686 in the
// reference): two frames cv2.resize()d to a common height and laid side by side under a header bar, in ONE launch.
//
// cv2.resize with the default interpolation on 8-bit data is integer arithmetic (cv: modules/imgproc/src/resize.cpp):
//   * equal sizes: a copy;
//   * both scale factors exactly 2: OpenCV switches INTER_LINEAR to INTER_AREA's fast path, (a + b + c + d + 2) >> 2;
//   * otherwise the fixed-point bilinear resizer: per-axis index + two 11-bit weights (host tables of
//     vis_linear_table), horizontal pass in int32, vertical pass
//         (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// All of it is evaluated with OpenCV's own integer operations, so the result equals cv2's bit for bit.
//
// A block owns a 128 x 32 tile of the canvas.  A tile that lies inside ONE bilinear panel takes the two-pass route
// OpenCV itself takes: the horizontal pass of every source row the tile needs (<= 56), `(r0 >> 4)` as 16-bit values in
// shared memory (planar, so both passes are conflict free), then the vertical pass — each source row is interpolated
// once however many output rows use it, and the per-pixel work is six IMAD.HI instead of twenty-four byte loads.  Every
// other tile (header, divider, panel borders, copies, the area path) is evaluated pixel by pixel: a thread owns 4
// consecutive canvas pixels (12 bytes, three 32-bit stores when the row is aligned).
// Bound: HBM (both sources read once, canvas written once).
#include <cfloat>
#include <cmath>

#include "vis_internal.h"

namespace {

constexpr int kMaxPanels = 4;
constexpr int kPx = 4;

struct PanelSet {
    VisPanel p[kMaxPanels];
    int n;
};

__device__ __forceinline__ void panel_pixel(const VisPanel& p, int x, int y, int* out) {
    // (x, y): position inside the panel
    if (p.mode == VIS_RESIZE_COPY) {
        const uint8_t* s = p.src + (int64_t)y * p.src_pitch + (int64_t)x * 3;
        out[0] = __ldg(s); out[1] = __ldg(s + 1); out[2] = __ldg(s + 2);
        return;
    }
    if (p.mode == VIS_RESIZE_AREA2) {
        const uint8_t* r0 = p.src + (int64_t)(2 * y) * p.src_pitch + (int64_t)(2 * x) * 3;
        const uint8_t* r1 = r0 + p.src_pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[c] = ((int)__ldg(r0 + c) + (int)__ldg(r0 + 3 + c) + (int)__ldg(r1 + c) + (int)__ldg(r1 + 3 + c) + 2) >> 2;
        return;
    }
    const int sx = __ldg(p.xofs + x), sx1 = min(sx + 1, p.src_w - 1);      // the weight of sx1 is 0 at the border
    const int a0 = __ldg(p.alpha + 2 * x), a1 = __ldg(p.alpha + 2 * x + 1);
    const int sy = __ldg(p.yofs + y);
    const int b0 = __ldg(p.beta + 2 * y), b1 = __ldg(p.beta + 2 * y + 1);
    const uint8_t* r0 = p.src + (int64_t)min(max(sy, 0), p.src_h - 1) * p.src_pitch;
    const uint8_t* r1 = p.src + (int64_t)min(max(sy + 1, 0), p.src_h - 1) * p.src_pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int h0 = (int)__ldg(r0 + sx * 3 + c) * a0 + (int)__ldg(r0 + sx1 * 3 + c) * a1;
        const int h1 = (int)__ldg(r1 + sx * 3 + c) * a0 + (int)__ldg(r1 + sx1 * 3 + c) * a1;
        out[c] = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    }
}

// one thread = four consecutive canvas pixels of a row
__device__ __forceinline__ void compose_row4(uint8_t* __restrict__ canvas, int64_t pitch, int h, int w, int fill,
                                             const VisPanel* __restrict__ panels, int n_panels, int x0, int y) {
    if (y >= h || x0 >= w) return;
    int v[kPx][3];
#pragma unroll
    for (int j = 0; j < kPx; ++j) {
        const int x = x0 + j;
        v[j][0] = v[j][1] = v[j][2] = fill;
        for (int k = 0; k < n_panels; ++k) {
            const VisPanel& p = panels[k];
            const int px = x - p.org_x, py = y - p.org_y;
            if (px >= 0 && px < p.dst_w && py >= 0 && py < p.dst_h) { panel_pixel(p, px, py, v[j]); break; }
        }
    }
    uint8_t* d = canvas + (int64_t)y * pitch + (int64_t)x0 * 3;
    if (x0 + kPx <= w && (((uintptr_t)d) & 3) == 0) {
        uint32_t* q = reinterpret_cast<uint32_t*>(d);
        q[0] = (uint32_t)v[0][0] | ((uint32_t)v[0][1] << 8) | ((uint32_t)v[0][2] << 16) | ((uint32_t)v[1][0] << 24);
        q[1] = (uint32_t)v[1][1] | ((uint32_t)v[1][2] << 8) | ((uint32_t)v[2][0] << 16) | ((uint32_t)v[2][1] << 24);
        q[2] = (uint32_t)v[2][2] | ((uint32_t)v[3][0] << 8) | ((uint32_t)v[3][1] << 16) | ((uint32_t)v[3][2] << 24);
    } else {
        for (int j = 0; j < kPx && x0 + j < w; ++j)
#pragma unroll
            for (int c = 0; c < 3; ++c) d[j * 3 + c] = (uint8_t)v[j][c];
    }
}

#ifndef VIS_PANEL_TILE_H
#define VIS_PANEL_TILE_H 32
#endif
#ifndef VIS_PANEL_SRC_ROWS
#define VIS_PANEL_SRC_ROWS 56
#endif
constexpr int kTileW = 128, kTileH = VIS_PANEL_TILE_H, kMaxSrcRows = VIS_PANEL_SRC_ROWS;
static_assert(kMaxSrcRows * 3 * kTileW * 2 <= 46 * 1024, "static shared memory");

// the tile at (x0, y0) of a canvas; `hs` = kMaxSrcRows x 3 x kTileW uint16 of shared memory
__device__ __forceinline__ void compose_tile(uint8_t* __restrict__ canvas, int64_t pitch, int h, int w, int fill,
                                             const VisPanel* __restrict__ panels, int n_panels, int x0, int y0, uint16_t* hs) {
    const int tid = threadIdx.x;
    // ---- is the whole tile inside one bilinear panel, with few enough source rows?  (uniform)
    int which = -1;
    if (x0 + kTileW <= w && y0 + kTileH <= h) {
        for (int k = 0; k < n_panels; ++k) {
            const VisPanel& p = panels[k];
            if (p.mode == VIS_RESIZE_BILINEAR && x0 >= p.org_x && x0 + kTileW <= p.org_x + p.dst_w && y0 >= p.org_y &&
                y0 + kTileH <= p.org_y + p.dst_h)
                which = k;
        }
    }
    int sy_first = 0, n_src = 0;
    if (which >= 0) {
        const VisPanel& p = panels[which];
        const int py0 = y0 - p.org_y;
        sy_first = min(max(__ldg(p.yofs + py0), 0), p.src_h - 1);
        n_src = min(max(__ldg(p.yofs + py0 + kTileH - 1) + 1, 0), p.src_h - 1) - sy_first + 1;
        if (n_src > kMaxSrcRows) which = -1;
    }
    if (which < 0) {                              // pixel by pixel: 32 threads x 4 pixels per row, 8 rows per step
        for (int r = tid >> 5; r < kTileH; r += 8) compose_row4(canvas, pitch, h, w, fill, panels, n_panels, x0 + (tid & 31) * kPx, y0 + r);
        return;
    }
    const VisPanel& p = panels[which];
    const int px0 = x0 - p.org_x, py0 = y0 - p.org_y;
    {   // ---- horizontal pass: thread = output column, two threads share a column's source rows
        const int cx = tid & (kTileW - 1);
        const int sx = __ldg(p.xofs + px0 + cx), sx1 = min(sx + 1, p.src_w - 1);      // the weight of sx1 is 0 at the border
        const int a0 = __ldg(p.alpha + 2 * (px0 + cx)), a1 = __ldg(p.alpha + 2 * (px0 + cx) + 1);
        const uint8_t* c0 = p.src + (int64_t)sx * 3 + (int64_t)(sy_first + (tid >> 7)) * p.src_pitch;
        const uint8_t* c1 = p.src + (int64_t)sx1 * 3 + (int64_t)(sy_first + (tid >> 7)) * p.src_pitch;
        const int64_t step = 2 * p.src_pitch;
        uint16_t* o = hs + (tid >> 7) * 3 * kTileW + cx;
#pragma unroll 2
        for (int r = tid >> 7; r < n_src; r += 2, c0 += step, c1 += step, o += 6 * kTileW) {
#pragma unroll
            for (int c = 0; c < 3; ++c) o[c * kTileW] = (uint16_t)(((int)__ldg(c0 + c) * a0 + (int)__ldg(c1 + c) * a1) >> 4);
        }
    }
    __syncthreads();
    // ---- vertical pass: thread = 4 consecutive pixels of a row; (b * v) >> 16 of non-negative values = mulhi(b << 16, v)
    const int gx = tid & 31;
    for (int ry = tid >> 5; ry < kTileH; ry += 8) {
        const int py = py0 + ry;
        const int sy = __ldg(p.yofs + py);
        const int r0 = min(max(sy, 0), p.src_h - 1) - sy_first, r1 = min(max(sy + 1, 0), p.src_h - 1) - sy_first;
        const unsigned b0 = (unsigned)__ldg(p.beta + 2 * py) << 16, b1 = (unsigned)__ldg(p.beta + 2 * py + 1) << 16;
        int v[kPx][3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const uint2 u0 = *reinterpret_cast<const uint2*>(hs + (r0 * 3 + c) * kTileW + gx * kPx);
            const uint2 u1 = *reinterpret_cast<const uint2*>(hs + (r1 * 3 + c) * kTileW + gx * kPx);
            const unsigned t0[4] = {u0.x & 0xffffu, u0.x >> 16, u0.y & 0xffffu, u0.y >> 16};
            const unsigned t1[4] = {u1.x & 0xffffu, u1.x >> 16, u1.y & 0xffffu, u1.y >> 16};
#pragma unroll
            for (int j = 0; j < kPx; ++j) v[j][c] = (int)(__umulhi(b0, t0[j]) + __umulhi(b1, t1[j]) + 2u) >> 2;
        }
        uint8_t* d = canvas + (int64_t)(y0 + ry) * pitch + (int64_t)(x0 + gx * kPx) * 3;
        if ((((uintptr_t)d) & 3) == 0) {
            uint32_t* q = reinterpret_cast<uint32_t*>(d);
            q[0] = (uint32_t)v[0][0] | ((uint32_t)v[0][1] << 8) | ((uint32_t)v[0][2] << 16) | ((uint32_t)v[1][0] << 24);
            q[1] = (uint32_t)v[1][1] | ((uint32_t)v[1][2] << 8) | ((uint32_t)v[2][0] << 16) | ((uint32_t)v[2][1] << 24);
            q[2] = (uint32_t)v[2][2] | ((uint32_t)v[3][0] << 8) | ((uint32_t)v[3][1] << 16) | ((uint32_t)v[3][2] << 24);
        } else {
#pragma unroll
            for (int j = 0; j < kPx; ++j)
#pragma unroll
                for (int c = 0; c < 3; ++c) d[j * 3 + c] = (uint8_t)v[j][c];
        }
    }
}

__global__ void __launch_bounds__(256)
k_compose_panels(uint8_t* __restrict__ canvas, int64_t pitch, int h, int w, int fill, const __grid_constant__ PanelSet ps) {
    __shared__ __align__(16) uint16_t hs[kMaxSrcRows * 3 * kTileW];
    compose_tile(canvas, pitch, h, w, fill, ps.p, ps.n, blockIdx.x * kTileW, blockIdx.y * kTileH, hs);
}

// a batch of canvases, one per blockIdx.z: the canvas record is read into shared memory once per block
__global__ void __launch_bounds__(256)
k_compose_panels_batch(const VisPanelCanvas* __restrict__ canvases) {
    __shared__ __align__(16) uint16_t hs[kMaxSrcRows * 3 * kTileW];
    __shared__ VisPanelCanvas cv;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(canvases + blockIdx.z);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&cv);
        for (int i = threadIdx.x; i < (int)(sizeof(VisPanelCanvas) / 4); i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if ((int)blockIdx.y * kTileH >= cv.h || (int)blockIdx.x * kTileW >= cv.w) return;
    compose_tile(cv.canvas, cv.pitch, cv.h, cv.w, cv.fill, cv.panels, min(max(cv.n_panels, 0), kMaxPanels),
                 blockIdx.x * kTileW, blockIdx.y * kTileH, hs);
}

inline int16_t sat_short(float v) {
    const long r = lrintf(v);                     // cvRound: half to even
    return (int16_t)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
}

}  // namespace

extern "C" int vis_resize_linear_mode(int src_h, int src_w, int dst_h, int dst_w) {
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) {
        vis::set_error("vis_resize_linear_mode: bad sizes (%dx%d -> %dx%d)", src_w, src_h, dst_w, dst_h);
        return VIS_E_INVALID;
    }
    if (src_h == dst_h && src_w == dst_w) return VIS_RESIZE_COPY;
    const double scale_x = 1. / ((double)dst_w / src_w), scale_y = 1. / ((double)dst_h / src_h);
    const long ix = lrint(scale_x), iy = lrint(scale_y);
    const bool area_fast = std::fabs(scale_x - ix) < DBL_EPSILON && std::fabs(scale_y - iy) < DBL_EPSILON;
    return area_fast && ix == 2 && iy == 2 ? VIS_RESIZE_AREA2 : VIS_RESIZE_BILINEAR;
}

extern "C" int vis_linear_table(int src_size, int dst_size, int is_x, int32_t* ofs, int16_t* coef) {
    if (src_size <= 0 || dst_size <= 0 || !ofs || !coef) {
        vis::set_error("vis_linear_table: bad arguments (src=%d dst=%d)", src_size, dst_size);
        return VIS_E_INVALID;
    }
    const double scale = 1. / ((double)dst_size / src_size);
    for (int d = 0; d < dst_size; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (is_x) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= src_size - 1) { f = 0.f; s = src_size - 1; }
        }
        ofs[d] = s;
        coef[2 * d] = sat_short((1.f - f) * 2048.f);
        coef[2 * d + 1] = sat_short(f * 2048.f);
    }
    return VIS_OK;
}

extern "C" int vis_compose_panels(uint8_t* canvas, int64_t canvas_pitch, int h, int w, int fill,
                                  const VisPanel* panels, int n_panels, void* stream) {
    if (!canvas || h <= 0 || w <= 0 || canvas_pitch < (int64_t)w * 3 || n_panels < 0 || n_panels > kMaxPanels ||
        (n_panels && !panels) || fill < 0 || fill > 255) {
        vis::set_error("vis_compose_panels: bad arguments (h=%d w=%d panels=%d)", h, w, n_panels);
        return VIS_E_INVALID;
    }
    PanelSet ps;
    ps.n = n_panels;
    for (int i = 0; i < n_panels; ++i) {
        const VisPanel& p = panels[i];
        const int mode = vis_resize_linear_mode(p.src_h, p.src_w, p.dst_h, p.dst_w);
        if (!p.src || mode < 0 || mode != p.mode || p.src_pitch < (int64_t)p.src_w * 3 ||
            (mode == VIS_RESIZE_BILINEAR && (!p.xofs || !p.alpha || !p.yofs || !p.beta))) {
            vis::set_error("vis_compose_panels: panel %d is inconsistent (mode %d, expected %d)", i, p.mode, mode);
            return VIS_E_INVALID;
        }
        ps.p[i] = p;
    }
    const dim3 grid((w + kTileW - 1) / kTileW, (h + kTileH - 1) / kTileH);
    if (grid.y > 65535) {
        vis::set_error("vis_compose_panels: a canvas of %d rows is beyond the grid", h);
        return VIS_E_UNSUPPORTED;
    }
    k_compose_panels<<<grid, 256, 0, (cudaStream_t)stream>>>(canvas, canvas_pitch, h, w, fill, ps);
    return vis::check_launch("vis_compose_panels");
}

extern "C" int vis_compose_panels_batch(const VisPanelCanvas* canvases, int n_canvases, int max_h, int max_w, void* stream) {
    if (!canvases || n_canvases <= 0 || n_canvases > 65535 || max_h <= 0 || max_w <= 0) {
        vis::set_error("vis_compose_panels_batch: bad arguments (canvases=%d max %dx%d)", n_canvases, max_w, max_h);
        return VIS_E_INVALID;
    }
    const dim3 grid((max_w + kTileW - 1) / kTileW, (max_h + kTileH - 1) / kTileH, n_canvases);
    if (grid.y > 65535) {
        vis::set_error("vis_compose_panels_batch: canvases of %d rows are beyond the grid", max_h);
        return VIS_E_UNSUPPORTED;
    }
    k_compose_panels_batch<<<grid, 256, 0, (cudaStream_t)stream>>>(canvases);
    return vis::check_launch("vis_compose_panels_batch");
}
