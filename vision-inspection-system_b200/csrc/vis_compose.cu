// vis_compose.cu — comparison panel of create_side_by_side_comparison (utils/image_utils.py:608-686 in the
// reference): two frames cv2.resize()d to a common height and laid side by side under a header bar, in ONE launch.
//
// cv2.resize with the default interpolation on 8-bit data is integer arithmetic (cv: modules/imgproc/src/resize.cpp):
//   * equal sizes: a copy;
//   * both scale factors exactly 2: OpenCV switches INTER_LINEAR to INTER_AREA's fast path, (a + b + c + d + 2) >> 2;
//   * otherwise the fixed-point bilinear resizer: per-axis index + two 11-bit weights (host tables of
//     vis_linear_table), horizontal pass in int32, vertical pass
//         (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2.
// The kernel evaluates all of it per output pixel, so the result equals OpenCV's bit for bit.
//
// Bound: HBM (both sources read once, canvas written once); a thread owns 4 consecutive canvas pixels (12 bytes, three
// 32-bit stores when the row is aligned) and the four taps of neighbouring threads share L1 lines.
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
__device__ __forceinline__ void compose_row4(uint8_t* __restrict__ canvas, int64_t pitch, int h, int w, int fill, const PanelSet& ps,
                                             int x0, int y) {
    if (y >= h || x0 >= w) return;
    int v[kPx][3];
#pragma unroll
    for (int j = 0; j < kPx; ++j) {
        const int x = x0 + j;
        v[j][0] = v[j][1] = v[j][2] = fill;
        for (int k = 0; k < ps.n; ++k) {
            const VisPanel& p = ps.p[k];
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

__global__ void __launch_bounds__(256)
k_compose_panels(uint8_t* __restrict__ canvas, int64_t pitch, int h, int w, int fill, const PanelSet ps) {
    compose_row4(canvas, pitch, h, w, fill, ps, (blockIdx.x * 32 + (threadIdx.x & 31)) * kPx, blockIdx.y * 8 + (threadIdx.x >> 5));
}

// a batch of canvases, one per blockIdx.z: the canvas record is read into shared memory once per block
__global__ void __launch_bounds__(256)
k_compose_panels_batch(const VisPanelCanvas* __restrict__ canvases) {
    __shared__ VisPanelCanvas cv;
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(canvases + blockIdx.z);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&cv);
        for (int i = threadIdx.x; i < (int)(sizeof(VisPanelCanvas) / 4); i += 256) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    if ((int)blockIdx.y * 8 >= cv.h || (int)blockIdx.x * 32 * kPx >= cv.w) return;
    PanelSet ps;
    ps.n = min(max(cv.n_panels, 0), kMaxPanels);
#pragma unroll
    for (int k = 0; k < kMaxPanels; ++k) ps.p[k] = cv.panels[k];
    compose_row4(cv.canvas, cv.pitch, cv.h, cv.w, cv.fill, ps, (blockIdx.x * 32 + (threadIdx.x & 31)) * kPx,
                 blockIdx.y * 8 + (threadIdx.x >> 5));
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
    const dim3 grid((w + 32 * kPx - 1) / (32 * kPx), (h + 7) / 8);
    k_compose_panels<<<grid, 256, 0, (cudaStream_t)stream>>>(canvas, canvas_pitch, h, w, fill, ps);
    return vis::check_launch("vis_compose_panels");
}

extern "C" int vis_compose_panels_batch(const VisPanelCanvas* canvases, int n_canvases, int max_h, int max_w, void* stream) {
    if (!canvases || n_canvases <= 0 || n_canvases > 65535 || max_h <= 0 || max_w <= 0) {
        vis::set_error("vis_compose_panels_batch: bad arguments (canvases=%d max %dx%d)", n_canvases, max_w, max_h);
        return VIS_E_INVALID;
    }
    const dim3 grid((max_w + 32 * kPx - 1) / (32 * kPx), (max_h + 7) / 8, n_canvases);
    if (grid.y > 65535) {
        vis::set_error("vis_compose_panels_batch: canvases of %d rows are beyond the grid", max_h);
        return VIS_E_UNSUPPORTED;
    }
    k_compose_panels_batch<<<grid, 256, 0, (cudaStream_t)stream>>>(canvases);
    return vis::check_launch("vis_compose_panels_batch");
}
