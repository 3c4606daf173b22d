// vis_overlay.cu — device half of the defect-overlay rasteriser (cv2 drawing calls of
// utils/image_utils.py:259-313 in the reference, reproduced pixel for pixel).
//
// One CTA owns a 64x16 pixel tile of one frame (256 threads x 4 pixels, 12 bytes per thread kept in
// registers).  The tile scans the frame's GROUP headers (one per box) in parallel, then for every box whose
// bounding box touches the tile scans that box's leaves 256 at a time, compacts the touching ones IN ORDER
// into shared memory (warp ballots + a block prefix), and every thread applies them in list order to its own
// pixels.  Per-pixel in-order application is what makes the result identical to OpenCV's sequential drawing:
// fills and LINE_8 points overwrite, LineAA pixels blend (twice, 8-bit alpha) with whatever is there.
// Tiles that no box touches are a straight 12-byte-per-thread copy (or nothing at all when drawing in place).
#include "vis_internal.h"
#include "vis_overlay_leaf.h"

namespace {

constexpr int kTileW = 64, kTileH = 16, kThreads = 256, kPx = 4;

__constant__ int c_filter[64] = {
    168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
    158, 149, 140, 131, 122, 114, 105, 97, 89, 82, 75, 68, 62, 56, 50, 45,
    40, 36, 32, 28, 25, 22, 19, 16, 14, 12, 11, 9, 8, 7, 5, 5};

struct Tile { int x0, y0, x1, y1; };

__device__ __forceinline__ bool touches(const Tile& t, int wx, int wy) {
    const int bx0 = wx & 0xffff, bx1 = (unsigned)wx >> 16, by0 = wy & 0xffff, by1 = (unsigned)wy >> 16;
    return bx0 <= t.x1 && bx1 >= t.x0 && by0 <= t.y1 && by1 >= t.y0;
}

__device__ __forceinline__ void set_px(int* c, int col) {
    c[0] = col & 0xff; c[1] = (col >> 8) & 0xff; c[2] = (col >> 16) & 0xff;
}
__device__ __forceinline__ void blend_px(int* c, int col, int a) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int cc = (col >> (8 * k)) & 0xff;
        int v = c[k];
        v += ((cc - v) * a + 127) >> 8;
        v += ((cc - v) * a + 127) >> 8;       // OpenCV's ICV_PUT_POINT applies the blend twice
        c[k] = v;
    }
}
// coverage of the AA line at step s (scount) with e steps left (ecount); d = 0,1,2 selects the three pixels across
__device__ __forceinline__ int aa_alpha(const int* w, const int* filt, int s, int e, int64_t minor, int d) {
    const int idx = (((s >= 2) + 1) & (s | 2)) * 3 + (((e >= 2) + 1) & (e | 2));
    const int ep = (w[6 + idx / 3] >> (10 * (idx % 3))) & 0x3ff;
    const int dist = (int)((minor >> 11) & 31);
    const int f = d == 0 ? filt[dist + 32] : d == 1 ? filt[dist] : filt[63 - dist];
    return ((ep * f) >> 8) & 0xff;
}

// apply one leaf to the thread's 4 pixels (x .. x+3, row y); returns true when the leaf could have written
__device__ __forceinline__ bool apply_leaf(const int* w, const int* filt, int x, int y, int (*c)[3]) {
    const int bx0 = w[10] & 0xffff, bx1 = (unsigned)w[10] >> 16, by0 = w[11] & 0xffff, by1 = (unsigned)w[11] >> 16;
    if (y < by0 || y > by1 || x + kPx - 1 < bx0 || x > bx1) return false;
    const int kind = w[0] & LEAF_KIND_MASK, col = w[1];
    switch (kind) {
        case LEAF_TRAP: {
            const int d = y - w[2];
            if (d < 0 || y > w[3]) return false;
            const int64_t xa = (int64_t)w[4] + (int64_t)w[5] * d, xb = (int64_t)w[6] + (int64_t)w[7] * d;
            const int64_t lo = xa < xb ? xa : xb, hi = xa < xb ? xb : xa;
            const bool aa = w[0] & LEAF_FLAG_AA;
            const int xx1 = (int)((lo + (aa ? 65535 : 32768)) >> 16), xx2 = (int)((hi + (aa ? 0 : 32768)) >> 16);
#pragma unroll
            for (int j = 0; j < kPx; ++j)
                if (x + j >= xx1 && x + j <= xx2) set_px(c[j], col);
            return true;
        }
        case LEAF_SPANS: {
            const int r = y - w[3];
            if (r < 0 || r >= w[4]) return false;
            const int hw = (w[5 + (r >> 2)] >> (8 * (r & 3))) & 0xff;
            if (hw == 0xff) return false;
#pragma unroll
            for (int j = 0; j < kPx; ++j)
                if (abs(x + j - w[2]) <= hw) set_px(c[j], col);
            return true;
        }
        case LEAF_LINE8: {
            const int m0 = w[2], ecount = w[3];
            if (w[0] & LEAF_FLAG_XMAJOR) {
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const int i = x + j - m0;
                    const bool on = i >= 0 && i <= ecount && (int)(((int64_t)w[4] + (int64_t)w[5] * i) >> 16) == y;
                    if (on || (x + j == w[6] && y == w[7])) set_px(c[j], col);
                }
            } else {
                const int i = y - m0;
                const bool in = i >= 0 && i <= ecount;
                const int xx = (int)(((int64_t)w[4] + (int64_t)w[5] * i) >> 16);
#pragma unroll
                for (int j = 0; j < kPx; ++j)
                    if ((in && x + j == xx) || (x + j == w[6] && y == w[7])) set_px(c[j], col);
            }
            return true;
        }
        case LEAF_LINEAA: {
            const int m0 = w[2], e0 = w[3];
            if (w[0] & LEAF_FLAG_XMAJOR) {
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const int s = x + j - m0;
                    if (s < 0 || s > e0) continue;
                    const int64_t minor = (int64_t)w[4] + (int64_t)w[5] * s;
                    const int d = y - ((int)(minor >> 16) - 1);
                    if (d < 0 || d > 2) continue;
                    blend_px(c[j], col, aa_alpha(w, filt, s, e0 - s, minor, d));
                }
            } else {
                const int s = y - m0;
                if (s < 0 || s > e0) return false;
                const int64_t minor = (int64_t)w[4] + (int64_t)w[5] * s;
                const int base = (int)(minor >> 16) - 1;
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const int d = x + j - base;
                    if (d < 0 || d > 2) continue;
                    blend_px(c[j], col, aa_alpha(w, filt, s, e0 - s, minor, d));
                }
            }
            return true;
        }
        default:
            return false;
    }
}

// ordered compaction: threads with `hit` get consecutive slots in thread order; returns the slot (or -1) and total
__device__ __forceinline__ int compact(bool hit, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const int n = s_warp[k];
        before += k < warp ? n : 0;
        all += n;
    }
    total = all;
    __syncthreads();                              // s_warp may be reused right after
    return hit ? before + __popc(m & ((1u << lane) - 1)) : -1;
}

__global__ void __launch_bounds__(kThreads)
k_overlay(const VisOverlayFrame* __restrict__ frames, const VisLeaf* __restrict__ leaves) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ int s_groups[kThreads];
    __shared__ int s_leaf[kThreads][VIS_LEAF_WORDS];
    __shared__ int s_filter[64];

    const VisOverlayFrame f = frames[blockIdx.z];
    Tile t;
    t.x0 = blockIdx.x * kTileW;
    t.y0 = blockIdx.y * kTileH;
    if (t.x0 >= f.w || t.y0 >= f.h) return;
    t.x1 = min(t.x0 + kTileW, f.w) - 1;
    t.y1 = min(t.y0 + kTileH, f.h) - 1;
    const int tid = threadIdx.x;
    const int x = t.x0 + (tid & 15) * kPx, y = t.y0 + (tid >> 4);
    const int nv = y < f.h ? max(0, min(kPx, f.w - x)) : 0;      // valid pixels of this thread
    if (tid < 64) s_filter[tid] = c_filter[tid];

    // ---- load ----
    int c[kPx][3];
    const uint8_t* sp = f.src + (size_t)y * f.src_pitch + (size_t)x * 3;
    const bool in_place = f.src == f.dst;
    const bool vec_in = ((f.src_pitch | (int64_t)(uintptr_t)f.src) & 3) == 0;
    if (nv == kPx && vec_in) {
        const uint32_t* q = reinterpret_cast<const uint32_t*>(sp);
        const uint32_t a = __ldg(q), b = __ldg(q + 1), d = __ldg(q + 2);
        c[0][0] = a & 0xff; c[0][1] = (a >> 8) & 0xff; c[0][2] = (a >> 16) & 0xff;
        c[1][0] = a >> 24;  c[1][1] = b & 0xff;        c[1][2] = (b >> 8) & 0xff;
        c[2][0] = (b >> 16) & 0xff; c[2][1] = b >> 24; c[2][2] = d & 0xff;
        c[3][0] = (d >> 8) & 0xff;  c[3][1] = (d >> 16) & 0xff; c[3][2] = d >> 24;
    } else {
#pragma unroll
        for (int j = 0; j < kPx; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) c[j][k] = j < nv ? (int)sp[j * 3 + k] : 0;
    }

    // ---- group headers touching this tile, in order ----
    const VisLeaf* fl = leaves + f.group_begin;       // the frame's leaf array; header indices are relative to it
    const int n_groups = f.group_end - f.group_begin;
    bool dirty = false;
    for (int g0 = 0; g0 < n_groups; g0 += kThreads) {
        const int gi = g0 + tid;
        bool hit = false;
        if (gi < n_groups) {
            const int2 bb = __ldg(reinterpret_cast<const int2*>(&fl[gi].w[10]));
            hit = touches(t, bb.x, bb.y);
        }
        int n_hit;
        const int slot = compact(hit, s_warp, n_hit);
        if (slot >= 0) s_groups[slot] = gi;
        __syncthreads();
        for (int k = 0; k < n_hit; ++k) {
            const int g = s_groups[k];
            const int lb = __ldg(&fl[g].w[2]), le = __ldg(&fl[g].w[3]);
            for (int l0 = lb; l0 < le; l0 += kThreads) {
                const int li = l0 + tid;
                bool lhit = false;
                if (li < le) {
                    const int2 bb = __ldg(reinterpret_cast<const int2*>(&fl[li].w[10]));
                    lhit = touches(t, bb.x, bb.y) && (__ldg(&fl[li].w[0]) & LEAF_KIND_MASK) > LEAF_GROUP;
                }
                int n_leaf;
                const int ls = compact(lhit, s_warp, n_leaf);
                if (ls >= 0) {
                    const int4* src = reinterpret_cast<const int4*>(&fl[li]);
                    int4* dst = reinterpret_cast<int4*>(s_leaf[ls]);
                    dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);
                }
                __syncthreads();
                if (nv > 0)
                    for (int q = 0; q < n_leaf; ++q) dirty |= apply_leaf(s_leaf[q], s_filter, x, y, c);
                __syncthreads();
            }
        }
        __syncthreads();
    }

    // ---- store ----
    if (nv == 0 || (in_place && !dirty)) return;
    uint8_t* dp = f.dst + (size_t)y * f.dst_pitch + (size_t)x * 3;
    const bool vec_out = ((f.dst_pitch | (int64_t)(uintptr_t)f.dst) & 3) == 0;
    if (nv == kPx && vec_out) {
        uint32_t* q = reinterpret_cast<uint32_t*>(dp);
        q[0] = (uint32_t)c[0][0] | ((uint32_t)c[0][1] << 8) | ((uint32_t)c[0][2] << 16) | ((uint32_t)c[1][0] << 24);
        q[1] = (uint32_t)c[1][1] | ((uint32_t)c[1][2] << 8) | ((uint32_t)c[2][0] << 16) | ((uint32_t)c[2][1] << 24);
        q[2] = (uint32_t)c[2][2] | ((uint32_t)c[3][0] << 8) | ((uint32_t)c[3][1] << 16) | ((uint32_t)c[3][2] << 24);
    } else {
        for (int j = 0; j < nv; ++j)
#pragma unroll
            for (int k = 0; k < 3; ++k) dp[j * 3 + k] = (uint8_t)c[j][k];
    }
}

}  // namespace

extern "C" int vis_overlay_draw(const VisOverlayFrame* frames, int n_frames, int max_h, int max_w,
                                const VisLeaf* leaves, void* stream) {
    if (!frames || n_frames <= 0 || max_h <= 0 || max_w <= 0 || n_frames > 65535) {
        vis::set_error("vis_overlay_draw: bad arguments (frames=%d max %dx%d)", n_frames, max_w, max_h);
        return VIS_E_INVALID;
    }
    dim3 grid((max_w + kTileW - 1) / kTileW, (max_h + kTileH - 1) / kTileH, n_frames);
    k_overlay<<<grid, kThreads, 0, (cudaStream_t)stream>>>(frames, leaves);
    return vis::check_launch("vis_overlay_draw");
}
