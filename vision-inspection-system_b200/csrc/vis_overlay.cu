// vis_overlay.cu — device half of the defect-overlay rasteriser (cv2 drawing calls of
// utils/image_utils.py:259-313 in the reference, reproduced pixel for pixel).
//
// Only tiles that some leaf can touch are visited: the host bins the leaves into tiles (vis_overlay_tiles),
// everything else is a plain vectorised frame copy (out of place) or nothing at all (in place).  A CTA owns one
// listed 64x16 tile; each warp owns a 32x4 sub-tile (lane = 4 pixels, 12 bytes in registers) and culls on its own
// with ballots over the tile's leaves, fetched 32 at a time; touching leaves are
// staged IN ORDER in shared memory and every lane applies them in list order to its own pixels.  Per-pixel in-order
// application is what makes the result identical to OpenCV's sequential drawing: fills and LINE_8 points
// overwrite, LineAA pixels blend (twice, 8-bit alpha) with whatever is there.
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

// CN == 8 is the RECORD mode used to build dash stamps: a "pixel" is 8 bytes describing what the leaves did to it instead
// of a colour — byte 0: number of blends recorded (bits 0..5), 0x40 = more than seven (stamp unusable), 0x80 = an opaque
// write came first; bytes 1..7: the 8-bit alphas of the blends that followed, in order.  Replaying that chain on a frame
// pixel (LEAF_STAMP) gives exactly what applying the leaves one by one gives.
constexpr int kRecord = 8;
template <int CN>
__device__ __forceinline__ void set_px(int* c, int col) {
    if (CN == kRecord) {
        c[0] = 0x80;
#pragma unroll
        for (int k = 1; k < kRecord; ++k) c[k] = 0;
        return;
    }
    c[0] = col & 0xff; c[1] = (col >> 8) & 0xff; c[2] = (col >> 16) & 0xff;
    if (CN == 4) c[3] = (unsigned)col >> 24;
}
template <int CN>
__device__ __forceinline__ void blend_px(int* c, int col, int a) {
    if (CN == kRecord) {
        if (a == 0) return;                                   // ((cc - v) * 0 + 127) >> 8 == 0: no effect on any pixel
        const int n = c[0] & 0x3f;
        if (n >= kRecord - 1) { c[0] |= 0x40; return; }
#pragma unroll
        for (int k = 1; k < kRecord; ++k)
            if (k == n + 1) c[k] = a;
        c[0] += 1;
        return;
    }
#pragma unroll
    for (int k = 0; k < CN; ++k) {
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
template <int CN>
__device__ __forceinline__ bool apply_leaf(const int* w, const int* filt, int x, int y, int (*c)[CN]) {
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
                if (x + j >= xx1 && x + j <= xx2) set_px<CN>(c[j], col);
            return true;
        }
        case LEAF_SPANS: {
            const int r = y - w[3];
            if (r < 0 || r >= w[4]) return false;
            const int hw = (w[5 + (r >> 2)] >> (8 * (r & 3))) & 0xff;
            if (hw == 0xff) return false;
#pragma unroll
            for (int j = 0; j < kPx; ++j)
                if (abs(x + j - w[2]) <= hw) set_px<CN>(c[j], col);
            return true;
        }
        case LEAF_LINE8: {
            const int m0 = w[2], ecount = w[3];
            if (w[0] & LEAF_FLAG_XMAJOR) {
#pragma unroll
                for (int j = 0; j < kPx; ++j) {
                    const int i = x + j - m0;
                    const bool on = i >= 0 && i <= ecount && (int)(((int64_t)w[4] + (int64_t)w[5] * i) >> 16) == y;
                    if (on || (x + j == w[6] && y == w[7])) set_px<CN>(c[j], col);
                }
            } else {
                const int i = y - m0;
                const bool in = i >= 0 && i <= ecount;
                const int xx = (int)(((int64_t)w[4] + (int64_t)w[5] * i) >> 16);
#pragma unroll
                for (int j = 0; j < kPx; ++j)
                    if ((in && x + j == xx) || (x + j == w[6] && y == w[7])) set_px<CN>(c[j], col);
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
                    blend_px<CN>(c[j], col, aa_alpha(w, filt, s, e0 - s, minor, d));
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
                    blend_px<CN>(c[j], col, aa_alpha(w, filt, s, e0 - s, minor, d));
                }
            }
            return true;
        }
        case LEAF_STAMP: {                        // replay the recorded blend chain of a dash with this leaf's colour
            if (CN == kRecord) return false;
            const int sy = y - w[5];
            if (sy < 0 || sy >= w[7]) return false;
            const uint2* sp = reinterpret_cast<const uint2*>(((uint64_t)(uint32_t)w[3] << 32) | (uint32_t)w[2]) +
                              (size_t)sy * w[6];
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int sx = x + j - w[4];
                if (sx < 0 || sx >= w[6]) continue;
                const uint2 rec = __ldg(sp + sx);
                if (!(rec.x & 0xff)) continue;
                if (rec.x & 0x80) set_px<CN>(c[j], col);
                const int n = rec.x & 0x3f;
#pragma unroll
                for (int i = 0; i < kRecord - 1; ++i) {
                    if (i >= n) break;
                    const unsigned word = i < 3 ? rec.x : rec.y;
                    blend_px<CN>(c[j], col, (int)((word >> (8 * ((i + 1) & 3))) & 0xff));
                }
            }
            return true;
        }
        case LEAF_SPRITE: {                       // opaque marker rasterised once: copy the pixels it covers
            if (CN == kRecord) return false;
            const int sy = y - w[5];
            if (sy < 0 || sy >= w[7]) return false;
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(((uint64_t)(uint32_t)w[3] << 32) | (uint32_t)w[2]) +
                                 (size_t)sy * w[6];
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const int sx = x + j - w[4];
                if (sx < 0 || sx >= w[6]) continue;
                const uint32_t v = __ldg(sp + sx);
                if (v >> 24) set_px<CN>(c[j], (int)v);
            }
            return true;
        }
        default:
            return false;
    }
}

// frame copy for out-of-place drawing: rows of `row_bytes` bytes, 16-byte vectors when everything is aligned
__global__ void __launch_bounds__(256)
k_overlay_copy(const VisOverlayFrame* __restrict__ frames, int channels) {
    const VisOverlayFrame f = frames[blockIdx.y];
    if (f.src == f.dst) return;
    int64_t rows = f.h, row_bytes = (int64_t)f.w * channels;
    if (f.src_pitch == row_bytes && f.dst_pitch == row_bytes) { row_bytes *= rows; rows = 1; }     // one long row
    const bool vec = (((uintptr_t)f.src | (uintptr_t)f.dst | (uint64_t)f.src_pitch | (uint64_t)f.dst_pitch) & 15) == 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const int64_t per_row = row_bytes / 16, n = rows * per_row;
        for (int64_t i = t0; i < n; i += 4 * stride) {
            uint4 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t j = i + k * stride;
                if (j < n) v[k] = __ldcs(reinterpret_cast<const uint4*>(f.src + (j / per_row) * f.src_pitch) + j % per_row);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t j = i + k * stride;
                if (j < n) __stcs(reinterpret_cast<uint4*>(f.dst + (j / per_row) * f.dst_pitch) + j % per_row, v[k]);
            }
        }
        const int64_t tail0 = per_row * 16, tail = row_bytes - tail0;
        for (int64_t i = t0; i < rows * tail; i += stride)
            f.dst[(i / tail) * f.dst_pitch + tail0 + i % tail] = f.src[(i / tail) * f.src_pitch + tail0 + i % tail];
    } else {
        for (int64_t i = t0; i < rows * row_bytes; i += stride)
            f.dst[(i / row_bytes) * f.dst_pitch + i % row_bytes] = f.src[(i / row_bytes) * f.src_pitch + i % row_bytes];
    }
}

// In-place drawing of the touched tiles.  CTA = one 64x16 tile of the host-built tile list; each of its 8 warps owns a
// 32x4 pixel sub-tile (lane = 4 pixels of one row) and works on its own, without block-wide barriers: the leaves the
// host binned into this tile (in leaf order) are fetched 32 at a time, one per lane, tested against the sub-tile with
// one ballot, and the touching ones are staged IN ORDER in the warp's shared-memory slots and applied per pixel in list
// order.
#ifndef VIS_OVERLAY_MIN_BLOCKS
#define VIS_OVERLAY_MIN_BLOCKS 6          // 40 registers: the kernel is latency bound (dependent loads), occupancy pays for a few spills
#endif
#ifndef VIS_OVERLAY_WAVES
#define VIS_OVERLAY_WAVES 16              // > 0: a persistent grid of (resident CTAs x this) walks the tile list; 0: one CTA per tile
#endif
// one listed tile; the eight warps of the CTA are autonomous (a warp owns a 32 x 4 sub-tile and never waits for another)
template <int CN>
__device__ __forceinline__ void draw_tile(const VisOverlayFrame* __restrict__ frames, const VisOverlayTile tl,
                                          const VisOverlayRef* __restrict__ refs, const VisLeaf* __restrict__ leaves,
                                          int (*my_leaf)[VIS_LEAF_WORDS], const int* s_filter, int lane, int warp) {
    const VisOverlayFrame f = frames[tl.frame];
    Tile t;
    t.x0 = (tl.txy & 0xffff) * kTileW + (warp & 1) * 32;
    t.y0 = ((unsigned)tl.txy >> 16) * kTileH + (warp >> 1) * 4;
    if (t.x0 >= f.w || t.y0 >= f.h) return;
    t.x1 = min(t.x0 + 32, f.w) - 1;
    t.y1 = min(t.y0 + 4, f.h) - 1;
    const int x = t.x0 + (lane & 7) * kPx, y = t.y0 + (lane >> 3);
    const int nv = y < f.h ? max(0, min(kPx, f.w - x)) : 0;      // valid pixels of this lane

    const VisLeaf* fl = leaves + f.group_begin;       // the frame's leaf array; leaf indices are relative to it
    int c[kPx][CN];
    bool loaded = false, dirty = false;
    uint8_t* const px = f.dst + (size_t)y * f.dst_pitch + (size_t)x * CN;
    const bool vec = ((f.dst_pitch | (int64_t)(uintptr_t)f.dst) & 3) == 0;

    auto fetch = [&]() {
        if (CN == 4 && nv == kPx && vec) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(px);
#pragma unroll
            for (int j = 0; j < kPx; ++j) {
                const uint32_t a = q[j];
#pragma unroll
                for (int k = 0; k < CN; ++k) c[j][k] = (a >> (8 * k)) & 0xff;
            }
        } else if (CN == 3 && nv == kPx && vec) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(px);
            const uint32_t a = q[0], b = q[1], d = q[2];
            c[0][0] = a & 0xff; c[0][1] = (a >> 8) & 0xff; c[0][2] = (a >> 16) & 0xff;
            c[1][0] = a >> 24;  c[1][1] = b & 0xff;        c[1][2] = (b >> 8) & 0xff;
            c[2][0] = (b >> 16) & 0xff; c[2][1] = b >> 24; c[2][2] = d & 0xff;
            c[3][0] = (d >> 8) & 0xff;  c[3][1] = (d >> 16) & 0xff; c[3][2] = d >> 24;
        } else {
#pragma unroll
            for (int j = 0; j < kPx; ++j)
#pragma unroll
                for (int q = 0; q < CN; ++q) c[j][q] = j < nv ? (int)px[j * CN + q] : 0;
        }
    };

    // the refs of this tile, 32 at a time in registers (lane i holds ref r0 + i), consumed in order
    for (int r0 = tl.ref_begin; r0 < tl.ref_end; r0 += 32) {
        const int nr = min(32, tl.ref_end - r0);
        int2 mine = make_int2(0, 0);
        if (lane < nr) mine = __ldg(reinterpret_cast<const int2*>(refs + r0 + lane));
        // refs that name single leaves (what vis_overlay_tiles emits: the host culls per leaf): every lane fetches its
        // whole leaf at once and the 32 of them are culled against the sub-tile with one ballot — two dependent loads
        // per 32 leaves instead of two per ref
        if (__all_sync(0xffffffffu, lane >= nr || mine.y == mine.x + 1)) {
            bool lhit = false;
            int4 b0 = make_int4(0, 0, 0, 0), b1 = b0, b2 = b0;
            if (lane < nr) {
                const int4* src = reinterpret_cast<const int4*>(&fl[mine.x]);
                b0 = __ldg(src); b1 = __ldg(src + 1); b2 = __ldg(src + 2);
                lhit = touches(t, b2.z, b2.w);
            }
            const unsigned lm = __ballot_sync(0xffffffffu, lhit);
            if (!lm) continue;
            if (!loaded) { loaded = true; fetch(); }
            if (lhit) {
                int4* dst = reinterpret_cast<int4*>(my_leaf[__popc(lm & ((1u << lane) - 1))]);
                dst[0] = b0; dst[1] = b1; dst[2] = b2;
            }
            __syncwarp();
            const int n_leaf = __popc(lm);
            if (nv > 0)
                for (int q = 0; q < n_leaf; ++q) dirty |= apply_leaf<CN>(my_leaf[q], s_filter, x, y, c);
            __syncwarp();
            continue;
        }
        for (int k = 0; k < nr; ++k) {
            const int lb = __shfl_sync(0xffffffffu, mine.x, k), le = __shfl_sync(0xffffffffu, mine.y, k);
            const int li = lb + lane;
            bool lhit = false;
            if (li < le) {
                const int2 bb = __ldg(reinterpret_cast<const int2*>(&fl[li].w[10]));
                lhit = touches(t, bb.x, bb.y);
            }
            const unsigned lm = __ballot_sync(0xffffffffu, lhit);
            if (!lm) continue;
            if (!loaded) { loaded = true; fetch(); }          // first leaf that reaches this sub-tile: fetch the pixels
            if (lhit) {
                const int4* src = reinterpret_cast<const int4*>(&fl[li]);
                int4* dst = reinterpret_cast<int4*>(my_leaf[__popc(lm & ((1u << lane) - 1))]);
                dst[0] = __ldg(src); dst[1] = __ldg(src + 1); dst[2] = __ldg(src + 2);
            }
            __syncwarp();
            const int n_leaf = __popc(lm);
            if (nv > 0)
                for (int q = 0; q < n_leaf; ++q) dirty |= apply_leaf<CN>(my_leaf[q], s_filter, x, y, c);
            __syncwarp();
        }
    }

    if (!dirty) return;
    if (CN == 4 && nv == kPx && vec) {
        uint32_t* q = reinterpret_cast<uint32_t*>(px);
#pragma unroll
        for (int j = 0; j < kPx; ++j)
            q[j] = (uint32_t)c[j][0] | ((uint32_t)c[j][1] << 8) | ((uint32_t)c[j][2] << 16) | ((uint32_t)c[j][CN - 1] << 24);
    } else if (CN == 3 && nv == kPx && vec) {
        uint32_t* q = reinterpret_cast<uint32_t*>(px);
        q[0] = (uint32_t)c[0][0] | ((uint32_t)c[0][1] << 8) | ((uint32_t)c[0][2] << 16) | ((uint32_t)c[1][0] << 24);
        q[1] = (uint32_t)c[1][1] | ((uint32_t)c[1][2] << 8) | ((uint32_t)c[2][0] << 16) | ((uint32_t)c[2][1] << 24);
        q[2] = (uint32_t)c[2][2] | ((uint32_t)c[3][0] << 8) | ((uint32_t)c[3][1] << 16) | ((uint32_t)c[3][2] << 24);
    } else {
        for (int j = 0; j < nv; ++j)
#pragma unroll
            for (int q = 0; q < CN; ++q) px[j * CN + q] = (uint8_t)c[j][q];
    }
}

template <int CN>
__global__ void __launch_bounds__(kThreads, VIS_OVERLAY_MIN_BLOCKS)
k_overlay_tiles(const VisOverlayFrame* __restrict__ frames, const VisOverlayTile* __restrict__ tiles, int n_tiles,
                const VisOverlayRef* __restrict__ refs, const VisLeaf* __restrict__ leaves) {
    __shared__ int s_leaf[kThreads / 32][32][VIS_LEAF_WORDS];
    __shared__ int s_filter[64];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 64) s_filter[tid] = c_filter[tid];
    __syncthreads();
    for (int ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
        draw_tile<CN>(frames, tiles[ti], refs, leaves, s_leaf[warp], s_filter, lane, warp);
        __syncwarp();                             // the warp's leaf staging area is reused by its next tile
    }
}

}  // namespace

extern "C" int vis_overlay_draw_cn(const VisOverlayFrame* frames, int n_frames, int channels, int copy_frames,
                                   const VisOverlayTile* tiles, int n_tiles, const VisOverlayRef* refs,
                                   const VisLeaf* leaves, void* stream) {
    if (!frames || n_frames <= 0 || n_frames > 65535 || n_tiles < 0 || (n_tiles && (!tiles || !refs || !leaves)) ||
        (channels != 3 && channels != 4 && channels != kRecord)) {
        vis::set_error("vis_overlay_draw: bad arguments (frames=%d tiles=%d channels=%d)", n_frames, n_tiles, channels);
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (copy_frames) {                            // out of place: dst = src first, then draw in place on dst
        k_overlay_copy<<<dim3(64, n_frames), 256, 0, st>>>(frames, channels);
        const int rc = vis::check_launch("vis_overlay_draw(copy)");
        if (rc != VIS_OK) return rc;
    }
    if (n_tiles > 0) {
        int grid = n_tiles;
#if VIS_OVERLAY_WAVES > 0
        static int sms = 0;
        if (sms == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        }
        grid = n_tiles < sms * VIS_OVERLAY_MIN_BLOCKS * VIS_OVERLAY_WAVES ? n_tiles : sms * VIS_OVERLAY_MIN_BLOCKS * VIS_OVERLAY_WAVES;
#endif
        if (channels == 3)      k_overlay_tiles<3><<<grid, kThreads, 0, st>>>(frames, tiles, n_tiles, refs, leaves);
        else if (channels == 4) k_overlay_tiles<4><<<grid, kThreads, 0, st>>>(frames, tiles, n_tiles, refs, leaves);
        else                    k_overlay_tiles<kRecord><<<grid, kThreads, 0, st>>>(frames, tiles, n_tiles, refs, leaves);   // stamp recording
        return vis::check_launch("vis_overlay_draw");
    }
    return VIS_OK;
}

extern "C" int vis_overlay_draw(const VisOverlayFrame* frames, int n_frames, int copy_frames,
                                const VisOverlayTile* tiles, int n_tiles, const VisOverlayRef* refs,
                                const VisLeaf* leaves, void* stream) {
    return vis_overlay_draw_cn(frames, n_frames, 3, copy_frames, tiles, n_tiles, refs, leaves, stream);
}
