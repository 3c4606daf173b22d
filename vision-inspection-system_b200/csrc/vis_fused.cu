// vis_fused.cu — the hot path: inspection frame (RGB uint8 HWC) -> Qwen2-VL pixel_values rows, one launch per batch.
//
// Replaces, per frame, tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:164-214:
//   Pillow 8bpc bicubic resample (horizontal pass -> uint8 -> vertical pass -> uint8), rescale 1/255,
//   mean/std normalisation (exact 768-entry table), temporal duplication, 14x14 patches in 2x2 merge order.
//
// Work decomposition (no tensor cores: the path is integer fixed point, HBM/INT32 bound):
//   one CTA = one column strip of one frame (<= 336 output columns, all or part of the output rows).
//   The CTA streams the strip's input rows top to bottom in chunks of 32 rows:
//     stage   : 32 row segments global -> shared with cp.async.bulk (UBLKCP), completion on an mbarrier
//     phase H : lane = input row, warp = sub-range of output columns.  Each lane walks along its row in
//               steps of 16 pixels (3 x LDS.128), keeps the last RING pixels per channel in registers
//               (static slots), and emits an output pixel whenever the input index reaches the end of that
//               output's tap window ("push" order) -> 3 x KT IMAD -> clip -> uint8 into the H ring (planar).
//     phase V : thread = 4 consecutive output columns of one channel.  Walks down the 32 fresh rows, keeps
//               the last RING rows in registers, emits an output row when the window completes -> KT IMAD
//               per pixel -> clip -> table lookup -> 8-byte stores straight into the patch layout (both
//               temporal copies).
//   The next chunk's bulk copies are issued between the two phases, so they overlap phase V and the other
//   resident CTA.  Coefficient records are host-built (vis_pack_records): newest-tap-first, zero padded to KT.
#include "vis_fused_common.cuh"

#include <cstdlib>
#include <cstring>

using namespace visf;

namespace visf {   // warp-specialised variant, vis_fused_ws.cu
int ws_layout_bytes(int span_bytes, int strip_w, int cls);
int ws_smem_max();
int ws_launch(int cls, const VisFrame* frames, const VisStrip* strips, int n_strips, int span_bytes, int strip_w,
              const float* lut768, float* pixel_values, cudaStream_t st);
}

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kChunk = 32;          // input rows per chunk == warp width (lane = row in phase H)
constexpr int kStepPx = 16;         // input pixels per unrolled step (48 B)
constexpr int kMaxStripW = 336;     // output columns per strip (3 * 336 / 4 = 252 phase-V threads)
constexpr int kPitch = kMaxStripW + 4;          // H-ring / output-tile row pitch: 340 = 4 * 85 (odd -> conflict free)
constexpr int kHPlane = kChunk * kPitch;        // one channel plane of the H ring
constexpr int kOPlane = VIS_PATCH * kPitch;     // one channel plane of the output tile (14 rows)
constexpr int kVCap = 48;           // vertical records staged per chunk
constexpr int kSmemBudget = 113 * 1024;   // two CTAs per SM

struct Layout {                     // shared-memory carve-up, fixed per launch (host-computed maxima)
    int stage_pitch;                // bytes per staged input row, odd multiple of 16
    int off_hring, off_otile, off_hrec, off_vrec, off_lut, off_mbar, total;
};

inline Layout make_layout(int span_bytes, int strip_w, int stride) {
    Layout L;
    L.stage_pitch = align_up(span_bytes, 16);
    if ((L.stage_pitch / 16) % 2 == 0) L.stage_pitch += 16;
    int off = kChunk * L.stage_pitch;
    L.off_hring = off;  off += 3 * kHPlane;
    L.off_otile = off;  off += 3 * kOPlane;
    off = align_up(off, 16);
    L.off_hrec = off;   off += (strip_w + 1) * stride * 4;
    L.off_vrec = off;   off += kVCap * stride * 4;
    L.off_lut = off;    off += 768 * 4;
    L.off_mbar = off;   off += 16;
    L.total = off;
    return L;
}

// One finished band of 14 output rows (a row of patches) -> pixel_values.  Thread t < 147 owns the 16-byte chunk
// q = t % 49 of channel plane c = t / 49 in EVERY patch of the strip; it looks the four values up in the table
// and writes the chunk to both temporal copies.  A warp therefore writes 512 contiguous bytes per store.
__device__ __noinline__ void store_band(const unsigned char* __restrict__ otile, const float* __restrict__ lut,
                                        float* __restrict__ band_out, int n_patches, int gx0) {
    const int t = threadIdx.x;
    if (t >= 147) return;
    const int c = t / 49, q = t - c * 49;
    const int f0 = 4 * q, f2 = f0 + 2;
    const int pya = f0 / VIS_PATCH, pxa = f0 - pya * VIS_PATCH;
    const int pyb = f2 / VIS_PATCH, pxb = f2 - pyb * VIS_PATCH;
    const unsigned char* sa = otile + c * kOPlane + pya * kPitch + pxa;
    const unsigned char* sb = otile + c * kOPlane + pyb * kPitch + pxb;
    const float* l = lut + c * 256;
    float* dst = band_out + c * 392 + f0;
    for (int g = 0; g < n_patches; ++g) {
        const int gx = gx0 + g;
        const unsigned a = *reinterpret_cast<const unsigned short*>(sa + g * VIS_PATCH);
        const unsigned b = *reinterpret_cast<const unsigned short*>(sb + g * VIS_PATCH);
        const float v0 = l[a & 0xff], v1 = l[a >> 8], v2 = l[b & 0xff], v3 = l[b >> 8];
        float* o = dst + (size_t)((gx >> 1) * 4 + (gx & 1)) * VIS_ROW_FLOATS;
        stg128(o, v0, v1, v2, v3);
        stg128(o + 196, v0, v1, v2, v3);
    }
}

// KT: taps per record (both axes), RING: register window (power of two >= KT), STRIDE: int32 slots per record
template <int KT, int RING, int STRIDE>
__global__ void __launch_bounds__(kThreads, RING == 8 ? 2 : 1)   // 16-slot windows need > 128 registers
k_fused(const VisFrame* __restrict__ frames, const VisStrip* __restrict__ strips, Layout L,
        const float* __restrict__ lut768, float* __restrict__ pixel_values) {
    extern __shared__ __align__(128) unsigned char smem[];
    const VisStrip sp = strips[blockIdx.x];
    const VisFrame fr = frames[sp.frame];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int x0 = sp.x0, x1 = sp.x1, y0 = sp.y0, y1 = sp.y1;
    const int sw = x1 - x0;

    unsigned char* stage = smem;
    unsigned char* hring = smem + L.off_hring;
    unsigned char* otile = smem + L.off_otile;
    int* hrec = reinterpret_cast<int*>(smem + L.off_hrec);
    int* vrec_s = reinterpret_cast<int*>(smem + L.off_vrec);
    float* lut = reinterpret_cast<float*>(smem + L.off_lut);          // transposed: lut[c * 256 + v]
    const uint32_t mbar = smem_u32(smem + L.off_mbar);

    // ---- prologue: records of this strip, table, barrier ----
    {
        const int4* src = reinterpret_cast<const int4*>(fr.hrec + (size_t)x0 * STRIDE);
        int4* dst = reinterpret_cast<int4*>(hrec);
        const int n4 = (sw + 1) * STRIDE / 4;
        for (int i = tid; i < n4; i += kThreads) dst[i] = __ldg(src + i);
        for (int i = tid; i < 768; i += kThreads) lut[(i % 3) * 256 + i / 3] = __ldg(lut768 + i);
        if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    }
    __syncthreads();

    // input columns staged per row: [px0, ...) with px0 a multiple of 16 pixels (48 B)
    const int px0 = hrec[STRIDE - 2] & ~(kStepPx - 1);
    const int px_last = hrec[(sw - 1) * STRIDE + STRIDE - 1];
    int row_bytes = align_up((px_last + 1) * 3, 16) - px0 * 3;
    if ((int64_t)px0 * 3 + row_bytes > fr.src_pitch) row_bytes = (int)(fr.src_pitch - (int64_t)px0 * 3);

    // input rows needed: [first row of window(y0), last row of window(y1-1)], chunk base multiple of 16
    const int* vrec = fr.vrec;
    const int r_first = __ldg(vrec + (size_t)y0 * STRIDE + STRIDE - 2) & ~15;
    const int r_end = __ldg(vrec + (size_t)(y1 - 1) * STRIDE + STRIDE - 1) + 1;      // exclusive
    const int n_chunks = (r_end - r_first + kChunk - 1) / kChunk;

    auto issue_chunk = [&](int chunk) {          // executed by warp 0
        const int r0 = r_first + chunk * kChunk;
        const int rows = min(kChunk, r_end - r0);
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(mbar, (uint32_t)rows * (uint32_t)row_bytes);
        }
        __syncwarp();
        if (lane < rows)
            bulk_g2s(smem_u32(stage + lane * L.stage_pitch),
                     fr.src + (size_t)(r0 + lane) * fr.src_pitch + (size_t)px0 * 3, (uint32_t)row_bytes, mbar);
    };
    if (warp == 0) issue_chunk(0);

    // ---- phase-H constants: this warp's output columns ----
    const int xa = x0 + (int)((int64_t)sw * warp / kWarps);
    const int xb = x0 + (int)((int64_t)sw * (warp + 1) / kWarps);

    // ---- phase-V constants: this thread's 4 output columns of one channel (idle threads shadow thread 0) ----
    const int wpr = sw / 4;                       // words per plane row
    const bool v_active = tid < 3 * wpr;
    const int vc = v_active ? tid / wpr : 0;
    const int vwx = v_active ? tid - vc * wpr : 0;
    const int half_gw = fr.dst_w / (2 * VIS_PATCH);
    const int n_patches = sw / VIS_PATCH, gx0 = x0 / VIS_PATCH;
    float* const frame_out = pixel_values + (size_t)fr.row0 * VIS_ROW_FLOATS;
    int vring[RING][4];
#pragma unroll
    for (int s = 0; s < RING; ++s) { vring[s][0] = vring[s][1] = vring[s][2] = vring[s][3] = 0; }
    int yo = y0;                                  // next output row (uniform across the CTA)
    int gy = y0 / VIS_PATCH, py = 0;              // its patch row and row inside the patch (y0 is a multiple of 14)
    Rec<KT> vr;

    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        const int r0 = r_first + chunk * kChunk;
        // vertical records for the output rows this chunk can complete: [yo, yo + kVCap), asynchronously
        const int yo_base = yo;
        for (int i = tid; i < kVCap * STRIDE / 4; i += kThreads) {
            const int rec = min(yo_base + i / (STRIDE / 4), fr.dst_h);            // clamp to the sentinel record
            cp_async16(smem_u32(vrec_s + i * 4), vrec + (size_t)rec * STRIDE + (i % (STRIDE / 4)) * 4);
        }
        mbar_wait(mbar, chunk & 1);

        // ================= phase H =================
        if (xa < xb) {
            int xo = xa;
            Rec<KT> hr;
            const int* hp = hrec + (xo - x0) * STRIDE;
            load_rec<KT, STRIDE>(hr, hp);
            int p = hp[STRIDE - 2] & ~(kStepPx - 1);
            uint32_t saddr = smem_u32(stage + lane * L.stage_pitch) + (uint32_t)(p - px0) * 3;
            unsigned char* hdst = hring + lane * kPitch + (xo - x0);
            int ring[3][RING];
#pragma unroll
            for (int s = 0; s < RING; ++s) { ring[0][s] = ring[1][s] = ring[2][s] = 0; }
            while (xo < xb) {
                uint32_t w[12];
                {
                    const uint4 q0 = lds128(saddr), q1 = lds128(saddr + 16), q2 = lds128(saddr + 32);
                    w[0] = q0.x; w[1] = q0.y; w[2] = q0.z; w[3] = q0.w;
                    w[4] = q1.x; w[5] = q1.y; w[6] = q1.z; w[7] = q1.w;
                    w[8] = q2.x; w[9] = q2.y; w[10] = q2.z; w[11] = q2.w;
                }
                saddr += 48;
#pragma unroll
                for (int j = 0; j < kStepPx; ++j) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const int b = 3 * j + c;
                        ring[c][j & (RING - 1)] = (int)__byte_perm(w[b >> 2], 0, 0x4440 + (b & 3));
                    }
                    while (hr.last == p + j) {
                        int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
                        for (int t = 0; t < KT; ++t) {
                            const int s = (j - t) & (RING - 1);
                            a0 += ring[0][s] * hr.k[t];
                            a1 += ring[1][s] * hr.k[t];
                            a2 += ring[2][s] * hr.k[t];
                        }
                        hdst[0] = (unsigned char)clip8i(a0);
                        hdst[kHPlane] = (unsigned char)clip8i(a1);
                        hdst[2 * kHPlane] = (unsigned char)clip8i(a2);
                        ++hdst;
                        ++xo;
                        hp += STRIDE;
                        if (xo < xb) load_rec<KT, STRIDE>(hr, hp);
                        else hr.last = INT_MAX;
                    }
                }
                p += kStepPx;
            }
        }
        cp_async_wait_all();
        __syncthreads();                           // H ring + vertical records complete, stage buffer free
        if (warp == 0 && chunk + 1 < n_chunks) issue_chunk(chunk + 1);

        // ================= phase V (all threads, uniform control flow; idle threads do not store) =================
        {
            auto fetch_vrec = [&]() {
                const int rel = yo - yo_base;
                if (rel < kVCap) load_rec<KT, STRIDE>(vr, vrec_s + rel * STRIDE);
                else load_rec<KT, STRIDE>(vr, vrec + (size_t)min(yo, fr.dst_h) * STRIDE);
                if (yo >= y1) vr.last = INT_MAX;
            };
            fetch_vrec();
            const unsigned char* hsrc = hring + vc * kHPlane + vwx * 4;
            unsigned char* odst = otile + vc * kOPlane + vwx * 4;
#pragma unroll 1
            for (int g = 0; g < kChunk / RING; ++g) {
                if (r0 + g * RING >= r_end) break;
                uint32_t words[RING];
#pragma unroll
                for (int u = 0; u < RING; ++u)
                    words[u] = *reinterpret_cast<const uint32_t*>(hsrc + (g * RING + u) * kPitch);
#pragma unroll
                for (int u = 0; u < RING; ++u) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) vring[u][e] = (int)__byte_perm(words[u], 0, 0x4440 + e);
                    while (vr.last == r0 + g * RING + u) {
                        int acc[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e] = 1 << (VIS_PRECISION_BITS - 1);
#pragma unroll
                        for (int t = 0; t < KT; ++t) {
                            const int s = (u - t) & (RING - 1);
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[e] += vring[s][e] * vr.k[t];
                        }
                        const uint32_t lo = __byte_perm(clip8i(acc[0]), clip8i(acc[1]), 0x0040);
                        const uint32_t hi = __byte_perm(clip8i(acc[2]), clip8i(acc[3]), 0x0040);
                        if (v_active) *reinterpret_cast<uint32_t*>(odst + py * kPitch) = __byte_perm(lo, hi, 0x5410);
                        ++yo;
                        if (++py == VIS_PATCH) {               // a row of patches is complete: write it out
                            __syncthreads();
                            store_band(otile, lut,
                                       frame_out + (size_t)((gy >> 1) * half_gw * 4 + (gy & 1) * 2) * VIS_ROW_FLOATS,
                                       n_patches, gx0);
                            __syncthreads();
                            py = 0;
                            ++gy;
                        }
                        fetch_vrec();
                    }
                }
            }
        }
        __syncthreads();                           // H ring consumed before the next phase H overwrites it
    }
}

template <int KT, int RING, int STRIDE>
int launch(const VisFrame* frames, const VisStrip* strips, int n_strips, const Layout& L,
           const float* lut768, float* pixel_values, cudaStream_t st) {
    auto kern = k_fused<KT, RING, STRIDE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_preprocess_fused: cudaFuncSetAttribute");
    kern<<<n_strips, kThreads, L.total, st>>>(frames, strips, L, lut768, pixel_values);
    return vis::check_launch("vis_preprocess_fused");
}

// taps -> kernel class
inline int kt_class(int kt) { return kt <= 6 ? 6 : kt <= 8 ? 8 : kt <= 12 ? 12 : kt <= 16 ? 16 : 0; }

// The warp-specialised persistent kernel serves the 8-slot tap classes; VIS_B200_FUSED=phased selects the
// phase-synchronous kernel instead (developer A/B switch, read per call).
inline bool use_ws(int cls) {
    if (cls != 6 && cls != 8) return false;
    const char* e = std::getenv("VIS_B200_FUSED");
    return !(e && std::strcmp(e, "phased") == 0);
}

inline int span_bytes_for(const int32_t* hbounds, int x0, int x1) {
    const int px0 = hbounds[2 * x0] & ~(kStepPx - 1);
    const int px_last = hbounds[2 * (x1 - 1)] + hbounds[2 * (x1 - 1) + 1] - 1;
    return align_up((px_last + 1) * 3, 16) - px0 * 3;
}

}  // namespace

extern "C" {

int vis_fused_kt_class(int kt) { return kt_class(kt); }

int vis_fused_supported(int64_t src_addr, int64_t src_pitch, int src_h, int src_w,
                        int dst_h, int dst_w, int hkt, int vkt) {
    if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || dst_h % 28 || dst_w % 28) return VIS_E_UNSUPPORTED;
    if ((src_addr % 16) || (src_pitch % 16) || src_pitch < (int64_t)src_w * 3) return VIS_E_UNSUPPORTED;
    if (kt_class(hkt > vkt ? hkt : vkt) == 0) return VIS_E_UNSUPPORTED;
    if ((int64_t)src_h > 100 * (int64_t)src_w && dst_h < src_h) return VIS_E_UNSUPPORTED;   // vertical-first branch
    return VIS_OK;
}

int vis_plan_strips_max(int dst_h, int dst_w) {
    if (dst_h <= 0 || dst_w <= 0) return VIS_E_INVALID;
    return (dst_w / 28 + 1) * (dst_h / 14 + 1);
}

// Splits one frame into column strips (width chosen so the CTA fits the two-per-SM shared-memory budget) and
// `vsplit` row segments.  hbounds: host copy of the horizontal bounds table.  Outputs the widest input span
// (bytes) and strip width over the emitted strips, which size the launch's shared memory.
int vis_plan_strips(int frame_index, int dst_h, int dst_w, const int32_t* hbounds, int kt, int vsplit,
                    VisStrip* strips, int capacity, int* span_bytes_out, int* strip_w_out) {
    const int cls = kt_class(kt);
    if (dst_h <= 0 || dst_w <= 0 || dst_h % 28 || dst_w % 28 || !hbounds || !strips || cls == 0 || vsplit < 1) {
        vis::set_error("vis_plan_strips: bad arguments");
        return VIS_E_INVALID;
    }
    const int hstride = vis_record_stride(cls);
    const int blocks = dst_w / 28;
    int best_n = 0;
    for (int per = kMaxStripW / 28; per >= 1; --per) {       // widest strips that fit the budget
        const int n = (blocks + per - 1) / per;
        int worst_span = 0, worst_w = 0;
        for (int s = 0; s < n; ++s) {
            const int b0 = (int)((int64_t)blocks * s / n), b1 = (int)((int64_t)blocks * (s + 1) / n);
            const int span = span_bytes_for(hbounds, b0 * 28, b1 * 28);
            worst_span = span > worst_span ? span : worst_span;
            worst_w = (b1 - b0) * 28 > worst_w ? (b1 - b0) * 28 : worst_w;
        }
        const bool fits = use_ws(cls) ? ws_layout_bytes(worst_span, worst_w, cls) <= ws_smem_max()
                                      : make_layout(worst_span, worst_w, hstride).total <= kSmemBudget;
        if (fits || per == 1) {
            best_n = n;
            *span_bytes_out = worst_span;
            *strip_w_out = worst_w;
            break;
        }
    }
    const int prow = dst_h / 14;
    if (vsplit > prow) vsplit = prow;
    if (best_n * vsplit > capacity) {
        vis::set_error("vis_plan_strips: capacity %d < %d", capacity, best_n * vsplit);
        return VIS_E_CAPACITY;
    }
    int n_out = 0;
    for (int v = 0; v < vsplit; ++v) {
        const int ya = (int)((int64_t)prow * v / vsplit) * 14, yb = (int)((int64_t)prow * (v + 1) / vsplit) * 14;
        for (int s = 0; s < best_n; ++s) {
            VisStrip& o = strips[n_out++];
            o.frame = frame_index;
            o.x0 = (int)((int64_t)blocks * s / best_n) * 28;
            o.x1 = (int)((int64_t)blocks * (s + 1) / best_n) * 28;
            o.y0 = ya;
            o.y1 = yb;
        }
    }
    return n_out;
}

int vis_preprocess_fused(const VisFrame* frames, int n_frames, const VisStrip* strips, int n_strips,
                         int max_kt, int max_span_bytes, int max_strip_w,
                         const float* lut768, float* pixel_values, void* stream) {
    const int cls = kt_class(max_kt);
    if (!frames || !strips || !lut768 || !pixel_values || n_frames <= 0 || n_strips <= 0 || cls == 0 ||
        max_span_bytes <= 0 || max_strip_w <= 0 || max_strip_w > kMaxStripW || max_strip_w % 28) {
        vis::set_error("vis_preprocess_fused: bad arguments (kt=%d span=%d strip_w=%d)", max_kt, max_span_bytes, max_strip_w);
        return cls == 0 ? VIS_E_UNSUPPORTED : VIS_E_INVALID;
    }
    if (use_ws(cls))
        return ws_launch(cls, frames, strips, n_strips, max_span_bytes, max_strip_w, lut768, pixel_values,
                         (cudaStream_t)stream);
    const Layout L = make_layout(max_span_bytes, max_strip_w, vis_record_stride(cls));
    if (L.total > 227 * 1024) {
        vis::set_error("vis_preprocess_fused: %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (cls) {
        case 6:  return launch<6, 8, 8>(frames, strips, n_strips, L, lut768, pixel_values, st);
        case 8:  return launch<8, 8, 12>(frames, strips, n_strips, L, lut768, pixel_values, st);
        case 12: return launch<12, 16, 16>(frames, strips, n_strips, L, lut768, pixel_values, st);
        default: return launch<16, 16, 20>(frames, strips, n_strips, L, lut768, pixel_values, st);
    }
}

}  // extern "C"
