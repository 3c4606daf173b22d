// vis_fused_sched16.cu — statically scheduled, warp-specialised kernel for 9..16-tap windows (16-slot register window).
//
// Same schedule-driven design as vis_fused_sched.cu (VisSched in the kernel parameter block, uniform branches, roles
// over shared-memory rings), sized for the strong downscales of the path:
//   * 4K frames at the processor's default max_pixels (3840x2160 -> 1316x728 bicubic: 13 taps)       -> pixel_values
//   * the agents' LANCZOS thumbnails and resize_image (src/agents/vlm_inspector.py:64, vlm_auditor.py:91,
//     utils/image_utils.py:75; 4K -> 2048x1152 and 1080p -> 1024x576: 13 taps)                       -> uint8 HWC
// Differences from the 8-slot kernel:
//   H  (11 warps) 16 input pixels per step (3 x LDS.128), 16-slot register window, 16-bit step masks.
//   V  (6 warps)  PULL order: for every output row of the schedule the thread reads its KT tap words straight from
//                 the H ring (the slot holds KT-1 carry rows in front of the 32 fresh rows, copied over from the
//                 previous chunk by the same thread), so no register ring and ONE emit body (a 16-slot register
//                 ring would need 16 unrolled bodies of ~150 instructions: far beyond the instruction cache).
//   Windows of 17..32 taps (4K -> 1024 LANCZOS thumbnails: 25) keep the same V / store roles; their H role PULLS: for
//   every output column the lane (= input row) reads the window's bytes from the staged row with 32-bit loads at the
//   window's (uniform) alignment and multiplies them by the record — no register window, 2x the instructions per MAC of
//   the push scheme, still far ahead of the generic per-pixel pass.
//   store (2 warps) pixel_values mode: LUT + 16-byte stores as in the 8-slot kernel; uint8 mode: the planar band is
//                 interleaved back to RGB with byte permutes and written as coalesced 32-bit words.
#include "vis_fused_common.cuh"

using namespace visf;

namespace {

#ifndef VIS_S16_HWARPS
#define VIS_S16_HWARPS 12
#endif
#ifndef VIS_S16_SWARPS
#define VIS_S16_SWARPS 2
#endif
constexpr int kHWarps = VIS_S16_HWARPS, kSWarps = VIS_S16_SWARPS;    // H is the heavy role at these scales (profiles/r01_fused_sched16_4k.txt)
// warp ranges in priority order (the scheduler prefers the highest ready warp id): H < loader < S < V
constexpr int kHBase = 0, kLBase = kHWarps, kSBase = kHWarps + 1, kVBase = kHWarps + 1 + kSWarps;
// vertical-pass warps NV: 6 for the mild downscales (V and H work comparable), 4 for the strong ones (V is light: fewer,
// fuller warps issue a third fewer instructions and every SMSP holds 3 H + 1 V warp), 3 from 3.4x on — picked by
// vis_sched_build (A/B on 4K frames: 2.97x: 6 -> 56 k, 4 -> 63 k, 3 -> 48 k images/s; 3.75x: 6 -> 29 k, 4 -> 31 k, 3 -> 34 k)
constexpr int threads16(int nv) { return (kHWarps + nv + kSWarps + 1) * 32; }      // 672 / 608: <= 96 registers per thread
constexpr int kChunk = 32, kStepPx = 16, kRing = 16;
constexpr int max_strip_w16(int nv) { return nv * 32 / 3 * 4; }         // 256 / 168: one V thread per 4 columns of one channel
constexpr int kVRecs = kChunk + 1;                // vertical records a chunk can touch (scale >= 1): 32 emits + 1 look-ahead
constexpr int kSmemMax = 227 * 1024;

enum Bar { SF = 0, SE = 2, HF = 4, HE = 6, VF = 8, OF = 10, OE = 12, kBars = 14 };   // full/empty pairs, two slots each

struct Layout16 {
    int stage_pitch, stage_slot, hrec_slot, vrec_slot;
    int hpitch, hplane;          // H ring: bytes per row (strip width + pad, an odd number of words: conflict-free
                                 // lane = row byte stores), bytes per channel plane (carry rows + fresh rows)
    int opitch, oplane;          // band tile: bytes per row (strip width), bytes per channel plane (14 rows)
    int off_stage, off_hring, off_otile, off_hrec, off_vrec, off_lut, off_bar, total;
};

inline Layout16 make_layout16(int stage_pitch, int strip_w, int cls) {
    const int stride = vis_record_stride(cls);
    Layout16 L;
    L.stage_pitch = stage_pitch;
    L.stage_slot = kChunk * stage_pitch;
    L.hrec_slot = align_up((strip_w + 1) * stride * 4, 16);
    L.vrec_slot = align_up(kVRecs * stride * 4, 16);
    L.hpitch = strip_w + 4 + ((strip_w / 4) % 2 ? 4 : 0);
    L.hplane = (cls - 1 + kChunk) * L.hpitch;
    L.opitch = strip_w;
    L.oplane = VIS_PATCH * L.opitch;
    int off = 128;                                 // pull-order H reads up to (kt-1)*3 bytes in front of a row's window
    L.off_stage = off; off += 2 * L.stage_slot;
    L.off_hring = off; off += 2 * 3 * L.hplane;
    L.off_otile = off; off += 2 * 3 * L.oplane;
    off = align_up(off, 16);
    L.off_hrec = off;  off += 2 * L.hrec_slot;
    L.off_vrec = off;  off += 2 * L.vrec_slot;
    L.off_lut = off;   off += 768 * 4;
    L.off_bar = off;   off += kBars * 8;
    L.total = off;
    return L;
}

template <int KT>
__device__ __forceinline__ void load_coeffs16(int (&k)[KT], uint32_t addr) {
    static_assert(KT >= 12 && KT <= 32, "tap classes 12..32");
    // a record holds KT coefficient slots + (first, end) in (KT + 5) & ~3 words: the last 16-byte load of an odd class
    // reads into those trailing words and the surplus is dropped
#pragma unroll
    for (int q = 0; q < (KT + 3) / 4; ++q) {
        const uint4 a = lds128(addr + 16 * q);
        k[4 * q] = (int)a.x;
        if (4 * q + 1 < KT) k[4 * q + 1] = (int)a.y;
        if (4 * q + 2 < KT) k[4 * q + 2] = (int)a.z;
        if (4 * q + 3 < KT) k[4 * q + 3] = (int)a.w;
    }
}

__device__ __noinline__ void band_done16(uint32_t bar0, int nb, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + (uint32_t)(OF + (nb & 1)) * 8);
    const int nx = nb + 1;
    if (nx >= 2) mbar_wait(bar0 + (uint32_t)(OE + (nx & 1)) * 8, ((nx >> 1) - 1) & 1);
}

// pull-order horizontal sample: window of KT pixels ending at the record's (virtual) end, bytes [A, A + 3*KT) of the
// words at `wbase` (A = alignment of the window start, compile time); slot s of the record <-> pixel KT-1-s of the window
template <int KT, int A>
__device__ __forceinline__ void hpull(uint32_t wbase, const int (&kf)[KT], int& a0, int& a1, int& a2) {
    constexpr int W = (A + 3 * KT + 3) / 4;
    uint32_t wv[W];
#pragma unroll
    for (int q = 0; q < W; ++q) wv[q] = lds32(wbase + 4 * q);
#pragma unroll
    for (int s = 0; s < KT; ++s) {
        const int b = A + 3 * (KT - 1 - s);
        a0 += (int)__byte_perm(wv[b >> 2], 0, 0x4440 + (b & 3)) * kf[s];
        a1 += (int)__byte_perm(wv[(b + 1) >> 2], 0, 0x4440 + ((b + 1) & 3)) * kf[s];
        a2 += (int)__byte_perm(wv[(b + 2) >> 2], 0, 0x4440 + ((b + 2) & 3)) * kf[s];
    }
}

struct FramePtrs { const unsigned char* src; long long second; };      // VisFrameRef / VisResizeRef: same layout

template <int KT, int STRIDE, bool U8, int NV>
__global__ void __launch_bounds__(threads16(NV), 1)
k_fused_sched16(const __grid_constant__ VisSched sc, const FramePtrs* __restrict__ frames, int n_items,
                const __grid_constant__ Layout16 L, long long dst_pitch, const int* __restrict__ hrec_g,
                const int* __restrict__ vrec_g, const float* __restrict__ lut768, float* __restrict__ pixel_values) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int CARRY = KT - 1;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);               // warp-uniform for the compiler
    float* lut = reinterpret_cast<float*>(smem + L.off_lut);              // transposed: lut[c * 256 + v]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    auto bar = [&](int which, int slot) { return bar0 + (uint32_t)(which + slot) * 8; };
    const int per_frame = sc.n_strips * sc.n_segs;

    if (!U8)
        for (int i = tid; i < 768; i += (int)blockDim.x) lut[(i % 3) * 256 + i / 3] = __ldg(lut768 + i);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(SF, s), 1);
            mbar_init(bar(SE, s), kHWarps);
            mbar_init(bar(HF, s), kHWarps);
            mbar_init(bar(HE, s), NV);
            mbar_init(bar(VF, s), 1);
            mbar_init(bar(OF, s), NV);
            mbar_init(bar(OE, s), kSWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();                                   // the only CTA-wide barrier

    if (warp == kLBase) {
        // ============================== loader ==============================
        int k = 0, sl = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++sl) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const unsigned char* src = frames[f].src + (size_t)S.px0 * 3;
            const uint32_t rec_bytes = (uint32_t)(S.x1 - S.x0 + 1) * STRIDE * 4;
            const int n_chunks = (G.r_end - G.r_first + kChunk - 1) / kChunk;
            int yo = G.y0;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1;
                const uint32_t prev = ((k >> 1) - 1) & 1;
                if (k >= 2) mbar_wait(bar(SE, slot), prev);                 // H is done with the stage slot
                const int r0 = G.r_first + c * kChunk;
                const int rows = max(0, min(kChunk, sc.src_h - r0));        // r_end may include virtual rows past the image
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(bar(SF, slot), (uint32_t)rows * (uint32_t)S.row_bytes + (c == 0 ? rec_bytes : 0u));
                }
                __syncwarp();
                unsigned char* stage = smem + L.off_stage + slot * L.stage_slot;
                if (lane < rows)
                    bulk_g2s(smem_u32(stage + lane * L.stage_pitch), src + (size_t)(r0 + lane) * sc.src_pitch,
                             (uint32_t)S.row_bytes, bar(SF, slot));
                if (c == 0 && lane == 0)
                    bulk_g2s(smem_u32(smem + L.off_hrec + (sl & 1) * L.hrec_slot), hrec_g + (size_t)S.x0 * STRIDE,
                             rec_bytes, bar(SF, slot));
                if (k >= 2) mbar_wait(bar(HE, slot), prev);                 // V is done with the record slot
                if (lane == 0) {
                    const uint32_t vbytes = (uint32_t)min(kVRecs, sc.dst_h + 1 - yo) * STRIDE * 4;
                    fence_proxy_async();
                    mbar_expect_tx(bar(VF, slot), vbytes);
                    bulk_g2s(smem_u32(smem + L.off_vrec + slot * L.vrec_slot), vrec_g + (size_t)yo * STRIDE, vbytes,
                             bar(VF, slot));
                }
                // 2 groups x (16-bit first-sample mask, 16-bit second-sample mask: never set here) per chunk
                const uint32_t* m8 = reinterpret_cast<const uint32_t*>(sc.mask + G.mask_off + c * 8);
                yo += __popc(m8[0]) + __popc(m8[1]);
            }
        }
    } else if (warp < kHBase + kHWarps) {
        // ============================== horizontal pass ==============================
        const int sub = warp - kHBase;
        int k = 0, sl = 0;
        int rg[3][kRing];                                  // the last 16 input pixels per channel (static slots)
#pragma unroll
        for (int q = 0; q < kRing; ++q) rg[0][q] = rg[1][q] = rg[2][q] = 0;
        if (KT > 16) {
            // ---- pull order (17..32 taps) ----
            for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++sl) {
                const int f = w / per_frame, r = w - f * per_frame;
                const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
                const VisSchedStrip S = sc.strip[st];
                const VisSchedSeg G = sc.seg[sg];
                const VisSchedSub U = sc.sub[st][sub];
                const int n_chunks = (G.r_end - G.r_first + kChunk - 1) / kChunk;
                const uint32_t hrec0 = smem_u32(smem + L.off_hrec + (sl & 1) * L.hrec_slot) + (uint32_t)(U.xa - S.x0) * STRIDE * 4;
                for (int c = 0; c < n_chunks; ++c, ++k) {
                    const int slot = k & 1, j = k >> 1;
                    mbar_wait(bar(SF, slot), j & 1);
                    if (k >= 2) mbar_wait(bar(HE, slot), (j - 1) & 1);
                    const uint32_t row = smem_u32(smem + L.off_stage + slot * L.stage_slot + lane * L.stage_pitch);
                    unsigned char* hdst = smem + L.off_hring + slot * 3 * L.hplane + (CARRY + lane) * L.hpitch + (U.xa - S.x0);
                    unsigned char* const hdst1 = hdst + L.hplane;
                    unsigned char* const hdst2 = hdst + 2 * L.hplane;
                    uint32_t hp = hrec0;
#pragma unroll 1
                    for (int xi = 0; xi < U.xb - U.xa; ++xi, hp += STRIDE * 4) {
                        int kf[KT];
                        load_coeffs16<KT>(kf, hp);
                        const int end = (int)lds32(hp + (STRIDE - 1) * 4);              // (virtual) window end of this column
                        const int o = (end - (KT - 1) - S.px0) * 3;                      // byte offset of the window start, may be < 0
                        const int al = o & 3;
                        const uint32_t wbase = row + (uint32_t)(o - al);
                        int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
                        if (al == 0) hpull<KT, 0>(wbase, kf, a0, a1, a2);
                        else if (al == 1) hpull<KT, 1>(wbase, kf, a0, a1, a2);
                        else if (al == 2) hpull<KT, 2>(wbase, kf, a0, a1, a2);
                        else hpull<KT, 3>(wbase, kf, a0, a1, a2);
                        hdst[xi] = (unsigned char)clip8i(a0);
                        hdst1[xi] = (unsigned char)clip8i(a1);
                        hdst2[xi] = (unsigned char)clip8i(a2);
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar(SE, slot));
                        mbar_arrive(bar(HF, slot));
                    }
                }
            }
        } else
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++sl) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const VisSchedSub U = sc.sub[st][sub];
            const int n_chunks = (G.r_end - G.r_first + kChunk - 1) / kChunk;
            const uint32_t hrec0 = smem_u32(smem + L.off_hrec + (sl & 1) * L.hrec_slot) + (uint32_t)(U.xa - S.x0) * STRIDE * 4;
            const uint8_t* const um = sc.mask + U.mask_off;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(SF, slot), j & 1);
                if (k >= 2) mbar_wait(bar(HE, slot), (j - 1) & 1);
                uint32_t sa = smem_u32(smem + L.off_stage + slot * L.stage_slot + lane * L.stage_pitch) + (uint32_t)(U.p0 - S.px0) * 3;
                unsigned char* hdst = smem + L.off_hring + slot * 3 * L.hplane + (CARRY + lane) * L.hpitch + (U.xa - S.x0);
                unsigned char* const hdst1 = hdst + L.hplane;
                unsigned char* const hdst2 = hdst + 2 * L.hplane;
                int xi = 0;
                uint32_t hp = hrec0;
                int kf[KT];
                load_coeffs16<KT>(kf, hp);
#pragma unroll 1
                for (int i = 0; i < U.nsteps; ++i) {
                    const uint32_t m = (uint32_t)um[4 * i] | ((uint32_t)um[4 * i + 1] << 8);
                    uint32_t wv[12];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint4 d = lds128(sa + 16 * q);
                        wv[4 * q] = d.x; wv[4 * q + 1] = d.y; wv[4 * q + 2] = d.z; wv[4 * q + 3] = d.w;
                    }
                    sa += kStepPx * 3;
#pragma unroll
                    for (int jj = 0; jj < kStepPx; ++jj) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const int b = 3 * jj + ch;
                            rg[ch][jj] = (int)__byte_perm(wv[b >> 2], 0, 0x4440 + (b & 3));
                        }
                        if (m & (1u << jj)) {
                            int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
                            for (int tt = 0; tt < KT; ++tt) {
                                const int q = (jj - tt) & (kRing - 1);
                                a0 += rg[0][q] * kf[tt];
                                a1 += rg[1][q] * kf[tt];
                                a2 += rg[2][q] * kf[tt];
                            }
                            hdst[xi] = (unsigned char)clip8i(a0);
                            hdst1[xi] = (unsigned char)clip8i(a1);
                            hdst2[xi] = (unsigned char)clip8i(a2);
                            ++xi;
                            hp += STRIDE * 4;
                            load_coeffs16<KT>(kf, hp);                // the slot holds sw + 1 records: always readable
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(SE, slot));          // stage slot may be refilled
                    mbar_arrive(bar(HF, slot));          // H-ring slot is complete
                }
            }
        }
    } else if (warp >= kVBase) {
        // ============================== vertical pass (pull order) ==============================
        const int v = tid - kVBase * 32;
        int k = 0, nb = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const int n_chunks = (G.r_end - G.r_first + kChunk - 1) / kChunk;
            const int wpr = (S.x1 - S.x0) / 4;
            const bool v_active = v < 3 * wpr;
            const int vc = v_active ? v / wpr : 0;
            const int vwx = v_active ? v - vc * wpr : 0;
            const uint32_t thr_off = (uint32_t)(vc * L.oplane + vwx * 4);
            int py = 0;
            uint32_t otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * L.oplane) + thr_off;
            const uint8_t* const gm = sc.mask + G.mask_off;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(VF, slot), j & 1);
                mbar_wait(bar(HF, slot), j & 1);
                uint32_t vaddr = smem_u32(smem + L.off_vrec + slot * L.vrec_slot);
                // address of this thread's word in fresh row 0 of the slot; carry rows sit right above it
                const uint32_t hcol = smem_u32(smem + L.off_hring + slot * 3 * L.hplane + vc * L.hplane + vwx * 4);
                const uint32_t h0 = hcol + CARRY * L.hpitch;
#pragma unroll 1
                for (int g = 0; g < kChunk / kRing; ++g) {
                    uint32_t m = (uint32_t)gm[4 * (c * (kChunk / kRing) + g)] | ((uint32_t)gm[4 * (c * (kChunk / kRing) + g) + 1] << 8);
#pragma unroll 1
                    while (m) {
                        const int u = __ffs(m) - 1;
                        m &= m - 1;
                        const uint32_t hrow = h0 + (uint32_t)((g * kRing + u) * L.hpitch);
                        int kf[KT];
                        load_coeffs16<KT>(kf, vaddr);
                        vaddr += STRIDE * 4;
                        int acc[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e] = 1 << (VIS_PRECISION_BITS - 1);
#pragma unroll
                        for (int tt = 0; tt < KT; ++tt) {
                            const uint32_t wd = lds32(hrow - (uint32_t)(tt * L.hpitch));
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[e] += (int)__byte_perm(wd, 0, 0x4440 + e) * kf[tt];
                        }
                        const uint32_t lo = __byte_perm(clip8i(acc[0]), clip8i(acc[1]), 0x0040);
                        const uint32_t hi = __byte_perm(clip8i(acc[2]), clip8i(acc[3]), 0x0040);
                        if (v_active) sts32(otile_thr, __byte_perm(lo, hi, 0x5410));
                        otile_thr += L.opitch;
                        if (++py == VIS_PATCH) {                  // band complete: hand it to the store warps
                            band_done16(bar0, nb, lane);
                            ++nb;
                            py = 0;
                            otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * L.oplane) + thr_off;
                        }
                    }
                }
                if (c + 1 < n_chunks && v_active) {               // carry: last KT-1 fresh rows -> front of the other slot
                    const uint32_t src = h0 + (uint32_t)((kChunk - CARRY) * L.hpitch);
                    const uint32_t dst = smem_u32(smem + L.off_hring + (slot ^ 1) * 3 * L.hplane + vc * L.hplane + vwx * 4);
#pragma unroll
                    for (int i = 0; i < CARRY; ++i) sts32(dst + i * L.hpitch, lds32(src + i * L.hpitch));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(HE, slot));       // H-ring slot and record slot consumed
            }
            if (py) {                                             // a segment that ends inside a band (uint8 mode only)
                band_done16(bar0, nb, lane);
                ++nb;
            }
        }
    } else {
        // ============================== band store ==============================
        const int sw_i = warp - kSBase;
        int nb = 0;
        if (U8) {
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const int f = w / per_frame, r = w - f * per_frame;
                const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
                const VisSchedStrip S = sc.strip[st];
                const VisSchedSeg G = sc.seg[sg];
                const int wpr = (S.x1 - S.x0) / 4;                 // 4-pixel groups per row of the strip
                unsigned char* const dst0 = reinterpret_cast<unsigned char*>(frames[f].second) + (size_t)S.x0 * 3;
                for (int y = G.y0; y < G.y1; y += VIS_PATCH, ++nb) {
                    const int os = nb & 1, rows = min(VIS_PATCH, G.y1 - y);
                    mbar_wait(bar(OF, os), (nb >> 1) & 1);
                    const uint32_t otile = smem_u32(smem + L.off_otile + os * 3 * L.oplane);
                    for (int i = sw_i * 32 + lane; i < rows * wpr; i += kSWarps * 32) {
                        const int rr = i / wpr, q = i - rr * wpr;
                        const uint32_t at = otile + (uint32_t)(rr * L.opitch + q * 4);
                        const uint32_t A = lds32(at), B = lds32(at + L.oplane), C = lds32(at + 2 * L.oplane);
                        const uint32_t ab = __byte_perm(A, B, 0x5140), ab2 = __byte_perm(A, B, 0x7362);   // a0 b0 a1 b1 / a2 b2 a3 b3
                        uint32_t* o = reinterpret_cast<uint32_t*>(dst0 + (size_t)(y + rr) * dst_pitch + (size_t)q * 12);
                        o[0] = __byte_perm(ab, C, 0x2410);                                      // a0 b0 c0 a1
                        o[1] = __byte_perm(__byte_perm(ab, C, 0x0053), ab2, 0x5410);             // b1 c1 a2 b2
                        o[2] = __byte_perm(ab2, C, 0x7326);                                     // c2 a3 b3 c3
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(OE, os));
                }
            }
        } else {
            // lane-constant description of up to five 16-byte chunks (c, q) of a patch row: item = lane + 32 * i < 147
            int sa[5], sb[5], go[5], lo[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int item = min(lane + 32 * i, 146);
                const int c = item / 49, q = item - c * 49;
                const int f0 = 4 * q, f2 = f0 + 2;
                const int pya = f0 / VIS_PATCH, pyb = f2 / VIS_PATCH;
                sa[i] = c * L.oplane + pya * L.opitch + (f0 - pya * VIS_PATCH);
                sb[i] = c * L.oplane + pyb * L.opitch + (f2 - pyb * VIS_PATCH);
                go[i] = c * 392 + f0;
                lo[i] = c * 256;
            }
            const int half_gw = sc.dst_w / (2 * VIS_PATCH);
            for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
                const int f = w / per_frame, r = w - f * per_frame;
                const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
                const VisSchedStrip S = sc.strip[st];
                const VisSchedSeg G = sc.seg[sg];
                const int n_patches = (S.x1 - S.x0) / VIS_PATCH, gx0 = S.x0 / VIS_PATCH;
                float* const frame_out = pixel_values + (size_t)frames[f].second * VIS_ROW_FLOATS;
                for (int gy = G.y0 / VIS_PATCH; gy < G.y1 / VIS_PATCH; ++gy, ++nb) {
                    const int os = nb & 1;
                    mbar_wait(bar(OF, os), (nb >> 1) & 1);
                    const unsigned char* otile = smem + L.off_otile + os * 3 * L.oplane;
                    float* band = frame_out + (size_t)((gy >> 1) * half_gw * 4 + (gy & 1) * 2) * VIS_ROW_FLOATS;
                    for (int g = sw_i; g < n_patches; g += kSWarps) {
                        const int gx = gx0 + g;
                        float* prow = band + (size_t)((gx >> 1) * 4 + (gx & 1)) * VIS_ROW_FLOATS;
                        const unsigned char* pt = otile + g * VIS_PATCH;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (lane + 32 * i < 147) {
                                const unsigned a = *reinterpret_cast<const unsigned short*>(pt + sa[i]);
                                const unsigned b = *reinterpret_cast<const unsigned short*>(pt + sb[i]);
                                const float* l = lut + lo[i];
                                const float v0 = l[a & 0xff], v1 = l[a >> 8], v2 = l[b & 0xff], v3 = l[b >> 8];
                                stg128(prow + go[i], v0, v1, v2, v3);
                                stg128(prow + go[i] + 196, v0, v1, v2, v3);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(OE, os));
                }
            }
        }
    }
}

template <int KT, int STRIDE, bool U8, int NV>
int launch16(const VisSched& sc, const void* frames, int n_frames, const Layout16& L, int64_t dst_pitch, const int* hrec,
             const int* vrec, const float* lut768, float* pixel_values, cudaStream_t st) {
    auto kern = k_fused_sched16<KT, STRIDE, U8, NV>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_fused_sched16: cudaFuncSetAttribute");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_items = n_frames * sc.n_strips * sc.n_segs;
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, threads16(NV), L.total, st>>>(sc, reinterpret_cast<const FramePtrs*>(frames), n_items, L, (long long)dst_pitch,
                                            hrec, vrec, lut768, pixel_values);
    return vis::check_launch("vis_fused_sched16");
}

}  // namespace

namespace visf {

int sched16_subs() { return kHWarps; }

int sched16_max_strip_w(int nv) { return max_strip_w16(nv); }

int sched16_layout_bytes(int stage_pitch, int strip_w, int cls) { return make_layout16(stage_pitch, strip_w, cls).total; }

int sched16_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
                   const float* lut768, float* pixel_values, cudaStream_t st) {
    if (sc.kt < 12 || sc.kt > 32 || sc.ring != 16 || sc.n_subs != kHWarps || sc.per_index != 1 || (sc.kt <= 16 ? (sc.n_vwarps != 4 && sc.n_vwarps != 6) : (sc.n_vwarps != 3 && sc.n_vwarps != 4))) {
        vis::set_error("vis_fused_sched16: schedule of another kernel class (ring %d, %d taps, %d sub-ranges)", sc.ring, sc.kt, sc.n_subs);
        return VIS_E_INVALID;
    }
    const Layout16 L = make_layout16(sc.stage_pitch, sc.max_strip_w, sc.kt);
    if (L.total > kSmemMax) {
        vis::set_error("vis_fused_sched16: %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
    const bool u8 = sc.out_mode == VIS_SCHED_OUT_U8;
#define VIS_L16N(KT, NV) (u8 ? launch16<KT, ((KT + 5) & ~3), true, NV>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st) \
                            : launch16<KT, ((KT + 5) & ~3), false, NV>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st))
#define VIS_L16(KT) (sc.n_vwarps == 4 ? VIS_L16N(KT, 4) : VIS_L16N(KT, 6))              /* push-order H: 9..16 taps */
#define VIS_L16P(KT) (sc.n_vwarps == 3 ? VIS_L16N(KT, 3) : VIS_L16N(KT, 4))             /* pull-order H: 17..32 taps */
#define VIS_L16PU(KT, ST) (sc.n_vwarps == 3 ? launch16<KT, ST, true, 3>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st) \
                                            : launch16<KT, ST, true, 4>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st))
    switch (sc.kt) {
        case 12: return VIS_L16(12);
        case 13: return VIS_L16(13);          // exact classes of the strong downscales (4K -> 1316x728, the 1.875x and
        case 14: return VIS_L16(14);          // 2x LANCZOS thumbnails): 13 taps cost 13 MACs, not 16
        case 16: return VIS_L16(16);
        case 24: return VIS_L16P(24);
        case 32: return VIS_L16P(32);
        case 20: if (u8) return VIS_L16PU(20, 24); break;
        case 28: if (u8) return VIS_L16PU(28, 32); break;
        default: break;
    }
#undef VIS_L16P
#undef VIS_L16PU
#undef VIS_L16N
#undef VIS_L16
    vis::set_error("vis_fused_sched16: no instantiation for %d taps in this output mode", sc.kt);
    return VIS_E_UNSUPPORTED;
}

}  // namespace visf
