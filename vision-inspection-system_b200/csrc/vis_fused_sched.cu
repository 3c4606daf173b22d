// vis_fused_sched.cu — statically scheduled, warp-specialised hot kernel (frame -> Qwen2-VL pixel_values).
//
// Same arithmetic as vis_fused_ws.cu (Pillow 8bpc horizontal pass -> uint8 -> vertical pass -> uint8 ->
// exact LUT -> patch layout; tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:164-214), same pipeline of roles over
// shared-memory rings with full/empty mbarriers:
//
//   loader (1 warp)   cp.async.bulk: 32 input row segments per chunk, the strip's horizontal records, and the vertical
//                     records each chunk will consume                                   -> stage[2], hrec[2], vrec[2]
//   H      (12 warps) lane = input row, warp = column sub-range; push order: 8 input pixels per step are unpacked into a
//                     register window, an output pixel is emitted where the schedule says one ends   -> hring[2] (u8 planar)
//   V      (8 warps)  thread = 4 output columns of one channel; every H-ring row is unpacked once into a register ring,
//                     an output row is emitted where the schedule says one ends                       -> otile[2] (14-row band)
//   store  (3 warps)  band -> LUT -> 16-byte stores, both temporal copies                             -> pixel_values
//
// What is different: ONE geometry per launch, and the emission pattern ("an output sample's tap window ends at this
// input index") is precomputed on the host into bit masks that travel in the kernel parameter block (constant bank).
// Strip, segment, sub-range and mask data are therefore warp-uniform values; the compiler keeps them in uniform
// registers, every branch in the resampling loops is a uniform branch (no BSSY/BSYNC reconvergence pairs), and the
// per-sample window bookkeeping (compare against `last`, update, reload) of the general kernels disappears.  The
// general kernels were instruction-issue bound with ~45 % of their instructions being such bookkeeping
// (profiles/r01_fused_ws4.txt).
#include "vis_fused_common.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace visf;

namespace visf {   // 16-slot kernel, vis_fused_sched16.cu
int sched16_subs();
int sched16_max_strip_w(int nv);
int sched16_layout_bytes(int stage_pitch, int strip_w, int cls);
int sched16_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
                   const float* lut768, float* pixel_values, cudaStream_t st);
}

namespace visf {   // packed-byte (IDP.4A) kernel, vis_fused_dp.cu
int dp_subs();
int dp_max_strip_w(int nv);
int dp_layout_bytes(int stage_pitch, int strip_w, int words);
int dp_record_stride(int words);
int dp_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
              const float* lut768, float* pixel_values, cudaStream_t st);
}

namespace visf {   // integer tensor-path (IMMA) kernel, vis_fused_mma.cu
int mma_max_strip_w(int ksteps);
int mma_max_ksteps();
int mma_layout_bytes(int stage_pitch, int strip_w, int words);
int mma_record_stride(int words);
int mma_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
               const float* lut768, float* pixel_values, cudaStream_t st);
}

#ifndef VIS_MMA_MIN_CHUNK_ROWS
#define VIS_MMA_MIN_CHUNK_ROWS 24     // tensor-path kernel: smallest chunk advance (the horizontal pass computes 32 rows per chunk anyway)
#endif
#ifndef VIS_DP_NV_SPLIT
#define VIS_DP_NV_SPLIT 2.4       // packed-byte kernel: 4 vertical-pass warps from this vertical scale on, 6 below
#endif

namespace {

constexpr int kHWarps = VIS_SCHED_SUBS, kVWarps = 8, kSWarps = 3;
// warp ranges in priority order (the scheduler prefers the highest ready warp id): H < loader < S < V
constexpr int kHBase = 0, kLBase = kHWarps, kSBase = kHWarps + 1, kVBase = kHWarps + 1 + kSWarps;
constexpr int kThreadsS = (kHWarps + kVWarps + kSWarps + 1) * 32;       // 768
constexpr int kChunk = 32, kStepPx = 8, kRing = 8, kMaxStripW = 336;
constexpr int kPitch = kMaxStripW + 4;            // 340 = 4 * 85: conflict-free lane = row byte stores
constexpr int kOPitch = kMaxStripW;
constexpr int kOPlane = VIS_PATCH * kOPitch;
constexpr int kHPlane = kChunk * kPitch;
constexpr int kVRecsMax = 2 * kChunk + 1;         // vertical records a chunk can touch: <= 2 emits per input row + 1 look-ahead
constexpr int kSmemMax = 227 * 1024;

enum Bar { SF = 0, SE = 2, HF = 4, HE = 6, VF = 8, OF = 10, OE = 12, kBars = 14 };   // full/empty pairs, two slots each

struct LayoutS {
    int stage_pitch, stage_slot, hrec_slot, vrec_slot;
    int off_stage, off_hring, off_otile, off_hrec, off_vrec, off_lut, off_bar, total;
};

inline LayoutS make_layout_s(int stage_pitch, int strip_w, int stride) {
    LayoutS L;
    L.stage_pitch = stage_pitch;
    L.stage_slot = kChunk * stage_pitch;
    L.hrec_slot = align_up((strip_w + 1) * stride * 4, 16);
    L.vrec_slot = align_up(kVRecsMax * stride * 4, 16);
    int off = 0;
    L.off_stage = off; off += 2 * L.stage_slot;
    L.off_hring = off; off += 2 * 3 * kHPlane;
    L.off_otile = off; off += 2 * 3 * kOPlane;
    off = align_up(off, 16);
    L.off_hrec = off;  off += 2 * L.hrec_slot;
    L.off_vrec = off;  off += 2 * L.vrec_slot;
    L.off_lut = off;   off += 768 * 4;
    L.off_bar = off;   off += kBars * 8;
    L.total = off;
    return L;
}

// coefficient part of a record (the window bounds are not needed on the device: the schedule replaces them)
template <int KT, int STRIDE>
__device__ __forceinline__ void load_coeffs(int (&k)[KT], uint32_t addr) {
    static_assert(KT == 6 || KT == 8, "tap classes 6 and 8");
    const uint4 a = lds128(addr);
    k[0] = (int)a.x; k[1] = (int)a.y; k[2] = (int)a.z; k[3] = (int)a.w;
    if (KT == 6) {
        const uint2 b = lds64(addr + 16);
        k[4] = (int)b.x; k[5] = (int)b.y;
    } else {
        const uint4 b = lds128(addr + 16);
        k[4] = (int)b.x; k[5] = (int)b.y; k[KT - 2] = (int)b.z; k[KT - 1] = (int)b.w;
    }
}

// band `nb` of the output tile ring is complete: publish it to the store warps, then make sure the tile the next
// band goes to has been drained (out of line: once per 14 output rows, keeps the unrolled emit bodies small)
__device__ __noinline__ void band_done(uint32_t bar0, int nb, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + (uint32_t)(OF + (nb & 1)) * 8);
    const int nx = nb + 1;
    if (nx >= 2) mbar_wait(bar0 + (uint32_t)(OE + (nx & 1)) * 8, ((nx >> 1) - 1) & 1);
}

// UP: up to two output samples may end at one input index (mild upscaling, scale > 0.5): a second mask byte per step
// DUP: frames may name a second destination for their rows (dual Inspector + Auditor inputs, vis_preprocess_fused_sched_dup)
template <int KT, int STRIDE, bool UP, bool DUP>
__global__ void __launch_bounds__(kThreadsS, 1)
k_fused_sched(const __grid_constant__ VisSched sc, const VisFrameRef* __restrict__ frames, int n_items,
              const __grid_constant__ LayoutS L, const int* __restrict__ hrec_g, const int* __restrict__ vrec_g,
              const float* __restrict__ lut768, float* __restrict__ pixel_values,
              const long long* __restrict__ dup_rows) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);               // warp-uniform for the compiler
    float* lut = reinterpret_cast<float*>(smem + L.off_lut);              // transposed: lut[c * 256 + v]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    auto bar = [&](int which, int slot) { return bar0 + (uint32_t)(which + slot) * 8; };
    const int per_frame = sc.n_strips * sc.n_segs;

    for (int i = tid; i < 768; i += kThreadsS) lut[(i % 3) * 256 + i / 3] = __ldg(lut768 + i);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(SF, s), 1);
            mbar_init(bar(SE, s), kHWarps);
            mbar_init(bar(HF, s), kHWarps);
            mbar_init(bar(HE, s), kVWarps);
            mbar_init(bar(VF, s), 1);
            mbar_init(bar(OF, s), kVWarps);
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
                // vertical records this chunk consumes: [yo, yo + 33) (the table ends with a sentinel at dst_h)
                if (k >= 2) mbar_wait(bar(HE, slot), prev);                 // V is done with the record slot
                if (lane == 0) {
                    const uint32_t vbytes = (uint32_t)min(UP ? kVRecsMax : kChunk + 1, sc.dst_h + 1 - yo) * STRIDE * 4;
                    fence_proxy_async();
                    mbar_expect_tx(bar(VF, slot), vbytes);
                    bulk_g2s(smem_u32(smem + L.off_vrec + slot * L.vrec_slot), vrec_g + (size_t)yo * STRIDE, vbytes,
                             bar(VF, slot));
                }
                const uint32_t* m8 = reinterpret_cast<const uint32_t*>(sc.mask + G.mask_off + c * 8);   // 4 groups x (m1, m2)
                yo += __popc(m8[0]) + __popc(m8[1]);
            }
        }
    } else if (warp < kHBase + kHWarps) {
        // ============================== horizontal pass ==============================
        const int sub = warp - kHBase;
        int k = 0, sl = 0;
        int rg[3][kRing];                                  // the last 8 input pixels per channel (static slots)
#pragma unroll
        for (int q = 0; q < kRing; ++q) rg[0][q] = rg[1][q] = rg[2][q] = 0;
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
                unsigned char* hdst = smem + L.off_hring + slot * 3 * kHPlane + lane * kPitch + (U.xa - S.x0);
                uint32_t hp = hrec0;
                int kf[KT];
                load_coeffs<KT, STRIDE>(kf, hp);
#pragma unroll 1
                for (int i = 0; i < U.nsteps; ++i) {
                    const uint32_t m = um[2 * i], m2 = UP ? um[2 * i + 1] : 0u;
                    uint32_t wv[6];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint2 d = lds64(sa + 8 * q);
                        wv[2 * q] = d.x; wv[2 * q + 1] = d.y;
                    }
                    sa += kStepPx * 3;
#pragma unroll
                    for (int jj = 0; jj < kStepPx; ++jj) {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const int b = 3 * jj + ch;
                            rg[ch][jj] = (int)__byte_perm(wv[b >> 2], 0, 0x4440 + (b & 3));
                        }
                        auto emit = [&]() {
                            int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
                            for (int tt = 0; tt < KT; ++tt) {
                                const int q = (jj - tt) & (kRing - 1);
                                a0 += rg[0][q] * kf[tt];
                                a1 += rg[1][q] * kf[tt];
                                a2 += rg[2][q] * kf[tt];
                            }
                            hdst[0] = (unsigned char)clip8i(a0);
                            hdst[kHPlane] = (unsigned char)clip8i(a1);
                            hdst[2 * kHPlane] = (unsigned char)clip8i(a2);
                            ++hdst;
                            hp += STRIDE * 4;
                            load_coeffs<KT, STRIDE>(kf, hp);          // the slot holds sw + 1 records: always readable
                        };
                        if (m & (1u << jj)) emit();
                        if (UP && (m2 & (1u << jj))) emit();
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
        // ============================== vertical pass ==============================
        const int v = tid - kVBase * 32;
        int k = 0, nb = 0;
        int ring[kRing][4];                                // the last 8 H-ring rows of this thread's 4 columns
#pragma unroll
        for (int q = 0; q < kRing; ++q) ring[q][0] = ring[q][1] = ring[q][2] = ring[q][3] = 0;
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
            const uint32_t thr_off = (uint32_t)(vc * kOPlane + vwx * 4);
            int py = 0;
            uint32_t otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * kOPlane) + thr_off;
            const uint8_t* const gm = sc.mask + G.mask_off;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(VF, slot), j & 1);
                mbar_wait(bar(HF, slot), j & 1);
                uint32_t vaddr = smem_u32(smem + L.off_vrec + slot * L.vrec_slot);
                int kf[KT];
                load_coeffs<KT, STRIDE>(kf, vaddr);
                const uint32_t hsrc = smem_u32(smem + L.off_hring + slot * 3 * kHPlane + vc * kHPlane + vwx * 4);
                const int groups = min(kChunk / kRing, (G.r_end - (G.r_first + c * kChunk) + kRing - 1) / kRing);
#pragma unroll 1
                for (int g = 0; g < groups; ++g) {
                    const uint32_t m = gm[2 * (c * (kChunk / kRing) + g)], m2 = UP ? gm[2 * (c * (kChunk / kRing) + g) + 1] : 0u;
                    uint32_t words[kRing];
#pragma unroll
                    for (int u = 0; u < kRing; ++u) words[u] = lds32(hsrc + (uint32_t)((g * kRing + u) * kPitch));
#pragma unroll
                    for (int u = 0; u < kRing; ++u) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) ring[u][e] = (int)__byte_perm(words[u], 0, 0x4440 + e);
                        auto emit = [&]() {
                            int acc[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[e] = 1 << (VIS_PRECISION_BITS - 1);
#pragma unroll
                            for (int tt = 0; tt < KT; ++tt) {
                                const int q = (u - tt) & (kRing - 1);
#pragma unroll
                                for (int e = 0; e < 4; ++e) acc[e] += ring[q][e] * kf[tt];
                            }
                            const uint32_t lo = __byte_perm(clip8i(acc[0]), clip8i(acc[1]), 0x0040);
                            const uint32_t hi = __byte_perm(clip8i(acc[2]), clip8i(acc[3]), 0x0040);
                            if (v_active) sts32(otile_thr, __byte_perm(lo, hi, 0x5410));
                            otile_thr += kOPitch;
                            if (++py == VIS_PATCH) {                  // band complete: hand it to the store warps
                                band_done(bar0, nb, lane);
                                ++nb;
                                py = 0;
                                otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * kOPlane) + thr_off;
                            }
                            vaddr += STRIDE * 4;
                            load_coeffs<KT, STRIDE>(kf, vaddr);
                        };
                        if (m & (1u << u)) emit();
                        if (UP && (m2 & (1u << u))) emit();
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(HE, slot));       // H-ring slot and record slot consumed
            }
        }
    } else {
        // ============================== band store ==============================
        const int sw_i = warp - kSBase;
        // lane-constant description of up to five 16-byte chunks (c, q) of a patch row: item = lane + 32 * i < 147
        int sa[5], sb[5], go[5], lo[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int item = min(lane + 32 * i, 146);
            const int c = item / 49, q = item - c * 49;
            const int f0 = 4 * q, f2 = f0 + 2;
            const int pya = f0 / VIS_PATCH, pyb = f2 / VIS_PATCH;
            sa[i] = c * kOPlane + pya * kOPitch + (f0 - pya * VIS_PATCH);
            sb[i] = c * kOPlane + pyb * kOPitch + (f2 - pyb * VIS_PATCH);
            go[i] = c * 392 + f0;
            lo[i] = c * 256;
        }
        int nb = 0;
        const int half_gw = sc.dst_w / (2 * VIS_PATCH);
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const int n_patches = (S.x1 - S.x0) / VIS_PATCH, gx0 = S.x0 / VIS_PATCH;
            float* const frame_out = pixel_values + (size_t)frames[f].row0 * VIS_ROW_FLOATS;
            // dual Inspector + Auditor inputs: a frame neither agent thumbnails gives both the SAME rows — they are
            // written to the second tensor from the same registers (dup_rows[f] >= 0), not computed or copied again
            const long long dup = DUP ? dup_rows[f] : -1;
            const long long dup_off = DUP && dup >= 0 ? (dup - frames[f].row0) * (long long)VIS_ROW_FLOATS : 0;
            for (int gy = G.y0 / VIS_PATCH; gy < G.y1 / VIS_PATCH; ++gy, ++nb) {
                const int os = nb & 1;
                mbar_wait(bar(OF, os), (nb >> 1) & 1);
                const unsigned char* otile = smem + L.off_otile + os * 3 * kOPlane;
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
                            if (DUP && dup >= 0) {
                                stg128(prow + dup_off + go[i], v0, v1, v2, v3);
                                stg128(prow + dup_off + go[i] + 196, v0, v1, v2, v3);
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(OE, os));
            }
        }
    }
}

template <int KT, int STRIDE, bool UP, bool DUP>
int launch_sched(const VisSched& sc, const VisFrameRef* frames, int n_frames, const LayoutS& L, const int* hrec,
                 const int* vrec, const float* lut768, float* pixel_values, const long long* dup_rows, cudaStream_t st) {
    auto kern = k_fused_sched<KT, STRIDE, UP, DUP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_preprocess_fused_sched: cudaFuncSetAttribute");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_items = n_frames * sc.n_strips * sc.n_segs;
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, kThreadsS, L.total, st>>>(sc, frames, n_items, L, hrec, vrec, lut768, pixel_values, dup_rows);
    return vis::check_launch("vis_preprocess_fused_sched");
}

// Window ends as the schedule uses them: non-decreasing, at most `per_index` samples ending at one input index.
// With scale >= 1 the true ends (first + taps - 1) strictly increase (per_index = 1), except at the far border where
// Pillow clamps the window to the image and the last few samples all end at the last input index; with a mild
// upscale (scale > 0.5) two samples may share an end (per_index = 2).  Samples beyond the capacity of their index are
// moved to the next (possibly virtual, past the border) index; their records are packed with as many leading zero
// coefficients, so the extra samples (whatever the staging buffer holds there) get weight 0.
// Returns false when some sample would need more than kt slots: not expressible as a schedule of this width.
inline bool schedule_ends(const int32_t* b, int n, int kt, int per_index, std::vector<int>& ends) {
    ends.resize(n);
    int used = 0;
    for (int i = 0; i < n; ++i) {
        int e = b[2 * i] + b[2 * i + 1] - 1;
        if (i > 0) {
            if (b[2 * i] < b[2 * i - 2]) return false;
            if (e < ends[i - 1]) e = ends[i - 1];
            if (e == ends[i - 1]) {
                if (used >= per_index) { e += 1; used = 1; } else ++used;
            } else {
                used = 1;
            }
        } else {
            used = 1;
        }
        if (e - b[2 * i] + 1 > kt) return false;
        ends[i] = e;
    }
    return true;
}

inline int stage_pitch_for(int span_bytes) {
    int p = align_up(span_bytes, 16);
    if ((p / 16) % 2 == 0) p += 16;               // odd multiple of 16: conflict-free lane = row wide loads
    return p;
}

}  // namespace

extern "C" {

int vis_sched_sizeof(void) { return (int)sizeof(VisSched); }

int vis_sched_build(int src_h, int src_w, int dst_h, int dst_w, int64_t src_pitch,
                    const int32_t* hb, const int32_t* vb, int vsplit, int out_mode, VisSched* out) {
    if (!hb || !vb || !out || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || vsplit < 1 ||
        ((out_mode & 0xff) != VIS_SCHED_OUT_PIXEL_VALUES && (out_mode & 0xff) != VIS_SCHED_OUT_U8)) {
        vis::set_error("vis_sched_build: bad arguments");
        return VIS_E_INVALID;
    }
    auto unsupported = [](const char* why) { vis::set_error("vis_sched_build: %s", why); return VIS_E_UNSUPPORTED; };
    const bool want_mma = (out_mode & VIS_SCHED_FLAG_MMA) != 0;
    const bool want_dp = want_mma || (out_mode & VIS_SCHED_FLAG_DP4A) != 0;
    const int mode_in = out_mode;
    out_mode &= 0xff;
    const bool u8 = out_mode == VIS_SCHED_OUT_U8;
    if (!u8 && (dst_h % 28 || dst_w % 28)) return unsupported("output size is not a multiple of 28");
    if (u8 && dst_w % 4) return unsupported("output width is not a multiple of 4");
    if (src_pitch % 16 || src_pitch < (int64_t)src_w * 3) return unsupported("row pitch must be a multiple of 16");
    if ((int64_t)src_h > 100 * (int64_t)src_w && dst_h < src_h) return unsupported("vertical-first pass order");
    if (src_w > 65528 || dst_w > 65528) return unsupported("image too wide");
    const int hkt = vis_max_taps(hb, dst_w), vkt = vis_max_taps(vb, dst_h);
    const int mk = hkt > vkt ? hkt : vkt;
    // tap class -> kernel: <= 8 taps: 8-slot register windows (pixel_values only); <= 16 taps: 16-slot kernel
    // and, past 16 taps, the same kernel with a pull-order horizontal role (classes 20/24/28/32; 24/32 for pixel_values)
    int cls = mk <= 12 ? 12 : mk <= 13 ? 13 : mk <= 14 ? 14 : mk <= 16 ? 16 : mk <= 20 ? 20 : mk <= 24 ? 24 : mk <= 28 ? 28 : mk <= 32 ? 32 : 0;
    if (!u8) cls = mk <= 6 ? 6 : mk <= 8 ? 8 : cls == 20 ? 24 : cls == 28 ? 32 : cls;
    if (!cls) return unsupported("more than 32 taps");
    const int ring = cls <= 8 ? 8 : 16;
    std::vector<int> hl, vl;                      // scheduled window ends (virtual past the far border)
    // packed-byte kernel: a window is W words of 4 taps ending in the word of its (virtual) end: any window of up to
    // 4W - 3 taps fits whatever its alignment; W grows until the far-border samples (virtual ends) fit as well
    int dp_words = 0;
    if (want_dp && ring == 16) {
        for (int w = (mk + 3 + 3) / 4 < 4 ? 4 : (mk + 3 + 3) / 4; w <= 9 && !dp_words; ++w)
            if (schedule_ends(hb, dst_w, 4 * w - 3, 1, hl) && schedule_ends(vb, dst_h, 4 * w - 3, 1, vl)) dp_words = w;
        if (dp_words) cls = 4 * dp_words - 3;
    }
    const bool h_pull = !dp_words && cls > 16;
    // same limb records, both passes on the integer tensor path.  Its vertical pass takes ONE M-tile of 16 output rows per
    // chunk at full efficiency, so a chunk advances by as many input rows (32, 28 or 24: whole 4-row words) as never emit
    // more than 16; the horizontal pass still computes 32 rows per chunk (the rest is unused), so below 24 rows per chunk
    // the packed-byte kernel is the better one.
    bool mma = want_mma && dp_words > 0;
    int chunk_rows = kChunk;
    const int prow = (dst_h + 13) / 14;
    const int n_segs_eff = std::max(1, std::min(std::min(vsplit, prow), (int)VIS_SCHED_MAX_SEGS));
    auto seg_rows = [&](int g, int* y0, int* y1) {
        *y0 = (int)((int64_t)prow * g / n_segs_eff) * 14;
        *y1 = std::min((int)((int64_t)prow * (g + 1) / n_segs_eff) * 14, dst_h);
    };
    if (mma) {
        // window ends inside ANY cr consecutive input rows, wherever the segments put their chunk bases: the choice is a
        // property of the geometry (the records are cached per geometry, whatever the segment count).  More than 16
        // (a second, nearly empty M-tile) is allowed for a few windows: the samples Pillow clamps at the far border
        // crowd into its last rows.
        auto crowded = [&](int cr) {
            int over = 0;
            for (int y = 0, z = 0; y < dst_h; ++y) {
                while (z < dst_h && vl[z] < vl[y] + cr) ++z;
                over += z - y > 16;
            }
            return over * 20 > dst_h;                      // more than 5 % of the windows
        };
        chunk_rows = 0;
        for (int cr = kChunk; cr >= VIS_MMA_MIN_CHUNK_ROWS && !chunk_rows; cr -= 4)
            if (!crowded(cr)) chunk_rows = cr;
        if (!chunk_rows) { mma = false; chunk_rows = kChunk; }
    }
    int mma_ks = 0;
    if (mma) {
        // k-steps of 32 input pixels that the window of a tile of 16 outputs spans, counted from the 4-pixel word of the
        // tile's first tap.  Tiles start at multiples of 4 columns wherever the strips fall, so this is a property of the
        // geometry (the records are cached per geometry); beyond the kernel's register budget: the packed-byte kernel.
        for (int x = 0; x < dst_w; x += 4) {
            int lastpx = 0;
            for (int o = x; o < x + 16 && o < dst_w; ++o) lastpx = std::max(lastpx, hb[2 * o] + hb[2 * o + 1] - 1);
            mma_ks = std::max(mma_ks, (lastpx + 1 - (hb[2 * x] & ~3) + 31) / 32);
        }
        if (mma_ks > visf::mma_max_ksteps())
            return vis_sched_build(src_h, src_w, dst_h, dst_w, src_pitch, hb, vb, vsplit, (mode_in & ~VIS_SCHED_FLAG_MMA) | VIS_SCHED_FLAG_DP4A, out);
    }
    const int n_subs = ring == 8 ? 12 : dp_words ? visf::dp_subs() : visf::sched16_subs();
    // 16-slot kernels: fewer vertical-pass warps the stronger the vertical downscale (the V role gets lighter)
    const double vscale = (double)src_h / dst_h;
    const int n_vwarps = ring == 8 ? 0 : dp_words ? (vscale >= VIS_DP_NV_SPLIT ? 4 : 6) : cls > 16 ? (vscale >= 3.4 ? 3 : 4) : (vscale >= 2.4 ? 4 : 6);
    const int max_w = ring == 8 ? kMaxStripW : mma ? visf::mma_max_strip_w(mma_ks) : dp_words ? visf::dp_max_strip_w(n_vwarps) : visf::sched16_max_strip_w(n_vwarps);
    int per_index = 1;
    // an exact class (13, 14) may be too tight for the virtual ends of the far-border samples: widen it
    while ((cls == 13 || cls == 14) && (!schedule_ends(hb, dst_w, cls, 1, hl) || !schedule_ends(vb, dst_h, cls, 1, vl)))
        cls = cls == 13 ? 14 : 16;
    if (!schedule_ends(hb, dst_w, cls, 1, hl) || !schedule_ends(vb, dst_h, cls, 1, vl)) {
        per_index = 2;                            // mild upscale on some axis: two samples per input index
        if (ring != 8) return unsupported("upscaling with more than 8 taps");
        if (!schedule_ends(hb, dst_w, cls, 2, hl)) return unsupported("horizontal upscale beyond two samples per input column");
        if (!schedule_ends(vb, dst_h, cls, 2, vl)) return unsupported("vertical upscale beyond two samples per input row");
    }

    VisSched& s = *out;
    std::memset(&s, 0, sizeof(s));
    s.src_h = src_h; s.src_w = src_w; s.dst_h = dst_h; s.dst_w = dst_w; s.src_pitch = src_pitch; s.kt = cls;
    s.per_index = per_index; s.ring = ring; s.n_subs = n_subs; s.out_mode = out_mode; s.h_pull = h_pull ? 1 : 0;
    s.n_vwarps = n_vwarps; s.dp_words = dp_words; s.mma_ks = mma_ks; s.chunk_rows = chunk_rows;
    const int stride = dp_words ? visf::dp_record_stride(dp_words) : vis_record_stride(cls);
    auto last = [](const int32_t* b, int i) { return b[2 * i] + b[2 * i + 1] - 1; };
    auto span_of = [&](int x0, int x1, int* px0) {
        *px0 = hb[2 * x0] & ~15;                                   // 48-byte aligned: bulk copies need 16
        int bytes = align_up((last(hb, x1 - 1) + 1) * 3, 16) - *px0 * 3;
        if ((int64_t)*px0 * 3 + bytes > src_pitch) bytes = (int)(src_pitch - (int64_t)*px0 * 3);
        return bytes;
    };
    auto layout_bytes = [&](int pitch, int strip_w) {
        return ring == 8 ? make_layout_s(pitch, strip_w, stride).total
                         : mma ? visf::mma_layout_bytes(pitch, strip_w, dp_words)
                         : dp_words ? visf::dp_layout_bytes(pitch, strip_w, dp_words) : visf::sched16_layout_bytes(pitch, strip_w, cls);
    };
    // column strips: the fewest (widest, <= 336 columns) whose shared-memory layout fits; strip edges are multiples
    // of 28 columns (pixel_values: whole patches) or 4 columns (uint8 rows: whole 32-bit words of a plane)
    const int unit = u8 ? 4 : 28;
    const int blocks = dst_w / unit;
    int n_strips = 0;
    for (int n = (dst_w + max_w - 1) / max_w; n <= VIS_SCHED_MAX_STRIPS && n <= blocks && !n_strips; ++n) {
        int worst_span = 0, worst_w = 0;
        for (int i = 0; i < n; ++i) {
            const int b0 = (int)((int64_t)blocks * i / n), b1 = (int)((int64_t)blocks * (i + 1) / n);
            int px0;
            const int span = span_of(b0 * unit, b1 * unit, &px0);
            worst_span = span > worst_span ? span : worst_span;
            worst_w = (b1 - b0) * unit > worst_w ? (b1 - b0) * unit : worst_w;
        }
        const int pitch_n = stage_pitch_for(worst_span);
        if (worst_w <= max_w && layout_bytes(pitch_n, worst_w) <= kSmemMax) {
            n_strips = n;
            s.stage_pitch = pitch_n;
            s.max_strip_w = worst_w;
        }
    }
    if (!n_strips) return unsupported("no strip width fits shared memory");
    s.n_strips = n_strips;
    const int step = ring, mbytes = ring / 8;                      // input pixels (rows) per mask word, bytes per mask
    int mask_at = 0;
    auto mask_room = [&](int bytes) { return mask_at + bytes <= VIS_SCHED_MASK_BYTES; };
    auto set_bit = [&](int base, int rel) {                        // per step: first-sample mask, then second-sample mask
        uint8_t* m = s.mask + base + 2 * mbytes * (rel / step) + (rel % step) / 8;
        const uint8_t bit = (uint8_t)(1u << (rel % 8));
        if (m[0] & bit) m[mbytes] |= bit; else m[0] |= bit;
    };
    for (int i = 0; i < n_strips; ++i) {
        VisSchedStrip& S = s.strip[i];
        S.x0 = (int)((int64_t)blocks * i / n_strips) * unit;
        S.x1 = (int)((int64_t)blocks * (i + 1) / n_strips) * unit;
        S.row_bytes = span_of(S.x0, S.x1, &S.px0);
        const int sw = S.x1 - S.x0;
        for (int u = 0; u < n_subs; ++u) {
            VisSchedSub& U = s.sub[i][u];
            const int xa = S.x0 + (int)((int64_t)sw * u / n_subs), xb = S.x0 + (int)((int64_t)sw * (u + 1) / n_subs);
            U.xa = (uint16_t)xa; U.xb = (uint16_t)xb;
            if (xa >= xb || h_pull || mma) continue;                   // nsteps = 0: nothing to walk (pull order / tiles need no masks)
            const int p0 = hb[2 * xa] & ~(step - 1);
            const int nsteps = (hl[xb - 1] - p0) / step + 1;
            if (!mask_room(2 * mbytes * nsteps) || nsteps > 65535) return unsupported("schedule too large");
            U.p0 = (uint16_t)p0; U.nsteps = (uint16_t)nsteps; U.mask_off = (uint16_t)mask_at;
            for (int x = xa; x < xb; ++x) set_bit(mask_at, hl[x] - p0);
            mask_at += 2 * mbytes * nsteps;
        }
    }
    // row segments: edges at multiples of 14 output rows (the last one ends at dst_h); chunk bases are multiples of 16
    vsplit = n_segs_eff;
    s.n_segs = vsplit;
    mask_at = align_up(mask_at, 8);
    for (int g = 0; g < vsplit; ++g) {
        VisSchedSeg& G = s.seg[g];
        seg_rows(g, &G.y0, &G.y1);
        G.r_first = vb[2 * G.y0] & ~15;
        G.r_end = vl[G.y1 - 1] + 1;                                  // may exceed src_h by the virtual rows
        const int n_chunks = (G.r_end - G.r_first + chunk_rows - 1) / chunk_rows;
        const int bytes = 2 * mbytes * n_chunks * (kChunk / step);
        if (!mask_room(bytes)) return unsupported("schedule too large");
        G.mask_off = mask_at;
        for (int y = G.y0; y < G.y1; ++y) {                          // chunk = rel / chunk_rows; bit = row inside the chunk
            const int rel = vl[y] - G.r_first;
            set_bit(mask_at, rel / chunk_rows * kChunk + rel % chunk_rows);
        }
        mask_at += bytes;
    }
    return VIS_OK;
}

int vis_sched_pack_records(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int kt, int per_index,
                           int32_t* rec, int64_t rec_capacity) {
    if (out_size <= 0 || !k || !bounds || !rec || ksize <= 0 || kt < 6 || kt > 33 ||
        per_index < 1 || per_index > 2) {
        vis::set_error("vis_sched_pack_records: bad arguments");
        return VIS_E_INVALID;
    }
    const int stride = vis_record_stride(kt);
    if (rec_capacity < (int64_t)(out_size + 1) * stride) {
        vis::set_error("vis_sched_pack_records: capacity too small");
        return VIS_E_CAPACITY;
    }
    std::vector<int> ends;
    if (!schedule_ends(bounds, out_size, kt, per_index, ends)) {
        vis::set_error("vis_sched_pack_records: table is not schedulable (upscale or too many taps)");
        return VIS_E_UNSUPPORTED;
    }
    std::memset(rec, 0, sizeof(int32_t) * (size_t)(out_size + 1) * stride);
    for (int o = 0; o < out_size; ++o) {
        const int first = bounds[2 * o], taps = bounds[2 * o + 1];
        const int shift = ends[o] - (first + taps - 1);             // leading zero slots for virtual samples
        int32_t* r = rec + (size_t)o * stride;
        for (int t = 0; t < taps; ++t) r[shift + t] = k[(size_t)o * ksize + (taps - 1 - t)];
        r[stride - 2] = first;
        r[stride - 1] = ends[o];
    }
    return VIS_OK;
}

int vis_sched_record_stride_dp(int words) {
    if (words < 4 || words > 9) return VIS_E_INVALID;
    return visf::dp_record_stride(words);
}

int vis_sched_pack_records_dp(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int words,
                              int32_t* rec, int64_t rec_capacity) {
    if (out_size <= 0 || !k || !bounds || !rec || ksize <= 0 || words < 4 || words > 9) {
        vis::set_error("vis_sched_pack_records_dp: bad arguments");
        return VIS_E_INVALID;
    }
    const int stride = visf::dp_record_stride(words);
    if (rec_capacity < (int64_t)(out_size + 1) * stride) {
        vis::set_error("vis_sched_pack_records_dp: capacity too small");
        return VIS_E_CAPACITY;
    }
    std::vector<int> ends;
    if (!schedule_ends(bounds, out_size, 4 * words - 3, 1, ends)) {
        vis::set_error("vis_sched_pack_records_dp: table is not schedulable with %d words (upscale or too many taps)", words);
        return VIS_E_UNSUPPORTED;
    }
    std::memset(rec, 0, sizeof(int32_t) * (size_t)(out_size + 1) * stride);
    for (int o = 0; o < out_size; ++o) {
        const int first = bounds[2 * o], taps = bounds[2 * o + 1], end = ends[o];
        const int base = 4 * ((end >> 2) - (words - 1));            // input index of byte 0 of word 0 (may be negative)
        if (end < first + taps - 1 || first < base) {
            vis::set_error("vis_sched_pack_records_dp: window of sample %d ([%d, %d), end %d) does not fit %d words", o, first,
                           first + taps, end, words);
            return VIS_E_UNSUPPORTED;
        }
        unsigned char* r = reinterpret_cast<unsigned char*>(rec + (size_t)o * stride);
        for (int t = 0; t < taps; ++t) {
            const int32_t c = k[(size_t)o * ksize + t];
            if (c < -(1 << 23) || c >= (1 << 23)) {
                vis::set_error("vis_sched_pack_records_dp: coefficient %d outside three byte limbs", c);
                return VIS_E_UNSUPPORTED;
            }
            const int at = first + t - base;                         // byte position inside the W-word window
            r[at] = (unsigned char)(c & 0xff);
            r[4 * words + at] = (unsigned char)((c >> 8) & 0xff);
            r[8 * words + at] = (unsigned char)((c >> 16) & 0xff);   // signed limb (arithmetic shift), two's complement byte
        }
    }
    return VIS_OK;
}

int vis_sched_record_stride_mma(int words) {
    if (words < 4 || words > 9) return VIS_E_INVALID;
    return visf::mma_record_stride(words);
}

int vis_sched_pack_records_mma(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int words,
                               int32_t* rec, int64_t rec_capacity) {
    if (out_size <= 0 || !k || !bounds || !rec || ksize <= 0 || words < 4 || words > 9) {
        vis::set_error("vis_sched_pack_records_mma: bad arguments");
        return VIS_E_INVALID;
    }
    const int stride = visf::mma_record_stride(words);
    if (rec_capacity < (int64_t)(out_size + 1) * stride) {
        vis::set_error("vis_sched_pack_records_mma: capacity too small");
        return VIS_E_CAPACITY;
    }
    std::vector<int> ends;
    if (!schedule_ends(bounds, out_size, 4 * words - 3, 1, ends)) {
        vis::set_error("vis_sched_pack_records_mma: table is not schedulable with %d words (upscale or too many taps)", words);
        return VIS_E_UNSUPPORTED;
    }
    std::memset(rec, 0, sizeof(int32_t) * (size_t)(out_size + 1) * stride);
    for (int o = 0; o < out_size; ++o) {
        const int first = bounds[2 * o], taps = bounds[2 * o + 1], end = ends[o];
        const int bw = (end >> 2) - (words - 1);                    // absolute word index of record byte 0 (may be negative)
        if (end < first + taps - 1 || first < 4 * bw) {
            vis::set_error("vis_sched_pack_records_mma: window of sample %d ([%d, %d), end %d) does not fit %d words", o, first,
                           first + taps, end, words);
            return VIS_E_UNSUPPORTED;
        }
        int32_t* r32 = rec + (size_t)o * stride;
        unsigned char* r = reinterpret_cast<unsigned char*>(r32);
        for (int t = 0; t < taps; ++t) {
            const int32_t c = k[(size_t)o * ksize + t];
            if (c < -(1 << 23) || c >= (1 << 23)) {
                vis::set_error("vis_sched_pack_records_mma: coefficient %d outside three byte limbs", c);
                return VIS_E_UNSUPPORTED;
            }
            const int at = first + t - 4 * bw;                       // byte position inside the W-word window
            unsigned char* q = r + 12 * (at >> 2) + (at & 3);          // word at >> 2: limbs 0, 1, 2 in consecutive 32-bit words
            q[0] = (unsigned char)(c & 0xff);
            q[4] = (unsigned char)((c >> 8) & 0xff);
            q[8] = (unsigned char)((c >> 16) & 0xff);                // signed limb (arithmetic shift), two's complement byte
        }
        r32[3 * words] = bw;
        r32[3 * words + 1] = first >> 2;
    }
    return VIS_OK;
}

int vis_preprocess_fused_sched_dup(const VisSched* sched, const VisFrameRef* frames, int n_frames,
                                   const int32_t* hrec, const int32_t* vrec,
                                   const float* lut768, float* pixel_values, const int64_t* dup_rows, void* stream) {
    if (!sched || !frames || !hrec || !vrec || !lut768 || !pixel_values || n_frames <= 0 ||
        sched->out_mode != VIS_SCHED_OUT_PIXEL_VALUES || sched->n_strips <= 0 || sched->n_segs <= 0) {
        vis::set_error("vis_preprocess_fused_sched: bad arguments");
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long* dup = reinterpret_cast<const long long*>(dup_rows);
    if (sched->ring == 16) {
        if (dup) {
            vis::set_error("vis_preprocess_fused_sched_dup: duplicate rows are served by the 8-slot kernel only (<= 8 taps)");
            return VIS_E_UNSUPPORTED;
        }
        if (sched->mma_ks) return mma_launch(*sched, frames, n_frames, 0, hrec, vrec, lut768, pixel_values, st);
        if (sched->dp_words) return dp_launch(*sched, frames, n_frames, 0, hrec, vrec, lut768, pixel_values, st);
        return sched16_launch(*sched, frames, n_frames, 0, hrec, vrec, lut768, pixel_values, st);
    }
    if (sched->ring != 8 || (sched->kt != 6 && sched->kt != 8)) {
        vis::set_error("vis_preprocess_fused_sched: schedule of an unknown kernel class (ring %d, %d taps)", sched->ring, sched->kt);
        return VIS_E_INVALID;
    }
    const LayoutS L = make_layout_s(sched->stage_pitch, sched->max_strip_w, vis_record_stride(sched->kt));
    if (L.total > kSmemMax) {
        vis::set_error("vis_preprocess_fused_sched: %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
#define VIS_LS(KT, ST, UP) (dup ? launch_sched<KT, ST, UP, true>(*sched, frames, n_frames, L, hrec, vrec, lut768, pixel_values, dup, st) \
                                : launch_sched<KT, ST, UP, false>(*sched, frames, n_frames, L, hrec, vrec, lut768, pixel_values, dup, st))
    if (sched->per_index > 1) return sched->kt == 6 ? VIS_LS(6, 8, true) : VIS_LS(8, 12, true);
    return sched->kt == 6 ? VIS_LS(6, 8, false) : VIS_LS(8, 12, false);
#undef VIS_LS
}

int vis_preprocess_fused_sched(const VisSched* sched, const VisFrameRef* frames, int n_frames,
                               const int32_t* hrec, const int32_t* vrec,
                               const float* lut768, float* pixel_values, void* stream) {
    return vis_preprocess_fused_sched_dup(sched, frames, n_frames, hrec, vrec, lut768, pixel_values, nullptr, stream);
}

int vis_resize_fused_sched(const VisSched* sched, const VisResizeRef* frames, int n_frames, int64_t dst_pitch,
                           const int32_t* hrec, const int32_t* vrec, void* stream) {
    if (!sched || !frames || !hrec || !vrec || n_frames <= 0 || sched->out_mode != VIS_SCHED_OUT_U8 || sched->ring != 16 ||
        sched->n_strips <= 0 || sched->n_segs <= 0 || dst_pitch % 4 || dst_pitch < (int64_t)sched->dst_w * 3) {
        vis::set_error("vis_resize_fused_sched: bad arguments");
        return VIS_E_INVALID;
    }
    if (sched->mma_ks) return mma_launch(*sched, frames, n_frames, dst_pitch, hrec, vrec, nullptr, nullptr, (cudaStream_t)stream);
    if (sched->dp_words) return dp_launch(*sched, frames, n_frames, dst_pitch, hrec, vrec, nullptr, nullptr, (cudaStream_t)stream);
    return sched16_launch(*sched, frames, n_frames, dst_pitch, hrec, vrec, nullptr, nullptr, (cudaStream_t)stream);
}

}  // extern "C"
