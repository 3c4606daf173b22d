// vis_fused_dp.cu — statically scheduled, warp-specialised kernel for 9..33-tap windows on packed bytes (IDP.4A).
//
// Same path and arithmetic as vis_fused_sched16.cu (Pillow 8bpc horizontal pass -> uint8 -> vertical pass -> uint8, then
// either LUT + Qwen2-VL patch layout or RGB uint8 rows), same loader / store roles and mbarrier rings, same VisSched.
// What changes is how the two resampling roles multiply.  The 16-slot kernel spends one IMAD per tap on a pixel that a PRMT
// first unpacked to 32 bits, is bound by instruction issue (profiles/r02_sched16_4k_a.txt: issue slots 70 % busy, 16 %
// of the stalls are instruction-cache misses of its 990-instruction unrolled step) and cannot go faster than one MAC per
// lane per IMAD.  The micro-benchmarks (profiles/r02_ubench_pipes.jsonl) say: IDP.4A issues at the IMAD rate (62 lanes /
// clk / SM) and co-issues with PRMT better than IMAD does (3.26 vs 2.42 warp-instructions / clk / SM); FFMA next to IMAD
// buys nothing.  So here pixels STAY packed, four to a word, and a 22-bit Pillow coefficient is three byte limbs
// k = k0 + 2^8 k1 + 2^16 k2 (k0, k1 unsigned, k2 signed):
//     sum p*k = sum p*k0 + 2^8 sum p*k1 + 2^16 sum p*k2      (mod 2^32 = Pillow's int32 accumulator, exactly)
// = three IDP.4A per FOUR taps instead of four IMAD + four PRMT.  The host pads every record to whole words at the
// alignment of its window (vis_sched_pack_records_dp), so the device never shifts anything:
//   H  (12 warps) lane = input row.  16 pixels per step: 3 x LDS.128, 24 PRMT de-interleave them into 4 planar words per
//                 channel; a window of W words per channel lives in registers.  An output pixel whose window ends in word
//                 group g of the step is 9W IDP on win[g..g+W), whatever its alignment: FOUR emit bodies per step instead
//                 of sixteen (~300 instructions instead of ~990), and every tap class 9..33 in push order (the 16-slot
//                 kernel needs a pull-order role past 16 taps).
//   V  (NV warps) the H ring is TRANSPOSED — [channel][column][row], a column's rows are consecutive bytes — so a window is
//                 W aligned LDS.32 and 3W IDP per output byte; lanes walk consecutive columns (pitch = odd number of
//                 words: conflict free), no byte unpacking at all.
#include "vis_fused_common.cuh"

#include <cstring>
#include <vector>

using namespace visf;

namespace {

#ifndef VIS_DP_SWARPS
#define VIS_DP_SWARPS 2
#endif
#ifndef VIS_DP_HWARPS
#define VIS_DP_HWARPS 12
#endif
constexpr int kHWarps = VIS_DP_HWARPS, kSWarps = VIS_DP_SWARPS;     // A/B on 4K frames: 12 -> 70.4 k, 16 -> 66.2 k, 20 -> 65.4 k images/s (profiles/r02_dp_hwarps.txt)
static_assert(kHWarps <= VIS_SCHED_MAX_SUBS, "schedule holds at most VIS_SCHED_MAX_SUBS sub-ranges per strip");
// warp ranges in priority order (the scheduler prefers the highest ready warp id): H < loader < S < V
constexpr int kHBase = 0, kLBase = kHWarps, kSBase = kHWarps + 1, kVBase = kHWarps + 1 + kSWarps;
constexpr int threads_dp(int nv) { return (kHWarps + nv + kSWarps + 1) * 32; }
constexpr int kChunk = 32, kStepPx = 16;
constexpr int max_strip_w_dp(int nv) { return nv * 32 / 3 * 4; }        // one V thread per 4 columns of one channel
constexpr int kVRecs = kChunk + 1;                // vertical records a chunk can touch (scale >= 1): 32 emits + 1 look-ahead
constexpr int kSmemMax = 227 * 1024;

enum Bar { SF = 0, SE = 2, HF = 4, HE = 6, VF = 8, OF = 10, OE = 12, kBars = 14 };   // full/empty pairs, two slots each

__host__ __device__ constexpr int rec_stride_dp(int W) { return (3 * W + 3) & ~3; }         // words per record: 3 limbs x W words, 16-byte rows

struct LayoutDP {
    int stage_pitch, stage_slot, hrec_slot, vrec_slot;
    int cpitch, hplane;          // H ring: bytes per (channel, column) = carry rows + 32 fresh rows (an odd number of
                                 // words: lanes on consecutive columns never share a bank), bytes per channel plane
    int opitch, oplane;          // band tile: bytes per row (strip width), bytes per channel plane (14 rows)
    int off_stage, off_hring, off_otile, off_hrec, off_vrec, off_lut, off_bar, total;
};

inline LayoutDP make_layout_dp(int stage_pitch, int strip_w, int W) {
    const int stride = rec_stride_dp(W);
    LayoutDP L;
    L.stage_pitch = stage_pitch;
    L.stage_slot = kChunk * stage_pitch;
    L.hrec_slot = align_up((strip_w + 1) * stride * 4, 16);
    L.vrec_slot = align_up(kVRecs * stride * 4, 16);
    int cw = W - 1 + kChunk / 4;                   // carry words + fresh words
    if (cw % 2 == 0) ++cw;
    L.cpitch = 4 * cw;
    L.hplane = strip_w * L.cpitch;
    L.opitch = strip_w;
    L.oplane = VIS_PATCH * L.opitch;
    int off = 0;
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

__device__ __forceinline__ int dp4a_uu(uint32_t px, uint32_t k, int acc) {       // 4 x (u8 pixel * u8 limb)
    int d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(k), "r"(acc));
    return d;
}
__device__ __forceinline__ int dp4a_us(uint32_t px, uint32_t k, int acc) {       // 4 x (u8 pixel * s8 limb)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(px), "r"(k), "r"(acc));
    return d;
}

template <int W>
__device__ __forceinline__ void load_rec_dp(uint32_t (&k)[3 * W], uint32_t addr) {
#pragma unroll
    for (int q = 0; q < (3 * W + 3) / 4; ++q) {
        const uint4 a = lds128(addr + 16 * q);
        k[4 * q] = a.x;
        if (4 * q + 1 < 3 * W) k[4 * q + 1] = a.y;
        if (4 * q + 2 < 3 * W) k[4 * q + 2] = a.z;
        if (4 * q + 3 < 3 * W) k[4 * q + 3] = a.w;
    }
}

// one output byte: W packed words (4 taps each) against the record's three limb rows; Pillow's (acc + 2^21) >> 22, clip8
#ifndef VIS_DP_LAZY
#define VIS_DP_LAZY 0
#endif
#ifndef VIS_DP_PREFETCH
#define VIS_DP_PREFETCH 0      // A/B: ping-pong prefetch of the next emit's coefficients: 70.4 k -> 70.0 k images/s on 4K (profiles/r02_dp_prefetch_rejected.txt)
#endif
#ifndef VIS_DP_SERIAL_LIMBS
#define VIS_DP_SERIAL_LIMBS 0
#endif
__device__ __forceinline__ int shl8(int v) {        // funnel shift: stays on the ALU pipe (a plain << 8 becomes IMAD.SHL,
    int d;                                          // which competes with IDP.4A for the one integer-MAC pipe)
    asm("shf.l.clamp.b32 %0, %1, %2, 8;" : "=r"(d) : "r"(0), "r"(v));
    return d;
}
template <int W>
__device__ __forceinline__ int mac_dp(const uint32_t* px, const uint32_t (&k)[3 * W]) {
#if VIS_DP_SERIAL_LIMBS
    // Horner over the limbs: the accumulator of a limb starts from the finished higher limb << 8; the rounding constant
    // 2^21 enters as 32 in the 2^16 limb.  No recombination on the MAC pipe.
    int a = 1 << (VIS_PRECISION_BITS - 1 - 16);
#pragma unroll
    for (int q = 0; q < W; ++q) a = dp4a_us(px[q], k[2 * W + q], a);
    a = shl8(a);
#pragma unroll
    for (int q = 0; q < W; ++q) a = dp4a_uu(px[q], k[W + q], a);
    a = shl8(a);
#pragma unroll
    for (int q = 0; q < W; ++q) a = dp4a_uu(px[q], k[q], a);
    return clip8i(a);
#else
    int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = 0, a2 = 0;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        a0 = dp4a_uu(px[q], k[q], a0);
        a1 = dp4a_uu(px[q], k[W + q], a1);
        a2 = dp4a_us(px[q], k[2 * W + q], a2);
    }
    return clip8i(a0 + (a1 << 8) + (a2 << 16));
#endif
}

__device__ __noinline__ void band_done_dp(uint32_t bar0, int nb, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar0 + (uint32_t)(OF + (nb & 1)) * 8);
    const int nx = nb + 1;
    if (nx >= 2) mbar_wait(bar0 + (uint32_t)(OE + (nx & 1)) * 8, ((nx >> 1) - 1) & 1);
}

struct FramePtrs { const unsigned char* src; long long second; };      // VisFrameRef / VisResizeRef: same layout

template <int W, bool U8, int NV>
__global__ void __launch_bounds__(threads_dp(NV), 1)
k_fused_dp(const __grid_constant__ VisSched sc, const FramePtrs* __restrict__ frames, int n_items,
           const __grid_constant__ LayoutDP L, long long dst_pitch, const int* __restrict__ hrec_g,
           const int* __restrict__ vrec_g, const float* __restrict__ lut768, float* __restrict__ pixel_values) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int STRIDE = rec_stride_dp(W);
    constexpr int CARRY = 4 * (W - 1);                 // rows of the previous chunk a window may reach back to
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
        uint32_t win[3][W + 3];                            // per channel: W-1 words of history + the step's 4 planar words
#pragma unroll
        for (int q = 0; q < W + 3; ++q) win[0][q] = win[1][q] = win[2][q] = 0;
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
                // this lane's byte in column (U.xa - S.x0) of channel 0; a column is cpitch bytes, a channel hplane
                uint32_t hdst = smem_u32(smem + L.off_hring + slot * 3 * L.hplane + (U.xa - S.x0) * L.cpitch + CARRY + lane);
                uint32_t hp = hrec0;
#if VIS_DP_PREFETCH
                // the coefficients of the NEXT emit are requested while the current one multiplies: two register sets
                // that swap roles every emit (two copies of each emit body instead of 3W moves per emit)
                uint32_t ka[3 * W], kb[3 * W];
                load_rec_dp<W>(ka, hp);
                bool odd = false;
#endif
#pragma unroll 1
                for (int i = 0; i < U.nsteps; ++i) {
                    const uint32_t m = (uint32_t)um[4 * i] | ((uint32_t)um[4 * i + 1] << 8);
                    uint32_t raw[12];
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const uint4 d = lds128(sa + 16 * q);
                        raw[4 * q] = d.x; raw[4 * q + 1] = d.y; raw[4 * q + 2] = d.z; raw[4 * q + 3] = d.w;
                    }
                    sa += kStepPx * 3;
                    // de-interleave: pixels 4q..4q+3 of a channel are bytes c, c+3, c+6, c+9 of raw[3q..3q+2]
#if !VIS_DP_LAZY
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t r0 = raw[3 * q], r1 = raw[3 * q + 1], r2 = raw[3 * q + 2];
                        win[0][W - 1 + q] = __byte_perm(__byte_perm(r0, r1, 0x0630), r2, 0x5210);
                        win[1][W - 1 + q] = __byte_perm(__byte_perm(r0, r1, 0x0741), r2, 0x6210);
                        win[2][W - 1 + q] = __byte_perm(__byte_perm(r0, r1, 0x0052), r2, 0x7410);
                    }
#endif
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#if VIS_DP_LAZY
                        {   // word group g is de-interleaved right before its emits: permutes between the dot-product bursts
                            const uint32_t r0 = raw[3 * g], r1 = raw[3 * g + 1], r2 = raw[3 * g + 2];
                            win[0][W - 1 + g] = __byte_perm(__byte_perm(r0, r1, 0x0630), r2, 0x5210);
                            win[1][W - 1 + g] = __byte_perm(__byte_perm(r0, r1, 0x0741), r2, 0x6210);
                            win[2][W - 1 + g] = __byte_perm(__byte_perm(r0, r1, 0x0052), r2, 0x7410);
                        }
#endif
                        const int cnt = __popc((m >> (4 * g)) & 0xfu);   // windows ending in word group g (uniform)
                        auto emit = [&](const uint32_t (&kw)[3 * W]) {
                            const int v0 = mac_dp<W>(&win[0][g], kw);
                            const int v1 = mac_dp<W>(&win[1][g], kw);
                            const int v2 = mac_dp<W>(&win[2][g], kw);
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(hdst), "r"(v0) : "memory");
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(hdst + (uint32_t)L.hplane), "r"(v1) : "memory");
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(hdst + 2u * (uint32_t)L.hplane), "r"(v2) : "memory");
                            hdst += (uint32_t)L.cpitch;
                        };
#pragma unroll 1
                        for (int e = 0; e < cnt; ++e) {
#if VIS_DP_PREFETCH
                            hp += STRIDE * 4;                  // the slot holds sw + 1 records: always readable
                            if (!odd) { load_rec_dp<W>(kb, hp); emit(ka); } else { load_rec_dp<W>(ka, hp); emit(kb); }
                            odd = !odd;
#else
                            uint32_t kw[3 * W];
                            load_rec_dp<W>(kw, hp);
                            hp += STRIDE * 4;
                            emit(kw);
#endif
                        }
                    }
#pragma unroll
                    for (int q = 0; q < W - 1; ++q) {                    // the last W-1 words become the history
                        win[0][q] = win[0][q + 4]; win[1][q] = win[1][q + 4]; win[2][q] = win[2][q + 4];
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
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const int n_chunks = (G.r_end - G.r_first + kChunk - 1) / kChunk;
            const int wpr = (S.x1 - S.x0) / 4;                 // threads per channel; a thread owns columns t, t+wpr, t+2wpr, t+3wpr
            const bool v_active = v < 3 * wpr;
            const int vc = v_active ? v / wpr : 0;
            const int vt = v_active ? v - vc * wpr : 0;
            const uint32_t col_step = (uint32_t)(wpr * L.cpitch);
            const uint32_t thr_off = (uint32_t)(vc * L.oplane + vt);
            int py = 0;
            uint32_t otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * L.oplane) + thr_off;
            const uint8_t* const gm = sc.mask + G.mask_off;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(VF, slot), j & 1);
                mbar_wait(bar(HF, slot), j & 1);
                uint32_t vaddr = smem_u32(smem + L.off_vrec + slot * L.vrec_slot);
                // this thread's first column in the slot (byte 0 = oldest carry row)
                const uint32_t hcol = smem_u32(smem + L.off_hring + slot * 3 * L.hplane + vc * L.hplane + vt * L.cpitch);
#pragma unroll 1
                for (int g = 0; g < kChunk / 16; ++g) {
                    uint32_t m = (uint32_t)gm[4 * (c * (kChunk / 16) + g)] | ((uint32_t)gm[4 * (c * (kChunk / 16) + g) + 1] << 8);
#pragma unroll 1
                    while (m) {
                        const int u = __ffs(m) - 1;
                        m &= m - 1;
                        // the window ends at row (g*16 + u) of the chunk = byte CARRY + g*16 + u of the column: its W words
                        const uint32_t wbase = hcol + (uint32_t)((((CARRY + g * 16 + u) >> 2) - (W - 1)) * 4);
                        uint32_t kw[3 * W];
                        load_rec_dp<W>(kw, vaddr);
                        vaddr += STRIDE * 4;
                        int out[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            uint32_t px[W];
#pragma unroll
                            for (int q = 0; q < W; ++q) px[q] = lds32(wbase + (uint32_t)e * col_step + 4u * q);
                            out[e] = mac_dp<W>(px, kw);
                        }
                        if (v_active) {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                asm volatile("st.shared.u8 [%0], %1;" ::"r"(otile_thr + (uint32_t)(e * wpr)), "r"(out[e]) : "memory");
                        }
                        otile_thr += L.opitch;
                        if (++py == VIS_PATCH) {                  // band complete: hand it to the store warps
                            band_done_dp(bar0, nb, lane);
                            ++nb;
                            py = 0;
                            otile_thr = smem_u32(smem + L.off_otile + (nb & 1) * 3 * L.oplane) + thr_off;
                        }
                    }
                }
                if (c + 1 < n_chunks && v_active) {               // carry: the last W-1 words of every column -> front of the other slot
                    const uint32_t dst = smem_u32(smem + L.off_hring + (slot ^ 1) * 3 * L.hplane + vc * L.hplane + vt * L.cpitch);
#pragma unroll
                    for (int e = 0; e < 4; ++e)
#pragma unroll
                        for (int q = 0; q < W - 1; ++q)
                            sts32(dst + (uint32_t)e * col_step + 4u * q, lds32(hcol + (uint32_t)e * col_step + (uint32_t)(kChunk + 4 * q)));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(HE, slot));       // H-ring slot and record slot consumed
            }
            if (py) {                                             // a segment that ends inside a band (uint8 mode only)
                band_done_dp(bar0, nb, lane);
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

template <int W, bool U8, int NV>
int launch_dp(const VisSched& sc, const void* frames, int n_frames, const LayoutDP& L, int64_t dst_pitch, const int* hrec,
              const int* vrec, const float* lut768, float* pixel_values, cudaStream_t st) {
    auto kern = k_fused_dp<W, U8, NV>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_fused_dp: cudaFuncSetAttribute");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_items = n_frames * sc.n_strips * sc.n_segs;
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, threads_dp(NV), L.total, st>>>(sc, reinterpret_cast<const FramePtrs*>(frames), n_items, L, (long long)dst_pitch,
                                                hrec, vrec, lut768, pixel_values);
    return vis::check_launch("vis_fused_dp");
}

}  // namespace

namespace visf {

int dp_subs() { return kHWarps; }
int dp_max_strip_w(int nv) { return max_strip_w_dp(nv); }
int dp_layout_bytes(int stage_pitch, int strip_w, int words) { return make_layout_dp(stage_pitch, strip_w, words).total; }
int dp_record_stride(int words) { return rec_stride_dp(words); }

int dp_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
              const float* lut768, float* pixel_values, cudaStream_t st) {
    const int W = sc.dp_words;
    if (W < 4 || W > 9 || sc.ring != 16 || sc.n_subs != kHWarps || sc.per_index != 1 || (sc.n_vwarps != 4 && sc.n_vwarps != 6)) {
        vis::set_error("vis_fused_dp: schedule of another kernel class (ring %d, %d words, %d sub-ranges, %d V warps)", sc.ring, W,
                       sc.n_subs, sc.n_vwarps);
        return VIS_E_INVALID;
    }
    const LayoutDP L = make_layout_dp(sc.stage_pitch, sc.max_strip_w, W);
    if (L.total > kSmemMax) {
        vis::set_error("vis_fused_dp: %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
    const bool u8 = sc.out_mode == VIS_SCHED_OUT_U8;
#define VIS_LDP(WW) (u8 ? (sc.n_vwarps == 4 ? launch_dp<WW, true, 4>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st)  \
                                            : launch_dp<WW, true, 6>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st)) \
                        : (sc.n_vwarps == 4 ? launch_dp<WW, false, 4>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st) \
                                            : launch_dp<WW, false, 6>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st)))
    switch (W) {
        case 4: return VIS_LDP(4);
        case 5: return VIS_LDP(5);
        case 6: return VIS_LDP(6);
        case 7: return VIS_LDP(7);
        case 8: return VIS_LDP(8);
        default: return VIS_LDP(9);
    }
#undef VIS_LDP
}

}  // namespace visf
