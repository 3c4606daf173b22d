// vis_fused_mma.cu — statically scheduled, warp-specialised kernel for 9..33-tap windows: both Pillow passes as banded
// u8 x 8-bit-limb matrix products on the integer tensor path (mma.sync.m16n8k32 .s32.u8.{u8,s8}, SASS IMMA.16832).
//
// Same path and arithmetic as vis_fused_dp.cu (Pillow 8bpc horizontal pass -> uint8 -> vertical pass -> uint8, then either
// LUT + Qwen2-VL patch layout or RGB uint8 rows), same loader / store roles, mbarrier rings, VisSched and coefficient
// limbs k = k0 + 2^8 k1 + 2^16 k2 (k0, k1 unsigned, k2 signed), so that
//     sum p*k = sum p*k0 + 2^8 sum p*k1 + 2^16 sum p*k2      (mod 2^32 = Pillow's int32 accumulator, exactly).
// Why it exists (tools/ubench_imma.cu -> profiles/r02_ubench_imma.jsonl): strong downscales are not HBM bound on CUDA
// cores — IDP.4A issues at 62 lanes / clk / SM = 248 byte MACs, the packed-byte kernel sits at 0.52 of that and at 0.2-0.5
// of the HBM roofline — while IMMA.16832 issues every 2 clocks per SM = 2037 byte MACs / clk / SM, 8.2x, with LDS free next
// to it.  A resampling pass IS a (banded) matrix product, out = C x in; per 16 outputs the band is 2-3 k-steps of 32 wide.
//   H  (12 / 9 warps)  warp = a tile of 16 output columns x the 32 rows of the chunk, per channel
//                 D[16 outputs][8 rows] = A[16 outputs][32 k] x B[32 k][8 rows].  A = coefficient limbs, gathered by predicated
//                 LDS.32 from compact per-output records (W words x 3 limbs, limb-minor, + the word index of the record and
//                 of its first tap; record stride = 4 mod 8 words: the eight records of a gather start in eight bank groups),
//                 once per item when the warp owns one tile of the strip.  B = pixels: 3 x LDS.32 + 6 PRMT give the words of
//                 all three channels (de-interleave in registers); N-tile c holds chunk rows 4g + c, and the stage rows are
//                 skewed by 16 bytes per 4 rows so that the eight rows of a fragment load never share a bank.  A 16-output
//                 window is hardly wider than an 8-output one, so with the outputs on the M side the pixels are fetched half
//                 as often as with the rows there.  The four N-tiles leave a thread with rows 8t..8t+3 / 8t+4..8t+7 of columns
//                 g and g+8: a 4x4 byte transpose and four STS.32 per channel into the TRANSPOSED H ring
//                 [channel * sw + column][row].
//   V  (9 / 8 warps)  warp = every n-th tile of 8 ring columns, D[16 output rows][8 columns].  A = coefficient limbs of the
//                 rows the chunk emits (gathered per chunk from the vertical records relative to ring byte 0), B = ring
//                 words (4 consecutive rows of a column: exactly the fragment layout, LDS.32, column pitch = 4 mod 8 words:
//                 conflict free), 6 IMMA per tile, STS.U16 into a three-slot band tile [row][channel * sw + x].  One M-tile
//                 per chunk is the efficient case: a chunk advances by 32, 28 or 24 input rows (VisSched.chunk_rows).
//   epilogue      acc0 (2^21 as the C operand of its first IMMA) + (acc1 << 8) + (acc2 << 16), >> 22, I2IP.U8.S32.SAT packs
//                 and clips two samples per instruction.
// DESIGN.md 4.1d has the measurements, the build history and what bounds the kernel now.
#include "vis_fused_common.cuh"

#include <cstring>
#include <vector>

using namespace visf;

namespace {

// Warp split of the roles, per k-step class.  Up to 24 warps keep 80 registers per thread, up to 20 keep 96: the two
// k-step kernels (12-13 taps at 4K, the 2048 / 1024 thumbnails of 16:9 frames) run 12 H + 9 V + 2 store + loader; the three
// k-step kernels hold 36 registers of coefficient fragments and are faster unspilled with 9 + 8 + 2 + 1
// (profiles/r02_mma_ab_log.txt, steps 11-13).
#ifndef VIS_MMA_HWARPS
#define VIS_MMA_HWARPS 12
#endif
#ifndef VIS_MMA_VWARPS
#define VIS_MMA_VWARPS 9
#endif
#ifndef VIS_MMA_HWARPS3
#define VIS_MMA_HWARPS3 9
#endif
#ifndef VIS_MMA_VWARPS3
#define VIS_MMA_VWARPS3 8
#endif
#ifndef VIS_MMA_SWARPS
#define VIS_MMA_SWARPS 2
#endif
constexpr int kSWarps = VIS_MMA_SWARPS;
__host__ __device__ constexpr int h_warps(int ks) { return ks <= 2 ? VIS_MMA_HWARPS : VIS_MMA_HWARPS3; }
__host__ __device__ constexpr int v_warps(int ks) { return ks <= 2 ? VIS_MMA_VWARPS : VIS_MMA_VWARPS3; }
__host__ __device__ constexpr int n_threads(int ks) { return (h_warps(ks) + v_warps(ks) + kSWarps + 1) * 32; }
#ifndef VIS_MMA_HOIST3
#define VIS_MMA_HOIST3 1           // three k-step kernels too: strips of one tile per H warp, coefficient fragments gathered once per item (0: 256-column strips; 80.3 k vs 84.0 k images/s on the 4K -> 1024 thumbnail)
#endif
#ifndef VIS_MMA_VPIPE
#define VIS_MMA_VPIPE 0            // A/B: 1 = software-pipelined vertical tile loop (tile i's IMMA between the pieces of tile
                                   // i-1's epilogue): 85.8 k -> 77.1 k images/s on 4K, rejected (profiles/r02_mma_ab_log.txt, step 15)
#endif
#ifndef VIS_MMA_VUNROLL
#define VIS_MMA_VUNROLL 1
#endif
constexpr int kVUnroll = VIS_MMA_VUNROLL;         // independent column tiles a V warp keeps in flight
constexpr int kChunk = 32;
constexpr int kVRecs = kChunk + 1;                // vertical records a chunk can touch (scale >= 1): 32 emits + 1 look-ahead
constexpr int kSmemMax = 227 * 1024;
constexpr int kStageSkew = 128;                   // 16 bytes x (row >> 2): 112 bytes of slack per slot

#ifndef VIS_MMA_OSLOTS
#define VIS_MMA_OSLOTS 3
#endif
constexpr int kOSlots = VIS_MMA_OSLOTS;           // band tiles in flight between the vertical pass and the store warps
// full/empty pairs, two slots each; band tiles: kOSlots each
enum Bar { SF = 0, SE = 2, HF = 4, HE = 6, VF = 8, OF = 10, OE = 10 + kOSlots, kBars = 10 + 2 * kOSlots };

// words per record: W words x 3 limbs (limb-minor), then the absolute word index of record byte 0 and of the window's first tap
// (4 mod 8 words: the eight records a fragment gather touches start in eight different bank groups)
__host__ __device__ constexpr int rec_stride_mma(int W) { return ((3 * W + 2 + 3) & ~7) + 4 >= 3 * W + 2 ? ((3 * W + 2 + 3) & ~7) + 4 : ((3 * W + 2 + 3) & ~7) + 12; }

struct LayoutM {
    int stage_pitch, stage_slot, hrec_slot, vrec_slot;
    int cpitch, hplane;          // H ring: bytes per (channel, column) = carry rows + 32 fresh rows, padded to 4 mod 8 words
    int opitch, oplane;          // band tile: bytes per row (strip width), bytes per channel plane (14 rows)
    int off_stage, off_hring, off_otile, off_hrec, off_vrec, off_lut, off_bar, total;
};

inline LayoutM make_layout_m(int stage_pitch, int strip_w, int W) {
    const int stride = rec_stride_mma(W);
    LayoutM L;
    L.stage_pitch = stage_pitch;
    L.stage_slot = kChunk * stage_pitch + kStageSkew;
    L.hrec_slot = align_up((strip_w + 1) * stride * 4, 16);
    L.vrec_slot = align_up(kVRecs * stride * 4, 16);
    int cw = W - 1 + kChunk / 4;                   // carry words + fresh words
    while (cw % 8 != 4) ++cw;
    L.cpitch = 4 * cw;
    L.hplane = strip_w * L.cpitch;
    L.opitch = strip_w;
    L.oplane = VIS_PATCH * L.opitch;
    int off = 0;
    L.off_stage = off; off += 2 * L.stage_slot;
    L.off_hring = off; off += 2 * 3 * L.hplane;           // tiles past the strip read (and discard) whatever follows: the ring is never the last region
    L.off_otile = off; off += kOSlots * 3 * L.oplane;
    off = align_up(off, 16);
    L.off_hrec = off;  off += 2 * L.hrec_slot;
    L.off_vrec = off;  off += 2 * L.vrec_slot;
    L.off_lut = off;   off += 768 * 4;
    L.off_bar = off;   off += kBars * 8;
    L.total = off;
    return L;
}

__device__ __forceinline__ void imma_uu(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void imma_us(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// coefficient limbs as the A operand (vertical pass): unsigned / signed limb rows x u8 ring bytes
__device__ __forceinline__ void imma_uu_a(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) { imma_uu(d, a, b0, b1); }
__device__ __forceinline__ void imma_su_a(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// first k-step of a tile: D = A x B + C with C a loop-invariant register quad (the rounding constant) or zero, so no
// accumulator is initialised by moves
#define VIS_IMMA_C(NAME, TA, TB)                                                                                              \
    __device__ __forceinline__ void NAME(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, const int (&c)[4]) {  \
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32." TA "." TB ".s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "          \
                     "{%10,%11,%12,%13};"                                                                                     \
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3])                                                         \
                     : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])); \
    }
VIS_IMMA_C(imma_uu_c, "u8", "u8")
VIS_IMMA_C(imma_su_c, "s8", "u8")
#undef VIS_IMMA_C
// (c << 16) | (sat_u8(hi) << 8) | sat_u8(lo)      (I2IP.U8.S32.SAT)
__device__ __forceinline__ uint32_t pack_sat(int hi, int lo, uint32_t c) {
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(hi), "r"(lo), "r"(c));
    return d;
}
// Pillow's (acc + 2^21) >> 22 with the rounding constant already inside limb 0's accumulator
__device__ __forceinline__ int recombine(int a0, int a1, int a2) { return (a0 + (a1 << 8) + (a2 << 16)) >> VIS_PRECISION_BITS; }

__device__ __forceinline__ void sts16_if(bool p, uint32_t addr, uint32_t v) {     // predicated, no branch
    asm volatile("{\n.reg .pred q;\nsetp.ne.u32 q, %0, 0;\n@q st.shared.u16 [%1], %2;\n}\n" ::"r"((uint32_t)p), "r"(addr), "h"((unsigned short)v) : "memory");
}

struct FramePtrs { const unsigned char* src; long long second; };      // VisFrameRef / VisResizeRef: same layout

template <int KS, bool U8>
__global__ void __launch_bounds__(n_threads(KS), 1)
k_fused_mma(const __grid_constant__ VisSched sc, const FramePtrs* __restrict__ frames, int n_items,
            const __grid_constant__ LayoutM L, long long dst_pitch, const int* __restrict__ hrec_g,
            const int* __restrict__ vrec_g, const float* __restrict__ lut768, float* __restrict__ pixel_values) {
    extern __shared__ __align__(128) unsigned char smem[];
    // warp ranges in priority order (the scheduler prefers the highest ready warp id): H < loader < S < V
    constexpr int kHWarps = h_warps(KS), kVWarps = v_warps(KS);
    constexpr int kHBase = 0, kLBase = kHWarps, kSBase = kHWarps + 1, kVBase = kSBase + kSWarps;
    const int W = sc.dp_words;
    const int STRIDE = rec_stride_mma(W);
    const int CARRY = 4 * (W - 1);                     // rows of the previous chunk a window may reach back to
    const int CR = sc.chunk_rows;                      // input rows a chunk advances by (<= kChunk: the stage / ring slots hold kChunk)
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;             // fragment coordinates of mma.m16n8k32
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);               // warp-uniform for the compiler
    float* lut = reinterpret_cast<float*>(smem + L.off_lut);              // transposed: lut[c * 256 + v]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    auto bar = [&](int which, int slot) { return bar0 + (uint32_t)(which + slot) * 8; };
    const int per_frame = sc.n_strips * sc.n_segs;
    const int kRound[4] = {1 << (VIS_PRECISION_BITS - 1), 1 << (VIS_PRECISION_BITS - 1), 1 << (VIS_PRECISION_BITS - 1),
                           1 << (VIS_PRECISION_BITS - 1)};
    const int kZero[4] = {0, 0, 0, 0};

    if (!U8)
        for (int i = tid; i < 768; i += (int)blockDim.x) lut[(i % 3) * 256 + i / 3] = __ldg(lut768 + i);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(SF, s), 1);
            mbar_init(bar(SE, s), kHWarps);
            mbar_init(bar(HF, s), kHWarps);
            mbar_init(bar(HE, s), kVWarps);
            mbar_init(bar(VF, s), 1);
        }
        for (int s = 0; s < kOSlots; ++s) {
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
            const int n_chunks = (G.r_end - G.r_first + CR - 1) / CR;
            int yo = G.y0;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1;
                const uint32_t prev = ((k >> 1) - 1) & 1;
                if (k >= 2) mbar_wait(bar(SE, slot), prev);                 // H is done with the stage slot
                const int r0 = G.r_first + c * CR;
                const int rows = max(0, min(CR, sc.src_h - r0));        // r_end may include virtual rows past the image
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(bar(SF, slot), (uint32_t)rows * (uint32_t)S.row_bytes + (c == 0 ? rec_bytes : 0u));
                }
                __syncwarp();
                unsigned char* stage = smem + L.off_stage + slot * L.stage_slot;
                if (lane < rows)                                            // row r sits 16 * (r >> 2) bytes to the right
                    bulk_g2s(smem_u32(stage + lane * L.stage_pitch + 16 * (lane >> 2)), src + (size_t)(r0 + lane) * sc.src_pitch,
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
        // D[16 output columns][8 rows] = A[16 outputs][32 k] (coefficient limbs) x B[32 k][8 rows] (pixels of one channel).
        // A 16-output window is hardly wider than an 8-output one (60 vs 36 of the 64 pixels two k-steps hold at 4K), so
        // the pixels are fetched half as often as with the rows on the M side, and they are de-interleaved in registers
        // (3 LDS.32 + 6 PRMT give the B words of all three channels): no separate pass over the staged chunk.
        // N-tile c holds rows 4g + c (n = g): with the 16-byte skew per 4 rows the eight rows of a fragment load sit in
        // eight different bank groups; the four N-tiles of a chunk give a thread rows 8t..8t+3 and 8t+4..8t+7 of columns
        // g and g + 8: four STS.32 per channel into the TRANSPOSED H ring [channel * sw + column][row].
        const int hw = warp - kHBase;
        int k = 0, sl = 0;
        for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++sl) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const int sw = S.x1 - S.x0;
            const int n_tiles = (sw + 15) >> 4;
            const int n_chunks = (G.r_end - G.r_first + CR - 1) / CR;
            const uint32_t hrec0 = smem_u32(smem + L.off_hrec + (sl & 1) * L.hrec_slot);
            const bool one_tile = (KS <= 2 || VIS_MMA_HOIST3) && n_tiles <= kHWarps;   // (three k-steps: 36 registers of limbs to keep)
            uint32_t a[KS][3][4];
            int kw = 0;
#pragma unroll
            for (int s = 0; s < KS; ++s)
#pragma unroll
                for (int l = 0; l < 3; ++l) a[s][l][0] = a[s][l][1] = a[s][l][2] = a[s][l][3] = 0u;
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(SF, slot), j & 1);
                if (k >= 2) mbar_wait(bar(HE, slot), (j - 1) & 1);
                // row 4g of the stage slot (skewed by 16 bytes per 4 rows); rows 8t.. of the ring columns this lane stores
                const uint32_t srow = smem_u32(smem + L.off_stage + slot * L.stage_slot) + (uint32_t)(4 * g * L.stage_pitch + 16 * g);
                const uint32_t hring = smem_u32(smem + L.off_hring + slot * 3 * L.hplane) + (uint32_t)(CARRY + 8 * t);
#pragma unroll 1
                for (int jt = hw; jt < n_tiles; jt += kHWarps) {
                    // ---- A: coefficient limbs of outputs 16 jt + g and 16 jt + g + 8, 32 input pixels per k-step ----
                    // (they do not depend on the chunk: a warp that owns ONE tile of the strip gathers them once per item)
                    if (!one_tile || c == 0) {
                        const int xa = min(16 * jt + g, sw - 1), xb = min(16 * jt + g + 8, sw - 1);
                        const uint32_t reca = hrec0 + (uint32_t)(xa * STRIDE * 4), recb = hrec0 + (uint32_t)(xb * STRIDE * 4);
                        const int bwa = (int)lds32(reca + (uint32_t)(3 * W * 4)), bwb = (int)lds32(recb + (uint32_t)(3 * W * 4));
                        kw = __shfl_sync(0xffffffffu, (int)lds32(reca + (uint32_t)((3 * W + 1) * 4)), 0);     // first word of the tile's window
#pragma unroll
                        for (int s = 0; s < KS; ++s)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int qa = kw + 8 * s + 4 * h + t - bwa, qb = kw + 8 * s + 4 * h + t - bwb;
                                const bool oka = (unsigned)qa < (unsigned)W, okb = (unsigned)qb < (unsigned)W;
                                const uint32_t pa = reca + (uint32_t)(12 * qa), pb = recb + (uint32_t)(12 * qb);   // limbs side by side
#pragma unroll
                                for (int l = 0; l < 3; ++l) {
                                    a[s][l][2 * h] = oka ? lds32(pa + 4 * l) : 0u;
                                    a[s][l][2 * h + 1] = okb ? lds32(pb + 4 * l) : 0u;
                                }
                            }
                    }
                    // pixel group (kw + t) of a stage row: 12 bytes = 4 RGB pixels
                    const uint32_t g_off = (uint32_t)(12 * (kw + t - (S.px0 >> 2)));
                    uint32_t P[3][4];                                   // per channel and N-tile: the four samples of D, saturated
#pragma unroll
                    for (int cn = 0; cn < 4; ++cn) {                    // N-tile cn: rows 4g + cn
                        const uint32_t ra = srow + (uint32_t)(cn * L.stage_pitch) + g_off;
                        uint32_t b[3][KS][2];
#pragma unroll
                        for (int s = 0; s < KS; ++s)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const uint32_t o = (uint32_t)(96 * s + 48 * h);
                                const uint32_t x0 = lds32(ra + o), x1 = lds32(ra + o + 4), x2 = lds32(ra + o + 8);
                                b[0][s][h] = __byte_perm(__byte_perm(x0, x1, 0x0630), x2, 0x5210);
                                b[1][s][h] = __byte_perm(__byte_perm(x0, x1, 0x0741), x2, 0x6210);
                                b[2][s][h] = __byte_perm(__byte_perm(x0, x1, 0x0052), x2, 0x7410);
                            }
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            int acc[3][4];
                            imma_uu_c(acc[0], a[0][0], b[ch][0][0], b[ch][0][1], kRound);
                            imma_uu_c(acc[1], a[0][1], b[ch][0][0], b[ch][0][1], kZero);
                            imma_su_c(acc[2], a[0][2], b[ch][0][0], b[ch][0][1], kZero);
#pragma unroll
                            for (int s = 1; s < KS; ++s) {
                                imma_uu_a(acc[0], a[s][0], b[ch][s][0], b[ch][s][1]);
                                imma_uu_a(acc[1], a[s][1], b[ch][s][0], b[ch][s][1]);
                                imma_su_a(acc[2], a[s][2], b[ch][s][0], b[ch][s][1]);
                            }
                            // D: e = 0/1 -> (column g, rows 8t+cn / 8t+4+cn), e = 2/3 -> (column g+8, same rows)
                            int v[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) v[e] = recombine(acc[0][e], acc[1][e], acc[2][e]);
                            P[ch][cn] = pack_sat(v[1], v[0], pack_sat(v[3], v[2], 0u));              // byte e = sample e
                        }
                    }
                    // 4 x 4 byte transpose per channel: word e = samples e of N-tiles 0..3 = four consecutive rows
                    const bool oka = 16 * jt + g < sw, okb = 16 * jt + g + 8 < sw;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) {
                        const uint32_t t0 = __byte_perm(P[ch][0], P[ch][1], 0x5140), t1 = __byte_perm(P[ch][2], P[ch][3], 0x5140);
                        const uint32_t t2 = __byte_perm(P[ch][0], P[ch][1], 0x7362), t3 = __byte_perm(P[ch][2], P[ch][3], 0x7362);
                        const uint32_t at = hring + (uint32_t)((ch * sw + 16 * jt + g) * L.cpitch);       // ring column = channel * sw + x
                        if (oka) {
                            sts32(at, __byte_perm(t0, t1, 0x5410));
                            sts32(at + 4, __byte_perm(t0, t1, 0x7632));
                        }
                        if (okb) {
                            sts32(at + (uint32_t)(8 * L.cpitch), __byte_perm(t2, t3, 0x5410));
                            sts32(at + (uint32_t)(8 * L.cpitch) + 4, __byte_perm(t2, t3, 0x7632));
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
        // ============================== vertical pass ==============================
        const int wv = warp - kVBase;
        int k = 0, nb = 0, ready = kOSlots - 1;            // the first kOSlots bands find their slots empty
        for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int f = w / per_frame, r = w - f * per_frame;
            const int sg = r / sc.n_strips, st = r - sg * sc.n_strips;
            const VisSchedStrip S = sc.strip[st];
            const VisSchedSeg G = sc.seg[sg];
            const int n_chunks = (G.r_end - G.r_first + CR - 1) / CR;
            const int sw = S.x1 - S.x0, NC = 3 * sw;
            const int n_tiles = (NC + 7) >> 3;
            const int seg_rows = G.y1 - G.y0;
            const uint8_t* const gm = sc.mask + G.mask_off;
            const uint32_t otile0 = smem_u32(smem + L.off_otile);
            int ycount = 0;                                // output rows of the segment emitted so far
            for (int c = 0; c < n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(VF, slot), j & 1);
                mbar_wait(bar(HF, slot), j & 1);
                const uint32_t vrec0 = smem_u32(smem + L.off_vrec + slot * L.vrec_slot);
                const uint32_t ring0 = smem_u32(smem + L.off_hring + slot * 3 * L.hplane);
                const uint32_t* m8 = reinterpret_cast<const uint32_t*>(gm + c * 8);
                const int n = __popc(m8[0]) + __popc(m8[1]);           // output rows whose window ends in this chunk
                const int cbw = (G.r_first + c * CR - CARRY) >> 2;  // absolute word index of ring byte 0
#pragma unroll 1
                for (int mt = 0; 16 * mt < n; ++mt) {
                    const int ihi = min(n, 16 * mt + 16);
                    // ---- A: coefficient limbs [16 output rows][64 ring rows] of this M-tile ----
                    uint32_t a[2][3][4];
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {                     // rows 16 mt + g and 16 mt + g + 8
                        const int i = 16 * mt + g + 8 * hh;
                        const uint32_t rec = vrec0 + (uint32_t)(min(i, kVRecs - 1) * STRIDE * 4);
                        const int rel = (int)lds32(rec + (uint32_t)(3 * W * 4)) - cbw;
#pragma unroll
                        for (int s = 0; s < 2; ++s)
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const int q = 8 * s + 4 * h + t - rel;
                                const bool ok = i < ihi && (unsigned)q < (unsigned)W;
                                const uint32_t qa = rec + (uint32_t)(12 * q);
#pragma unroll
                                for (int l = 0; l < 3; ++l) a[s][l][2 * h + hh] = ok ? lds32(qa + 4 * l) : 0u;
                            }
                    }
                    const int bfirst = (ycount + 16 * mt) / VIS_PATCH, blast = (ycount + ihi - 1) / VIS_PATCH;
#pragma unroll 1
                    for (int bs = bfirst; bs <= blast; bs += 2) {        // at most two bands per pass: never more slots than there are
                        const int be = min(bs + 1, blast);
                        while (ready < nb + be) {
                            ++ready;
                            mbar_wait(bar(OE, ready % kOSlots), (ready / kOSlots - 1) & 1);
                        }
                        // where this thread's two rows go (byte offset into the band tile of column 0, channel 0)
                        bool ok_r[2];
                        uint32_t off_r[2];
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int i = 16 * mt + g + 8 * hh, yr = ycount + i, bl = yr / VIS_PATCH;
                            ok_r[hh] = i < ihi && bl >= bs && bl <= be;
                            off_r[hh] = (uint32_t)(((nb + bl) % kOSlots) * 3 * L.oplane + (yr - bl * VIS_PATCH) * NC + 2 * t);
                        }
                        // ring column 8 jt + g: linear in the tile index (columns past the strip read slack, results unused)
                        uint32_t ba = ring0 + (uint32_t)((8 * wv + g) * L.cpitch + 4 * t);
                        uint32_t oa0 = otile0 + (uint32_t)(8 * wv) + off_r[0], oa1 = otile0 + (uint32_t)(8 * wv) + off_r[1];
                        const uint32_t ba_step = (uint32_t)(8 * kVWarps * L.cpitch);
                        const int lim = (NC - 2 * t + 7) >> 3;                            // tiles whose columns 2t, 2t+1 lie inside the strip
                        const int lim0 = ok_r[0] ? lim : 0, lim1 = ok_r[1] ? lim : 0;
#if VIS_MMA_VPIPE
                        // software pipeline: the six IMMA of tile i (one per 8 clocks and sub-core at best) are issued
                        // between the pieces of tile i-1's epilogue, and tile i+1's ring words are requested behind them
                        {
                            int accp[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
                            int jtp = 0x40000000;                                         // "no previous tile": stores nothing
                            uint32_t b0 = lds32(ba), b1 = lds32(ba + 16), b2 = lds32(ba + 32), b3 = lds32(ba + 48);
#pragma unroll 1
                            for (int jt = wv; jt < n_tiles; jt += kVWarps) {
                                int acc[3][4];
                                imma_uu_c(acc[0], a[0][0], b0, b1, kRound);
                                const int v0 = recombine(accp[0][0], accp[1][0], accp[2][0]);
                                const int v1 = recombine(accp[0][1], accp[1][1], accp[2][1]);
                                imma_uu_c(acc[1], a[0][1], b0, b1, kZero);
                                const uint32_t p0 = pack_sat(v1, v0, 0u);
                                imma_su_c(acc[2], a[0][2], b0, b1, kZero);
                                const int v2 = recombine(accp[0][2], accp[1][2], accp[2][2]);
                                const int v3 = recombine(accp[0][3], accp[1][3], accp[2][3]);
                                imma_uu_a(acc[0], a[1][0], b2, b3);
                                const uint32_t p1 = pack_sat(v3, v2, 0u);
                                imma_uu_a(acc[1], a[1][1], b2, b3);
                                if (jtp < lim0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa0 - 8 * kVWarps), "h"((unsigned short)p0) : "memory");
                                if (jtp < lim1) asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa1 - 8 * kVWarps), "h"((unsigned short)p1) : "memory");
                                imma_su_a(acc[2], a[1][2], b2, b3);
                                ba += ba_step;
                                b0 = lds32(ba); b1 = lds32(ba + 16); b2 = lds32(ba + 32); b3 = lds32(ba + 48);   // next tile (or slack)
                                oa0 += 8 * kVWarps; oa1 += 8 * kVWarps;
                                jtp = jt;
#pragma unroll
                                for (int l = 0; l < 3; ++l)
#pragma unroll
                                    for (int e = 0; e < 4; ++e) accp[l][e] = acc[l][e];
                            }
                            // epilogue of the last tile
                            const uint32_t p0 = pack_sat(recombine(accp[0][1], accp[1][1], accp[2][1]), recombine(accp[0][0], accp[1][0], accp[2][0]), 0u);
                            const uint32_t p1 = pack_sat(recombine(accp[0][3], accp[1][3], accp[2][3]), recombine(accp[0][2], accp[1][2], accp[2][2]), 0u);
                            if (jtp < lim0) asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa0 - 8 * kVWarps), "h"((unsigned short)p0) : "memory");
                            if (jtp < lim1) asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa1 - 8 * kVWarps), "h"((unsigned short)p1) : "memory");
                        }
#else
#pragma unroll 1
                        for (int jt = wv; jt < n_tiles; jt += kVUnroll * kVWarps, ba += kVUnroll * ba_step,
                                 oa0 += kVUnroll * 8 * kVWarps, oa1 += kVUnroll * 8 * kVWarps) {
                            // kVUnroll independent tiles in flight (tiles past the strip read slack and store nothing)
                            uint32_t bq[kVUnroll][4];
                            int acc[kVUnroll][3][4];
#pragma unroll
                            for (int u = 0; u < kVUnroll; ++u) {
                                const uint32_t bu = ba + u * ba_step;
                                bq[u][0] = lds32(bu); bq[u][1] = lds32(bu + 16); bq[u][2] = lds32(bu + 32); bq[u][3] = lds32(bu + 48);
                            }
#pragma unroll
                            for (int u = 0; u < kVUnroll; ++u) {
                                imma_uu_c(acc[u][0], a[0][0], bq[u][0], bq[u][1], kRound);
                                imma_uu_c(acc[u][1], a[0][1], bq[u][0], bq[u][1], kZero);
                                imma_su_c(acc[u][2], a[0][2], bq[u][0], bq[u][1], kZero);
                            }
#pragma unroll
                            for (int u = 0; u < kVUnroll; ++u) {
                                imma_uu_a(acc[u][0], a[1][0], bq[u][2], bq[u][3]);
                                imma_uu_a(acc[u][1], a[1][1], bq[u][2], bq[u][3]);
                                imma_su_a(acc[u][2], a[1][2], bq[u][2], bq[u][3]);
                            }
                            // D: e = 0/1 -> (row g, columns 2t / 2t+1), e = 2/3 -> (row g+8, columns 2t / 2t+1)
#pragma unroll
                            for (int u = 0; u < kVUnroll; ++u) {
                                const uint32_t p0 = pack_sat(recombine(acc[u][0][1], acc[u][1][1], acc[u][2][1]),
                                                             recombine(acc[u][0][0], acc[u][1][0], acc[u][2][0]), 0u);
                                const uint32_t p1 = pack_sat(recombine(acc[u][0][3], acc[u][1][3], acc[u][2][3]),
                                                             recombine(acc[u][0][2], acc[u][1][2], acc[u][2][2]), 0u);
                                if (jt + u * kVWarps < lim0)
                                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa0 + (uint32_t)(u * 8 * kVWarps)), "h"((unsigned short)p0) : "memory");
                                if (jt + u * kVWarps < lim1)
                                    asm volatile("st.shared.u16 [%0], %1;" ::"r"(oa1 + (uint32_t)(u * 8 * kVWarps)), "h"((unsigned short)p1) : "memory");
                            }
                        }
#endif
                        __syncwarp();
                        for (int bb = bs; bb <= be; ++bb)                                // bands completed by the rows stored so far
                            if (min(VIS_PATCH * (bb + 1), seg_rows) <= ycount + ihi && lane == 0)
                                mbar_arrive(bar(OF, (nb + bb) % kOSlots));
                    }
                }
                ycount += n;
                if (c + 1 < n_chunks) {                   // carry: the last W-1 words of every column -> front of the other slot
                    const uint32_t dst0 = smem_u32(smem + L.off_hring + (slot ^ 1) * 3 * L.hplane);
                    for (int jt = wv; jt < n_tiles; jt += kVWarps) {                     // lane (g, t): words t and t + 4 of column 8 jt + g
                        const int cc = 8 * jt + g;
                        const uint32_t o = (uint32_t)(cc * L.cpitch + 4 * t);
                        if (cc < NC && t < W - 1) sts32(dst0 + o, lds32(ring0 + o + (uint32_t)CR));
                        if (cc < NC && t + 4 < W - 1) sts32(dst0 + o + 16, lds32(ring0 + o + (uint32_t)CR + 16));
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(HE, slot));       // H-ring slot and record slot consumed
            }
            nb += (seg_rows + VIS_PATCH - 1) / VIS_PATCH;
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
                const uint32_t oplane = (uint32_t)(S.x1 - S.x0), opitch = 3 * oplane;     // band tile: [row][channel * sw + x]
                unsigned char* const dst0 = reinterpret_cast<unsigned char*>(frames[f].second) + (size_t)S.x0 * 3;
                for (int y = G.y0; y < G.y1; y += VIS_PATCH, ++nb) {
                    const int os = nb % kOSlots, rows = min(VIS_PATCH, G.y1 - y);
                    mbar_wait(bar(OF, os), (nb / kOSlots) & 1);
                    const uint32_t otile = smem_u32(smem + L.off_otile + os * 3 * L.oplane);
                    for (int i = sw_i * 32 + lane; i < rows * wpr; i += kSWarps * 32) {
                        const int rr = i / wpr, q = i - rr * wpr;
                        const uint32_t at = otile + (uint32_t)rr * opitch + (uint32_t)(q * 4);
                        const uint32_t A = lds32(at), B = lds32(at + oplane), C = lds32(at + 2 * oplane);
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
            int ca[5], ra[5], rb[5], xa[5], xb[5], go[5], lo[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int item = min(lane + 32 * i, 146);
                const int c = item / 49, q = item - c * 49;
                const int f0 = 4 * q, f2 = f0 + 2;
                ca[i] = c; ra[i] = f0 / VIS_PATCH; rb[i] = f2 / VIS_PATCH;
                xa[i] = f0 - ra[i] * VIS_PATCH; xb[i] = f2 - rb[i] * VIS_PATCH;
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
                const int oplane = S.x1 - S.x0, opitch = 3 * oplane;                    // band tile: [row][channel * sw + x]
                int sa[5], sb[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    sa[i] = ca[i] * oplane + ra[i] * opitch + xa[i];
                    sb[i] = ca[i] * oplane + rb[i] * opitch + xb[i];
                }
                float* const frame_out = pixel_values + (size_t)frames[f].second * VIS_ROW_FLOATS;
                for (int gy = G.y0 / VIS_PATCH; gy < G.y1 / VIS_PATCH; ++gy, ++nb) {
                    const int os = nb % kOSlots;
                    mbar_wait(bar(OF, os), (nb / kOSlots) & 1);
                    const unsigned char* otile = smem + L.off_otile + os * 3 * L.oplane;
                    float* band = frame_out + (size_t)((gy >> 1) * half_gw * 4 + (gy & 1) * 2) * VIS_ROW_FLOATS;
                    for (int gp = sw_i; gp < n_patches; gp += kSWarps) {
                        const int gx = gx0 + gp;
                        float* prow = band + (size_t)((gx >> 1) * 4 + (gx & 1)) * VIS_ROW_FLOATS;
                        const unsigned char* pt = otile + gp * VIS_PATCH;
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

template <int KS, bool U8>
int launch_mma(const VisSched& sc, const void* frames, int n_frames, const LayoutM& L, int64_t dst_pitch, const int* hrec,
               const int* vrec, const float* lut768, float* pixel_values, cudaStream_t st) {
    auto kern = k_fused_mma<KS, U8>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_fused_mma: cudaFuncSetAttribute");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_items = n_frames * sc.n_strips * sc.n_segs;
    const int grid = n_items < sms ? n_items : sms;
    kern<<<grid, n_threads(KS), L.total, st>>>(sc, reinterpret_cast<const FramePtrs*>(frames), n_items, L, (long long)dst_pitch,
                                          hrec, vrec, lut768, pixel_values);
    return vis::check_launch("vis_fused_mma");
}

}  // namespace

namespace visf {

// up to two k-steps: one tile of 16 columns per horizontal-pass warp, whose coefficient fragments then stay in registers
int mma_max_strip_w(int ksteps) { return ksteps <= 2 ? 16 * h_warps(2) : VIS_MMA_HOIST3 ? 16 * h_warps(3) : 256; }
int mma_max_ksteps() { return 3; }
int mma_layout_bytes(int stage_pitch, int strip_w, int words) { return make_layout_m(stage_pitch, strip_w, words).total; }
int mma_record_stride(int words) { return rec_stride_mma(words); }

int mma_launch(const VisSched& sc, const void* frames, int n_frames, int64_t dst_pitch, const int* hrec, const int* vrec,
               const float* lut768, float* pixel_values, cudaStream_t st) {
    const int W = sc.dp_words;
    if (W < 4 || W > 9 || sc.ring != 16 || sc.per_index != 1 || sc.mma_ks < 1 || sc.mma_ks > 3 || sc.chunk_rows < 4 ||
        sc.chunk_rows > kChunk || sc.chunk_rows % 4) {
        vis::set_error("vis_fused_mma: schedule of another kernel class (ring %d, %d words, %d k-steps)", sc.ring, W, sc.mma_ks);
        return VIS_E_INVALID;
    }
    const LayoutM L = make_layout_m(sc.stage_pitch, sc.max_strip_w, W);
    if (L.total > kSmemMax) {
        vis::set_error("vis_fused_mma: %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
    const bool u8 = sc.out_mode == VIS_SCHED_OUT_U8;
#define VIS_LM(KK) (u8 ? launch_mma<KK, true>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st) \
                       : launch_mma<KK, false>(sc, frames, n_frames, L, dst_pitch, hrec, vrec, lut768, pixel_values, st))
    return sc.mma_ks <= 2 ? VIS_LM(2) : VIS_LM(3);
#undef VIS_LM
}

}  // namespace visf
