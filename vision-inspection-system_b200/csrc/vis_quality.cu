// vis_quality.cu — image-quality statistics of inspection frames ("next" row 3).
//
// Replaces the array work of src/safety/image_quality.py:42-125 of the reference:
//   cv2.cvtColor(BGR2GRAY)            -> gray = (3735*B + 19235*G + 9798*R + 2^14) >> 15   (cv: RGB2Gray<uchar>, 15-bit)
//   cv2.Laplacian(gray, CV_64F).var() -> 3x3 kernel [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101; integer valued
//   np.mean(gray)
// One pass over the BGR frame.  A CTA owns up to eight adjacent 128-pixel strips of a band of rows (18 .. 180, by batch size); a producer warp
// streams the band's rows (one cp.async.bulk of up to 3 KB per row, plus 16 bytes of halo either side) through a ring
// of shared-memory stages guarded by mbarriers, so the bytes in flight do not depend on what the compute warps are
// doing.  A compute WARP owns one strip and walks down it; a lane owns four pixels (12 bytes = three words, read
// conflict-free from the stage) of every row and keeps the gray rows above / at / below in registers, in two packed
// 16-bit lanes each (even pixels, odd pixels).  The gray conversion is eight IDP.2A on the raw words
// with weights doubled, so that the result is byte 2 of the accumulator (no shift, no unaligned pixel extraction); the
// horizontal neighbours of a lane's first / last pixel come from the adjacent lanes by shuffle (the strip's own halo
// columns are converted once per 32 rows, one row per lane, and broadcast); the four Laplacians of a lane are two
// packed 16-bit sums biased by 1024.  Sums are kept biased (sum v, sum v^2, count) and unbiased once per band in int64;
// the three sums the caller needs — sum(gray), sum(lap), sum(lap^2) — are reduced exactly (warp shuffles, one
// atomicAdd triple per CTA of eight strips).  Variance and scores are finished on the host from these exact sums.
// Frames whose rows are not 16-byte aligned take the same walk with per-lane 32-bit global loads (4-byte aligned) or
// byte loads (anything else, and the partial strip at the right edge).  Bound: HBM (H*W*3 bytes read per frame, + 2
// halo rows per band).
#include "vis_fused_common.cuh"

namespace {
using namespace visf;

#ifndef VIS_Q_BAND
#define VIS_Q_BAND 0                              // > 0: a fixed band height instead of the host's choice (A/B builds)
#endif
#ifndef VIS_Q_AHEAD
#define VIS_Q_AHEAD 4
#endif
#ifndef VIS_Q_MINB
#define VIS_Q_MINB 4
#endif
#ifndef VIS_Q_K
#define VIS_Q_K 8
#endif
#ifndef VIS_Q_S
#define VIS_Q_S 2
#endif
constexpr int kStripW = 128;                      // pixels per warp-row: 32 lanes x 4 pixels
constexpr int kStripBytes = kStripW * 3;
constexpr int kMaxBandRows = 256;                 // rows a warp walks, chosen per launch: sum v^2 <= 2044^2 * 4 * 256 < 2^32
constexpr int kAhead = VIS_Q_AHEAD;               // global-load walk: rows in flight per warp (divides 32)
constexpr int kWarps = 8;                         // compute warps = strips of a CTA; warp 8 is the producer
constexpr int kThreads = (kWarps + 1) * 32;
constexpr int kK = VIS_Q_K, kS = VIS_Q_S;         // rows per stage (<= 32), stages of the ring
constexpr int kPad = 16;                          // halo bytes either side of a staged row (one pixel is used)
constexpr int kRowPitch = kPad + kWarps * kStripBytes + kPad;
constexpr int kStageBytes = kK * kRowPitch;
constexpr int kSmemBytes = kS * kStageBytes;
static_assert(kK <= 16 && kAhead <= 16 && 32 % kAhead == 0, "a lane converts the halo pixels of one row of a stage / of 32 rows");

// cv: BORDER_REFLECT_101 for -n < i < 2n (one reflection), n >= 1; the clamp covers n = 1 (every index -> 0) and the
// far side of frames narrower than a strip, whose pixels are masked out anyway
__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// cv: RGB2Gray<uchar> with the weights doubled: gray = byte 2 of (2*3735*B + 2*19235*G + 2*9798*R + 2^15), byte 3 = 0
constexpr unsigned kWB = 2 * 3735, kWG = 2 * 19235, kWR = 2 * 9798, kRound = 1u << 15;
constexpr unsigned kW_BG = kWB | (kWG << 16), kW_R0 = kWR, kW_0B = kWB << 16, kW_GR = kWG | (kWR << 16);

__device__ __forceinline__ unsigned gray2(unsigned b, unsigned g, unsigned r) { return kWB * b + kWG * g + kWR * r + kRound; }
__device__ __forceinline__ unsigned gray2_bytes(const unsigned char* p) { return gray2(p[0], p[1], p[2]); }

struct Row { unsigned e, o; };                    // gray of pixels (0, 2) and (1, 3) of the lane, 16-bit lanes

// three aligned words = four BGR pixels -> the two packed gray lanes: 8 IDP.2A + 2 PRMT
__device__ __forceinline__ Row gray_row(unsigned a, unsigned b, unsigned c) {
    const unsigned p0 = __dp2a_hi(kW_R0, a, __dp2a_lo(kW_BG, a, kRound));            // B G R .
    const unsigned p1 = __dp2a_lo(kW_GR, b, __dp2a_hi(kW_0B, a, kRound));            // . . . B | G R
    const unsigned p2 = __dp2a_lo(kW_R0, c, __dp2a_hi(kW_BG, b, kRound));            // . . B G | R
    const unsigned p3 = __dp2a_hi(kW_GR, c, __dp2a_lo(kW_0B, c, kRound));            // . B G R
    return Row{__byte_perm(p0, p2, 0x7632), __byte_perm(p1, p3, 0x7632)};
}

// one row of the walk: the Laplacians of the lane's four pixels from the gray rows above / at / below, accumulated biased
// sums per lane and band: gray and v = lap + 1024 first in packed 16-bit lanes (pg, pv: flushed every <= 16 rows), v^2 as
// v * (v & 255) and v * (v >> 8) (two IDP.2A on the packed lanes against their own bytes; <= 2044 * 255 * 4 * 256 < 2^32)
struct Acc { unsigned pg, pv, sg, sv, qa, qb; };
__device__ __forceinline__ void flush(Acc& a) {
    a.sg += (a.pg & 0xffffu) + (a.pg >> 16);
    a.sv += (a.pv & 0xffffu) + (a.pv >> 16);
    a.pg = a.pv = 0;
}
__device__ __forceinline__ void lap_row(const Row& up, const Row& own, const Row& down, unsigned gl, unsigned gr, int lane,
                                        unsigned me, unsigned mo, bool masked, Acc& a) {
    unsigned o_prev = __shfl_up_sync(0xffffffffu, own.o, 1), e_next = __shfl_down_sync(0xffffffffu, own.e, 1);
    if (lane == 0) o_prev = gl;                                           // the strip's halo columns
    if (lane == 31) e_next = gr;
    const unsigned left_e = __byte_perm(o_prev, own.o, 0x5432);           // (pixel -1, pixel 1)
    const unsigned right_o = __byte_perm(own.e, e_next, 0x5432);          // (pixel 2, pixel 4)
    unsigned le = up.e + down.e + left_e + own.o + 0x04000400u - 4u * own.e;           // lap + 1024 per 16-bit lane
    unsigned lo = up.o + down.o + own.e + right_o + 0x04000400u - 4u * own.o;
    unsigned ge = own.e, go = own.o;
    if (masked) { le &= me; lo &= mo; ge &= me; go &= mo; }
    a.pg += ge + go;                                                      // <= 510 per lane and row
    a.pv += le + lo;                                                      // <= 4088 per lane and row
    const unsigned be = __byte_perm(le, 0, 0x3120), bo = __byte_perm(lo, 0, 0x3120);   // (low bytes | high bytes) of the two lanes
    a.qa = __dp2a_lo(le, be, __dp2a_lo(lo, bo, a.qa));
    a.qb = __dp2a_hi(le, be, __dp2a_hi(lo, bo, a.qb));
}

// a halo pixel's doubled accumulator -> the lane position lap_row() wants (left: high lane, right: low lane)
__device__ __forceinline__ unsigned halo_left(unsigned g2) { return g2 & 0x00ff0000u; }
__device__ __forceinline__ unsigned halo_right(unsigned g2) { return (g2 >> 16) & 0xffu; }

// the strip's halo columns of 32 rows: lane r converts row yg + r
__device__ __forceinline__ void halo_columns(const VisQualityFrame& f, int yg, int lane, int xl, int xr, unsigned& hl, unsigned& hr) {
    const unsigned char* row = f.src + (size_t)min(yg + lane, f.h - 1) * f.pitch;
    hl = halo_left(gray2_bytes(row + (size_t)xl * 3));
    hr = halo_right(gray2_bytes(row + (size_t)xr * 3));
}

__device__ __forceinline__ void finish_band(const Acc& a, long long n, long long& out_g, long long& out_l, long long& out_l2) {
    out_g = a.sg;
    out_l = (long long)a.sv - 1024 * n;
    out_l2 = (long long)a.qa + 256ll * a.qb - 2048ll * a.sv + 1048576ll * n;
}

// ---- the staged walk: rows come from the CTA's ring ------------------------------------------------------------------
// Sequence row i of a band is frame row reflect101(yb - 1 + i): i = 0 is the row above the band, i = rows + 1 the row
// below it; stage st of the ring holds sequence rows st*kK .. st*kK + kK - 1.
struct Ring { uint32_t data, full, empty; };      // shared-space addresses: stages, kS full barriers, kS empty barriers

__device__ __forceinline__ void produce_band(const VisQualityFrame& f, const Ring& ring, int yb, int ye, int x0c, int n_ring) {
    const int x_end = x0c + n_ring * kStripW;
    const bool has_left = x0c > 0, has_right = x_end < f.w;               // (then >= 6 more pixels: see ring_strip())
    const uint32_t bytes = (has_left ? kPad : 0) + n_ring * kStripBytes + (has_right ? kPad : 0);
    const unsigned char* col = f.src + (size_t)x0c * 3 - (has_left ? kPad : 0);
    const uint32_t dst0 = ring.data + (has_left ? 0 : kPad);
    const int n_seq = ye - yb + 2;
    for (int st = 0, i0 = 0; i0 < n_seq; ++st, i0 += kK) {
        const int slot = st % kS, cnt = min(kK, n_seq - i0);
        if (st >= kS) mbar_wait(ring.empty + 8 * slot, ((st / kS) - 1) & 1);
        mbar_expect_tx(ring.full + 8 * slot, cnt * bytes);
        for (int j = 0; j < cnt; ++j)
            bulk_g2s(dst0 + slot * kStageBytes + j * kRowPitch, col + (size_t)reflect101(yb - 1 + i0 + j, f.h) * f.pitch, bytes,
                     ring.full + 8 * slot);
    }
}

// one stage of the ring: FIRST = the band's first stage (its first two rows only fill `up` and `own`), FULL = all kK rows
template <bool FIRST, bool FULL>
__device__ __forceinline__ void walk_stage(uint32_t base, uint32_t lane_off, uint32_t off_l, uint32_t off_r, int cnt, int lane,
                                           Row& up, Row& own, unsigned& hl, unsigned& hr, Acc& acc) {
    const unsigned hl_prev = hl, hr_prev = hr;
    {                                             // lane j converts the halo pixels of row j of the stage
        const uint32_t row = base + (FULL ? min(lane, kK - 1) : min(lane, cnt - 1)) * kRowPitch;
        unsigned b0, b1, b2, c0, c1, c2;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b0) : "r"(row + off_l));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b1) : "r"(row + off_l + 1));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b2) : "r"(row + off_l + 2));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c0) : "r"(row + off_r));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c1) : "r"(row + off_r + 1));
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c2) : "r"(row + off_r + 2));
        hl = halo_left(gray2(b0, b1, b2));
        hr = halo_right(gray2(c0, c1, c2));
    }
#pragma unroll
    for (int j = 0; j < kK; ++j) {
        if (FULL || j < cnt) {                    // uniform
            const uint32_t p = base + j * kRowPitch + lane_off;
            const Row down = gray_row(lds32(p), lds32(p + 4), lds32(p + 8));
            if (!FIRST || j >= 2) {               // `own` is row j - 1 of this stage, or the last row of the previous one
                const unsigned gl = j ? __shfl_sync(0xffffffffu, hl, j ? j - 1 : 0) : __shfl_sync(0xffffffffu, hl_prev, kK - 1);
                const unsigned gr = j ? __shfl_sync(0xffffffffu, hr, j ? j - 1 : 0) : __shfl_sync(0xffffffffu, hr_prev, kK - 1);
                lap_row(up, own, down, gl, gr, lane, 0, 0, false, acc);
            }
            up = own;
            own = down;
        }
    }
}

__device__ __forceinline__ void walk_band_ring(const VisQualityFrame& f, const Ring& ring, int warp, int x0, int yb, int ye, int lane,
                                               long long& out_g, long long& out_l, long long& out_l2) {
    const uint32_t lane_off = kPad + warp * kStripBytes + lane * 12;
    // the halo pixels, as byte offsets in a staged row: pixel x0 - 1 / x0 + 128, reflected at the frame's edges
    const uint32_t off_l = kPad + warp * kStripBytes + (x0 > 0 ? -3 : 3);
    const uint32_t off_r = kPad + warp * kStripBytes + (x0 + kStripW < f.w ? kStripBytes : kStripBytes - 6);
    const int n_seq = ye - yb + 2;
    Row up{0, 0}, own{0, 0};
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned hl = 0, hr = 0;
    for (int st = 0, i0 = 0; i0 < n_seq; ++st, i0 += kK) {
        const int slot = st % kS, cnt = min(kK, n_seq - i0);
        const uint32_t base = ring.data + slot * kStageBytes;
        mbar_wait(ring.full + 8 * slot, (st / kS) & 1);
        if (st == 0) walk_stage<true, false>(base, lane_off, off_l, off_r, cnt, lane, up, own, hl, hr, acc);
        else if (cnt == kK) walk_stage<false, true>(base, lane_off, off_l, off_r, cnt, lane, up, own, hl, hr, acc);
        else walk_stage<false, false>(base, lane_off, off_l, off_r, cnt, lane, up, own, hl, hr, acc);
        flush(acc);
        __syncwarp();
        if (lane == 0) mbar_arrive(ring.empty + 8 * slot);
    }
    finish_band(acc, 4ll * (ye - yb), out_g, out_l, out_l2);
}

// ---- the same walk on global loads: frames whose rows are only 4-byte aligned ------------------------------------------
// three aligned 32-bit loads per lane and row, issued kAhead rows before they are converted (a rotating register window)
__device__ __forceinline__ void walk_band_ldg(const VisQualityFrame& f, int x0, int yb, int ye, int lane,
                                           long long& out_g, long long& out_l, long long& out_l2) {
    const unsigned char* lane_ptr = f.src + (size_t)(x0 + 4 * lane) * 3;
    const int xl = reflect101(x0 - 1, f.w), xr = reflect101(x0 + kStripW, f.w);
    auto row_ptr = [&](int y) { return reinterpret_cast<const uint32_t*>(lane_ptr + (size_t)(y < f.h ? y : reflect101(y, f.h)) * f.pitch); };
    uint32_t ra[kAhead], rb[kAhead], rc[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; ++k) {            // rows yb + 1 .. yb + kAhead: the rows below the first kAhead rows
        const uint32_t* p = row_ptr(min(yb + 1 + k, ye));
        ra[k] = __ldg(p); rb[k] = __ldg(p + 1); rc[k] = __ldg(p + 2);
    }
    Row up, own;
    {
        const uint32_t* p = row_ptr(reflect101(yb - 1, f.h));
        const uint32_t* q = row_ptr(yb);
        const uint32_t a0 = __ldg(p), a1 = __ldg(p + 1), a2 = __ldg(p + 2), b0 = __ldg(q), b1 = __ldg(q + 1), b2 = __ldg(q + 2);
        up = gray_row(a0, a1, a2);
        own = gray_row(b0, b1, b2);
    }
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned hl = 0, hr = 0;
    for (int y = yb; y < ye; y += kAhead) {
        if (((y - yb) & 31) == 0) halo_columns(f, y, lane, xl, xr, hl, hr);            // kAhead divides 32
#pragma unroll
        for (int k = 0; k < kAhead; ++k) {
            if (y + k < ye) {                     // uniform
                const Row down = gray_row(ra[k], rb[k], rc[k]);
                if (y + k + 1 + kAhead <= ye) {
                    const uint32_t* p = row_ptr(y + k + 1 + kAhead);
                    ra[k] = __ldg(p); rb[k] = __ldg(p + 1); rc[k] = __ldg(p + 2);
                }
                const int r = (y + k - yb) & 31;
                lap_row(up, own, down, __shfl_sync(0xffffffffu, hl, r), __shfl_sync(0xffffffffu, hr, r), lane, 0, 0, false, acc);
                up = own;
                own = down;
            }
        }
        flush(acc);
    }
    finish_band(acc, 4ll * (ye - yb), out_g, out_l, out_l2);
}

// ---- edge strips and unaligned frames: byte by byte, columns reflected, the pixels beyond the right edge masked out ----
__device__ __forceinline__ void walk_band_edge(const VisQualityFrame& f, int x0, int yb, int ye, int lane,
                                            long long& out_g, long long& out_l, long long& out_l2) {
    const int x = x0 + 4 * lane;
    const int vx = min(max(f.w - x, 0), 4);                               // valid pixels of this lane's word
    const unsigned me = (vx > 0 ? 0x0000ffffu : 0u) | (vx > 2 ? 0xffff0000u : 0u);
    const unsigned mo = (vx > 1 ? 0x0000ffffu : 0u) | (vx > 3 ? 0xffff0000u : 0u);
    const int xl = reflect101(x0 - 1, f.w), xr = reflect101(x0 + kStripW, f.w);
    auto load_row = [&](int y) {
        unsigned g[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            g[j] = gray2_bytes(f.src + (size_t)y * f.pitch + (size_t)reflect101(x + j, f.w) * 3);
        return Row{__byte_perm(g[0], g[2], 0x7632), __byte_perm(g[1], g[3], 0x7632)};
    };
    Row up = load_row(reflect101(yb - 1, f.h)), own = load_row(yb);
    Acc acc{0, 0, 0, 0, 0, 0};
    unsigned hl = 0, hr = 0;
    for (int y = yb; y < ye; ++y) {
        const int r = (y - yb) & 31;
        if (r == 0) halo_columns(f, y, lane, xl, xr, hl, hr);
        const Row down = load_row(reflect101(y + 1, f.h));
        lap_row(up, own, down, __shfl_sync(0xffffffffu, hl, r), __shfl_sync(0xffffffffu, hr, r), lane, me, mo, true, acc);
        flush(acc);
        up = own;
        own = down;
    }
    finish_band(acc, (long long)vx * (ye - yb), out_g, out_l, out_l2);
}

// a strip goes through the ring when it is whole and its right halo pixel is either the reflection of one of its own
// pixels (the strip ends the row) or comes with 16 whole bytes of the same row (>= 6 more pixels)
__device__ __forceinline__ bool ring_strip(int x0, int w) { return x0 + kStripW == w || x0 + kStripW + 6 <= w; }

// A CTA = (band, group): a band's strips are split into groups of <= 8 adjacent strips, as evenly as possible.  The first
// n_ring strips of the group go through the ring (k_quality_ring), the others (a row's last strips, unaligned frames)
// through k_quality_rest: two kernels, so that neither carries the other's code.
struct Task { int ns, n_ring, s0, yb, ye, align; bool valid; };
__device__ __forceinline__ Task task_of(const VisQualityFrame& f, int cta, int band_rows) {
    Task t{};
    const int strips = (f.w + kStripW - 1) / kStripW, bands = (f.h + band_rows - 1) / band_rows;
    const int groups = (strips + kWarps - 1) / kWarps, per = (strips + groups - 1) / groups;
    t.valid = cta < bands * groups;
    if (!t.valid) return t;
    const int band = cta / groups;
    t.s0 = (cta - band * groups) * per;
    t.ns = min(per, strips - t.s0);
    t.yb = band * band_rows;
    t.ye = min(t.yb + band_rows, f.h);
    t.align = (int)((f.pitch | (int64_t)(uintptr_t)f.src) & 15);
    if (t.align == 0)                             // the ring's strips are a prefix of the group's (only a row's last strips can fail)
        while (t.n_ring < t.ns && ring_strip((t.s0 + t.n_ring) * kStripW, f.w)) ++t.n_ring;
    return t;
}

// the three sums of the CTA's compute warps -> one atomicAdd triple on the frame's record
__device__ __forceinline__ void add_sums(long long (*red)[kWarps], int warp, int lane, long long sg, long long sl, long long sl2,
                                         long long* frame_sums) {
    if (warp < kWarps) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sg += __shfl_down_sync(0xffffffffu, sg, o);
            sl += __shfl_down_sync(0xffffffffu, sl, o);
            sl2 += __shfl_down_sync(0xffffffffu, sl2, o);
        }
        if (lane == 0) { red[0][warp] = sg; red[1][warp] = sl; red[2][warp] = sl2; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        long long t = 0;
#pragma unroll
        for (int k = 0; k < kWarps; ++k) t += red[threadIdx.x][k];
        atomicAdd(reinterpret_cast<unsigned long long*>(frame_sums + threadIdx.x), (unsigned long long)t);
    }
}

// The frame index is the FASTEST grid dimension of both kernels: the CTAs resident at any moment belong to many frames,
// so their final atomicAdds land on different sums (the CTAs of one frame adding to one 24-byte record serialise in L2).
__global__ void __launch_bounds__(kThreads, VIS_Q_MINB)
k_quality_ring(const VisQualityFrame* __restrict__ frames, long long* __restrict__ sums, int band_rows) {
    extern __shared__ __align__(128) unsigned char q_smem[];
    __shared__ __align__(8) uint64_t bars[2 * kS];
    __shared__ long long red[3][kWarps];
    const int fi = blockIdx.x;
    const VisQualityFrame f = frames[fi];
    const Task t = task_of(f, blockIdx.y, band_rows);
    if (!t.valid || t.n_ring == 0) return;        // uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ring ring{smem_u32(q_smem), smem_u32(&bars[0]), smem_u32(&bars[kS])};
    if (threadIdx.x == 0) {
        for (int s = 0; s < kS; ++s) {
            mbar_init(ring.full + 8 * s, 1);
            mbar_init(ring.empty + 8 * s, t.n_ring);
        }
        fence_mbar_init();
    }
    __syncthreads();
    long long sg = 0, sl = 0, sl2 = 0;
    if (warp == kWarps) {
        if (lane == 0) produce_band(f, ring, t.yb, t.ye, t.s0 * kStripW, t.n_ring);
    } else if (warp < t.n_ring) {
        walk_band_ring(f, ring, warp, (t.s0 + warp) * kStripW, t.yb, t.ye, lane, sg, sl, sl2);
    }
    add_sums(red, warp, lane, sg, sl, sl2, sums + 3 * (size_t)fi);
}

__global__ void __launch_bounds__(kWarps * 32)
k_quality_rest(const VisQualityFrame* __restrict__ frames, long long* __restrict__ sums, int band_rows) {
    __shared__ long long red[3][kWarps];
    const int fi = blockIdx.x;
    const VisQualityFrame f = frames[fi];
    const Task t = task_of(f, blockIdx.y, band_rows);
    if (!t.valid || t.n_ring == t.ns) return;     // uniform
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long sg = 0, sl = 0, sl2 = 0;
    if (warp >= t.n_ring && warp < t.ns) {
        const int x0 = (t.s0 + warp) * kStripW;
        if ((t.align & 3) == 0 && x0 + kStripW <= f.w) walk_band_ldg(f, x0, t.yb, t.ye, lane, sg, sl, sl2);
        else walk_band_edge(f, x0, t.yb, t.ye, lane, sg, sl, sl2);
    }
    add_sums(red, warp, lane, sg, sl, sl2, sums + 3 * (size_t)fi);
}

}  // namespace

extern "C" int vis_quality_stats(const VisQualityFrame* frames, int n_frames, int max_h, int max_w,
                                 int64_t* sums, void* stream) {
    if (!frames || !sums || n_frames <= 0 || n_frames > 65535 || max_h <= 0 || max_w <= 0) {
        vis::set_error("vis_quality_stats: bad arguments (frames=%d max %dx%d)", n_frames, max_w, max_h);
        return VIS_E_INVALID;
    }
    // Band height: tall bands re-read fewer halo rows and fill the ring less often, short ones give a small batch enough
    // CTAs; the tallest of the list that still yields ~4 CTAs per resident slot (4 per SM), else the shortest.
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const long long strips = (max_w + kStripW - 1) / kStripW, groups = (strips + kWarps - 1) / kWarps;
    int band_rows = VIS_Q_BAND;
    if (band_rows <= 0) {
        static const int kChoices[] = {180, 120, 72, 36, 18};
        band_rows = kChoices[4];
        for (int c : kChoices)
            if ((long long)n_frames * groups * ((max_h + c - 1) / c) >= 4ll * VIS_Q_MINB * sms) { band_rows = c; break; }
    }
    static_assert(VIS_Q_BAND <= kMaxBandRows, "band too tall for the 32-bit sums");
    const long long ctas = groups * ((max_h + band_rows - 1) / band_rows);
    if (ctas > 65535) {
        vis::set_error("vis_quality_stats: frames of %dx%d are beyond the grid (%lld CTAs per frame)", max_w, max_h, ctas);
        return VIS_E_UNSUPPORTED;
    }
    static bool attr_set = false;                 // idempotent; a race sets it twice
    if (!attr_set) {
        cudaError_t ea = cudaFuncSetAttribute(k_quality_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (ea != cudaSuccess) return vis::cuda_fail(ea, "vis_quality_stats: cudaFuncSetAttribute");
        attr_set = true;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(int64_t) * 3 * (size_t)n_frames, st);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_quality_stats: cudaMemsetAsync");
    dim3 grid(n_frames, (unsigned)ctas);
    k_quality_ring<<<grid, kThreads, kSmemBytes, st>>>(frames, reinterpret_cast<long long*>(sums), band_rows);
    cudaError_t el = cudaGetLastError();
    if (el != cudaSuccess) return vis::cuda_fail(el, "vis_quality_stats: k_quality_ring");
    k_quality_rest<<<grid, kWarps * 32, 0, st>>>(frames, reinterpret_cast<long long*>(sums), band_rows);
    return vis::check_launch("vis_quality_stats");
}
