// vis_quality.cu — image-quality statistics of inspection frames ("next" row 3).
//
// Replaces the array work of src/safety/image_quality.py:42-125 of the reference:
//   cv2.cvtColor(BGR2GRAY)            -> gray = (3735*B + 19235*G + 9798*R + 2^14) >> 15   (cv: RGB2Gray<uchar>, 15-bit)
//   cv2.Laplacian(gray, CV_64F).var() -> 3x3 kernel [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101; integer valued
//   np.mean(gray)
// One pass over the BGR frame: a CTA converts a 128x32 tile plus a one-pixel halo to gray in shared memory (packed, four
// pixels per word), every thread evaluates the Laplacians of four words in packed 16-bit lanes, and the three sums the
// caller needs — sum(gray), sum(lap), sum(lap^2) — are reduced exactly in int64 (warp shuffles, one atomicAdd per CTA).  Variance and scores are
// finished on the host from these exact sums.  Bound: HBM (H*W*3 bytes read per frame).
#include "vis_internal.h"

namespace {

constexpr int kTW = 128, kTH = 32, kThreads = 256;
constexpr int kTilesPerCta = 8;                   // a CTA walks 8 tiles down its column: launch overhead, reduction and atomics per 256 rows
constexpr int kRowWords = kTW / 4 + 2;            // one word of left halo (its last byte is used), 32 tile words, one of right halo

__device__ __forceinline__ int reflect101(int i, int n) {       // cv: BORDER_REFLECT_101, n >= 1
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

// cv: RGB2Gray<uchar> on a BGR pixel held in the low three bytes of `p`: two packed 16 x 8-bit dot products
__device__ __forceinline__ unsigned gray_of(unsigned p) {
    const unsigned acc = __dp2a_hi(9798u, p, 1u << 14);                    // R * 9798 (+ 0 * byte 3) + rounding
    return __dp2a_lo(3735u | (19235u << 16), p, acc) >> 15;                // + B * 3735 + G * 19235
}
__device__ __forceinline__ unsigned gray_bytes(const unsigned char* p) {
    return gray_of((unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16));
}

// Phase 1 converts the tile and its one-pixel halo to gray, packed four pixels per 32-bit word (tile column 4q .. 4q+3
// in word q + 1 of its row; the halos are the last byte of word 0 and the first byte of word 33).  Phase 2 works on
// whole words: the left / right neighbours come from funnel shifts, the four Laplacians of a word are evaluated in two
// packed 16-bit lanes (biased by 1024 so that no borrow crosses a lane), and the three sums stay in 32-bit registers
// until the warp reduction (|lap| <= 1020).
__global__ void __launch_bounds__(kThreads)
k_quality(const VisQualityFrame* __restrict__ frames, long long* __restrict__ sums) {
    __shared__ unsigned g[kTH + 2][kRowWords];
    __shared__ long long red[3][kThreads / 32];
    // the frame index is the FASTEST grid dimension: the CTAs resident at any moment belong to many frames, so their
    // final atomicAdds land on different sums (510 CTAs of one 1080p frame adding to one 24-byte record serialise in L2)
    const VisQualityFrame f = frames[blockIdx.x];
    const int x0 = blockIdx.y * kTW;
    if (x0 >= f.w || (int)blockIdx.z * kTilesPerCta * kTH >= f.h) return;
    const int tid = threadIdx.x;
    int sg = 0, sl = 0, sl2 = 0;                  // a thread sees <= 128 pixels: sum(lap^2) < 2^27
    for (int t = 0; t < kTilesPerCta; ++t) {
    const int y0 = ((int)blockIdx.z * kTilesPerCta + t) * kTH;
    if (y0 >= f.h) break;                         // uniform
    if (t) __syncthreads();                       // the previous tile's words are consumed
    const bool fast = ((f.pitch | (int64_t)(uintptr_t)f.src) & 3) == 0 && x0 + kTW <= f.w;
    if (fast) {                                   // interior columns: 4 pixels = three aligned 32-bit loads per thread
        for (int i = tid; i < (kTH + 2) * (kTW / 4); i += kThreads) {
            const int r = i / (kTW / 4), q = i - r * (kTW / 4);
            const int y = reflect101(y0 + r - 1, f.h);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(f.src + (size_t)y * f.pitch + (size_t)(x0 + 4 * q) * 3);
            const uint32_t a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
            const unsigned g0 = gray_of(a), g1 = gray_of(__byte_perm(a, b, 0x0543));
            const unsigned g2 = gray_of(__byte_perm(b, d, 0x0432)), g3 = gray_of(d >> 8);
            g[r][1 + q] = __byte_perm(__byte_perm(g0, g1, 0x0040), __byte_perm(g2, g3, 0x0040), 0x5410);
        }
        for (int i = tid; i < (kTH + 2) * 2; i += kThreads) {      // the two halo columns
            const int r = i >> 1, right = i & 1;
            const int y = reflect101(y0 + r - 1, f.h), x = reflect101(right ? x0 + kTW : x0 - 1, f.w);
            const unsigned v = gray_bytes(f.src + (size_t)y * f.pitch + (size_t)x * 3);
            g[r][right ? kRowWords - 1 : 0] = right ? v : v << 24;
        }
    } else {                                      // edge tiles / unaligned frames: byte by byte (indices reflected)
        unsigned char* gb = reinterpret_cast<unsigned char*>(&g[0][0]);
        for (int i = tid; i < (kTH + 2) * (kTW + 2); i += kThreads) {
            const int r = i / (kTW + 2), c = i - r * (kTW + 2);                 // c = 0 is the left halo
            const int y = reflect101(y0 + r - 1, f.h), x = reflect101(x0 + c - 1, f.w);
            gb[r * kRowWords * 4 + 3 + c] = (unsigned char)gray_bytes(f.src + (size_t)y * f.pitch + (size_t)x * 3);
        }
    }
    __syncthreads();
    const int q = tid & 31, rb = (tid >> 5) * 4;                              // word column, first of 4 consecutive rows
    const int vx = min(4, f.w - (x0 + 4 * q));                                // valid pixels of this word (<= 0: none)
    if (vx > 0) {
        unsigned up = g[rb][q + 1], own = g[rb + 1][q + 1];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = rb + k;
            const unsigned down = g[r + 2][q + 1], prev = g[r + 1][q], next = g[r + 1][q + 2];
            if (y0 + r < f.h) {
                const unsigned left = __funnelshift_l(prev, own, 8), right = __funnelshift_r(own, next, 8);
                // even (0, 2) and odd (1, 3) pixels of the word in 16-bit lanes
                const unsigned oe = own & 0x00ff00ffu, oo = (own >> 8) & 0x00ff00ffu;
                const unsigned se = (up & 0x00ff00ffu) + (down & 0x00ff00ffu) + (left & 0x00ff00ffu) + (right & 0x00ff00ffu);
                const unsigned so = ((up >> 8) & 0x00ff00ffu) + ((down >> 8) & 0x00ff00ffu) + ((left >> 8) & 0x00ff00ffu) +
                                    ((right >> 8) & 0x00ff00ffu);
                const unsigned le = se + 0x04000400u - 4u * oe, lo = so + 0x04000400u - 4u * oo;   // lap + 1024 per lane
                const int l0 = (int)(le & 0xffffu) - 1024, l2 = (int)(le >> 16) - 1024;
                const int l1 = (int)(lo & 0xffffu) - 1024, l3 = (int)(lo >> 16) - 1024;
                if (vx == 4) {
                    sg += (int)__dp4a(own, 0x01010101u, 0u);
                    sl += l0 + l1 + l2 + l3;
                    sl2 += l0 * l0 + l1 * l1 + l2 * l2 + l3 * l3;
                } else {                                                        // the word straddles the right image edge
                    const int l[4] = {l0, l1, l2, l3};
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        if (j < vx) { sg += (int)((own >> (8 * j)) & 0xffu); sl += l[j]; sl2 += l[j] * l[j]; }
                }
            }
            up = own;
            own = down;
        }
    }
    }                                             // tiles of this CTA
    long long lsg = sg, lsl = sl, lsl2 = sl2;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lsg += __shfl_down_sync(0xffffffffu, lsg, o);
        lsl += __shfl_down_sync(0xffffffffu, lsl, o);
        lsl2 += __shfl_down_sync(0xffffffffu, lsl2, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = lsg; red[1][tid >> 5] = lsl; red[2][tid >> 5] = lsl2; }
    __syncthreads();
    if (tid < 3) {
        long long t = 0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += red[tid][k];
        atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3 * (size_t)blockIdx.x + tid), (unsigned long long)t);
    }
}

}  // namespace

extern "C" int vis_quality_stats(const VisQualityFrame* frames, int n_frames, int max_h, int max_w,
                                 int64_t* sums, void* stream) {
    if (!frames || !sums || n_frames <= 0 || n_frames > 65535 || max_h <= 0 || max_w <= 0) {
        vis::set_error("vis_quality_stats: bad arguments (frames=%d max %dx%d)", n_frames, max_w, max_h);
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(int64_t) * 3 * (size_t)n_frames, st);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_quality_stats: cudaMemsetAsync");
    dim3 grid(n_frames, (max_w + kTW - 1) / kTW, (max_h + kTH * kTilesPerCta - 1) / (kTH * kTilesPerCta));
    k_quality<<<grid, kThreads, 0, st>>>(frames, reinterpret_cast<long long*>(sums));
    return vis::check_launch("vis_quality_stats");
}
