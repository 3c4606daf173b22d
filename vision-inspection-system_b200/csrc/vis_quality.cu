// vis_quality.cu — image-quality statistics of inspection frames ("next" row 3).
//
// Replaces the array work of src/safety/image_quality.py:42-125 of the reference:
//   cv2.cvtColor(BGR2GRAY)            -> gray = (3735*B + 19235*G + 9798*R + 2^14) >> 15   (cv: RGB2Gray<uchar>, 15-bit)
//   cv2.Laplacian(gray, CV_64F).var() -> 3x3 kernel [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101; integer valued
//   np.mean(gray)
// One pass over the BGR frame: a CTA converts a 128x32 tile plus a one-pixel halo to gray in shared memory, every
// thread evaluates the Laplacian of its pixels, and the three sums the caller needs — sum(gray), sum(lap),
// sum(lap^2) — are reduced exactly in int64 (warp shuffles, one atomicAdd per CTA).  Variance and scores are
// finished on the host from these exact sums.  Bound: HBM (H*W*3 bytes read per frame).
#include "vis_internal.h"

namespace {

constexpr int kTW = 128, kTH = 32, kThreads = 256;

__device__ __forceinline__ int reflect101(int i, int n) {       // cv: BORDER_REFLECT_101, n >= 1
    if (n == 1) return 0;
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

__global__ void __launch_bounds__(kThreads)
k_quality(const VisQualityFrame* __restrict__ frames, long long* __restrict__ sums) {
    __shared__ unsigned char g[kTH + 2][kTW + 4];
    __shared__ long long red[3][kThreads / 32];
    const VisQualityFrame f = frames[blockIdx.z];
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
    if (x0 >= f.w || y0 >= f.h) return;
    const int tid = threadIdx.x;
    // gray of the tile + halo (indices reflected at the image border, so the halo of an edge tile is real data)
    auto gray = [](int b, int gg, int r) { return (unsigned char)((3735 * b + 19235 * gg + 9798 * r + (1 << 14)) >> 15); };
    const bool fast = ((f.pitch | (int64_t)(uintptr_t)f.src) & 3) == 0 && x0 + kTW <= f.w;
    if (fast) {                                   // interior columns: 4 pixels = three aligned 32-bit loads per thread
        for (int i = tid; i < (kTH + 2) * (kTW / 4); i += kThreads) {
            const int r = i / (kTW / 4), q = i - r * (kTW / 4);
            const int y = reflect101(y0 + r - 1, f.h);
            const uint32_t* p = reinterpret_cast<const uint32_t*>(f.src + (size_t)y * f.pitch + (size_t)(x0 + 4 * q) * 3);
            const uint32_t a = __ldg(p), b = __ldg(p + 1), d = __ldg(p + 2);
            unsigned char* o = &g[r][1 + 4 * q];
            o[0] = gray(a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff);
            o[1] = gray(a >> 24, b & 0xff, (b >> 8) & 0xff);
            o[2] = gray((b >> 16) & 0xff, b >> 24, d & 0xff);
            o[3] = gray((d >> 8) & 0xff, (d >> 16) & 0xff, d >> 24);
        }
        for (int i = tid; i < (kTH + 2) * 2; i += kThreads) {      // the two halo columns
            const int r = i >> 1, c = (i & 1) ? kTW + 1 : 0;
            const int y = reflect101(y0 + r - 1, f.h), x = reflect101(x0 + c - 1, f.w);
            const unsigned char* p = f.src + (size_t)y * f.pitch + (size_t)x * 3;
            g[r][c] = gray(p[0], p[1], p[2]);
        }
    } else {
        for (int i = tid; i < (kTH + 2) * (kTW + 2); i += kThreads) {
            const int r = i / (kTW + 2), c = i - r * (kTW + 2);
            const int y = reflect101(y0 + r - 1, f.h), x = reflect101(x0 + c - 1, f.w);
            const unsigned char* p = f.src + (size_t)y * f.pitch + (size_t)x * 3;
            g[r][c] = gray(p[0], p[1], p[2]);
        }
    }
    __syncthreads();
    long long sg = 0, sl = 0, sl2 = 0;
    for (int i = tid; i < kTH * kTW; i += kThreads) {
        const int r = i / kTW, c = i - r * kTW;
        if (y0 + r < f.h && x0 + c < f.w) {
            const int v = g[r + 1][c + 1];
            const int lap = (int)g[r][c + 1] + (int)g[r + 2][c + 1] + (int)g[r + 1][c] + (int)g[r + 1][c + 2] - 4 * v;
            sg += v;
            sl += lap;
            sl2 += lap * lap;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sg += __shfl_down_sync(0xffffffffu, sg, o);
        sl += __shfl_down_sync(0xffffffffu, sl, o);
        sl2 += __shfl_down_sync(0xffffffffu, sl2, o);
    }
    if ((tid & 31) == 0) { red[0][tid >> 5] = sg; red[1][tid >> 5] = sl; red[2][tid >> 5] = sl2; }
    __syncthreads();
    if (tid < 3) {
        long long t = 0;
#pragma unroll
        for (int k = 0; k < kThreads / 32; ++k) t += red[tid][k];
        atomicAdd(reinterpret_cast<unsigned long long*>(sums + 3 * (size_t)blockIdx.z + tid), (unsigned long long)t);
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
    dim3 grid((max_w + kTW - 1) / kTW, (max_h + kTH - 1) / kTH, n_frames);
    k_quality<<<grid, kThreads, 0, st>>>(frames, reinterpret_cast<long long*>(sums));
    return vis::check_launch("vis_quality_stats");
}
