// ubench_imma.cu — is the legacy integer tensor path (mma.sync.m16n8k32 u8 x s8/u8 -> s32, SASS IMMA.16832) worth using for the
// banded 9..33-tap resampling passes on sm_100a?  Measures its issue rate alone and next to the epilogue the exact Pillow
// arithmetic needs (limb recombination, >> 22, saturating pack), and checks the fragment layout against a host product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_imma tools/ubench_imma.cu && gpurun_out/ubench_imma
//
// One CTA of 512 threads per SM, CH independent accumulator sets per warp; cycles = the CTA's own clock64() span.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 6

__device__ __forceinline__ void imma_us(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void imma_uu(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

enum Mix { IMMA, IMMA_EPI, EPI, CVTPACK, IMMA_LDS, N_MIX };
static const char* kNames[N_MIX] = {"imma.16832", "3imma+epilogue(8lea+4shf+2i2ip)", "epilogue alone", "cvt.pack.sat.u8.s32", "imma+2lds32"};
static const int kOps[N_MIX] = {1, 17, 14, 1, 3};

template <int MIX>
__global__ void __launch_bounds__(512, 1) k_mix(uint32_t* out, long long* cycles, uint32_t seed) {
    __shared__ uint32_t sm[2048];
    for (int i = threadIdx.x; i < 2048; i += 512) sm[i] = seed * i;
    int d[CH][4];
    uint32_t a[4], b[2], p[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { d[i][0] = d[i][1] = d[i][2] = d[i][3] = (int)(seed + i); p[i] = seed ^ i; }
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = seed * (threadIdx.x + i + 1);
    b[0] = seed ^ 0x01020304; b[1] = seed + 0x01010101;
    uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 4;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (MIX == IMMA) {
#pragma unroll
            for (int i = 0; i < CH; ++i) imma_us(d[i], a, b);
        }
        if (MIX == IMMA_LDS) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                uint32_t x, y;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(sa + 128u * i));
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(y) : "r"(sa + 128u * i + 1024u));
                const uint32_t bb[2] = {x, y};
                imma_us(d[i], a, bb);
            }
        }
        if (MIX == IMMA_EPI || MIX == EPI) {
            // CH / 3 tiles: three limb accumulators each, recombined, shifted, saturated and packed to one word per tile pair
#pragma unroll
            for (int i = 0; i + 2 < CH; i += 3) {
                if (MIX == IMMA_EPI) { imma_uu(d[i], a, b); imma_uu(d[i + 1], a, b); imma_us(d[i + 2], a, b); }
                int v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int t;
                    asm volatile("mad.lo.s32 %0, %1, 256, %2;" : "=r"(t) : "r"(d[i + 1][e]), "r"(d[i][e]));
                    asm volatile("mad.lo.s32 %0, %1, 65536, %2;" : "=r"(t) : "r"(d[i + 2][e]), "r"(t));
                    asm volatile("shr.s32 %0, %1, 22;" : "=r"(v[e]) : "r"(t));
                }
                uint32_t w;
                asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(w) : "r"(v[1]), "r"(v[0]));
                asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(w) : "r"(v[3]), "r"(v[2]), "r"(w));
                p[i] ^= w;
                if (MIX == EPI) { d[i][0] += (int)w; d[i + 1][1] ^= (int)w; d[i + 2][2] += it; }
            }
        }
        if (MIX == CVTPACK) {
#pragma unroll
            for (int i = 0; i < CH; ++i)
                asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %0;" : "+r"(p[i]) : "r"(d[i][0]), "r"(p[(i + 1) % CH]));
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3] + p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
void run(int sms, uint32_t* out, long long* cyc_d) {
    k_mix<MIX><<<sms, 512>>>(out, cyc_d, 12345u);
    k_mix<MIX><<<sms, 512>>>(out, cyc_d, 12345u);
    cudaDeviceSynchronize();
    long long cyc[1024];
    cudaMemcpy(cyc, cyc_d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double per_iter = (MIX == IMMA_EPI || MIX == EPI) ? (CH / 3) * (double)kOps[MIX] : CH * (double)kOps[MIX];
    const double warp_instr = 16.0 * ITERS * per_iter;
    const double immas = (MIX == IMMA || MIX == IMMA_LDS) ? 16.0 * ITERS * CH : (MIX == IMMA_EPI ? 16.0 * ITERS * CH : 0);
    printf("{\"mix\": \"%s\", \"cycles\": %lld, \"warp_instr_per_clk_per_sm\": %.3f, \"imma_per_clk_per_sm\": %.3f, \"int8_macs_per_clk_per_sm\": %.0f}\n",
           kNames[MIX], mx, warp_instr / mx, immas / mx, immas * 4096.0 / mx);
}

// ---- fragment layout check: D[16x8] = A[16x32] (u8, row) x B[32x8] (limb, "col" = k contiguous per column) ----
__global__ void k_layout(const uint8_t* A, const int8_t* Bs, const uint8_t* Bu, int* Ds, int* Du) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    uint32_t a[4], bs[2], bu[2];
    a[0] = *reinterpret_cast<const uint32_t*>(A + g * 32 + 4 * t);
    a[1] = *reinterpret_cast<const uint32_t*>(A + (g + 8) * 32 + 4 * t);
    a[2] = *reinterpret_cast<const uint32_t*>(A + g * 32 + 16 + 4 * t);
    a[3] = *reinterpret_cast<const uint32_t*>(A + (g + 8) * 32 + 16 + 4 * t);
    bs[0] = *reinterpret_cast<const uint32_t*>(Bs + g * 32 + 4 * t);          // column g, k = 4t..4t+3
    bs[1] = *reinterpret_cast<const uint32_t*>(Bs + g * 32 + 16 + 4 * t);
    bu[0] = *reinterpret_cast<const uint32_t*>(Bu + g * 32 + 4 * t);
    bu[1] = *reinterpret_cast<const uint32_t*>(Bu + g * 32 + 16 + 4 * t);
    int ds[4] = {0, 0, 0, 0}, du[4] = {0, 0, 0, 0};
    imma_us(ds, a, bs);
    imma_uu(du, a, bu);
    Ds[g * 8 + 2 * t] = ds[0]; Ds[g * 8 + 2 * t + 1] = ds[1]; Ds[(g + 8) * 8 + 2 * t] = ds[2]; Ds[(g + 8) * 8 + 2 * t + 1] = ds[3];
    Du[g * 8 + 2 * t] = du[0]; Du[g * 8 + 2 * t + 1] = du[1]; Du[(g + 8) * 8 + 2 * t] = du[2]; Du[(g + 8) * 8 + 2 * t + 1] = du[3];
}

static int layout_check() {
    uint8_t A[16 * 32], Bu[8 * 32];
    int8_t Bs[8 * 32];
    srand(7);
    for (auto& v : A) v = (uint8_t)(rand() & 255);
    for (auto& v : Bu) v = (uint8_t)(rand() & 255);
    for (auto& v : Bs) v = (int8_t)(rand() & 255);
    uint8_t *dA, *dBu; int8_t* dBs; int *dDs, *dDu;
    cudaMalloc(&dA, sizeof A); cudaMalloc(&dBu, sizeof Bu); cudaMalloc(&dBs, sizeof Bs);
    cudaMalloc(&dDs, 128 * 4); cudaMalloc(&dDu, 128 * 4);
    cudaMemcpy(dA, A, sizeof A, cudaMemcpyHostToDevice);
    cudaMemcpy(dBu, Bu, sizeof Bu, cudaMemcpyHostToDevice);
    cudaMemcpy(dBs, Bs, sizeof Bs, cudaMemcpyHostToDevice);
    k_layout<<<1, 32>>>(dA, dBs, dBu, dDs, dDu);
    int Ds[128], Du[128];
    cudaMemcpy(Ds, dDs, sizeof Ds, cudaMemcpyDeviceToHost);
    cudaMemcpy(Du, dDu, sizeof Du, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 16; ++m)
        for (int n = 0; n < 8; ++n) {
            int s = 0, u = 0;
            for (int k = 0; k < 32; ++k) { s += (int)A[m * 32 + k] * (int)Bs[n * 32 + k]; u += (int)A[m * 32 + k] * (int)Bu[n * 32 + k]; }
            bad += (s != Ds[m * 8 + n]) + (u != Du[m * 8 + n]);
        }
    printf("{\"layout_check\": \"%s\", \"mismatches\": %d}\n", bad ? "FAILED" : "ok", bad);
    return bad;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    const int bad = layout_check();
    run<IMMA>(sms, out, cyc);
    run<IMMA_EPI>(sms, out, cyc);
    run<EPI>(sms, out, cyc);
    run<CVTPACK>(sms, out, cyc);
    run<IMMA_LDS>(sms, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return bad ? 2 : 0;
}
