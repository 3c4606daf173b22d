// ubench_pipes.cu — issue-rate micro-benchmarks for the instruction mixes of the resampling roles (sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench_pipes tools/ubench_pipes.cu && gpurun_out/ubench_pipes
//
// One CTA of 512 threads per SM (4 warps per scheduler), 8 independent dependency chains per thread, every operation
// an `asm volatile` so nothing is folded.  Cycles are the CTA's own clock64() span (max over CTAs), so the result is
// in warp-instructions per clock per SM and does not depend on the SM clock.  Printed as one JSON object per line;
// the results decide which arithmetic the 9..32-tap kernel should use.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define CH 8

enum Mix { IMAD, IDP4A, FFMA, PRMT, IMAD_PRMT, IDP_PRMT, IMAD_FFMA, IMAD_2FFMA, IDP_FFMA, LOP3, IMAD_LOP3, SHF, IMAD_SHF_MNMX,
           IDP3_IMAD1, HFMA2, IMAD_HFMA2, I2F, IMAD_WIDE, N_MIX };
static const char* kNames[N_MIX] = {"imad", "idp4a", "ffma", "prmt", "imad+prmt", "idp4a+prmt", "imad+ffma", "imad+2ffma", "idp4a+ffma",
                                    "lop3", "imad+lop3", "shf", "imad+shf+vimnmx", "3idp4a+1imad", "hfma2", "imad+hfma2", "i2f",
                                    "imad.wide"};
static const int kOps[N_MIX] = {1, 1, 1, 1, 2, 2, 2, 3, 2, 1, 2, 1, 3, 4, 1, 2, 1, 1};

template <int MIX>
__global__ void __launch_bounds__(512, 1) k_mix(uint32_t* out, long long* cycles, uint32_t seed) {
    uint32_t a[CH], b[CH];
    float f[CH];
    unsigned long long w[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = a[i] ^ 0x5a5a5a5a; f[i] = (float)(a[i] & 255); w[i] = a[i]; }
    const uint32_t k = seed | 3, sel = 0x4440 + (seed & 3);
    const float fk = (float)(seed & 15) + 0.5f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MIX == IMAD || MIX == IMAD_PRMT || MIX == IMAD_FFMA || MIX == IMAD_2FFMA || MIX == IMAD_LOP3 || MIX == IMAD_SHF_MNMX ||
                MIX == IDP3_IMAD1 || MIX == IMAD_HFMA2)
                asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(k));
            if (MIX == IDP4A || MIX == IDP_PRMT || MIX == IDP_FFMA)
                asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(k));
            if (MIX == IDP3_IMAD1) {
                asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(b[i]), "r"(k));
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[(i + 1) % CH]) : "r"(b[i]), "r"(sel));
                asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[(i + 2) % CH]) : "r"(b[i]), "r"(seed));
            }
            if (MIX == FFMA || MIX == IMAD_FFMA || MIX == IMAD_2FFMA || MIX == IDP_FFMA)
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(f[(i + 1) % CH]), "f"(fk));
            if (MIX == IMAD_2FFMA)
                asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(f[(i + 3) % CH]), "f"(fk));
            if (MIX == PRMT || MIX == IMAD_PRMT || MIX == IDP_PRMT)
                asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(b[i]) : "r"(b[i]), "r"(a[(i + 1) % CH]), "r"(sel));
            if (MIX == LOP3 || MIX == IMAD_LOP3)
                asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(b[i]) : "r"(b[i]), "r"(a[(i + 1) % CH]), "r"(k));
            if (MIX == SHF || MIX == IMAD_SHF_MNMX)
                asm volatile("shf.r.clamp.b32 %0, %1, %2, %3;" : "=r"(b[i]) : "r"(b[i]), "r"(a[(i + 1) % CH]), "r"(k & 31));
            if (MIX == IMAD_SHF_MNMX)
                asm volatile("min.s32.relu %0, %1, %2;" : "=r"(b[i]) : "r"(b[i]), "r"(k));
            if (MIX == HFMA2 || MIX == IMAD_HFMA2)
                asm volatile("fma.rn.f16x2 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(b[(i + 1) % CH]), "r"(sel));
            if (MIX == I2F)
                asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(a[i]));
            if (MIX == IMAD_WIDE)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(k));
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i] + b[i] + __float_as_uint(f[i]) + (uint32_t)w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MIX>
void run(int sms, uint32_t* out, long long* cyc_d) {
    k_mix<MIX><<<sms, 512>>>(out, cyc_d, 12345u);          // warm-up
    k_mix<MIX><<<sms, 512>>>(out, cyc_d, 12345u);
    cudaDeviceSynchronize();
    long long cyc[1024];
    cudaMemcpy(cyc, cyc_d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double warp_instr = 16.0 * ITERS * CH * kOps[MIX];
    printf("{\"mix\": \"%s\", \"ops_per_slot\": %d, \"cycles\": %lld, \"warp_instr_per_clk_per_sm\": %.3f, \"lanes_per_clk_per_sm\": %.1f}\n",
           kNames[MIX], kOps[MIX], mx, warp_instr / mx, 32.0 * warp_instr / mx);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(uint32_t) * sms * 512);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    run<IMAD>(sms, out, cyc);
    run<IDP4A>(sms, out, cyc);
    run<FFMA>(sms, out, cyc);
    run<PRMT>(sms, out, cyc);
    run<LOP3>(sms, out, cyc);
    run<SHF>(sms, out, cyc);
    run<HFMA2>(sms, out, cyc);
    run<I2F>(sms, out, cyc);
    run<IMAD_WIDE>(sms, out, cyc);
    run<IMAD_PRMT>(sms, out, cyc);
    run<IDP_PRMT>(sms, out, cyc);
    run<IMAD_FFMA>(sms, out, cyc);
    run<IMAD_2FFMA>(sms, out, cyc);
    run<IDP_FFMA>(sms, out, cyc);
    run<IMAD_LOP3>(sms, out, cyc);
    run<IMAD_SHF_MNMX>(sms, out, cyc);
    run<IDP3_IMAD1>(sms, out, cyc);
    run<IMAD_HFMA2>(sms, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
