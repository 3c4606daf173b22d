// vis_fused_ws.cu — warp-specialised persistent variant of the hot kernel (frame -> Qwen2-VL pixel_values).
//
// Same arithmetic and the same three stages as vis_fused.cu (Pillow 8bpc horizontal pass -> uint8 -> vertical pass ->
// uint8 -> exact LUT -> patch layout), but instead of one CTA alternating between phases behind __syncthreads, ONE
// CTA per SM stays resident and its 25 warps run the stages concurrently as a pipeline over shared-memory rings:
//
//   loader (1 warp)  cp.async.bulk row segments + the strip's coefficient records  -> stage[2]      (mbarrier tx)
//   H      (12 warps) lane = input row, warp = 28-column sub-range, push-order MACs -> hring[2]     (uint8, planar)
//   V      (8 warps)  thread = 4 output columns of one channel, register row window -> otile[2]     (14-row band)
//   store  (4 warps)  band -> LUT -> 16-byte stores, both temporal copies           -> pixel_values
//
// Every hand-off is a full/empty mbarrier pair (one arrive per producing / consuming warp); there is no CTA-wide
// barrier after start-up, so a stage never waits for an unrelated one.  25 resident warps (vs 16 for the phased
// kernel) and per-role code paths that need < 80 registers give the schedulers more eligible warps, which is what
// the phased kernel's profile said it lacked (profiles/r01_fused_v2.txt: issue slots 56 % busy, 18 % barrier stalls).
// CTAs walk the strip list with a stride of gridDim.x; all roles iterate the same (strip, chunk, band) sequence.
#include "vis_fused_common.cuh"

using namespace visf;

namespace {

constexpr int kHWarps = 12, kVWarps = 8, kSWarps = 4;
constexpr int kThreadsWS = (kHWarps + kVWarps + kSWarps + 1) * 32;      // 800
constexpr int kVThreads = kVWarps * 32;
constexpr int kChunk = 32, kStepPx = 16, kMaxStripW = 336;
constexpr int kPitch = kMaxStripW + 4;            // 340 = 4 * 85: conflict-free lane = row byte stores
constexpr int kOPitch = kMaxStripW;               // band tile rows: word stores / u16 loads only, no padding needed
constexpr int kOPlane = VIS_PATCH * kOPitch;
constexpr int kVCap = 48;
constexpr int kSmemMax = 227 * 1024;

enum Bar { SF = 0, SE = 2, HF = 4, HE = 6, OF = 8, OE = 10, kBars = 12 };   // full/empty pairs, two slots each

struct LayoutWS {
    int stage_pitch;
    int hplane;                                    // bytes of one channel plane of an H-ring slot: (carry + 32) rows
    int off_stage, off_hring, off_otile, off_hrec, off_vrec, off_lut, off_bar, total;
    int stage_slot, hrec_slot, vrec_slot;          // bytes per ring slot
};

inline LayoutWS make_layout_ws(int span_bytes, int strip_w, int stride, int carry) {
    LayoutWS L;
    L.hplane = (carry + kChunk) * kPitch;
    L.stage_pitch = align_up(span_bytes, 16);
    if ((L.stage_pitch / 16) % 2 == 0) L.stage_pitch += 16;
    L.stage_slot = kChunk * L.stage_pitch;
    L.hrec_slot = align_up((strip_w + 1) * stride * 4, 16);
    L.vrec_slot = kVCap * stride * 4;
    int off = 0;
    L.off_stage = off; off += 2 * L.stage_slot;
    L.off_hring = off; off += 2 * 3 * L.hplane;
    L.off_otile = off; off += 2 * 3 * kOPlane;
    off = align_up(off, 16);
    L.off_hrec = off;  off += 2 * L.hrec_slot;
    L.off_vrec = off;  off += 2 * L.vrec_slot;
    L.off_lut = off;   off += 768 * 4;
    L.off_bar = off;   off += kBars * 8;
    L.total = off;
    return L;
}

struct Strip {                       // per-strip constants every role derives the same way
    VisFrame fr;
    int x0, x1, y0, y1, sw;
    int px0, row_bytes;              // staged columns start at px0 (multiple of 16 pixels); bytes per staged row
    int r_first, r_end, n_chunks;    // input rows [r_first, r_end) in chunks of 32
};

template <int STRIDE>
__device__ __forceinline__ Strip load_strip(const VisFrame* __restrict__ frames, const VisStrip* __restrict__ strips, int s) {
    Strip t;
    const VisStrip sp = strips[s];
    t.fr = frames[sp.frame];
    t.x0 = sp.x0; t.x1 = sp.x1; t.y0 = sp.y0; t.y1 = sp.y1;
    t.sw = t.x1 - t.x0;
    t.px0 = __ldg(t.fr.hrec + (size_t)t.x0 * STRIDE + STRIDE - 2) & ~(kStepPx - 1);
    const int px_last = __ldg(t.fr.hrec + (size_t)(t.x1 - 1) * STRIDE + STRIDE - 1);
    t.row_bytes = align_up((px_last + 1) * 3, 16) - t.px0 * 3;
    if ((int64_t)t.px0 * 3 + t.row_bytes > t.fr.src_pitch) t.row_bytes = (int)(t.fr.src_pitch - (int64_t)t.px0 * 3);
    t.r_first = __ldg(t.fr.vrec + (size_t)t.y0 * STRIDE + STRIDE - 2) & ~15;
    t.r_end = __ldg(t.fr.vrec + (size_t)(t.y1 - 1) * STRIDE + STRIDE - 1) + 1;
    t.n_chunks = (t.r_end - t.r_first + kChunk - 1) / kChunk;
    return t;
}

template <int KT, int RING, int STRIDE>
__global__ void __launch_bounds__(kThreadsWS, 1)
k_fused_ws(const VisFrame* __restrict__ frames, const VisStrip* __restrict__ strips, int n_strips, LayoutWS L,
           const float* __restrict__ lut768, float* __restrict__ pixel_values) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* lut = reinterpret_cast<float*>(smem + L.off_lut);              // transposed: lut[c * 256 + v]
    const uint32_t bar0 = smem_u32(smem + L.off_bar);
    auto bar = [&](int which, int slot) { return bar0 + (uint32_t)(which + slot) * 8; };

    for (int i = tid; i < 768; i += kThreadsWS) lut[(i % 3) * 256 + i / 3] = __ldg(lut768 + i);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar(SF, s), 1);
            mbar_init(bar(SE, s), kHWarps);
            mbar_init(bar(HF, s), kHWarps);
            mbar_init(bar(HE, s), kVWarps);
            mbar_init(bar(OF, s), kVWarps);
            mbar_init(bar(OE, s), kSWarps);
        }
        fence_mbar_init();
    }
    __syncthreads();                                   // the only CTA-wide barrier

    if (warp == kHWarps + kVWarps + kSWarps) {
        // ============================== loader ==============================
        int k = 0, sl = 0;
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x, ++sl) {
            const Strip t = load_strip<STRIDE>(frames, strips, s);
            const uint32_t rec_bytes = (uint32_t)(t.sw + 1) * STRIDE * 4;
            for (int c = 0; c < t.n_chunks; ++c, ++k) {
                const int slot = k & 1;
                if (k >= 2) mbar_wait(bar(SE, slot), ((k >> 1) - 1) & 1);
                const int r0 = t.r_first + c * kChunk;
                const int rows = min(kChunk, t.r_end - r0);
                if (lane == 0) {
                    fence_proxy_async();
                    mbar_expect_tx(bar(SF, slot), (uint32_t)rows * (uint32_t)t.row_bytes + (c == 0 ? rec_bytes : 0u));
                }
                __syncwarp();
                unsigned char* stage = smem + L.off_stage + slot * L.stage_slot;
                if (lane < rows)
                    bulk_g2s(smem_u32(stage + lane * L.stage_pitch),
                             t.fr.src + (size_t)(r0 + lane) * t.fr.src_pitch + (size_t)t.px0 * 3,
                             (uint32_t)t.row_bytes, bar(SF, slot));
                if (c == 0 && lane == 0)
                    bulk_g2s(smem_u32(smem + L.off_hrec + (sl & 1) * L.hrec_slot), t.fr.hrec + (size_t)t.x0 * STRIDE,
                             rec_bytes, bar(SF, slot));
            }
        }
    } else if (warp < kHWarps) {
        // ============================== horizontal pass ==============================
        int k = 0, sl = 0;
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x, ++sl) {
            const Strip t = load_strip<STRIDE>(frames, strips, s);
            const int* hrec = reinterpret_cast<const int*>(smem + L.off_hrec + (sl & 1) * L.hrec_slot);
            const int xa = t.x0 + (int)((int64_t)t.sw * warp / kHWarps);
            const int xb = t.x0 + (int)((int64_t)t.sw * (warp + 1) / kHWarps);
            for (int c = 0; c < t.n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                mbar_wait(bar(SF, slot), j & 1);
                if (k >= 2) mbar_wait(bar(HE, slot), (j - 1) & 1);
                const unsigned char* stage = smem + L.off_stage + slot * L.stage_slot;
                unsigned char* hring = smem + L.off_hring + slot * 3 * L.hplane + (KT - 1) * kPitch;   // rows after the carry area
                if (xa < xb) {
                    int xo = xa;
                    Rec<KT> hr;
                    const int* hp = hrec + (xo - t.x0) * STRIDE;
                    load_rec<KT, STRIDE>(hr, hp);
                    int p = hp[STRIDE - 2] & ~(kStepPx - 1);
                    uint32_t saddr = smem_u32(stage + lane * L.stage_pitch) + (uint32_t)(p - t.px0) * 3;
                    unsigned char* hdst = hring + lane * kPitch + (xo - t.x0);
                    int ring[3][RING];
#pragma unroll
                    for (int q = 0; q < RING; ++q) { ring[0][q] = ring[1][q] = ring[2][q] = 0; }
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
                        for (int jj = 0; jj < kStepPx; ++jj) {
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                const int b = 3 * jj + ch;
                                ring[ch][jj & (RING - 1)] = (int)__byte_perm(w[b >> 2], 0, 0x4440 + (b & 3));
                            }
                            while (hr.last == p + jj) {
                                int a0 = 1 << (VIS_PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
                                for (int tt = 0; tt < KT; ++tt) {
                                    const int q = (jj - tt) & (RING - 1);
                                    a0 += ring[0][q] * hr.k[tt];
                                    a1 += ring[1][q] * hr.k[tt];
                                    a2 += ring[2][q] * hr.k[tt];
                                }
                                hdst[0] = (unsigned char)clip8i(a0);
                                hdst[L.hplane] = (unsigned char)clip8i(a1);
                                hdst[2 * L.hplane] = (unsigned char)clip8i(a2);
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
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(bar(SE, slot));          // stage slot may be refilled
                    mbar_arrive(bar(HF, slot));          // H-ring slot is complete
                }
            }
        }
    } else if (warp < kHWarps + kVWarps) {
        // ============================== vertical pass (pull order) ==============================
        // Thread = 4 consecutive output columns of one channel.  For every output row whose tap window ends inside
        // this chunk it reads the KT tap rows straight from the H ring: the slot holds KT-1 carry rows (copies of the
        // previous chunk's last rows) directly in front of the 32 fresh rows, so a window never leaves the slot.
        constexpr int CARRY = KT - 1;
        const int v = tid - kHWarps * 32;
        int k = 0, nb = 0;
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x) {
            const Strip t = load_strip<STRIDE>(frames, strips, s);
            const int* vrec = t.fr.vrec;
            const int wpr = t.sw / 4;
            const bool v_active = v < 3 * wpr;
            const int vc = v_active ? v / wpr : 0;
            const int vwx = v_active ? v - vc * wpr : 0;
            int yo = t.y0, py = 0;
            auto stage_vrec = [&](int buf, int first) {          // records [first, first + kVCap) -> vrec ring
                int* dst = reinterpret_cast<int*>(smem + L.off_vrec + buf * L.vrec_slot);
                for (int i = v; i < kVCap * STRIDE / 4; i += kVThreads) {
                    const int rec = min(first + i / (STRIDE / 4), t.fr.dst_h);
                    cp_async16(smem_u32(dst + i * 4), vrec + (size_t)rec * STRIDE + (i % (STRIDE / 4)) * 4);
                }
            };
            stage_vrec(k & 1, yo);
            for (int c = 0; c < t.n_chunks; ++c, ++k) {
                const int slot = k & 1, j = k >> 1;
                const int r0 = t.r_first + c * kChunk;
                const int r_lim = r0 + kChunk;                   // windows ending before this row are complete
                cp_async_wait_all();
                named_bar_sync(1, kVThreads);                    // records + carry rows visible to all V warps
                mbar_wait(bar(HF, slot), j & 1);
                const int* vrec_s = reinterpret_cast<const int*>(smem + L.off_vrec + slot * L.vrec_slot);
                const int yo_base = yo;
                const unsigned char* hbase = smem + L.off_hring + slot * 3 * L.hplane + vc * L.hplane + vwx * 4
                                             + (CARRY - r0) * kPitch;       // + row * kPitch addresses input row `row`
                auto fetch = [&](Rec<KT>& r, int y) {
                    const int rel = y - yo_base;
                    if (rel < kVCap) load_rec<KT, STRIDE>(r, vrec_s + rel * STRIDE);
                    else load_rec<KT, STRIDE>(r, vrec + (size_t)min(y, t.fr.dst_h) * STRIDE);
                    if (y >= t.y1) r.last = INT_MAX;
                };
                auto emit = [&](const Rec<KT>& r) {
                    const unsigned char* hrow = hbase + r.last * kPitch;
                    uint32_t wv[KT];
#pragma unroll
                    for (int tt = 0; tt < KT; ++tt) wv[tt] = *reinterpret_cast<const uint32_t*>(hrow - tt * kPitch);
                    int acc[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[e] = 1 << (VIS_PRECISION_BITS - 1);
#pragma unroll
                    for (int tt = 0; tt < KT; ++tt) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[e] += (int)__byte_perm(wv[tt], 0, 0x4440 + e) * r.k[tt];
                    }
                    const uint32_t lo = __byte_perm(clip8i(acc[0]), clip8i(acc[1]), 0x0040);
                    const uint32_t hi = __byte_perm(clip8i(acc[2]), clip8i(acc[3]), 0x0040);
                    const int os = nb & 1;
                    if (py == 0 && nb >= 2) mbar_wait(bar(OE, os), ((nb >> 1) - 1) & 1);   // band tile free?
                    if (v_active)
                        *reinterpret_cast<uint32_t*>(smem + L.off_otile + os * 3 * kOPlane + vc * kOPlane +
                                                     py * kOPitch + vwx * 4) = __byte_perm(lo, hi, 0x5410);
                    ++yo;
                    if (++py == VIS_PATCH) {                      // band complete: hand it to the store warps
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar(OF, os));
                        ++nb;
                        py = 0;
                    }
                };
                Rec<KT> ra, rb;                                   // alternate: the next record is always in flight
                fetch(ra, yo);
                while (true) {
                    if (ra.last >= r_lim) break;
                    fetch(rb, yo + 1);
                    emit(ra);
                    if (rb.last >= r_lim) break;
                    fetch(ra, yo + 1);
                    emit(rb);
                }
                if (c + 1 < t.n_chunks) {
                    stage_vrec((k + 1) & 1, yo);
                    if (v_active) {                               // carry: last KT-1 rows -> front of the other slot
                        const unsigned char* src = smem + L.off_hring + slot * 3 * L.hplane + vc * L.hplane + vwx * 4 + kChunk * kPitch;
                        unsigned char* dst = smem + L.off_hring + (slot ^ 1) * 3 * L.hplane + vc * L.hplane + vwx * 4;
#pragma unroll
                        for (int i = 0; i < CARRY; ++i)
                            *reinterpret_cast<uint32_t*>(dst + i * kPitch) = *reinterpret_cast<const uint32_t*>(src + i * kPitch);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(HE, slot));       // H-ring slot consumed
            }
        }
    } else {
        // ============================== band store ==============================
        const int w = warp - (kHWarps + kVWarps);
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
        for (int s = blockIdx.x; s < n_strips; s += gridDim.x) {
            const Strip t = load_strip<STRIDE>(frames, strips, s);
            const int n_patches = t.sw / VIS_PATCH, gx0 = t.x0 / VIS_PATCH;
            const int half_gw = t.fr.dst_w / (2 * VIS_PATCH);
            float* const frame_out = pixel_values + (size_t)t.fr.row0 * VIS_ROW_FLOATS;
            for (int gy = t.y0 / VIS_PATCH; gy < t.y1 / VIS_PATCH; ++gy, ++nb) {
                const int os = nb & 1;
                mbar_wait(bar(OF, os), (nb >> 1) & 1);
                const unsigned char* otile = smem + L.off_otile + os * 3 * kOPlane;
                float* band = frame_out + (size_t)((gy >> 1) * half_gw * 4 + (gy & 1) * 2) * VIS_ROW_FLOATS;
                for (int g = w; g < n_patches; g += kSWarps) {
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

template <int KT, int RING, int STRIDE>
int launch_ws(const VisFrame* frames, const VisStrip* strips, int n_strips, const LayoutWS& L,
              const float* lut768, float* pixel_values, cudaStream_t st) {
    auto kern = k_fused_ws<KT, RING, STRIDE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total);
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_preprocess_fused(ws): cudaFuncSetAttribute");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_strips < sms ? n_strips : sms;
    kern<<<grid, kThreadsWS, L.total, st>>>(frames, strips, n_strips, L, lut768, pixel_values);
    return vis::check_launch("vis_preprocess_fused(ws)");
}

}  // namespace

namespace visf {

// host hooks used by vis_fused.cu (planning and dispatch)
int ws_layout_bytes(int span_bytes, int strip_w, int cls) {
    return make_layout_ws(span_bytes, strip_w, vis_record_stride(cls), cls - 1).total;
}
int ws_smem_max() { return kSmemMax; }

int ws_launch(int cls, const VisFrame* frames, const VisStrip* strips, int n_strips, int span_bytes, int strip_w,
              const float* lut768, float* pixel_values, cudaStream_t st) {
    const LayoutWS L = make_layout_ws(span_bytes, strip_w, vis_record_stride(cls), cls - 1);
    if (L.total > kSmemMax) {
        vis::set_error("vis_preprocess_fused(ws): %d bytes of shared memory needed", L.total);
        return VIS_E_UNSUPPORTED;
    }
    switch (cls) {
        case 6: return launch_ws<6, 8, 8>(frames, strips, n_strips, L, lut768, pixel_values, st);
        case 8: return launch_ws<8, 8, 12>(frames, strips, n_strips, L, lut768, pixel_values, st);
        default:
            vis::set_error("vis_preprocess_fused(ws): tap class %d has no warp-specialised instantiation", cls);
            return VIS_E_UNSUPPORTED;
    }
}

}  // namespace visf
