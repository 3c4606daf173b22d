/*
 * resample_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).
 *
 * Plain-C restatement of the arithmetic the reference reaches through third-party binaries that are
 * not vendored under /root/reference:
 *   - Pillow 12.2.0 (pin `pillow>=12.1.0`, reference pyproject.toml:24), src/libImaging/Resample.c:
 *       precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc,
 *       ImagingResampleVertical_8bpc, and the pass ordering of Image.resize (PIL:Image.py:2400-2440).
 *     Call sites in the reference: utils/image_utils.py:75 (resize_image, LANCZOS),
 *       src/agents/vlm_inspector.py:64 and src/agents/vlm_auditor.py:91 (thumbnail, LANCZOS).
 *   - transformers 5.5.0 Qwen2-VL PIL image processor (not a dependency of the reference; it is the
 *     server-side half of the Inspector/Auditor input path): tf:image_transforms.py:118-122 (rescale),
 *     :427-439 (normalize), tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:186-214 (patchify).
 *
 * Parity pin: the reference has no test or golden vector for this path, so this
 * oracle is pinned against (1) the installed Pillow / transformers binaries in tests/test_oracle_resample.py
 * and (2) the golden fixtures under tests/golden/ produced by tests/golden/make_goldens.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

#define ORC_LANCZOS 1
#define ORC_BICUBIC 3

/* ---- filters (Resample.c: bicubic_filter, sinc_filter, lanczos_filter) ---- */
static double orc_bicubic(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}
static double orc_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
static double orc_lanczos(double x) {
    if (-3.0 <= x && x < 3.0) return orc_sinc(x) * orc_sinc(x / 3);
    return 0.0;
}

static int orc_filter_support(int filter, double* support, double (**fn)(double)) {
    if (filter == ORC_BICUBIC) { *support = 2.0; *fn = orc_bicubic; return 0; }
    if (filter == ORC_LANCZOS) { *support = 3.0; *fn = orc_lanczos; return 0; }
    return -1;
}

/* ksize for (in -> out): Resample.c precompute_coeffs */
int orc_ksize(int in_size, int out_size, int filter) {
    double support; double (*fn)(double);
    if (orc_filter_support(filter, &support, &fn) || in_size <= 0 || out_size <= 0) return -1;
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    support = support * filterscale;
    return (int)ceil(support) * 2 + 1;
}

/* Resample.c precompute_coeffs + normalize_coeffs_8bpc.  k: out_size*ksize int32, bounds: out_size*2. */
int orc_coeffs(int in_size, int out_size, int filter, int32_t* k, int32_t* bounds) {
    double support; double (*fn)(double);
    if (orc_filter_support(filter, &support, &fn) || in_size <= 0 || out_size <= 0) return -1;
    double scale = (double)((float)in_size - 0.0f) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    support = support * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double* w = (double*)malloc(sizeof(double) * (size_t)ksize);
    if (!w) return -2;
    for (int xx = 0; xx < out_size; xx++) {
        double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            double v = fn((x + xmin - center + 0.5) * ss);
            w[x] = v;
            ww += v;
        }
        for (x = 0; x < xmax; x++) {
            if (ww != 0.0) w[x] /= ww;
        }
        for (; x < ksize; x++) w[x] = 0;
        for (x = 0; x < ksize; x++) {
            if (w[x] < 0) k[(size_t)xx * ksize + x] = (int)(-0.5 + w[x] * (1 << PRECISION_BITS));
            else          k[(size_t)xx * ksize + x] = (int)(0.5 + w[x] * (1 << PRECISION_BITS));
        }
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    free(w);
    return ksize;
}

static inline uint8_t orc_clip8(int32_t v) {
    v >>= PRECISION_BITS;              /* arithmetic shift, as clip8_lookups[in >> PRECISION_BITS] */
    if (v < 0) return 0;
    if (v > 255) return 255;
    return (uint8_t)v;
}

/* ImagingResampleHorizontal_8bpc on interleaved `ch`-channel uint8 rows */
int orc_resample_h(const uint8_t* src, int64_t src_pitch, int rows, int in_w, int ch,
                   uint8_t* dst, int64_t dst_pitch, int out_w, int filter) {
    int ksize = orc_ksize(in_w, out_w, filter);
    if (ksize < 0) return -1;
    int32_t* k = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_w * ksize);
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_w * 2);
    if (!k || !b) { free(k); free(b); return -2; }
    orc_coeffs(in_w, out_w, filter, k, b);
    for (int y = 0; y < rows; y++) {
        const uint8_t* s = src + (size_t)y * src_pitch;
        uint8_t* d = dst + (size_t)y * dst_pitch;
        for (int xx = 0; xx < out_w; xx++) {
            int xmin = b[xx * 2], cnt = b[xx * 2 + 1];
            const int32_t* kk = k + (size_t)xx * ksize;
            for (int c = 0; c < ch; c++) {
                uint32_t ss = 1u << (PRECISION_BITS - 1);   /* unsigned: int32 wrap semantics made explicit */
                for (int x = 0; x < cnt; x++) ss += (uint32_t)((int32_t)s[(size_t)(xmin + x) * ch + c] * kk[x]);
                d[(size_t)xx * ch + c] = orc_clip8((int32_t)ss);
            }
        }
    }
    free(k); free(b);
    return 0;
}

/* ImagingResampleVertical_8bpc: every byte column is resampled independently */
int orc_resample_v(const uint8_t* src, int64_t src_pitch, int in_h, int row_bytes,
                   uint8_t* dst, int64_t dst_pitch, int out_h, int filter) {
    int ksize = orc_ksize(in_h, out_h, filter);
    if (ksize < 0) return -1;
    int32_t* k = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_h * ksize);
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * (size_t)out_h * 2);
    if (!k || !b) { free(k); free(b); return -2; }
    orc_coeffs(in_h, out_h, filter, k, b);
    for (int yy = 0; yy < out_h; yy++) {
        int ymin = b[yy * 2], cnt = b[yy * 2 + 1];
        const int32_t* kk = k + (size_t)yy * ksize;
        uint8_t* d = dst + (size_t)yy * dst_pitch;
        for (int i = 0; i < row_bytes; i++) {
            uint32_t ss = 1u << (PRECISION_BITS - 1);
            for (int y = 0; y < cnt; y++) ss += (uint32_t)((int32_t)src[(size_t)(ymin + y) * src_pitch + i] * kk[y]);
            d[i] = orc_clip8((int32_t)ss);
        }
    }
    free(k); free(b);
    return 0;
}

/*
 * Image.resize((ow,oh), resample) for a `ch`-channel uint8 HWC image, reducing_gap=None
 * (PIL:Image.py:2400-2440 + ImagingResampleInner):
 *   same size -> copy; horizontal pass iff ow != w; vertical pass iff oh != h;
 *   order horizontal-then-vertical, EXCEPT the tall-image branch (h > 100*w and oh < h) which
 *   runs the vertical pass first (PIL:Image.py:2431-2435).
 * dst is a dense [oh][ow][ch] buffer.
 */
int orc_resize(const uint8_t* src, int h, int w, int ch, uint8_t* dst, int oh, int ow, int filter) {
    if (h <= 0 || w <= 0 || oh <= 0 || ow <= 0 || ch <= 0) return -1;
    int need_h = ow != w, need_v = oh != h;
    if (!need_h && !need_v) { memcpy(dst, src, (size_t)h * w * ch); return 0; }
    if (need_h && !need_v) return orc_resample_h(src, (int64_t)w * ch, h, w, ch, dst, (int64_t)ow * ch, ow, filter);
    if (!need_h && need_v) return orc_resample_v(src, (int64_t)w * ch, h, w * ch, dst, (int64_t)ow * ch, oh, filter);
    int rc;
    if ((int64_t)h > 100 * (int64_t)w && oh < h) {          /* tall image: vertical first */
        uint8_t* tmp = (uint8_t*)malloc((size_t)oh * w * ch);
        if (!tmp) return -2;
        rc = orc_resample_v(src, (int64_t)w * ch, h, w * ch, tmp, (int64_t)w * ch, oh, filter);
        if (!rc) rc = orc_resample_h(tmp, (int64_t)w * ch, oh, w, ch, dst, (int64_t)ow * ch, ow, filter);
        free(tmp);
        return rc;
    }
    uint8_t* tmp = (uint8_t*)malloc((size_t)h * ow * ch);
    if (!tmp) return -2;
    rc = orc_resample_h(src, (int64_t)w * ch, h, w, ch, tmp, (int64_t)ow * ch, ow, filter);
    if (!rc) rc = orc_resample_v(tmp, (int64_t)ow * ch, h, ow * ch, dst, (int64_t)ow * ch, oh, filter);
    free(tmp);
    return rc;
}

/*
 * rescale + normalize as a table (tf:image_transforms.py:118-122 and :427-439):
 *   rescaled = (float32)((double)v * rescale)       — `image.astype(np.float64) * scale` then `.astype(np.float32)`
 *   out      = (rescaled - (float32)mean[c]) / (float32)std[c]     in float32
 */
void orc_lut(const float mean[3], const float stdv[3], double rescale, float* lut768) {
    for (int v = 0; v < 256; v++)
        for (int c = 0; c < 3; c++) {
            volatile float r = (float)((double)v * rescale);
            volatile float d = r - mean[c];
            lut768[v * 3 + c] = d / stdv[c];
        }
}

/*
 * Patch layout of tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:186-214 for ONE frame
 * (grid_t = 1, frame repeated to temporal_patch_size = 2), patch 14, merge 2:
 *   row = ((bh*(gw/2) + bw)*2 + mh)*2 + mw ;  col = ((c*2 + t)*14 + py)*14 + px
 * img: resized RGB uint8 HWC [h][w][3], h and w multiples of 28.  out: [(h/14)*(w/14)][1176] float32.
 */
int orc_patchify(const uint8_t* img, int h, int w, const float* lut768, float* out) {
    if (h % 28 || w % 28) return -1;
    int gh = h / 14, gw = w / 14;
    for (int y = 0; y < h; y++) {
        int gy = y / 14, py = y % 14, bh = gy / 2, mh = gy % 2;
        for (int x = 0; x < w; x++) {
            int gx = x / 14, px = x % 14, bw = gx / 2, mw = gx % 2;
            size_t row = (((size_t)bh * (gw / 2) + bw) * 2 + mh) * 2 + mw;
            for (int c = 0; c < 3; c++) {
                float v = lut768[img[((size_t)y * w + x) * 3 + c] * 3 + c];
                for (int t = 0; t < 2; t++)
                    out[row * 1176 + ((size_t)(c * 2 + t) * 14 + py) * 14 + px] = v;
            }
        }
    }
    (void)gh;
    return 0;
}

/* =====================================================================================================
 * Image.thumbnail / Image.resize with reducing_gap (PIL:Image.py:2400-2440, :2831-2915) for frames that are
 * >= 4x larger than the target: Image.reduce (libImaging/Reduce.c) followed by a resample over a fractional
 * source box (Resample.c: precompute_coeffs with in0 / in1, ImagingResample's ybox window).
 * ===================================================================================================== */

/* Resample.c precompute_coeffs with a source box [in0, in1) given as FLOATS (ImagingResample takes float box[4]). */
int orc_ksize_box(float in0, float in1, int out_size, int filter) {
    double support; double (*fn)(double);
    if (orc_filter_support(filter, &support, &fn) || out_size <= 0 || !(in1 > in0)) return -1;
    double scale = (double)(in1 - in0) / out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(support * filterscale) * 2 + 1;
}

int orc_coeffs_box(int in_size, float in0, float in1, int out_size, int filter, int32_t* k, int32_t* bounds) {
    double support; double (*fn)(double);
    if (orc_filter_support(filter, &support, &fn) || in_size <= 0 || out_size <= 0 || !(in1 > in0)) return -1;
    double scale = (double)(in1 - in0) / out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    support = support * filterscale;
    int ksize = (int)ceil(support) * 2 + 1;
    double* w = (double*)malloc(sizeof(double) * (size_t)ksize);
    if (!w) return -2;
    for (int xx = 0; xx < out_size; xx++) {
        double center = in0 + (xx + 0.5) * scale;
        double ww = 0.0;
        double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        int x;
        for (x = 0; x < xmax; x++) {
            double v = fn((x + xmin - center + 0.5) * ss);
            w[x] = v;
            ww += v;
        }
        for (x = 0; x < xmax; x++) {
            if (ww != 0.0) w[x] /= ww;
        }
        for (; x < ksize; x++) w[x] = 0;
        for (x = 0; x < ksize; x++) {
            if (w[x] < 0) k[(size_t)xx * ksize + x] = (int)(-0.5 + w[x] * (1 << PRECISION_BITS));
            else          k[(size_t)xx * ksize + x] = (int)(0.5 + w[x] * (1 << PRECISION_BITS));
        }
        bounds[xx * 2 + 0] = xmin;
        bounds[xx * 2 + 1] = xmax;
    }
    free(w);
    return ksize;
}

/* ImagingResample (Resample.c) for a `ch`-channel uint8 HWC image with a float source box (x0, y0, x1, y1):
 * horizontal pass over the rows the vertical pass will read (ybox_first .. ybox_last), then the vertical pass. */
int orc_resize_box(const uint8_t* src, int h, int w, int ch, uint8_t* dst, int oh, int ow, int filter, const float* box) {
    if (h <= 0 || w <= 0 || oh <= 0 || ow <= 0 || ch <= 0 || !box) return -1;
    const int need_h = ow != w || box[0] != 0.f || box[2] != (float)ow;
    const int need_v = oh != h || box[1] != 0.f || box[3] != (float)oh;
    const int ksh = orc_ksize_box(box[0], box[2], ow, filter), ksv = orc_ksize_box(box[1], box[3], oh, filter);
    if (ksh < 0 || ksv < 0) return -1;
    int32_t* kh = (int32_t*)malloc(sizeof(int32_t) * (size_t)ow * ksh);
    int32_t* bh = (int32_t*)malloc(sizeof(int32_t) * (size_t)ow * 2);
    int32_t* kv = (int32_t*)malloc(sizeof(int32_t) * (size_t)oh * ksv);
    int32_t* bv = (int32_t*)malloc(sizeof(int32_t) * (size_t)oh * 2);
    if (!kh || !bh || !kv || !bv) return -2;
    orc_coeffs_box(w, box[0], box[2], ow, filter, kh, bh);
    orc_coeffs_box(h, box[1], box[3], oh, filter, kv, bv);
    const int yfirst = bv[0], ylast = bv[oh * 2 - 2] + bv[oh * 2 - 1];
    const uint8_t* cur = src;
    int cur_h = h, cur_w = w, row_off = 0;
    uint8_t* tmp = NULL;
    if (need_h) {
        const int rows = ylast - yfirst;
        tmp = (uint8_t*)malloc((size_t)rows * ow * ch);
        if (!tmp) return -2;
        for (int y = 0; y < rows; y++) {
            const uint8_t* s = src + (size_t)(y + yfirst) * w * ch;
            uint8_t* d = tmp + (size_t)y * ow * ch;
            for (int xx = 0; xx < ow; xx++) {
                const int xmin = bh[xx * 2], cnt = bh[xx * 2 + 1];
                const int32_t* kk = kh + (size_t)xx * ksh;
                for (int c = 0; c < ch; c++) {
                    uint32_t ss = 1u << (PRECISION_BITS - 1);
                    for (int x = 0; x < cnt; x++) ss += (uint32_t)((int32_t)s[(size_t)(xmin + x) * ch + c] * kk[x]);
                    d[(size_t)xx * ch + c] = orc_clip8((int32_t)ss);
                }
            }
        }
        cur = tmp; cur_h = rows; cur_w = ow; row_off = yfirst;
    }
    if (need_v) {
        for (int yy = 0; yy < oh; yy++) {
            const int ymin = bv[yy * 2] - row_off, cnt = bv[yy * 2 + 1];
            const int32_t* kk = kv + (size_t)yy * ksv;
            uint8_t* d = dst + (size_t)yy * cur_w * ch;
            for (int i = 0; i < cur_w * ch; i++) {
                uint32_t ss = 1u << (PRECISION_BITS - 1);
                for (int y = 0; y < cnt; y++) ss += (uint32_t)((int32_t)cur[(size_t)(ymin + y) * cur_w * ch + i] * kk[y]);
                d[i] = orc_clip8((int32_t)ss);
            }
        }
    } else {
        memcpy(dst, cur, (size_t)cur_h * cur_w * ch);
    }
    (void)cur_h;
    free(tmp); free(kh); free(bh); free(kv); free(bv);
    return 0;
}

/* Reduce.c division_UINT32: float division, truncated */
static inline uint32_t orc_division_u32(int divider, int result_bits) {
    uint32_t max_dividend = (uint32_t)(1 << result_bits) * (uint32_t)divider;
    float max_int = (1 << 30) * 4.0f;
    return (uint32_t)(max_int / (float)max_dividend);
}

/* Image.reduce((fx, fy), box) for a `ch`-channel uint8 HWC image (Reduce.c: ImagingReduce + ImagingReduceCorners):
 * every output sample is ((sum + n/2) * division_UINT32(n, 8)) >> 24 over the n source pixels of its (possibly
 * clipped, at the right / bottom edge of the box) fx x fy cell.  dst: dense [ceil(bh/fy)][ceil(bw/fx)][ch]. */
int orc_reduce(const uint8_t* src, int h, int w, int ch, int fx, int fy, const int* box, uint8_t* dst) {
    if (h <= 0 || w <= 0 || ch <= 0 || fx < 1 || fy < 1 || !box) return -1;
    const int x0 = box[0], y0 = box[1], x1 = box[2], y1 = box[3];
    if (x0 < 0 || y0 < 0 || x1 > w || y1 > h || x1 <= x0 || y1 <= y0) return -1;
    const int ow = (x1 - x0 + fx - 1) / fx, oh = (y1 - y0 + fy - 1) / fy;
    for (int oy = 0; oy < oh; oy++) {
        const int ys = y0 + oy * fy, ye = ys + fy < y1 ? ys + fy : y1;
        for (int ox = 0; ox < ow; ox++) {
            const int xs = x0 + ox * fx, xe = xs + fx < x1 ? xs + fx : x1;
            const int n = (ye - ys) * (xe - xs);
            const uint32_t mult = orc_division_u32(n, 8), amend = (uint32_t)n / 2;
            for (int c = 0; c < ch; c++) {
                uint32_t ss = amend;
                for (int y = ys; y < ye; y++)
                    for (int x = xs; x < xe; x++) ss += src[((size_t)y * w + x) * ch + c];
                dst[((size_t)oy * ow + ox) * ch + c] = (uint8_t)((ss * mult) >> 24);
            }
        }
    }
    return 0;
}

/* Image.resize(size, NEAREST, box) — what Pillow runs for palette ("P") and bilevel ("1") images whatever filter the
 * caller names (PIL:Image.py:2396-2397): _imaging.c _resize builds the affine a = {(x1-x0)/out, 0, x0, 0, (y1-y0)/out, y0}
 * and Geometry.c ImagingScaleAffine walks it with an ACCUMULATED double (xo = a[2] + a[0]*0.5; xo += a[0]),
 * COORD(v) = v < 0 ? -1 : (int)v; samples outside the image are 0.  tab[d] = source index or -1. */
int orc_nearest_table(int in_size, float in0, float in1, int out_size, int32_t* tab) {
    if (in_size <= 0 || out_size <= 0 || !tab) return -1;
    const double a = (double)(in1 - in0) / out_size;
    double xo = (double)in0 + a * 0.5;
    for (int x = 0; x < out_size; x++) {
        const int xin = xo < 0.0 ? -1 : (int)xo;
        tab[x] = (xin >= 0 && xin < in_size) ? xin : -1;
        xo += a;
    }
    return 0;
}
