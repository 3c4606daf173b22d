/*
 * cvresize_oracle.c — CPU ORACLE (test infrastructure only, never on the product path).
 *
 * Plain-C restatement of what `cv2.resize(img, (new_w, target_h))` computes for 8-bit images in OpenCV 4.13
 * (modules/imgproc/src/resize.cpp; source not in the container, restated from the published algorithm and pinned
 * bit-exact against opencv-python-headless 4.13.0.92 by tests/test_oracle_compare.py).  The reference calls it with
 * the default interpolation (INTER_LINEAR) at utils/image_utils.py:641 (create_side_by_side_comparison).
 *
 *   cv::resize, INTER_LINEAR, CV_8U
 *     - dsize == ssize: plain copy;
 *     - scale_x == scale_y == 2 exactly: OpenCV switches to INTER_AREA's 2x2 fast path, (a + b + c + d + 2) >> 2;
 *     - otherwise the fixed-point bilinear resizer: per axis fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx),
 *       fx -= s, border clamps, 11-bit coefficients saturate_cast<short>(w * 2048) (round half to even);
 *       horizontal pass into int32 rows, vertical pass ((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2 >> 2.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

static inline short sat_short_from_float(float v) {
    long r = lrintf(v);                       /* cvRound: round half to even */
    return (short)(r < -32768 ? -32768 : r > 32767 ? 32767 : r);
}

/* per-axis table: ofs[d] = first source index, coef[2d], coef[2d+1] = 11-bit weights of ofs[d] and ofs[d]+1.
 * `clamp_weights`: the x axis folds the border handling into the table (cv: the xofs/alpha loop); the y axis keeps the
 * raw index and clamps rows when they are fetched. */
static void linear_table(int ssize, int dsize, double scale, int clamp_weights, int* ofs, short* coef) {
    for (int d = 0; d < dsize; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= s;
        if (clamp_weights) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        }
        ofs[d] = s;
        coef[2 * d] = sat_short_from_float((1.f - f) * 2048.f);
        coef[2 * d + 1] = sat_short_from_float(f * 2048.f);
    }
}

/* mode chosen by cv::resize for INTER_LINEAR: 0 copy, 1 area 2x2, 2 bilinear */
int ocv_resize_linear_mode(int sh, int sw, int dh, int dw) {
    if (sh == dh && sw == dw) return 0;
    const double inv_x = (double)dw / sw, inv_y = (double)dh / sh;
    const double scale_x = 1. / inv_x, scale_y = 1. / inv_y;
    const int ix = (int)lrint(scale_x), iy = (int)lrint(scale_y);     /* saturate_cast<int>(double) = cvRound */
    const int area_fast = fabs(scale_x - ix) < DBL_EPSILON && fabs(scale_y - iy) < DBL_EPSILON;
    if (area_fast && ix == 2 && iy == 2) return 1;
    return 2;
}

/* src [sh, sw, cn] uint8 with row pitch sstep -> dst [dh, dw, cn] with row pitch dstep */
int ocv_resize_linear_u8(const uint8_t* src, int sh, int sw, int64_t sstep, int cn,
                         uint8_t* dst, int dh, int dw, int64_t dstep) {
    if (sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0 || cn <= 0 || cn > 4) return -1;
    const int mode = ocv_resize_linear_mode(sh, sw, dh, dw);
    if (mode == 0) {
        for (int y = 0; y < dh; ++y) memcpy(dst + y * dstep, src + y * sstep, (size_t)dw * cn);
        return 0;
    }
    if (mode == 1) {
        for (int y = 0; y < dh; ++y) {
            const uint8_t* r0 = src + (int64_t)(2 * y) * sstep;
            const uint8_t* r1 = r0 + sstep;
            uint8_t* d = dst + y * dstep;
            for (int x = 0; x < dw; ++x)
                for (int c = 0; c < cn; ++c)
                    d[x * cn + c] = (uint8_t)((r0[2 * x * cn + c] + r0[(2 * x + 1) * cn + c] +
                                               r1[2 * x * cn + c] + r1[(2 * x + 1) * cn + c] + 2) >> 2);
        }
        return 0;
    }
    const double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    int* xofs = (int*)malloc(sizeof(int) * dw);
    int* yofs = (int*)malloc(sizeof(int) * dh);
    short* alpha = (short*)malloc(sizeof(short) * 2 * dw);
    short* beta = (short*)malloc(sizeof(short) * 2 * dh);
    int* row0 = (int*)malloc(sizeof(int) * dw * cn);
    int* row1 = (int*)malloc(sizeof(int) * dw * cn);
    if (!xofs || !yofs || !alpha || !beta || !row0 || !row1) return -2;
    linear_table(sw, dw, scale_x, 1, xofs, alpha);
    linear_table(sh, dh, scale_y, 0, yofs, beta);
    for (int y = 0; y < dh; ++y) {
        for (int k = 0; k < 2; ++k) {
            int sy = yofs[y] + k;
            sy = sy < 0 ? 0 : sy >= sh ? sh - 1 : sy;
            const uint8_t* s = src + (int64_t)sy * sstep;
            int* r = k ? row1 : row0;
            for (int x = 0; x < dw; ++x) {
                const int sx = xofs[x], sx1 = sx + 1 < sw ? sx + 1 : sx;     /* weight of sx1 is 0 at the border */
                for (int c = 0; c < cn; ++c)
                    r[x * cn + c] = s[sx * cn + c] * alpha[2 * x] + s[sx1 * cn + c] * alpha[2 * x + 1];
            }
        }
        const int b0 = beta[2 * y], b1 = beta[2 * y + 1];
        uint8_t* d = dst + y * dstep;
        for (int i = 0; i < dw * cn; ++i)
            d[i] = (uint8_t)((((b0 * (row0[i] >> 4)) >> 16) + ((b1 * (row1[i] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(yofs); free(alpha); free(beta); free(row0); free(row1);
    return 0;
}
