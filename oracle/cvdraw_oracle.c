/*
 * cvdraw_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never imported by the product path).
 *
 * Sequential plain-C restatement of the OpenCV drawing routines that the reference's
 * draw_bounding_boxes (utils/image_utils.py:259-313) reaches through cv2:
 *     cv2.rectangle(.., 2, LINE_AA)  cv2.line(.., 2, LINE_AA)  cv2.circle(.., -1)  cv2.circle(.., 3)
 *     cv2.getTextSize / cv2.putText(FONT_HERSHEY_SIMPLEX)
 * OpenCV is an un-vendored dependency (opencv-python>=4.11.0.86, reference pyproject.toml:23; the build
 * image has opencv-python-headless 4.13.0.92).  The algorithm restated is modules/imgproc/src/drawing.cpp:
 * clipLine, Line2, LineAA, FillConvexPoly, Circle, ellipse2Poly, EllipseEx, ThickLine, PolyLine,
 * rectangle, line, circle, putText, getTextSize (8-bit 3-channel images only).
 *
 * Parity pin: tests/test_oracle_cvdraw.py compares every routine with the installed cv2 binary on
 * randomised geometry; tests/golden/ holds overlays produced by the real reference function.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)
#define LINE_8 8
#define LINE_AA 16

typedef struct { int64_t x, y; } P2l;
typedef struct { uint8_t* data; int h, w; int64_t step; } Img;
typedef struct { uint8_t c[3]; } Col;

static inline int cv_round(double v) { return (int)lrint(v); }   /* round-half-even (default FP mode) */

/* ---------------- tables (drawing.cpp: icvSlopeCorrTable, FilterTable, SinTable) ---------------- */
static const uint8_t SlopeCorrTable[32] = {
    181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
    203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};
static const int FilterTable[64] = {
    168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
    254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
    158, 149, 140, 131, 122, 114, 105, 97, 89, 82, 75, 68, 62, 56, 50, 45,
    40, 36, 32, 28, 25, 22, 19, 16, 14, 12, 11, 9, 8, 7, 5, 5};

static float SinTable[451];
static int sin_ready = 0;
/* the table in the binary holds 7-decimal literals (0.0174524f, ...): float32(round(sin(d deg), 7)) */
static void init_sin(void) {
    if (sin_ready) return;
    for (int d = 0; d <= 450; d++) {
        char buf[32];
        double s = sin(d * M_PI / 180.0);
        snprintf(buf, sizeof buf, "%.7f", s);
        SinTable[d] = strtof(buf, NULL);
    }
    /* exact landmarks as written in the source table */
    SinTable[0] = 0.f; SinTable[90] = 1.f; SinTable[180] = 0.f; SinTable[270] = -1.f;
    SinTable[360] = 0.f; SinTable[450] = 1.f;
    sin_ready = 1;
}

/* ---------------- clipLine (int64 variant on a scaled size) ---------------- */
static int clip_line(int64_t width, int64_t height, P2l* pt1, P2l* pt2) {
    int c1, c2;
    int64_t right = width - 1, bottom = height - 1;
    if (width <= 0 || height <= 0) return 0;
    int64_t x1 = pt1->x, y1 = pt1->y, x2 = pt2->x, y2 = pt2->y;
    c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        int64_t a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (int64_t)((double)(a - y1) * (x2 - x1) / (y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (int64_t)((double)(a - y2) * (x2 - x1) / (y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (int64_t)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (int64_t)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    pt1->x = x1; pt1->y = y1; pt2->x = x2; pt2->y = y2;
    return (c1 | c2) == 0;
}

static inline void put3(Img* im, int x, int y, const Col* col) {
    if (x >= 0 && x < im->w && y >= 0 && y < im->h) {
        uint8_t* p = im->data + (int64_t)y * im->step + (int64_t)x * 3;
        p[0] = col->c[0]; p[1] = col->c[1]; p[2] = col->c[2];
    }
}

static inline void hline(Img* im, int y, int x1, int x2, const Col* col) {   /* caller clipped y */
    uint8_t* row = im->data + (int64_t)y * im->step;
    for (int x = x1; x <= x2; x++) { row[x * 3] = col->c[0]; row[x * 3 + 1] = col->c[1]; row[x * 3 + 2] = col->c[2]; }
}

/* ---------------- Line2: 8-connected line with 16.16 end points ---------------- */
static void line2(Img* im, P2l pt1, P2l pt2, const Col* col) {
    int64_t dx, dy, ax, ay, i, j, x_step, y_step;
    int ecount;
    if (!clip_line((int64_t)im->w << XY_SHIFT, (int64_t)im->h << XY_SHIFT, &pt1, &pt2)) return;
    dx = pt2.x - pt1.x;
    dy = pt2.y - pt1.y;
    j = dx < 0 ? -1 : 0;
    ax = (dx ^ j) - j;
    i = dy < 0 ? -1 : 0;
    ay = (dy ^ i) - i;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { P2l t = pt1; pt1 = pt2; pt2 = t; }
        x_step = XY_ONE;
        y_step = (dy * XY_ONE) / (ax | 1);
        ecount = (int)((pt2.x - pt1.x) >> XY_SHIFT);
    } else {
        dx = (dx ^ i) - i;
        if (i) { P2l t = pt1; pt1 = pt2; pt2 = t; }
        x_step = (dx * XY_ONE) / (ay | 1);
        y_step = XY_ONE;
        ecount = (int)((pt2.y - pt1.y) >> XY_SHIFT);
    }
    pt1.x += (XY_ONE >> 1);
    pt1.y += (XY_ONE >> 1);
    put3(im, (int)((pt2.x + (XY_ONE >> 1)) >> XY_SHIFT), (int)((pt2.y + (XY_ONE >> 1)) >> XY_SHIFT), col);
    if (ax > ay) {
        pt1.x >>= XY_SHIFT;
        while (ecount >= 0) {
            put3(im, (int)pt1.x, (int)(pt1.y >> XY_SHIFT), col);
            pt1.x++;
            pt1.y += y_step;
            ecount--;
        }
    } else {
        pt1.y >>= XY_SHIFT;
        while (ecount >= 0) {
            put3(im, (int)(pt1.x >> XY_SHIFT), (int)pt1.y, col);
            pt1.x += x_step;
            pt1.y++;
            ecount--;
        }
    }
    (void)x_step; (void)y_step;
}

/* ---------------- LineAA: Wu-style antialiased line, 8-bit alpha, blend applied twice ---------------- */
static inline void blend3(Img* im, int x, int y, const Col* col, int a) {
    uint8_t* p = im->data + (int64_t)y * im->step + (int64_t)x * 3;
    for (int k = 0; k < 3; k++) {
        int c = p[k], cc = col->c[k];
        c += ((cc - c) * a + 127) >> 8;
        c += ((cc - c) * a + 127) >> 8;
        p[k] = (uint8_t)c;
    }
}

static void line_aa(Img* im, P2l pt1, P2l pt2, const Col* col) {
    int64_t dx, dy, ax, ay, x_step, y_step, i, j;
    int ecount, scount = 0, slope;
    int ep_table[9];
    if (!clip_line((int64_t)im->w << XY_SHIFT, (int64_t)im->h << XY_SHIFT, &pt1, &pt2)) return;
    dx = pt2.x - pt1.x;
    dy = pt2.y - pt1.y;
    j = dx < 0 ? -1 : 0;
    ax = (dx ^ j) - j;
    i = dy < 0 ? -1 : 0;
    ay = (dy ^ i) - i;
    if (ax > ay) {
        dy = (dy ^ j) - j;
        if (j) { P2l t = pt1; pt1 = pt2; pt2 = t; }
        x_step = XY_ONE;
        y_step = (dy * XY_ONE) / (ax | 1);
        pt2.x += XY_ONE;
        ecount = (int)((pt2.x >> XY_SHIFT) - (pt1.x >> XY_SHIFT));
        j = -(pt1.x & (XY_ONE - 1));
        pt1.y += ((y_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((y_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (y_step < 0 ? 0x3f : 0);
        i = (pt1.x >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.x >> (XY_SHIFT - 7)) & 0x78;
    } else {
        dx = (dx ^ i) - i;
        if (i) { P2l t = pt1; pt1 = pt2; pt2 = t; }
        x_step = (dx * XY_ONE) / (ay | 1);
        y_step = XY_ONE;
        pt2.y += XY_ONE;
        ecount = (int)((pt2.y >> XY_SHIFT) - (pt1.y >> XY_SHIFT));
        j = -(pt1.y & (XY_ONE - 1));
        pt1.x += ((x_step * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((x_step >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (x_step < 0 ? 0x3f : 0);
        i = (pt1.y >> (XY_SHIFT - 7)) & 0x78;
        j = (pt2.y >> (XY_SHIFT - 7)) & 0x78;
    }
    slope = (slope & 0x20) ? 0x100 : SlopeCorrTable[slope];
    {
        int t0 = slope << 7;
        int t1 = ((0x78 - (int)i) | 4) * slope;
        int t2 = ((int)j | 4) * slope;
        ep_table[0] = 0;
        ep_table[8] = slope;
        ep_table[1] = ep_table[3] = (int)((((((j - i) & 0x78) | 4) * slope) >> 8) & 0x1ff);
        ep_table[2] = (t1 >> 8) & 0x1ff;
        ep_table[4] = (int)((((((j - i) + 0x80) | 4) * slope) >> 8) & 0x1ff);
        ep_table[5] = ((t1 + t0) >> 8) & 0x1ff;
        ep_table[6] = (t2 >> 8) & 0x1ff;
        ep_table[7] = ((t2 + t0) >> 8) & 0x1ff;
    }
    if (ax > ay) {
        int x = (int)(pt1.x >> XY_SHIFT);
        for (; ecount >= 0; x++, pt1.y += y_step, scount++, ecount--) {
            if ((unsigned)x >= (unsigned)im->w) continue;
            int y = (int)((pt1.y >> XY_SHIFT) - 1);
            int ep_corr = ep_table[(((scount >= 2) + 1) & (scount | 2)) * 3 + (((ecount >= 2) + 1) & (ecount | 2))];
            int a, dist = (int)((pt1.y >> (XY_SHIFT - 5)) & 31);
            a = (ep_corr * FilterTable[dist + 32] >> 8) & 0xff;
            if ((unsigned)y < (unsigned)im->h) blend3(im, x, y, col, a);
            a = (ep_corr * FilterTable[dist] >> 8) & 0xff;
            if ((unsigned)(y + 1) < (unsigned)im->h) blend3(im, x, y + 1, col, a);
            a = (ep_corr * FilterTable[63 - dist] >> 8) & 0xff;
            if ((unsigned)(y + 2) < (unsigned)im->h) blend3(im, x, y + 2, col, a);
        }
    } else {
        int y = (int)(pt1.y >> XY_SHIFT);
        for (; ecount >= 0; y++, pt1.x += x_step, scount++, ecount--) {
            if ((unsigned)y >= (unsigned)im->h) continue;
            int x = (int)((pt1.x >> XY_SHIFT) - 1);
            int ep_corr = ep_table[(((scount >= 2) + 1) & (scount | 2)) * 3 + (((ecount >= 2) + 1) & (ecount | 2))];
            int a, dist = (int)((pt1.x >> (XY_SHIFT - 5)) & 31);
            a = (ep_corr * FilterTable[dist + 32] >> 8) & 0xff;
            if ((unsigned)x < (unsigned)im->w) blend3(im, x, y, col, a);
            a = (ep_corr * FilterTable[dist] >> 8) & 0xff;
            if ((unsigned)(x + 1) < (unsigned)im->w) blend3(im, x + 1, y, col, a);
            a = (ep_corr * FilterTable[63 - dist] >> 8) & 0xff;
            if ((unsigned)(x + 2) < (unsigned)im->w) blend3(im, x + 2, y, col, a);
        }
    }
    (void)x_step; (void)y_step;
}

/* ---------------- FillConvexPoly (vertices already 16.16; shift == XY_SHIFT) ---------------- */
static void fill_convex_poly(Img* im, const P2l* v, int npts, const Col* col, int line_type) {
    struct { int idx, di; int64_t x, dx; int ye; } edge[2];
    const int shift = XY_SHIFT;
    int delta = 1 << shift >> 1;
    int i, y, imin = 0;
    int edges = npts;
    int64_t xmin, xmax, ymin, ymax;
    int delta1, delta2;
    P2l p0;
    if (line_type < LINE_AA) delta1 = delta2 = XY_ONE >> 1;
    else { delta1 = XY_ONE - 1; delta2 = 0; }
    p0 = v[npts - 1];
    xmin = xmax = v[0].x;
    ymin = ymax = v[0].y;
    for (i = 0; i < npts; i++) {
        P2l p = v[i];
        if (p.y < ymin) { ymin = p.y; imin = i; }
        if (p.y > ymax) ymax = p.y;
        if (p.x > xmax) xmax = p.x;
        if (p.x < xmin) xmin = p.x;
        if (line_type <= 8) line2(im, p0, p, col);
        else line_aa(im, p0, p, col);
        p0 = p;
    }
    xmin = (xmin + delta) >> shift;
    xmax = (xmax + delta) >> shift;
    ymin = (ymin + delta) >> shift;
    ymax = (ymax + delta) >> shift;
    if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= im->w || (int)ymin >= im->h) return;
    if (ymax > im->h - 1) ymax = im->h - 1;
    edge[0].idx = edge[1].idx = imin;
    edge[0].ye = edge[1].ye = y = (int)ymin;
    edge[0].di = 1;
    edge[1].di = npts - 1;
    edge[0].x = edge[1].x = -XY_ONE;
    edge[0].dx = edge[1].dx = 0;
    do {
        if (line_type < LINE_AA || y < (int)ymax || y == (int)ymin) {
            for (i = 0; i < 2; i++) {
                if (y >= edge[i].ye) {
                    int idx0 = edge[i].idx, di = edge[i].di;
                    int idx = idx0 + di;
                    if (idx >= npts) idx -= npts;
                    int ty = 0;
                    for (; edges-- > 0;) {
                        ty = (int)((v[idx].y + delta) >> shift);
                        if (ty > y) {
                            int64_t xs = v[idx0].x;
                            int64_t xe = v[idx].x;
                            edge[i].ye = ty;
                            edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                            edge[i].x = xs;
                            edge[i].idx = idx;
                            break;
                        }
                        idx0 = idx;
                        idx += di;
                        if (idx >= npts) idx -= npts;
                    }
                }
            }
        }
        if (edges < 0) break;
        if (y >= 0) {
            int left = 0, right = 1;
            if (edge[0].x > edge[1].x) { left = 1; right = 0; }
            int xx1 = (int)((edge[left].x + delta1) >> XY_SHIFT);
            int xx2 = (int)((edge[right].x + delta2) >> XY_SHIFT);
            if (xx2 >= 0 && xx1 < im->w) {
                if (xx1 < 0) xx1 = 0;
                if (xx2 >= im->w) xx2 = im->w - 1;
                hline(im, y, xx1, xx2, col);
            }
        }
        edge[0].x += edge[0].dx;
        edge[1].x += edge[1].dx;
    } while (++y <= (int)ymax);
}

/* ---------------- Circle (midpoint), filled variant only ---------------- */
static void circle_fill(Img* im, int cx, int cy, int radius, const Col* col) {
    int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
    while (dx >= dy) {
        int mask;
        int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
        int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
        if (x11 < im->w && x12 >= 0 && y21 < im->h && y22 >= 0) {
            if (x11 < 0) x11 = 0;
            if (x12 > im->w - 1) x12 = im->w - 1;
            if ((unsigned)y11 < (unsigned)im->h) hline(im, y11, x11, x12, col);
            if ((unsigned)y12 < (unsigned)im->h) hline(im, y12, x11, x12, col);
            if (x21 < im->w && x22 >= 0) {
                if (x21 < 0) x21 = 0;
                if (x22 > im->w - 1) x22 = im->w - 1;
                if ((unsigned)y21 < (unsigned)im->h) hline(im, y21, x21, x22, col);
                if ((unsigned)y22 < (unsigned)im->h) hline(im, y22, x21, x22, col);
            }
        }
        dy++;
        err += plus;
        plus += 2;
        mask = (err <= 0) - 1;
        err -= minus & mask;
        dx += mask;
        minus -= mask & 2;
    }
}

static void thick_line(Img* im, P2l p0, P2l p1, const Col* col, int thickness, int line_type, int flags, int shift);

static void poly_line(Img* im, const P2l* v, int count, int is_closed, const Col* col, int thickness,
                      int line_type, int shift) {
    if (!v || count <= 0) return;
    int i = is_closed ? count - 1 : 0;
    int flags = 2 + !is_closed;
    P2l p0 = v[i];
    for (i = !is_closed; i < count; i++) {
        P2l p = v[i];
        thick_line(im, p0, p, col, thickness, line_type, flags, shift);
        p0 = p;
        flags = 2;
    }
}

/* ---------------- ellipse2Poly (angle 0, arc 0..360) + EllipseEx ---------------- */
static void ellipse_ex(Img* im, P2l center, int64_t aw, int64_t ah, const Col* col, int thickness, int line_type) {
    init_sin();
    if (aw < 0) aw = -aw;
    if (ah < 0) ah = -ah;
    int delta = (int)(((aw > ah ? aw : ah) + (XY_ONE >> 1)) >> XY_SHIFT);
    delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
    P2l v[400];
    int n = 0;
    P2l prev = {-1, -1};
    float alpha = SinTable[450], beta = SinTable[0];      /* sincos(0) */
    double cxd = (double)center.x, cyd = (double)center.y, awd = (double)aw, ahd = (double)ah;
    int npts2 = 0;
    for (int a = 0; a < 360 + delta; a += delta) {
        int angle = a;
        if (angle > 360) angle = 360;
        double x = awd * SinTable[450 - angle];
        double y = ahd * SinTable[angle];
        double px = cxd + x * alpha - y * beta;
        double py = cyd + x * beta + y * alpha;
        npts2++;
        P2l pt;
        pt.x = (int64_t)cv_round(px / (double)XY_ONE) << XY_SHIFT;
        pt.y = (int64_t)cv_round(py / (double)XY_ONE) << XY_SHIFT;
        pt.x += cv_round(px - pt.x);
        pt.y += cv_round(py - pt.y);
        if (pt.x != prev.x || pt.y != prev.y) { v[n++] = pt; prev = pt; }
    }
    if (n <= 1) { v[0] = center; v[1] = center; n = 2; }
    if (thickness >= 0) poly_line(im, v, n, 0, col, thickness, line_type, XY_SHIFT);
    else fill_convex_poly(im, v, n, col, line_type);
    (void)npts2;
}

/* ---------------- ThickLine (thickness > 1 only: all the reference ever uses) ---------------- */
static void thick_line(Img* im, P2l p0, P2l p1, const Col* col, int thickness, int line_type, int flags, int shift) {
    const double INV_XY_ONE = 1. / (double)XY_ONE;
    p0.x <<= XY_SHIFT - shift;
    p0.y <<= XY_SHIFT - shift;
    p1.x <<= XY_SHIFT - shift;
    p1.y <<= XY_SHIFT - shift;
    if (thickness <= 1) {
        if (line_type < LINE_AA) line2(im, p0, p1, col);   /* shift>0 path; shift==0 LINE_8 uses LineIterator (unused here) */
        else line_aa(im, p0, p1, col);
        return;
    }
    P2l pt[4], dp = {0, 0};
    double dx = (p0.x - p1.x) * INV_XY_ONE, dy = (p1.y - p0.y) * INV_XY_ONE;
    double r = dx * dx + dy * dy;
    int i, odd = thickness & 1;
    thickness <<= XY_SHIFT - 1;
    if (fabs(r) > DBL_EPSILON) {
        r = (thickness + odd * XY_ONE * 0.5) / sqrt(r);
        dp.x = cv_round(dy * r);
        dp.y = cv_round(dx * r);
        pt[0].x = p0.x + dp.x; pt[0].y = p0.y + dp.y;
        pt[1].x = p0.x - dp.x; pt[1].y = p0.y - dp.y;
        pt[2].x = p1.x - dp.x; pt[2].y = p1.y - dp.y;
        pt[3].x = p1.x + dp.x; pt[3].y = p1.y + dp.y;
        fill_convex_poly(im, pt, 4, col, line_type);
    }
    for (i = 0; i < 2; i++) {
        if (flags & (i + 1)) {
            if (line_type < LINE_AA) {
                int cx = (int)((p0.x + (XY_ONE >> 1)) >> XY_SHIFT);
                int cy = (int)((p0.y + (XY_ONE >> 1)) >> XY_SHIFT);
                circle_fill(im, cx, cy, (thickness + (XY_ONE >> 1)) >> XY_SHIFT, col);
            } else {
                ellipse_ex(im, p0, thickness, thickness, col, -1, line_type);
            }
        }
        p0 = p1;
    }
}

/* ---------------- Hershey simplex glyphs, printable ASCII 32..126 ----------------
 * cv: HersheySimplex[] -> g_HersheyGlyphs[] (drawing.cpp).  The strings are font DATA of the installed OpenCV 4.13
 * binary; tests/golden/find_glyphs.py recovers them by rendering every candidate string found in the binary through
 * cv2.polylines with putText's own geometry and keeping the one that reproduces cv2.putText at four scale /
 * thickness / line-type settings. */
static const char* const kSimplexGlyphs[95] = {
    "JZ",  /*   */
    "MWRFRT RYQZR[SZRY",  /* ! */
    "JZNFNM VFVM",  /* quote */
    "G]OFOb UFUb JQZQ JWZW",  /* # */
    "H\\PBP_ TBT_ YIWGTFPFMGKIKKLMMNOOUQWRXSYUYXWZT[P[MZKX",  /* $ */
    "F^[FYGVHSHPGNFLFJGIIIKKMMMOLPJPHNF [FI[ YTWTUUTWTYV[X[ZZ[X[VYT",  /* % */
    "E_\\O\\N[MZMYNXPVUTXRZP[L[JZIYHWHUISJRQNRMSKSIRGPFNGMIMKNNPQUXWZY[[[\\Z\\Y",  /* & */
    "NVRFRM",  /* ' */
    "KYVBTDRGPKOPOTPYR]T`Vb",  /* ( */
    "KYNBPDRGTKUPUTTYR]P`Nb",  /* ) */
    "JZRLRX MOWU WOMU",
    "E_RIR[ IR[R",  /* + */
    "MWSZR[QZRYSZS\\R^Q_",  /* , */
    "E_IR[R",  /* - */
    "MWRYQZR[SZRY",  /* . */
    "G][BIb",
    "H\\QFNGLJKOKRLWNZQ[S[VZXWYRYOXJVGSFQF",  /* 0 */
    "H\\NJPISFS[",  /* 1 */
    "H\\LKLJMHNGPFTFVGWHXJXLWNUQK[Y[",  /* 2 */
    "H\\MFXFRNUNWOXPYSYUXXVZS[P[MZLYKW",  /* 3 */
    "H\\UFKTZT UFU[",  /* 4 */
    "H\\WFMFLOMNPMSMVNXPYSYUXXVZS[P[MZLYKW",  /* 5 */
    "H\\XIWGTFRFOGMJLOLTMXOZR[S[VZXXYUYTXQVOSNRNOOMQLT",  /* 6 */
    "H\\YFO[ KFYF",  /* 7 */
    "H\\PFMGLILKMMONSOVPXRYTYWXYWZT[P[MZLYKWKTLRNPQOUNWMXKXIWGTFPF",  /* 8 */
    "H\\XMWPURRSQSNRLPKMKLLINGQFRFUGWIXMXRWWUZR[P[MZLX",  /* 9 */
    "MWRMQNROSNRM RYQZR[SZRY",  /* : */
    "MWRMQNROSNRM SZR[QZRYSZS\\R^Q_",  /* ; */
    "F^ZIJRZ[",  /* < */
    "E_IO[O IU[U",  /* = */
    "F^JIZRJ[",  /* > */
    "I[LKLJMHNGPFTFVGWHXJXLWNVORQRT RYQZR[SZRY",  /* ? */
    "DaWNVLTKQKOLNMMOMRNTOUQVTVVUWS WKWSXUYV[V\\U]S]O\\L[JYHWGTFQFNGLHJJILHOHRIUJWLYNZQ[T[WZYY",  /* @ */
    "I[RFJ[ RFZ[ MTWT",  /* A */
    "G\\KFK[ KFTFWGXHYJYLXNWOTP KPTPWQXRYTYWXYWZT[K[",  /* B */
    "H]ZKYIWGUFQFOGMILKKNKSLVMXOZQ[U[WZYXZV",  /* C */
    "G\\KFK[ KFRFUGWIXKYNYSXVWXUZR[K[",  /* D */
    "H[LFL[ LFYF LPTP L[Y[",  /* E */
    "HZLFL[ LFYF LPTP",  /* F */
    "H]ZKYIWGUFQFOGMILKKNKSLVMXOZQ[U[WZYXZVZS USZS",  /* G */
    "G]KFK[ YFY[ KPYP",  /* H */
    "NVRFR[",  /* I */
    "JZVFVVUYTZR[P[NZMYLVLT",  /* J */
    "G\\KFK[ YFKT POY[",  /* K */
    "HYLFL[ L[X[",  /* L */
    "F^JFJ[ JFR[ ZFR[ ZFZ[",  /* M */
    "G]KFK[ KFY[ YFY[",  /* N */
    "G]PFNGLIKKJNJSKVLXNZP[T[VZXXYVZSZNYKXIVGTFPF",  /* O */
    "G\\KFK[ KFTFWGXHYJYMXOWPTQKQ",  /* P */
    "G]PFNGLIKKJNJSKVLXNZP[T[VZXXYVZSZNYKXIVGTFPF SWY]",  /* Q */
    "G\\KFK[ KFTFWGXHYJYLXNWOTPKP RPY[",  /* R */
    "H\\YIWGTFPFMGKIKKLMMNOOUQWRXSYUYXWZT[P[MZKX",  /* S */
    "JZRFR[ KFYF",  /* T */
    "G]KFKULXNZQ[S[VZXXYUYF",  /* U */
    "I[JFR[ ZFR[",  /* V */
    "F^HFM[ RFM[ RFW[ \\FW[",  /* W */
    "H\\KFY[ YFK[",  /* X */
    "I[JFRPR[ ZFRP",  /* Y */
    "H\\YFK[ KFYF K[Y[",  /* Z */
    "KYOBOb OBVB ObVb",  /* [ */
    "G]IL[b",  /* backslash */
    "KYUBUb NBUB NbUb",  /* ] */
    "G]JTROZT JTRPZT",  /* ^ */
    "I[J[Z[",  /* _ */
    "LXPFUL PFOGUL",  /* ` */
    "I\\XMX[ XPVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* a */
    "H[LFL[ LPNNPMSMUNWPXSXUWXUZS[P[NZLX",  /* b */
    "I[XPVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* c */
    "I\\XFX[ XPVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* d */
    "I[LSXSXQWOVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* e */
    "MYWFUFSGRJR[ OMVM",  /* f */
    "I\\XMX]W`VaTbQbOa XPVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* g */
    "I\\MFM[ MQPNRMUMWNXQX[",  /* h */
    "NVQFRGSFREQF RMR[",  /* i */
    "MWRFSGTFSERF SMS^RaPbNb",  /* j */
    "IZMFM[ WMMW QSX[",  /* k */
    "NVRFR[",  /* l */
    "CaGMG[ GQJNLMOMQNRQR[ RQUNWMZM\\N]Q][",  /* m */
    "I\\MMM[ MQPNRMUMWNXQX[",  /* n */
    "I\\QMONMPLSLUMXOZQ[T[VZXXYUYSXPVNTMQM",  /* o */
    "H[LMLb LPNNPMSMUNWPXSXUWXUZS[P[NZLX",  /* p */
    "I\\XMXb XPVNTMQMONMPLSLUMXOZQ[T[VZXX",  /* q */
    "KXOMO[ OSPPRNTMWM",  /* r */
    "J[XPWNTMQMNNMPNRPSUTWUXWXXWZT[Q[NZMX",  /* s */
    "MYRFRWSZU[W[ OMVM",  /* t */
    "I\\MMMWNZP[S[UZXW XMX[",  /* u */
    "JZLMR[ XMR[",  /* v */
    "G]JMN[ RMN[ RMV[ ZMV[",  /* w */
    "J[MMX[ XMM[",  /* x */
    "JZLMR[ XMR[P_NaLbKb",  /* y */
    "J[XMM[ MMXM M[X[",  /* z */
    "KYTBQEPHPJQMSOSPORSTSUQWPZP\\Q_Tb",  /* { */
    "NVRBRb",  /* | */
    "KYPBSETHTJSMQOQPURQTQUSWTZT\\S_Pb",  /* } */
    "F^IUISJPLONOPPTSVTXTZS[Q ISJQLPNPPQTTVUXUZT[Q[O",  /* ~ */
};
/* cv: drawing.cpp readCheck(): with FONT_HERSHEY_SIMPLEX every byte outside 32..126 (so every byte of a multi-byte
 * UTF-8 character) is replaced by '?' — pinned against the installed cv2 in tests/test_oracle_cvdraw.py */
static const char* simplex_glyph(int c) {
    return kSimplexGlyphs[((c >= 32 && c <= 126) ? c : '?') - 32];
}

/* =============================== exported entry points =============================== */
static Img mk(uint8_t* data, int h, int w, int64_t step) { Img im = {data, h, w, step}; return im; }

void ocv_rectangle(uint8_t* data, int h, int w, int64_t step, int x1, int y1, int x2, int y2,
                   int b, int g, int r, int thickness, int line_type) {
    Img im = mk(data, h, w, step);
    Col col = {{(uint8_t)b, (uint8_t)g, (uint8_t)r}};
    P2l pt[4] = {{x1, y1}, {x2, y1}, {x2, y2}, {x1, y2}};
    poly_line(&im, pt, 4, 1, &col, thickness, line_type, 0);
}

void ocv_line(uint8_t* data, int h, int w, int64_t step, int x1, int y1, int x2, int y2,
              int b, int g, int r, int thickness, int line_type) {
    Img im = mk(data, h, w, step);
    Col col = {{(uint8_t)b, (uint8_t)g, (uint8_t)r}};
    P2l p0 = {x1, y1}, p1 = {x2, y2};
    /* cv::line pre-clips the centre line against the image rectangle grown by `thickness` on every side
     * (observed on the 4.13 binary: clipLine(Rect(-t,-t,w+2t,h+2t)), segment dropped when it misses it) */
    p0.x += thickness; p0.y += thickness; p1.x += thickness; p1.y += thickness;
    if (!clip_line((int64_t)w + 2 * thickness, (int64_t)h + 2 * thickness, &p0, &p1)) return;
    p0.x -= thickness; p0.y -= thickness; p1.x -= thickness; p1.y -= thickness;
    thick_line(&im, p0, p1, &col, thickness, line_type, 3, 0);
}

void ocv_circle(uint8_t* data, int h, int w, int64_t step, int cx, int cy, int radius,
                int b, int g, int r, int thickness, int line_type) {
    Img im = mk(data, h, w, step);
    Col col = {{(uint8_t)b, (uint8_t)g, (uint8_t)r}};
    if (thickness > 1 || line_type != LINE_8) {
        P2l c = {(int64_t)cx << XY_SHIFT, (int64_t)cy << XY_SHIFT};
        int64_t rr = (int64_t)radius << XY_SHIFT;
        ellipse_ex(&im, c, rr, rr, &col, thickness, line_type);
    } else if (thickness < 0) {
        circle_fill(&im, cx, cy, radius, &col);
    }
    /* thickness 0/1 outline with LINE_8 (midpoint outline) is never used by the reference */
}

int ocv_get_text_size(const char* text, double font_scale, int thickness, int* out_w, int* out_h) {
    const int base_line = 9, cap_line = 12;     /* HersheySimplex[0] = 9 + 12*16 */
    double view_x = 0;
    *out_h = cv_round((cap_line + base_line) * font_scale + (thickness + 1) / 2);
    for (const char* s = text; *s; s++) {
        const char* g = simplex_glyph((unsigned char)*s);
        if (!g) return -1;
        int px = (unsigned char)g[0] - 'R', py = (unsigned char)g[1] - 'R';
        view_x += (py - px) * font_scale;
    }
    *out_w = cv_round(view_x + thickness);
    return 0;
}

int ocv_put_text(uint8_t* data, int h, int w, int64_t step, const char* text, int org_x, int org_y,
                 double font_scale, int b, int g, int r, int thickness) {
    Img im = mk(data, h, w, step);
    Col col = {{(uint8_t)b, (uint8_t)g, (uint8_t)r}};
    int base_line = -9;
    int hscale = cv_round(font_scale * XY_ONE), vscale = hscale;
    int64_t view_x = (int64_t)org_x << XY_SHIFT;
    int64_t view_y = ((int64_t)org_y << XY_SHIFT) + (int64_t)base_line * vscale;
    P2l pts[256];
    for (const char* s = text; *s; s++) {
        const char* ptr = simplex_glyph((unsigned char)*s);
        if (!ptr) return -1;
        int64_t px = (unsigned char)ptr[0] - 'R', py = (unsigned char)ptr[1] - 'R';
        int64_t dx = py * hscale;
        view_x -= px * hscale;
        int n = 0;
        for (ptr += 2;;) {
            if (*ptr == ' ' || !*ptr) {
                if (n > 1) poly_line(&im, pts, n, 0, &col, thickness, LINE_8, XY_SHIFT);
                if (!*ptr++) break;
                n = 0;
            } else {
                px = (unsigned char)ptr[0] - 'R';
                py = (unsigned char)ptr[1] - 'R';
                ptr += 2;
                pts[n].x = px * hscale + view_x;
                pts[n].y = py * vscale + view_y;
                n++;
            }
        }
        view_x += dx;
    }
    return 0;
}

/* test hooks: leaf routines with 16.16 fixed-point vertices (cv2.fillConvexPoly(..., shift=16) etc.) */
void ocv_fill_convex_poly16(uint8_t* data, int h, int w, int64_t step, const int64_t* xy, int npts,
                            int b, int g, int r, int line_type) {
    Img im = mk(data, h, w, step);
    Col col = {{(uint8_t)b, (uint8_t)g, (uint8_t)r}};
    P2l v[64];
    if (npts > 64) return;
    for (int i = 0; i < npts; i++) { v[i].x = xy[2 * i]; v[i].y = xy[2 * i + 1]; }
    fill_convex_poly(&im, v, npts, &col, line_type);
}
