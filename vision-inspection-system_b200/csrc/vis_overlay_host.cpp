// vis_overlay_host.cpp — host half of the defect-overlay rasteriser of libvis_b200.so.
//
// vis_overlay_expand turns the validated pixel boxes of one frame (utils/image_utils.py:229-257 of the
// reference) into an ORDERED list of leaf primitives that reproduces, pixel for pixel, what the reference's
// cv2 calls (utils/image_utils.py:259-313) draw sequentially:
//     rectangle/line (thickness 2, LINE_AA) -> marker disc (filled) -> marker ring (thickness 3) -> label text.
// The decomposition follows OpenCV 4.13 modules/imgproc/src/drawing.cpp (ThickLine, PolyLine, EllipseEx,
// ellipse2Poly, FillConvexPoly, Circle, putText, getTextSize, clipLine); the device (vis_overlay.cu) only ever
// sees four leaf kinds, each O(1) to evaluate at a pixel:
//     LINE8   a clipped 8-connected fixed-point line        (cv: Line2)
//     LINEAA  a clipped antialiased line, 8-bit coverage     (cv: LineAA)
//     TRAP    rows [ya,yb] between two linear edge walkers   (cv: FillConvexPoly scan loop, one segment pair)
//     SPANS   up to 16 rows of spans symmetric about cx      (cv: Circle, filled midpoint circle)
// plus GROUP headers (bounding box + leaf range) that let a tile skip whole boxes.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "vis_internal.h"
#include "vis_overlay_leaf.h"

namespace {

constexpr int kShift = 16;
constexpr int64_t kOne = 1 << kShift;
constexpr int kLine8 = 8, kLineAA = 16;

struct Pt { int64_t x, y; };

// sine of whole degrees 0..450 as stored by OpenCV (7-decimal literals)
const float kSin[451] = {
    0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f, 0.1218693f,
    0.1391731f, 0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f, 0.2419219f, 0.2588190f,
    0.2756374f, 0.2923717f, 0.3090170f, 0.3255682f, 0.3420201f, 0.3583679f, 0.3746066f, 0.3907311f,
    0.4067366f, 0.4226183f, 0.4383711f, 0.4539905f, 0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f,
    0.5299193f, 0.5446390f, 0.5591929f, 0.5735764f, 0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f,
    0.6427876f, 0.6560590f, 0.6691306f, 0.6819984f, 0.6946584f, 0.7071068f, 0.7193398f, 0.7313537f,
    0.7431448f, 0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f, 0.8090170f, 0.8191520f,
    0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f, 0.8910065f,
    0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f, 0.9396926f, 0.9455186f,
    0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f, 0.9743701f, 0.9781476f, 0.9816272f,
    0.9848078f, 0.9876883f, 0.9902681f, 0.9925462f, 0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f,
    0.9993908f, 0.9998477f, 1.0000000f, 0.9998477f, 0.9993908f, 0.9986295f, 0.9975641f, 0.9961947f,
    0.9945219f, 0.9925462f, 0.9902681f, 0.9876883f, 0.9848078f, 0.9816272f, 0.9781476f, 0.9743701f,
    0.9702957f, 0.9659258f, 0.9612617f, 0.9563048f, 0.9510565f, 0.9455186f, 0.9396926f, 0.9335804f,
    0.9271839f, 0.9205049f, 0.9135455f, 0.9063078f, 0.8987940f, 0.8910065f, 0.8829476f, 0.8746197f,
    0.8660254f, 0.8571673f, 0.8480481f, 0.8386706f, 0.8290376f, 0.8191520f, 0.8090170f, 0.7986355f,
    0.7880108f, 0.7771460f, 0.7660444f, 0.7547096f, 0.7431448f, 0.7313537f, 0.7193398f, 0.7071068f,
    0.6946584f, 0.6819984f, 0.6691306f, 0.6560590f, 0.6427876f, 0.6293204f, 0.6156615f, 0.6018150f,
    0.5877853f, 0.5735764f, 0.5591929f, 0.5446390f, 0.5299193f, 0.5150381f, 0.5000000f, 0.4848096f,
    0.4694716f, 0.4539905f, 0.4383711f, 0.4226183f, 0.4067366f, 0.3907311f, 0.3746066f, 0.3583679f,
    0.3420201f, 0.3255682f, 0.3090170f, 0.2923717f, 0.2756374f, 0.2588190f, 0.2419219f, 0.2249511f,
    0.2079117f, 0.1908090f, 0.1736482f, 0.1564345f, 0.1391731f, 0.1218693f, 0.1045285f, 0.0871557f,
    0.0697565f, 0.0523360f, 0.0348995f, 0.0174524f, 0.0000000f, -0.0174524f, -0.0348995f, -0.0523360f,
    -0.0697565f, -0.0871557f, -0.1045285f, -0.1218693f, -0.1391731f, -0.1564345f, -0.1736482f, -0.1908090f,
    -0.2079117f, -0.2249511f, -0.2419219f, -0.2588190f, -0.2756374f, -0.2923717f, -0.3090170f, -0.3255682f,
    -0.3420201f, -0.3583679f, -0.3746066f, -0.3907311f, -0.4067366f, -0.4226183f, -0.4383711f, -0.4539905f,
    -0.4694716f, -0.4848096f, -0.5000000f, -0.5150381f, -0.5299193f, -0.5446390f, -0.5591929f, -0.5735764f,
    -0.5877853f, -0.6018150f, -0.6156615f, -0.6293204f, -0.6427876f, -0.6560590f, -0.6691306f, -0.6819984f,
    -0.6946584f, -0.7071068f, -0.7193398f, -0.7313537f, -0.7431448f, -0.7547096f, -0.7660444f, -0.7771460f,
    -0.7880108f, -0.7986355f, -0.8090170f, -0.8191520f, -0.8290376f, -0.8386706f, -0.8480481f, -0.8571673f,
    -0.8660254f, -0.8746197f, -0.8829476f, -0.8910065f, -0.8987940f, -0.9063078f, -0.9135455f, -0.9205049f,
    -0.9271839f, -0.9335804f, -0.9396926f, -0.9455186f, -0.9510565f, -0.9563048f, -0.9612617f, -0.9659258f,
    -0.9702957f, -0.9743701f, -0.9781476f, -0.9816272f, -0.9848078f, -0.9876883f, -0.9902681f, -0.9925462f,
    -0.9945219f, -0.9961947f, -0.9975641f, -0.9986295f, -0.9993908f, -0.9998477f, -1.0000000f, -0.9998477f,
    -0.9993908f, -0.9986295f, -0.9975641f, -0.9961947f, -0.9945219f, -0.9925462f, -0.9902681f, -0.9876883f,
    -0.9848078f, -0.9816272f, -0.9781476f, -0.9743701f, -0.9702957f, -0.9659258f, -0.9612617f, -0.9563048f,
    -0.9510565f, -0.9455186f, -0.9396926f, -0.9335804f, -0.9271839f, -0.9205049f, -0.9135455f, -0.9063078f,
    -0.8987940f, -0.8910065f, -0.8829476f, -0.8746197f, -0.8660254f, -0.8571673f, -0.8480481f, -0.8386706f,
    -0.8290376f, -0.8191520f, -0.8090170f, -0.7986355f, -0.7880108f, -0.7771460f, -0.7660444f, -0.7547096f,
    -0.7431448f, -0.7313537f, -0.7193398f, -0.7071068f, -0.6946584f, -0.6819984f, -0.6691306f, -0.6560590f,
    -0.6427876f, -0.6293204f, -0.6156615f, -0.6018150f, -0.5877853f, -0.5735764f, -0.5591929f, -0.5446390f,
    -0.5299193f, -0.5150381f, -0.5000000f, -0.4848096f, -0.4694716f, -0.4539905f, -0.4383711f, -0.4226183f,
    -0.4067366f, -0.3907311f, -0.3746066f, -0.3583679f, -0.3420201f, -0.3255682f, -0.3090170f, -0.2923717f,
    -0.2756374f, -0.2588190f, -0.2419219f, -0.2249511f, -0.2079117f, -0.1908090f, -0.1736482f, -0.1564345f,
    -0.1391731f, -0.1218693f, -0.1045285f, -0.0871557f, -0.0697565f, -0.0523360f, -0.0348995f, -0.0174524f,
    0.0000000f, 0.0174524f, 0.0348995f, 0.0523360f, 0.0697565f, 0.0871557f, 0.1045285f, 0.1218693f,
    0.1391731f, 0.1564345f, 0.1736482f, 0.1908090f, 0.2079117f, 0.2249511f, 0.2419219f, 0.2588190f,
    0.2756374f, 0.2923717f, 0.3090170f, 0.3255682f, 0.3420201f, 0.3583679f, 0.3746066f, 0.3907311f,
    0.4067366f, 0.4226183f, 0.4383711f, 0.4539905f, 0.4694716f, 0.4848096f, 0.5000000f, 0.5150381f,
    0.5299193f, 0.5446390f, 0.5591929f, 0.5735764f, 0.5877853f, 0.6018150f, 0.6156615f, 0.6293204f,
    0.6427876f, 0.6560590f, 0.6691306f, 0.6819984f, 0.6946584f, 0.7071068f, 0.7193398f, 0.7313537f,
    0.7431448f, 0.7547096f, 0.7660444f, 0.7771460f, 0.7880108f, 0.7986355f, 0.8090170f, 0.8191520f,
    0.8290376f, 0.8386706f, 0.8480481f, 0.8571673f, 0.8660254f, 0.8746197f, 0.8829476f, 0.8910065f,
    0.8987940f, 0.9063078f, 0.9135455f, 0.9205049f, 0.9271839f, 0.9335804f, 0.9396926f, 0.9455186f,
    0.9510565f, 0.9563048f, 0.9612617f, 0.9659258f, 0.9702957f, 0.9743701f, 0.9781476f, 0.9816272f,
    0.9848078f, 0.9876883f, 0.9902681f, 0.9925462f, 0.9945219f, 0.9961947f, 0.9975641f, 0.9986295f,
    0.9993908f, 0.9998477f, 1.0000000f,
};

const unsigned char kSlopeCorr[32] = {181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
                                      203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};

inline int round_half_even(double v) { return (int)std::lrint(v); }

// Hershey simplex glyphs of printable ASCII 32..126 (cv: HersheySimplex[] -> g_HersheyGlyphs[]): font data of the
// installed OpenCV 4.13 binary, recovered and verified by tests/golden/find_glyphs.py.  The reference only ever draws
// "#<int>" labels with the '#' removed (src/reporting/pdf_generator.py:1303, utils/image_utils.py:240-242), but
// draw_bounding_boxes accepts any label text.
const char* glyph_for(unsigned char ch) {
    static const char* const glyphs[95] = {
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
    // cv: readCheck() — with FONT_HERSHEY_SIMPLEX every byte outside 32..126 (UTF-8 continuation bytes included) is '?'
    return glyphs[(ch >= 32 && ch <= 126) ? ch - 32 : '?' - 32];
}

class Emitter {
  public:
    Emitter(int h, int w, VisLeaf* out, int cap) : h_(h), w_(w), out_(out), cap_(cap) {}

    int count() const { return n_; }
    bool ok() const { return glyph_ok_; }

    // one leaf that copies the covered pixels of a marker sprite placed with its anchor at (cx, cy)
    void sprite(const VisSprite& sp, int cx, int cy) {
        VisLeaf l;
        std::memset(&l, 0, sizeof l);
        l.w[0] = LEAF_SPRITE;
        l.w[2] = (int32_t)(uint32_t)(sp.pixels & 0xffffffffu);
        l.w[3] = (int32_t)(uint32_t)(sp.pixels >> 32);
        l.w[4] = cx - sp.ox;
        l.w[5] = cy - sp.oy;
        l.w[6] = sp.w;
        l.w[7] = sp.h;
        push_boxed(l, l.w[4], l.w[5], (int64_t)l.w[4] + sp.w - 1, (int64_t)l.w[5] + sp.h - 1);
    }

    // one leaf that replays the recorded blend chains of a dash stamp anchored at (x, y), with the current colour
    void stamp(const VisSprite& sp, int x, int y, int bx0, int by0, int bx1, int by1) {   // b*: touched box on the canvas
        VisLeaf l;
        std::memset(&l, 0, sizeof l);
        l.w[0] = LEAF_STAMP;
        l.w[2] = (int32_t)(uint32_t)(sp.pixels & 0xffffffffu);
        l.w[3] = (int32_t)(uint32_t)(sp.pixels >> 32);
        l.w[4] = x - sp.ox;
        l.w[5] = y - sp.oy;
        l.w[6] = sp.w;
        l.w[7] = sp.h;
        push_boxed(l, (int64_t)l.w[4] + bx0, (int64_t)l.w[5] + by0, (int64_t)l.w[4] + bx1, (int64_t)l.w[5] + by1);
    }

    // appends a copy of every leaf of `tpl` moved by (dx, dy) whole pixels: all leaf parameters are affine in the
    // pixel coordinates (16.16 walkers, pixel rows / columns, packed boxes), so this equals expanding the same calls
    // at the moved position as long as nothing there is clipped by the image border (the caller checks)
    void instantiate(const std::vector<VisLeaf>& tpl, int dx, int dy) {
        for (const VisLeaf& t : tpl) {
            VisLeaf l = t;
            const int kind = l.w[0] & LEAF_KIND_MASK;
            const bool xmajor = l.w[0] & LEAF_FLAG_XMAJOR;
            switch (kind) {
                case LEAF_LINE8:
                    l.w[2] += xmajor ? dx : dy;
                    l.w[4] += (xmajor ? dy : dx) * (int32_t)kOne;
                    l.w[6] += dx;
                    l.w[7] += dy;
                    break;
                case LEAF_LINEAA:
                    l.w[2] += xmajor ? dx : dy;
                    l.w[4] += (xmajor ? dy : dx) * (int32_t)kOne;
                    break;
                case LEAF_TRAP:
                    l.w[2] += dy; l.w[3] += dy;
                    l.w[4] += dx * (int32_t)kOne; l.w[6] += dx * (int32_t)kOne;
                    break;
                case LEAF_SPANS:
                    l.w[2] += dx; l.w[3] += dy;
                    break;
                default:
                    continue;
            }
            const int x0 = (l.w[10] & 0xffff) + dx, x1 = (int)((uint32_t)l.w[10] >> 16) + dx;
            const int y0 = (l.w[11] & 0xffff) + dy, y1 = (int)((uint32_t)l.w[11] >> 16) + dy;
            pack_bbox(l, x0, y0, x1, y1);
            gx0_ = std::min(gx0_, x0); gy0_ = std::min(gy0_, y0);
            gx1_ = std::max(gx1_, x1); gy1_ = std::max(gy1_, y1);
            push(l);
        }
    }

    void set_color(int b, int g, int r, int a = 0) {
        color_ = (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16) | ((uint32_t)a << 24);
    }

    // ---- group headers: slots [0, n) lead the leaf array, one per box -------------------------------
    void reserve_headers(int n) {
        VisLeaf l;
        std::memset(&l, 0, sizeof l);
        l.w[0] = LEAF_GROUP;
        for (int i = 0; i < n; ++i) push(l);
    }
    void begin_group() {
        group_first_ = n_;
        gx0_ = gy0_ = INT_MAX;
        gx1_ = gy1_ = INT_MIN;
    }
    void end_group(int slot) {                   // header = leaf range + bounding box of everything in it
        const int first = group_first_, last = n_;
        if (slot >= cap_) return;
        VisLeaf& l = out_[slot];
        l.w[2] = first;
        l.w[3] = last;
        if (gx0_ > gx1_ || gy0_ > gy1_) { l.w[3] = first; l.w[10] = 1; l.w[11] = 1; return; }   // nothing visible
        pack_bbox(l, gx0_, gy0_, gx1_, gy1_);
    }

    // ---- cv::line / cv::rectangle / cv::circle / cv::putText ---------------------------------------
    void line(int x1, int y1, int x2, int y2, int thickness, int line_type) {
        Pt a{x1 + thickness, y1 + thickness}, b{x2 + thickness, y2 + thickness};
        // cv::line first clips the centre line to the image grown by `thickness` (4.13 binary behaviour)
        if (!clip_line((int64_t)w_ + 2 * thickness, (int64_t)h_ + 2 * thickness, a, b)) return;
        a.x -= thickness; a.y -= thickness; b.x -= thickness; b.y -= thickness;
        thick_line(a, b, thickness, line_type, 3, 0);
    }
    void rectangle(int x1, int y1, int x2, int y2, int thickness, int line_type) {
        const Pt v[4] = {{x1, y1}, {x2, y1}, {x2, y2}, {x1, y2}};
        poly_line(v, 4, true, thickness, line_type, 0);
    }
    void circle_filled(int cx, int cy, int radius) { disc(cx, cy, radius); }
    void circle_outline(int cx, int cy, int radius, int thickness) {          // thickness > 1, LINE_8
        ellipse({(int64_t)cx << kShift, (int64_t)cy << kShift}, (int64_t)radius << kShift, thickness, kLine8);
    }
    bool text_size(const char* text, double scale, int thickness, int* tw, int* th) {
        double view_x = 0;
        *th = round_half_even((12 + 9) * scale + (thickness + 1) / 2);
        for (const char* s = text; *s; ++s) {
            const char* g = glyph_for((unsigned char)*s);
            if (!g) { glyph_ok_ = false; return false; }
            view_x += (((unsigned char)g[1] - 'R') - ((unsigned char)g[0] - 'R')) * scale;
        }
        *tw = round_half_even(view_x + thickness);
        return true;
    }
    void put_text(const char* text, int org_x, int org_y, double scale, int thickness) {
        const int hscale = round_half_even(scale * kOne), vscale = hscale;
        int64_t view_x = (int64_t)org_x << kShift;
        const int64_t view_y = ((int64_t)org_y << kShift) - 9 * (int64_t)vscale;
        std::vector<Pt> pts;
        for (const char* s = text; *s; ++s) {
            const char* g = glyph_for((unsigned char)*s);
            if (!g) { glyph_ok_ = false; return; }
            const int64_t left = (unsigned char)g[0] - 'R', right = (unsigned char)g[1] - 'R';
            view_x -= left * hscale;
            pts.clear();
            for (const char* p = g + 2;;) {
                if (*p == ' ' || !*p) {
                    if (pts.size() > 1) poly_line(pts.data(), (int)pts.size(), false, thickness, kLine8, kShift);
                    if (!*p++) break;
                    pts.clear();
                } else {
                    const int64_t gx = (unsigned char)p[0] - 'R', gy = (unsigned char)p[1] - 'R';
                    p += 2;
                    pts.push_back({gx * hscale + view_x, gy * vscale + view_y});
                }
            }
            view_x += right * hscale;
        }
    }

  private:
    int h_, w_;
    VisLeaf* out_;
    int cap_;
    int n_ = 0;
    uint32_t color_ = 0;
    bool glyph_ok_ = true;
    int gx0_ = INT_MAX, gy0_ = INT_MAX, gx1_ = INT_MIN, gy1_ = INT_MIN;
    int group_first_ = 0;

    void push(const VisLeaf& l) {
        if (n_ < cap_) out_[n_] = l;
        ++n_;
    }
    static void pack_bbox(VisLeaf& l, int x0, int y0, int x1, int y1) {
        l.w[10] = (x0 & 0xffff) | (x1 << 16);
        l.w[11] = (y0 & 0xffff) | (y1 << 16);
    }
    // clamps the box to the image, drops the leaf when nothing is visible
    void push_boxed(VisLeaf& l, int64_t x0, int64_t y0, int64_t x1, int64_t y1) {
        x0 = std::max<int64_t>(x0, 0); y0 = std::max<int64_t>(y0, 0);
        x1 = std::min<int64_t>(x1, w_ - 1); y1 = std::min<int64_t>(y1, h_ - 1);
        if (x0 > x1 || y0 > y1) return;
        pack_bbox(l, (int)x0, (int)y0, (int)x1, (int)y1);
        l.w[1] = (int32_t)color_;
        gx0_ = std::min(gx0_, (int)x0); gy0_ = std::min(gy0_, (int)y0);
        gx1_ = std::max(gx1_, (int)x1); gy1_ = std::max(gy1_, (int)y1);
        push(l);
    }

    // ---- cv: clipLine on a (scaled) size --------------------------------------------------------
    static bool clip_line(int64_t width, int64_t height, Pt& p1, Pt& p2) {
        if (width <= 0 || height <= 0) return false;
        const int64_t right = width - 1, bottom = height - 1;
        int64_t &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
        auto code = [&](int64_t x, int64_t y) { return (x < 0) + (x > right) * 2 + (y < 0) * 4 + (y > bottom) * 8; };
        int c1 = code(x1, y1), c2 = code(x2, y2);
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1 & 12) {
                const int64_t a = c1 < 8 ? 0 : bottom;
                x1 += (int64_t)((double)(a - y1) * (x2 - x1) / (y2 - y1));
                y1 = a;
                c1 = (x1 < 0) + (x1 > right) * 2;
            }
            if (c2 & 12) {
                const int64_t a = c2 < 8 ? 0 : bottom;
                x2 += (int64_t)((double)(a - y2) * (x2 - x1) / (y2 - y1));
                y2 = a;
                c2 = (x2 < 0) + (x2 > right) * 2;
            }
            if ((c1 & c2) == 0 && (c1 | c2) != 0) {
                if (c1) {
                    const int64_t a = c1 == 1 ? 0 : right;
                    y1 += (int64_t)((double)(a - x1) * (y2 - y1) / (x2 - x1));
                    x1 = a;
                    c1 = 0;
                }
                if (c2) {
                    const int64_t a = c2 == 1 ? 0 : right;
                    y2 += (int64_t)((double)(a - x2) * (y2 - y1) / (x2 - x1));
                    x2 = a;
                    c2 = 0;
                }
            }
        }
        return (c1 | c2) == 0;
    }

    // orientation shared by Line2 and LineAA: major axis, end points ordered along +major, minor step
    struct Oriented { bool xmajor; Pt a, b; int64_t step; };
    static Oriented orient(Pt p1, Pt p2) {
        Oriented o;
        int64_t dx = p2.x - p1.x, dy = p2.y - p1.y;
        const int64_t ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy;
        o.xmajor = ax > ay;
        if (o.xmajor) {
            if (dx < 0) { std::swap(p1, p2); dy = -dy; }
            o.step = (dy * kOne) / (ax | 1);
        } else {
            if (dy < 0) { std::swap(p1, p2); dx = -dx; }
            o.step = (dx * kOne) / (ay | 1);
        }
        o.a = p1; o.b = p2;
        return o;
    }

    // ---- cv: Line2 -> LINE8 leaf ------------------------------------------------------------------
    void line8(Pt p1, Pt p2) {
        if (!clip_line((int64_t)w_ << kShift, (int64_t)h_ << kShift, p1, p2)) return;
        const Oriented o = orient(p1, p2);
        VisLeaf l;
        std::memset(&l, 0, sizeof l);
        l.w[0] = LEAF_LINE8 | (o.xmajor ? LEAF_FLAG_XMAJOR : 0);
        const int64_t maj_a = o.xmajor ? o.a.x : o.a.y, maj_b = o.xmajor ? o.b.x : o.b.y;
        const int64_t min_a = (o.xmajor ? o.a.y : o.a.x) + (kOne >> 1);
        const int ecount = (int)((maj_b - maj_a) >> kShift);
        const int m0 = (int)((maj_a + (kOne >> 1)) >> kShift);
        l.w[2] = m0;
        l.w[3] = ecount;
        l.w[4] = (int32_t)min_a;
        l.w[5] = (int32_t)o.step;
        const int ex = (int)((o.b.x + (kOne >> 1)) >> kShift), ey = (int)((o.b.y + (kOne >> 1)) >> kShift);
        l.w[6] = ex;
        l.w[7] = ey;
        const int64_t min_end = min_a + o.step * ecount;
        int64_t lo = std::min(min_a, min_end) >> kShift, hi = std::max(min_a, min_end) >> kShift;
        int64_t x0, y0, x1, y1;
        if (o.xmajor) { x0 = m0; x1 = m0 + ecount; y0 = lo; y1 = hi; }
        else          { y0 = m0; y1 = m0 + ecount; x0 = lo; x1 = hi; }
        x0 = std::min<int64_t>(x0, ex); x1 = std::max<int64_t>(x1, ex);
        y0 = std::min<int64_t>(y0, ey); y1 = std::max<int64_t>(y1, ey);
        push_boxed(l, x0, y0, x1, y1);
    }

    // ---- cv: LineAA -> LINEAA leaf ----------------------------------------------------------------
    void line_aa(Pt p1, Pt p2) {
        if (!clip_line((int64_t)w_ << kShift, (int64_t)h_ << kShift, p1, p2)) return;
        Oriented o = orient(p1, p2);
        int64_t maj_a = o.xmajor ? o.a.x : o.a.y, maj_b = o.xmajor ? o.b.x : o.b.y;
        int64_t min_a = o.xmajor ? o.a.y : o.a.x;
        maj_b += kOne;
        const int ecount = (int)((maj_b >> kShift) - (maj_a >> kShift));
        const int64_t frac = -(maj_a & (kOne - 1));
        min_a += ((o.step * frac) >> kShift) + (kOne >> 1);
        int slope = (int)((o.step >> (kShift - 5)) & 0x3f);
        if (o.step < 0) slope ^= 0x3f;
        const int64_t i = (maj_a >> (kShift - 7)) & 0x78, j = (maj_b >> (kShift - 7)) & 0x78;
        slope = (slope & 0x20) ? 0x100 : kSlopeCorr[slope];
        int ep[9];
        const int t0 = slope << 7, t1 = ((0x78 - (int)i) | 4) * slope, t2 = ((int)j | 4) * slope;
        ep[0] = 0;
        ep[8] = slope;
        ep[1] = ep[3] = (int)((((((j - i) & 0x78) | 4) * slope) >> 8) & 0x1ff);
        ep[2] = (t1 >> 8) & 0x1ff;
        ep[4] = (int)((((((j - i) + 0x80) | 4) * slope) >> 8) & 0x1ff);
        ep[5] = ((t1 + t0) >> 8) & 0x1ff;
        ep[6] = (t2 >> 8) & 0x1ff;
        ep[7] = ((t2 + t0) >> 8) & 0x1ff;
        VisLeaf l;
        std::memset(&l, 0, sizeof l);
        l.w[0] = LEAF_LINEAA | (o.xmajor ? LEAF_FLAG_XMAJOR : 0);
        const int m0 = (int)(maj_a >> kShift);
        l.w[2] = m0;
        l.w[3] = ecount;
        l.w[4] = (int32_t)min_a;
        l.w[5] = (int32_t)o.step;
        l.w[6] = ep[0] | (ep[1] << 10) | (ep[2] << 20);
        l.w[7] = ep[3] | (ep[4] << 10) | (ep[5] << 20);
        l.w[8] = ep[6] | (ep[7] << 10) | (ep[8] << 20);
        const int64_t min_end = min_a + o.step * ecount;
        const int64_t lo = (std::min(min_a, min_end) >> kShift) - 1, hi = (std::max(min_a, min_end) >> kShift) + 1;
        if (o.xmajor) push_boxed(l, m0, lo, (int64_t)m0 + ecount, hi);
        else          push_boxed(l, lo, m0, hi, (int64_t)m0 + ecount);
    }

    // ---- cv: FillConvexPoly (vertices 16.16) -> edge lines + TRAP leaves -----------------------------
    void fill_convex(const Pt* v, int n, int line_type) {
        const int64_t delta = kOne >> 1;
        const int64_t d1 = line_type < kLineAA ? (kOne >> 1) : kOne - 1, d2 = line_type < kLineAA ? (kOne >> 1) : 0;
        int64_t xmin = v[0].x, xmax = v[0].x, ymin = v[0].y, ymax = v[0].y;
        int imin = 0;
        Pt prev = v[n - 1];
        for (int i = 0; i < n; ++i) {
            const Pt p = v[i];
            if (p.y < ymin) { ymin = p.y; imin = i; }
            ymax = std::max(ymax, p.y);
            xmax = std::max(xmax, p.x);
            xmin = std::min(xmin, p.x);
            if (line_type <= 8) line8(prev, p); else line_aa(prev, p);
            prev = p;
        }
        xmin = (xmin + delta) >> kShift; xmax = (xmax + delta) >> kShift;
        ymin = (ymin + delta) >> kShift; ymax = (ymax + delta) >> kShift;
        if (n < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= w_ || (int)ymin >= h_) return;
        ymax = std::min<int64_t>(ymax, h_ - 1);
        struct Walker { int idx, di, ye; int64_t x, dx; int y_set; int64_t x_set; } e[2];
        int y = (int)ymin, edges = n;
        for (int i = 0; i < 2; ++i) e[i] = {imin, i == 0 ? 1 : n - 1, y, -kOne, 0, y, -kOne};
        int run_start = y;                       // first row of the current (segment pair) run
        auto flush = [&](int y_last) {           // rows [run_start, y_last] share both walkers' segments
            const int ya = std::max(run_start, 0);
            if (ya > y_last) return;
            VisLeaf l;
            std::memset(&l, 0, sizeof l);
            l.w[0] = LEAF_TRAP | (line_type < kLineAA ? 0 : LEAF_FLAG_AA);
            l.w[2] = ya;
            l.w[3] = y_last;
            int64_t xs[2], xe[2];
            for (int i = 0; i < 2; ++i) {
                xs[i] = e[i].x_set + e[i].dx * (ya - e[i].y_set);
                xe[i] = e[i].x_set + e[i].dx * (y_last - e[i].y_set);
                l.w[4 + 2 * i] = (int32_t)xs[i];
                l.w[5 + 2 * i] = (int32_t)e[i].dx;
            }
            const int64_t lo = std::min(std::min(xs[0], xs[1]), std::min(xe[0], xe[1]));
            const int64_t hi = std::max(std::max(xs[0], xs[1]), std::max(xe[0], xe[1]));
            push_boxed(l, (lo + d1) >> kShift, ya, (hi + d2) >> kShift, y_last);
        };
        do {
            if (line_type < kLineAA || y < (int)ymax || y == (int)ymin) {
                for (int i = 0; i < 2; ++i) {
                    if (y < e[i].ye) continue;
                    int idx0 = e[i].idx, di = e[i].di, idx = idx0 + di;
                    if (idx >= n) idx -= n;
                    for (; edges-- > 0;) {
                        const int ty = (int)((v[idx].y + delta) >> kShift);
                        if (ty > y) {
                            if (y > run_start) flush(y - 1);     // close the run that used the old segment
                            run_start = y;
                            const int64_t xs = v[idx0].x, xe = v[idx].x;
                            e[i].ye = ty;
                            e[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
                            e[i].x = xs;
                            e[i].idx = idx;
                            e[i].y_set = y;
                            e[i].x_set = xs;
                            break;
                        }
                        idx0 = idx;
                        idx += di;
                        if (idx >= n) idx -= n;
                    }
                }
            }
            if (edges < 0) break;
            e[0].x += e[0].dx;
            e[1].x += e[1].dx;
        } while (++y <= (int)ymax);
        // rows [run_start, y-1] were scanned with the current pair (y stopped one past the last filled row)
        if (y - 1 >= run_start) flush(y - 1);
    }

    // ---- cv: Circle (filled midpoint circle) -> SPANS leaves -----------------------------------------
    void disc(int cx, int cy, int radius) {
        std::vector<int> half((size_t)radius + 1, -1);
        int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
        while (dx >= dy) {
            half[dy] = std::max(half[dy], dx);
            half[dx] = std::max(half[dx], dy);
            ++dy;
            err += plus;
            plus += 2;
            const int mask = (err <= 0) - 1;
            err -= minus & mask;
            dx += mask;
            minus -= mask & 2;
        }
        for (int base = -radius; base <= radius; base += 16) {
            const int rows = std::min(16, radius - base + 1);
            VisLeaf l;
            std::memset(&l, 0, sizeof l);
            l.w[0] = LEAF_SPANS;
            l.w[2] = cx;
            l.w[3] = cy + base;
            l.w[4] = rows;
            int widest = -1;
            for (int r = 0; r < rows; ++r) {
                const int off = base + r, hw = half[off < 0 ? -off : off];      // 0xff marks "no span"
                const int byte = hw < 0 ? 0xff : hw;
                l.w[5 + r / 4] |= byte << (8 * (r % 4));
                widest = std::max(widest, hw);
            }
            if (widest < 0) continue;
            push_boxed(l, (int64_t)cx - widest, (int64_t)cy + base, (int64_t)cx + widest, (int64_t)cy + base + rows - 1);
        }
    }

    // ---- cv: EllipseEx (full circle, axis-aligned) -------------------------------------------------
    void ellipse(Pt center, int64_t axis, int thickness, int line_type) {
        int delta = (int)((axis + (kOne >> 1)) >> kShift);
        delta = delta < 3 ? 90 : delta < 10 ? 30 : delta < 15 ? 18 : 5;
        std::vector<Pt> v;
        Pt prev{-1, -1};
        const double cxd = (double)center.x, cyd = (double)center.y, ad = (double)axis;
        const float alpha = kSin[450], beta = kSin[0];
        for (int a = 0; a < 360 + delta; a += delta) {
            const int ang = std::min(a, 360);
            const double x = ad * kSin[450 - ang], y = ad * kSin[ang];
            const double px = cxd + x * alpha - y * beta, py = cyd + x * beta + y * alpha;
            Pt q;
            q.x = (int64_t)round_half_even(px / (double)kOne) << kShift;
            q.y = (int64_t)round_half_even(py / (double)kOne) << kShift;
            q.x += round_half_even(px - q.x);
            q.y += round_half_even(py - q.y);
            if (q.x != prev.x || q.y != prev.y) { v.push_back(q); prev = q; }
        }
        if (v.size() <= 1) v.assign(2, center);
        if (thickness >= 0) poly_line(v.data(), (int)v.size(), false, thickness, line_type, kShift);
        else fill_convex(v.data(), (int)v.size(), line_type);
    }

    // ---- cv: PolyLine / ThickLine --------------------------------------------------------------------
    void poly_line(const Pt* v, int count, bool closed, int thickness, int line_type, int shift) {
        if (count <= 0) return;
        int flags = 2 + (closed ? 0 : 1);
        Pt p0 = v[closed ? count - 1 : 0];
        for (int i = closed ? 0 : 1; i < count; ++i) {
            thick_line(p0, v[i], thickness, line_type, flags, shift);
            p0 = v[i];
            flags = 2;
        }
    }
    void thick_line(Pt p0, Pt p1, int thickness, int line_type, int flags, int shift) {
        p0.x <<= kShift - shift; p0.y <<= kShift - shift;
        p1.x <<= kShift - shift; p1.y <<= kShift - shift;
        if (thickness <= 1) {
            if (line_type < kLineAA) line8(p0, p1); else line_aa(p0, p1);
            return;
        }
        const double inv = 1.0 / (double)kOne;
        const double dx = (p0.x - p1.x) * inv, dy = (p1.y - p0.y) * inv;
        double r = dx * dx + dy * dy;
        const int odd = thickness & 1;
        const int64_t half = (int64_t)thickness << (kShift - 1);
        if (std::fabs(r) > DBL_EPSILON) {
            r = (half + odd * kOne * 0.5) / std::sqrt(r);
            const int64_t ox = round_half_even(dy * r), oy = round_half_even(dx * r);
            const Pt quad[4] = {{p0.x + ox, p0.y + oy}, {p0.x - ox, p0.y - oy}, {p1.x - ox, p1.y - oy}, {p1.x + ox, p1.y + oy}};
            fill_convex(quad, 4, line_type);
        }
        for (int i = 0; i < 2; ++i) {
            if (flags & (i + 1)) {
                if (line_type < kLineAA)
                    disc((int)((p0.x + (kOne >> 1)) >> kShift), (int)((p0.y + (kOne >> 1)) >> kShift),
                         (int)((half + (kOne >> 1)) >> kShift));
                else
                    ellipse(p0, half, -1, line_type);
            }
            p0 = p1;
        }
    }
};


// ---- templates ------------------------------------------------------------------------------------------------------
// The marker (white disc, coloured ring, label) and an interior dash expand to the same leaves wherever they sit, up to
// a whole-pixel translation: the leaves are built once on a private canvas and copied with an offset afterwards
// (Emitter::instantiate) — ~600 leaves per marker and ~20 per dash that the host no longer recomputes per box.
struct Template {
    std::vector<VisLeaf> leaves;
    int ox = 0, oy = 0;              // canvas position of the anchor (marker centre / dash start)
    int ex = 0, ey = 0;              // marker: half extents (pixels) the instance must have free around its centre
    int bx0 = 0, by0 = 0, bx1 = -1, by1 = -1;   // union of the leaf boxes on the canvas
    bool ok = true;                  // false: no template (a label too wide for a private canvas)
    void bound() {
        bx0 = by0 = INT_MAX; bx1 = by1 = INT_MIN;
        for (const VisLeaf& l : leaves) {
            bx0 = std::min(bx0, l.w[10] & 0xffff); bx1 = std::max(bx1, (int)((uint32_t)l.w[10] >> 16));
            by0 = std::min(by0, l.w[11] & 0xffff); by1 = std::max(by1, (int)((uint32_t)l.w[11] >> 16));
        }
    }
};

struct TemplateCache {
    std::unordered_map<std::string, Template> map;
    const Template& get(const std::string& key, Template (*build)(const void*), const void* arg) {
        auto it = map.find(key);
        if (it != map.end()) return it->second;
        if (map.size() >= 512) map.clear();
        return map.emplace(key, build(arg)).first->second;
    }
};

struct MarkerSpec { int radius; uint8_t b, g, r; const char* label; uint8_t alpha; };
struct DashSpec { int dx, dy, thickness, line_type; uint8_t b, g, r; };

Template build_marker(const void* arg) {
    const MarkerSpec& m = *static_cast<const MarkerSpec*>(arg);
    Template t;
    // half extents: the ring, or the label when it is wider / taller than the ring (any printable ASCII is allowed)
    const double fs = m.radius / 20.0 * 0.7;
    const int tt = std::max(2, (int)(fs * 2));
    int lw = 0, lh = 0;
    {
        Emitter probe(1, 1, nullptr, 0);
        probe.text_size(m.label, fs, tt, &lw, &lh);
    }
    t.ex = std::max(m.radius + 4, lw / 2 + tt + 8);
    t.ey = std::max(m.radius + 4, lh + tt + 8);
    if (t.ex > 8000 || t.ey > 8000) { t.ok = false; return t; }      // a label wider than any frame it could sit in whole:
                                                                      // no template, the caller expands (and clips) in place
    const int pad = 8, cxc = t.ex + pad, cyc = t.ey + pad;           // nothing comes near the canvas border
    t.ox = cxc; t.oy = cyc;
    std::vector<VisLeaf> buf(4096);
    for (;;) {
        Emitter em(2 * cyc + 1, 2 * cxc + 1, buf.data(), (int)buf.size());
        em.begin_group();
        em.set_color(255, 255, 255, m.alpha);
        em.circle_filled(cxc, cyc, m.radius);                         // utils/image_utils.py:299
        em.set_color(m.b, m.g, m.r, m.alpha);
        em.circle_outline(cxc, cyc, m.radius, 3);                     // :302
        const double font_scale = m.radius / 20.0 * 0.7;              // :305-313
        const int text_thickness = std::max(2, (int)(font_scale * 2));
        int tw = 0, th = 0;
        if (!em.text_size(m.label, font_scale, text_thickness, &tw, &th)) { t.ok = false; return t; }
        em.set_color(0, 0, 0, m.alpha);
        em.put_text(m.label, (int)(cxc - tw / 2.0), (int)(cyc + th / 2.0), font_scale, text_thickness);
        if (em.count() <= (int)buf.size()) { buf.resize((size_t)em.count()); break; }
        buf.resize((size_t)em.count());
    }
    t.leaves = std::move(buf);
    return t;
}

Template build_dash(const void* arg) {
    const DashSpec& d = *static_cast<const DashSpec*>(arg);
    Template t;
    const int pad = d.thickness + 8;
    t.ox = t.oy = pad;
    const int w = d.dx + 2 * pad + 1, h = d.dy + 2 * pad + 1;
    std::vector<VisLeaf> buf(256);
    for (;;) {
        Emitter em(h, w, buf.data(), (int)buf.size());
        em.begin_group();
        em.set_color(d.b, d.g, d.r);
        em.line(pad, pad, pad + d.dx, pad + d.dy, d.thickness, d.line_type);
        if (em.count() <= (int)buf.size()) { buf.resize((size_t)em.count()); break; }
        buf.resize((size_t)em.count());
    }
    t.leaves = std::move(buf);
    t.bound();
    return t;
}

thread_local TemplateCache g_templates;

// sprites the current expansion may reference (set by vis_overlay_plan_batch_sprites around its per-frame work)
thread_local const VisSprite* g_sprites = nullptr;
thread_local int g_n_sprites = 0;

const VisSprite* find_stamp(int dx, int dy) {
    char key[24];
    std::snprintf(key, sizeof key, "%d,%d", dx, dy);
    for (int i = 0; i < g_n_sprites; ++i) {
        const VisSprite& s = g_sprites[i];
        if (s.radius == -1 && s.pixels && s.label && std::strcmp(s.label, key) == 0) return &s;
    }
    return nullptr;
}

const VisSprite* find_sprite(int radius, const VisBox& b, const char* label) {
    for (int i = 0; i < g_n_sprites; ++i) {
        const VisSprite& s = g_sprites[i];
        if (s.radius == radius && s.b == b.b && s.g == b.g && s.r == b.r && s.pixels && s.label && std::strcmp(s.label, label) == 0)
            return &s;
    }
    return nullptr;
}

}  // namespace

extern "C" int vis_overlay_expand(int img_h, int img_w, const VisBox* boxes, int n_boxes,
                                  VisLeaf* leaves, int capacity, int* needed) {
    if (img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767 || n_boxes < 0 || (n_boxes && !boxes) ||
        capacity < 0 || (capacity && !leaves)) {
        vis::set_error("vis_overlay_expand: bad arguments (h=%d w=%d boxes=%d)", img_h, img_w, n_boxes);
        return VIS_E_INVALID;
    }
    // group headers first (one per box), leaves after: a tile scans the headers and skips whole boxes
    Emitter em(img_h, img_w, leaves, capacity);
    em.reserve_headers(n_boxes);
    for (int i = 0; i < n_boxes; ++i) {
        const VisBox& b = boxes[i];
        const int x = b.x, y = b.y, w = b.w, h = b.h;
        em.begin_group();
        em.set_color(b.b, b.g, b.r);
        // a dash whose strokes stay clear of the image border is an instance of the dash template
        auto dash = [&](int x1, int y1, int x2, int y2) {
            const int m = 8;
            if (std::min(x1, x2) >= m && std::min(y1, y2) >= m && std::max(x1, x2) < img_w - m && std::max(y1, y2) < img_h - m) {
                DashSpec d{x2 - x1, y2 - y1, 2, 16, b.b, b.g, b.r};
                char key[64];
                std::snprintf(key, sizeof key, "d%d,%d,%d,%d,%d", d.dx, d.dy, b.b, b.g, b.r);
                const Template& t = g_templates.get(key, build_dash, &d);
                const VisSprite* sp = find_stamp(d.dx, d.dy);
                if (sp && sp->ox == t.ox && sp->oy == t.oy && sp->w == d.dx + 2 * t.ox + 1 && sp->h == d.dy + 2 * t.oy + 1)
                    em.stamp(*sp, x1, y1, t.bx0, t.by0, t.bx1, t.by1);   // blend chains recorded once on the device: one leaf
                else
                    em.instantiate(t.leaves, x1 - t.ox, y1 - t.oy);
            } else {
                em.line(x1, y1, x2, y2, 2, 16);
            }
        };
        if (b.dashed) {                          // utils/image_utils.py:260-283: 10 px dashes, 5 px gaps
            for (int k = 0; k < 2; ++k) {
                const int yy = k == 0 ? y : y + h;
                for (int px = x; px < x + w; px += 15) {
                    const int ex = std::min(px + 10, x + w);
                    if (ex > px) dash(px, yy, ex, yy);
                }
            }
            for (int k = 0; k < 2; ++k) {
                const int xx = k == 0 ? x : x + w;
                for (int py = y; py < y + h; py += 15) {
                    const int ey = std::min(py + 10, y + h);
                    if (ey > py) dash(xx, py, xx, ey);
                }
            }
        } else {
            em.rectangle(x, y, x + w, y + h, 2, 16);                      // :286
        }
        // marker (:290-302)
        int radius = (int)(std::max(img_w, img_h) * 0.04);
        radius = std::max(25, std::min(radius, 60));
        const int cx = std::max(radius + 5, std::min(x + radius + 5, img_w - radius - 5));
        const int cy = std::max(radius + 5, std::min(y + radius + 5, img_h - radius - 5));
        // disc, ring and label (:299-313): an instance of the marker template when the ring and the label stay clear of
        // the image border (the usual case: the centre is clamped radius + 5 away from it and '#<n>' fits the ring);
        // expanded in place otherwise (wide free-text labels near an edge, images smaller than the marker)
        const char* label = b.label ? b.label : "";                  // any length, any bytes (cv2.putText draws them all)
        MarkerSpec m{radius, b.b, b.g, b.r, label, 0};
        char key[64];
        std::snprintf(key, sizeof key, "m%d,%d,%d,%d,", radius, b.b, b.g, b.r);
        const Template& t = g_templates.get(std::string(key) + label, build_marker, &m);
        if (t.ok && cx - t.ex >= 0 && cy - t.ey >= 0 && cx + t.ex < img_w && cy + t.ey < img_h) {
            const VisSprite* sp = find_sprite(radius, b, label);
            if (sp && sp->w == 2 * t.ox + 1 && sp->h == 2 * t.oy + 1 && sp->ox == t.ox && sp->oy == t.oy)
                em.sprite(*sp, cx, cy);                  // rasterised once on the device: one leaf
            else
                em.instantiate(t.leaves, cx - t.ox, cy - t.oy);
        } else {
            em.set_color(255, 255, 255);
            em.circle_filled(cx, cy, radius);
            em.set_color(b.b, b.g, b.r);
            em.circle_outline(cx, cy, radius, 3);
            const double font_scale = radius / 20.0 * 0.7;
            const int text_thickness = std::max(2, (int)(font_scale * 2));
            int tw = 0, th = 0;
            em.text_size(label, font_scale, text_thickness, &tw, &th);
            em.set_color(0, 0, 0);
            em.put_text(label, (int)(cx - tw / 2.0), (int)(cy + th / 2.0), font_scale, text_thickness);
        }
        em.end_group(i);
    }
    if (needed) *needed = em.count();
    if (em.count() > capacity) {
        vis::set_error("vis_overlay_expand: %d leaves needed, capacity %d", em.count(), capacity);
        return VIS_E_CAPACITY;
    }
    return em.count();
}

// cv2.getTextSize(text, FONT_HERSHEY_SIMPLEX, font_scale, thickness)[0] — used by the callers of the draw list to
// centre their labels (utils/image_utils.py:657, :665, :731 in the reference).
extern "C" int vis_text_size(const char* text, double font_scale, int thickness, int* width, int* height) {
    if (!text || !width || !height) {
        vis::set_error("vis_text_size: null argument");
        return VIS_E_INVALID;
    }
    Emitter em(1, 1, nullptr, 0);
    if (!em.text_size(text, font_scale, thickness, width, height)) {
        vis::set_error("vis_text_size: '%s' has a character outside printable ASCII", text);
        return VIS_E_UNSUPPORTED;
    }
    return VIS_OK;
}

// Draw list -> leaves: one group per command, so tile binning and the draw kernel treat a command like a box.
// The commands are the cv2 calls of create_side_by_side_comparison (:658, :666) and create_status_stamp (:726, :733).
extern "C" int vis_draw_expand(int img_h, int img_w, const VisDrawCmd* cmds, int n_cmds,
                               VisLeaf* leaves, int capacity, int* needed) {
    if (img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767 || n_cmds < 0 || (n_cmds && !cmds) ||
        capacity < 0 || (capacity && !leaves)) {
        vis::set_error("vis_draw_expand: bad arguments (h=%d w=%d cmds=%d)", img_h, img_w, n_cmds);
        return VIS_E_INVALID;
    }
    Emitter em(img_h, img_w, leaves, capacity);
    em.reserve_headers(n_cmds);
    for (int i = 0; i < n_cmds; ++i) {
        const VisDrawCmd& c = cmds[i];
        em.begin_group();
        em.set_color(c.color[0], c.color[1], c.color[2], c.color[3]);
        const bool lt_ok = c.line_type == 8 || c.line_type == 16;
        switch (c.kind) {
            case VIS_DRAW_LINE:
                if (!lt_ok || c.thickness < 1 || c.thickness > 255) goto bad;
                em.line(c.x1, c.y1, c.x2, c.y2, c.thickness, c.line_type);
                break;
            case VIS_DRAW_RECTANGLE:
                if (!lt_ok || c.thickness < 1 || c.thickness > 255) goto bad;
                em.rectangle(c.x1, c.y1, c.x2, c.y2, c.thickness, c.line_type);
                break;
            case VIS_DRAW_CIRCLE:
                if (c.x2 < 0 || c.x2 > 16383 || !(c.thickness < 0 || (c.thickness > 1 && c.thickness <= 255))) goto bad;
                if (c.thickness < 0) em.circle_filled(c.x1, c.y1, c.x2);
                else em.circle_outline(c.x1, c.y1, c.x2, c.thickness);
                break;
            case VIS_DRAW_TEXT: {
                const char* text = c.text ? c.text : "";
                int tw = 0, th = 0;
                if (c.thickness < 2 || c.thickness > 255 || !(c.font_scale > 0)) goto bad;   // 1-px strokes are not pinned against cv2 (the reference never draws them)
                if (!em.text_size(text, c.font_scale, c.thickness, &tw, &th)) {
                    vis::set_error("vis_draw_expand: text '%s' has a character outside printable ASCII", text);
                    return VIS_E_UNSUPPORTED;
                }
                em.put_text(text, c.x1, c.y1, c.font_scale, c.thickness);
                break;
            }
            default:
            bad:
                vis::set_error("vis_draw_expand: command %d (kind %d, thickness %d, line type %d) is not drawable",
                               i, c.kind, c.thickness, c.line_type);
                return VIS_E_INVALID;
        }
        em.end_group(i);
    }
    if (needed) *needed = em.count();
    if (em.count() > capacity) {
        vis::set_error("vis_draw_expand: %d leaves needed, capacity %d", em.count(), capacity);
        return VIS_E_CAPACITY;
    }
    return em.count();
}

// Leaves of one marker on its own canvas, colours with alpha 255: drawn once on a zeroed BGRA canvas (on the device) they
// give the sprite every whole marker of that (radius, colour, label) is copied from.
extern "C" int vis_overlay_sprite_expand(int radius, int b, int g, int r, const char* label, VisLeaf* leaves, int capacity,
                                         int* needed, int* w, int* h, int* ox, int* oy) {
    if (radius < 1 || radius > 4096 || !label || capacity < 0 || (capacity && !leaves) || !w || !h || !ox || !oy ||
        (b | g | r) < 0 || (b | g | r) > 255) {
        vis::set_error("vis_overlay_sprite_expand: bad arguments (radius %d)", radius);
        return VIS_E_INVALID;
    }
    MarkerSpec m{radius, (uint8_t)b, (uint8_t)g, (uint8_t)r, label, 255};
    const Template t = build_marker(&m);
    if (!t.ok) {
        vis::set_error("vis_overlay_sprite_expand: the label is too wide for a sprite (it is expanded in place instead)");
        return VIS_E_UNSUPPORTED;
    }
    *w = 2 * t.ox + 1; *h = 2 * t.oy + 1; *ox = t.ox; *oy = t.oy;
    Emitter em(*h, *w, leaves, capacity);
    em.reserve_headers(1);
    em.begin_group();
    em.instantiate(t.leaves, 0, 0);
    em.end_group(0);
    if (needed) *needed = em.count();
    if (em.count() > capacity) {
        vis::set_error("vis_overlay_sprite_expand: %d leaves needed, capacity %d", em.count(), capacity);
        return VIS_E_CAPACITY;
    }
    return em.count();
}

// Leaves of one dash (0,0)-(dx,dy), thickness 2, LINE_AA, on its own canvas: drawn in record mode (channels = 8) they give
// the per-pixel blend chains a stamp leaf replays.
extern "C" int vis_overlay_stamp_expand(int dx, int dy, VisLeaf* leaves, int capacity, int* needed, int* w, int* h,
                                        int* ox, int* oy) {
    if (dx < 0 || dy < 0 || dx > 64 || dy > 64 || (dx == 0) == (dy == 0) || capacity < 0 || (capacity && !leaves) || !w || !h ||
        !ox || !oy) {
        vis::set_error("vis_overlay_stamp_expand: bad arguments (dash %d,%d)", dx, dy);
        return VIS_E_INVALID;
    }
    DashSpec d{dx, dy, 2, 16, 0, 0, 0};
    const Template t = build_dash(&d);
    *w = dx + 2 * t.ox + 1; *h = dy + 2 * t.oy + 1; *ox = t.ox; *oy = t.oy;
    Emitter em(*h, *w, leaves, capacity);
    em.reserve_headers(1);
    em.begin_group();
    em.instantiate(t.leaves, 0, 0);
    em.end_group(0);
    if (needed) *needed = em.count();
    if (em.count() > capacity) {
        vis::set_error("vis_overlay_stamp_expand: %d leaves needed, capacity %d", em.count(), capacity);
        return VIS_E_CAPACITY;
    }
    return em.count();
}

// Bins the frame's leaves (culled one by one, in order) into the 64x16-pixel CTA tiles of
// vis_overlay.cu.  tiles_out: one record per touched tile, row-major: {tx | ty << 16, first ref, one past last ref};
// refs_out: per tile, IN LEAF ORDER, {leaf, leaf + 1} of every leaf whose box touches the tile.
extern "C" int vis_overlay_tiles(int img_h, int img_w, const VisLeaf* leaves, int n_boxes,
                                 int32_t* tiles_out, int tile_capacity, int32_t* refs_out, int ref_capacity,
                                 int* tiles_needed, int* refs_needed) {
    if (img_h <= 0 || img_w <= 0 || img_h > 32767 || img_w > 32767 || n_boxes < 0 || (n_boxes && !leaves) ||
        tile_capacity < 0 || ref_capacity < 0 || (tile_capacity && !tiles_out) || (ref_capacity && !refs_out)) {
        vis::set_error("vis_overlay_tiles: bad arguments (h=%d w=%d boxes=%d)", img_h, img_w, n_boxes);
        return VIS_E_INVALID;
    }
    constexpr int kTileW = 64, kTileH = 16;
    const int tw = (img_w + kTileW - 1) / kTileW, th = (img_h + kTileH - 1) / kTileH;
    std::vector<int> count((size_t)tw * th, 0);
    // every leaf of every group, in order, into the tiles its box touches: the refs of a tile name single leaves
    auto each = [&](auto&& fn) {
        for (int g = 0; g < n_boxes; ++g) {
            const VisLeaf& h = leaves[g];
            for (int li = h.w[2]; li < h.w[3]; ++li) {
                const VisLeaf& leaf = leaves[li];
                const int x0 = leaf.w[10] & 0xffff, x1 = (int)((uint32_t)leaf.w[10] >> 16);
                const int y0 = leaf.w[11] & 0xffff, y1 = (int)((uint32_t)leaf.w[11] >> 16);
                if (x0 > x1 || y0 > y1) continue;
                for (int ty = y0 / kTileH; ty <= std::min(y1 / kTileH, th - 1); ++ty)
                    for (int tx = x0 / kTileW; tx <= std::min(x1 / kTileW, tw - 1); ++tx) fn(ty * tw + tx, li);
            }
        }
    };
    each([&](int t, int) { ++count[t]; });
    std::vector<int> at((size_t)tw * th, 0);
    int n_tiles = 0, n_refs = 0;
    for (int t = 0; t < tw * th; ++t) {
        at[t] = n_refs;
        if (!count[t]) continue;
        if (n_tiles < tile_capacity) {
            tiles_out[3 * n_tiles] = (t % tw) | ((t / tw) << 16);
            tiles_out[3 * n_tiles + 1] = n_refs;
            tiles_out[3 * n_tiles + 2] = n_refs + count[t];
        }
        ++n_tiles;
        n_refs += count[t];
    }
    if (tiles_needed) *tiles_needed = n_tiles;
    if (refs_needed) *refs_needed = n_refs;
    if (n_tiles > tile_capacity || n_refs > ref_capacity) {
        vis::set_error("vis_overlay_tiles: %d tiles / %d refs, capacity %d / %d", n_tiles, n_refs, tile_capacity, ref_capacity);
        return VIS_E_CAPACITY;
    }
    each([&](int t, int li) {
        refs_out[2 * at[t]] = li;
        refs_out[2 * at[t] + 1] = li + 1;
        ++at[t];
    });
    return n_tiles;
}

// Batch form of vis_overlay_expand + vis_overlay_tiles: every frame is independent, so the frames are spread over host
// threads (the per-frame entry points are reentrant; only the error text is thread local).  Outputs are the
// concatenated, batch-indexed arrays vis_overlay_draw consumes.
extern "C" int vis_overlay_plan_batch(int n_frames, const int32_t* hw, const VisBox* boxes, const int32_t* box_begin,
                                      VisLeaf* leaves, int64_t leaf_capacity, int32_t* leaf_begin,
                                      VisOverlayTile* tiles, int64_t tile_capacity,
                                      VisOverlayRef* refs, int64_t ref_capacity, int64_t* needed, int n_threads) {
    return vis_overlay_plan_batch_sprites(n_frames, hw, boxes, box_begin, leaves, leaf_capacity, leaf_begin, tiles,
                                          tile_capacity, refs, ref_capacity, needed, n_threads, nullptr, 0);
}

extern "C" int vis_overlay_plan_batch_sprites(int n_frames, const int32_t* hw, const VisBox* boxes, const int32_t* box_begin,
                                              VisLeaf* leaves, int64_t leaf_capacity, int32_t* leaf_begin,
                                              VisOverlayTile* tiles, int64_t tile_capacity,
                                              VisOverlayRef* refs, int64_t ref_capacity, int64_t* needed, int n_threads,
                                              const VisSprite* sprites, int n_sprites) {
    if (n_sprites < 0 || (n_sprites && !sprites)) {
        vis::set_error("vis_overlay_plan_batch: bad sprite table");
        return VIS_E_INVALID;
    }
    if (n_frames <= 0 || !hw || !box_begin || !leaf_begin || !needed || leaf_capacity < 0 || tile_capacity < 0 ||
        ref_capacity < 0 || (leaf_capacity && !leaves) || (tile_capacity && !tiles) || (ref_capacity && !refs)) {
        vis::set_error("vis_overlay_plan_batch: bad arguments");
        return VIS_E_INVALID;
    }
    struct Frame { std::vector<VisLeaf> leaves; std::vector<int32_t> tiles, refs; int rc = VIS_OK; std::string err; };
    std::vector<Frame> out((size_t)n_frames);
    std::atomic<int> next(0);
    auto work = [&]() {
        g_sprites = sprites;                      // thread-local: the expansion of this thread may emit sprite leaves
        g_n_sprites = n_sprites;
        struct Reset { ~Reset() { g_sprites = nullptr; g_n_sprites = 0; } } reset;
        static thread_local std::vector<VisLeaf> scratch_l;
        static thread_local std::vector<int32_t> scratch_t, scratch_r;
        for (int i = next.fetch_add(1); i < n_frames; i = next.fetch_add(1)) {
            Frame& fr = out[(size_t)i];
            const int h = hw[2 * i], w = hw[2 * i + 1], nb = box_begin[i + 1] - box_begin[i];
            const VisBox* b = boxes + box_begin[i];
            if (nb <= 0) continue;
            // worst-case scratch lives once per thread and is never zero-filled (value-initialising 1400 leaves per box
            // and a full tile grid per frame used to cost more than the expansion itself); a frame keeps exact-size copies
            int need = 0;
            if (scratch_l.size() < (size_t)nb * 1400) scratch_l.resize((size_t)nb * 1400);
            int rc = vis_overlay_expand(h, w, b, nb, scratch_l.data(), (int)scratch_l.size(), &need);
            if (rc == VIS_E_CAPACITY) {
                scratch_l.resize((size_t)need);
                rc = vis_overlay_expand(h, w, b, nb, scratch_l.data(), (int)scratch_l.size(), &need);
            }
            if (rc < 0) { fr.rc = rc; fr.err = vis_last_error(); continue; }
            fr.leaves.assign(scratch_l.begin(), scratch_l.begin() + rc);
            int nt = 0, nr = 0;
            const int tcap = ((w + 63) / 64) * ((h + 15) / 16);
            if (scratch_t.size() < (size_t)tcap * 3) scratch_t.resize((size_t)tcap * 3);
            if (scratch_r.size() < (size_t)tcap * 8) scratch_r.resize((size_t)tcap * 8);
            rc = vis_overlay_tiles(h, w, fr.leaves.data(), nb, scratch_t.data(), tcap, scratch_r.data(), (int)scratch_r.size() / 2, &nt, &nr);
            if (rc == VIS_E_CAPACITY) {
                scratch_r.resize((size_t)nr * 2);
                rc = vis_overlay_tiles(h, w, fr.leaves.data(), nb, scratch_t.data(), tcap, scratch_r.data(), nr, &nt, &nr);
            }
            if (rc < 0) { fr.rc = rc; fr.err = vis_last_error(); continue; }
            fr.tiles.assign(scratch_t.begin(), scratch_t.begin() + (size_t)nt * 3);
            fr.refs.assign(scratch_r.begin(), scratch_r.begin() + (size_t)nr * 2);
        }
    };
    int nth = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nth = std::max(1, std::min(nth, std::min(n_frames, 64)));
    std::vector<std::thread> pool;
    for (int t = 1; t < nth; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    int64_t nl = 0, nt = 0, nr = 0;
    for (int i = 0; i < n_frames; ++i) {
        const Frame& fr = out[(size_t)i];
        if (fr.rc < 0) { vis::set_error("vis_overlay_plan_batch: frame %d: %s", i, fr.err.c_str()); return fr.rc; }
        nl += (int64_t)fr.leaves.size(); nt += (int64_t)fr.tiles.size() / 3; nr += (int64_t)fr.refs.size() / 2;
    }
    needed[0] = nl; needed[1] = nt; needed[2] = nr;
    if (nl > leaf_capacity || nt > tile_capacity || nr > ref_capacity || nl > INT32_MAX || nr > INT32_MAX) {
        vis::set_error("vis_overlay_plan_batch: needs %lld leaves / %lld tiles / %lld refs", (long long)nl, (long long)nt, (long long)nr);
        return VIS_E_CAPACITY;
    }
    // offsets of every frame in the three output arrays, then the copy-out spread over the same threads (the leaves
    // alone are ~280 KB per annotated 1080p frame: a single-threaded merge would cost as much as the expansion)
    std::vector<int64_t> off_l((size_t)n_frames + 1, 0), off_t((size_t)n_frames + 1, 0), off_r((size_t)n_frames + 1, 0);
    for (int i = 0; i < n_frames; ++i) {
        const Frame& fr = out[(size_t)i];
        off_l[i + 1] = off_l[i] + (int64_t)fr.leaves.size();
        off_t[i + 1] = off_t[i] + (int64_t)fr.tiles.size() / 3;
        off_r[i + 1] = off_r[i] + (int64_t)fr.refs.size() / 2;
        leaf_begin[i] = (int32_t)off_l[i];
    }
    const int64_t al = off_l[n_frames];
    next.store(0);
    auto merge = [&]() {
        for (int i = next.fetch_add(1); i < n_frames; i = next.fetch_add(1)) {
            const Frame& fr = out[(size_t)i];
            if (!fr.leaves.empty()) std::memcpy(leaves + off_l[i], fr.leaves.data(), fr.leaves.size() * sizeof(VisLeaf));
            const int32_t rbase = (int32_t)off_r[i];
            VisOverlayTile* t = tiles + off_t[i];
            for (size_t k = 0; k < fr.tiles.size() / 3; ++k, ++t) {
                t->frame = i;
                t->txy = fr.tiles[3 * k];
                t->ref_begin = fr.tiles[3 * k + 1] + rbase;
                t->ref_end = fr.tiles[3 * k + 2] + rbase;
            }
            VisOverlayRef* r = refs + off_r[i];
            for (size_t k = 0; k < fr.refs.size() / 2; ++k, ++r) {
                r->leaf_begin = fr.refs[2 * k];
                r->leaf_end = fr.refs[2 * k + 1];
            }
        }
    };
    pool.clear();
    for (int t = 1; t < nth; ++t) pool.emplace_back(merge);
    merge();
    for (auto& t : pool) t.join();
    leaf_begin[n_frames] = (int32_t)al;
    return (int)nt;
}
