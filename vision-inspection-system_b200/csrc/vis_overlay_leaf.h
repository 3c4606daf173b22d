// vis_overlay_leaf.h — encoding of the 48-byte leaf primitives shared by vis_overlay_host.cpp (producer)
// and vis_overlay.cu (consumer).  Every leaf: w[0] = kind | flags, w[1] = colour (B | G<<8 | R<<16),
// w[10] = x0 | x1<<16, w[11] = y0 | y1<<16 (inclusive bounding box, clamped to the image).
#pragma once


enum : int {
    LEAF_NOP = 0,
    LEAF_GROUP = 1,    // header of a box: w[2] = first leaf, w[3] = one past last leaf (indices inside the frame's array)
    LEAF_LINE8 = 2,    // w[2] major start px, w[3] ecount, w[4] minor start (16.16, +0.5), w[5] minor step, w[6],w[7] end pixel
    LEAF_LINEAA = 3,   // w[2] major start px, w[3] ecount, w[4] minor start (16.16), w[5] step, w[6..8] nine 10-bit end-point factors
    LEAF_TRAP = 4,     // w[2] ya, w[3] yb, w[4] x of walker 0 at ya, w[5] its step, w[6] x of walker 1 at ya, w[7] its step
    LEAF_SPANS = 5,    // w[2] cx, w[3] first row, w[4] rows (<= 16), w[5..8] half-widths, one byte per row (0xff = none)
    LEAF_STAMP = 7,    // w[1] colour, w[2],w[3] device pointer to the blend-chain records (8 bytes per pixel) of a w[6] x w[7] stamp at (w[4], w[5])
    LEAF_SPRITE = 6,   // w[2],w[3] device pointer (lo, hi) to a BGRA sprite of w[6] x w[7] pixels placed at (w[4], w[5]): pixels with alpha are copied
    LEAF_KIND_MASK = 0xff,
    LEAF_FLAG_XMAJOR = 0x100,   // LINE8 / LINEAA: x is the major axis
    LEAF_FLAG_AA = 0x200,       // TRAP: antialiased polygon rounding (left +ONE-1, right +0) instead of +ONE/2
};
