/*
 * vis_b200.h — C ABI of the B200 inspection-frame preprocessing engine (libvis_b200.so).
 *
 * This is the drop-in boundary for the ONE hot path of Aditya-Somasi/Vision-Inspection-System:
 *   inspection frame -> Qwen2-VL pixel_values / image_grid_thw, thumbnail/resize, defect overlay.
 * Every entry point names the reference interface it replaces (file:line under the reference
 * tree; `tf:` = transformers 5.5.0, `PIL:` = Pillow 12.2.0, `cv:` = OpenCV 4.13 drawing.cpp).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.
 *   - return 0 on success, negative VIS_E_* on failure; vis_last_error() gives a thread-local text.
 *   - functions marked [host] never touch the GPU; functions marked [device] enqueue work on the
 *     given CUDA stream (a cudaStream_t passed as void*), never synchronise, never allocate:
 *     every device buffer (inputs, outputs, tables, scratch) is owned by the caller.
 *   - thread-safe: no global mutable state except the thread-local error string.
 */
#ifndef VIS_B200_H
#define VIS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VIS_B200_ABI_VERSION 20

/* status codes */
#define VIS_OK            0
#define VIS_E_INVALID    -1   /* bad argument */
#define VIS_E_CUDA       -2   /* CUDA runtime error (text in vis_last_error) */
#define VIS_E_UNSUPPORTED -3  /* geometry outside what the fast path handles; caller must use the generic path */
#define VIS_E_CAPACITY   -4   /* caller-provided buffer too small */

/* resampling filters (PIL:Image.py Resampling enum values are kept) */
#define VIS_FILTER_LANCZOS 1  /* Image.Resampling.LANCZOS, support 3.0 — utils/image_utils.py:75, src/agents/vlm_inspector.py:64, src/agents/vlm_auditor.py:91 */
#define VIS_FILTER_BICUBIC 3  /* Image.Resampling.BICUBIC, a=-0.5, support 2.0 — tf:image_transforms.py:367 */

int         vis_abi_version(void);      /* [host] */
const char* vis_last_error(void);       /* [host] thread-local, never NULL */
const char* vis_source_hash(void);      /* [host] hex16 of the sources, header and flags this binary was built from
                                           (build.py compares it with the tree: a stale binary is rebuilt, never tested) */

/* ------------------------------------------------------------------------------------------
 * Coefficient tables                                                              [host]
 * Replaces Pillow's precompute_coeffs + normalize_coeffs_8bpc (libImaging/Resample.c), reached
 * from Image.resize at utils/image_utils.py:75 and tf:image_transforms.py:367.
 * IEEE double arithmetic, no FMA contraction, coefficients truncated to 22-bit fixed point.
 * ------------------------------------------------------------------------------------------ */
/* number of coefficient slots per output sample for (in_size -> out_size) */
int vis_coeff_ksize(int in_size, int out_size, int filter);
/* k[out_size*ksize] (zero padded), bounds[out_size*2] = (first input index, tap count) */
int vis_build_coeffs(int in_size, int out_size, int filter, int32_t* k, int32_t* bounds, int* ksize_out);

/* the same tables for a fractional source box [in0, in1) of the axis (Resample.c precompute_coeffs with in0 / in1;
 * ImagingResample takes a float box): what Image.resize(..., box=...) and the thumbnail's reduce pre-pass need.  [host] */
int vis_coeff_ksize_box(float in0, float in1, int out_size, int filter);
int vis_build_coeffs_box(int in_size, float in0, float in1, int out_size, int filter, int32_t* k, int32_t* bounds,
                         int* ksize_out);

/* 768-entry normalisation table: lut[v*3+c] = (f32(f64(v)*rescale) - f32(mean[c])) / f32(std[c])
 * Replaces tf:image_transforms.py:118-122 (rescale) + :427-439 (normalize) — exact because a
 * uint8 input admits only 256x3 distinct results.                                  [host] */
int vis_build_lut(const float mean[3], const float stdv[3], double rescale, float* lut768);

/* ------------------------------------------------------------------------------------------
 * Generic (any geometry) device passes — one image per call.                      [device]
 * Pillow ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc; the caller chooses the
 * pass order exactly as PIL:Image.py:2431-2435 does (horizontal first unless the tall-image branch).
 * src/dst: interleaved uint8, `channels` bytes per pixel, row pitches in bytes.
 * k/bounds: device copies of the vis_build_coeffs tables.
 * ------------------------------------------------------------------------------------------ */
int vis_resample_h_u8(const uint8_t* src, int64_t src_pitch, int rows, int in_w, int channels,
                      uint8_t* dst, int64_t dst_pitch, int out_w,
                      const int32_t* k, const int32_t* bounds, int ksize, void* stream);
int vis_resample_v_u8(const uint8_t* src, int64_t src_pitch, int in_h, int row_bytes,
                      uint8_t* dst, int64_t dst_pitch, int out_h,
                      const int32_t* k, const int32_t* bounds, int ksize, void* stream);

/* Image.reduce((fx, fy), box) — libImaging/Reduce.c, the integer box-average pre-pass Image.thumbnail / Image.resize
 * insert (reducing_gap = 2.0) when the frame is >= 4x larger than the target (src/agents/vlm_inspector.py:64,
 * src/agents/vlm_auditor.py:91): every output sample is ((sum + n/2) * M(n)) >> 24 over the n source pixels of its
 * fx x fy cell (clipped at the right / bottom edge of the box), M(n) = (uint32)(2^32f / (256 n)) in float arithmetic.
 * box: (x0, y0, x1, y1) source region; dst: [ceil((y1-y0)/fy), ceil((x1-x0)/fx), channels].          [device] */
int vis_reduce_u8(const uint8_t* src, int64_t src_pitch, int h, int w, int channels, int fx, int fy,
                  int x0, int y0, int x1, int y1, uint8_t* dst, int64_t dst_pitch, void* stream);

/* Image.resize(size, NEAREST, box): what Pillow runs for palette ("P") and bilevel ("1") frames whatever filter the
 * caller names (PIL:Image.py:2396-2397; _imaging.c _resize + Geometry.c ImagingScaleAffine) — e.g. a palette PNG
 * going through the agents' thumbnail (src/agents/vlm_inspector.py:64) before its RGB conversion (:68).
 * vis_nearest_table [host]: tab[d] = source index of output sample d or -1 (outside: written as 0), walked with the
 * accumulated double of ImagingScaleAffine.  vis_gather_u8 [device]: dst[y][x] = src[ytab[y]][xtab[x]].            */
int vis_nearest_table(int in_size, float in0, float in1, int out_size, int32_t* tab);
int vis_gather_u8(const uint8_t* src, int64_t src_pitch, int h, int w, int channels, uint8_t* dst, int64_t dst_pitch,
                  int out_h, int out_w, const int32_t* xtab, const int32_t* ytab, void* stream);

/* Alpha premultiplication around the resample of "RGBA" / "LA" frames, which Pillow resamples in premultiplied form
 * (PIL:Image.py:2399-2402 -> libImaging/Convert.c rgbA2rgba / rgba2rgbA, la2lA / lA2la).  In place; channels = 4 (RGBA)
 * or 2 (LA), alpha last.  forward != 0: c = MULDIV255(c, a) = ((t = c*a + 128) + (t >> 8)) >> 8; forward == 0:
 * c = min(255, 255*c / a) unless a is 0 or 255 (copied).                                               [device] */
int vis_alpha_premultiply_u8(uint8_t* img, int64_t pitch, int h, int w, int channels, int forward, void* stream);

/* Modes Pillow does not resample as 8 bits per channel — "I;16" / "I;16L" / "I;16B", "I" (int32), "F" (float32) — go
 * through its double-precision passes (libImaging/Resample.c ImagingResampleHorizontal/Vertical_16bpc / _32bpc): that is
 * what img.resize(new_size, LANCZOS) at utils/image_utils.py:75 runs for such frames.  vis_build_coeffs_f64 [host]: the
 * normalised weights as doubles, k[out_size * ksize] (ksize = vis_coeff_ksize), bounds as vis_build_coeffs.
 * vis_resample_hp [device]: one pass (vertical = 0: along rows, h rows of w samples -> out_size samples; vertical = 1:
 * along columns), sum of pixel * k in tap order with one rounding per multiply and add, then ROUND_UP / the two CLIP8 byte
 * writes of the 16-bit path / the float conversion, exactly as Pillow.  Single channel.                              */
#define VIS_HP_U16LE 0
#define VIS_HP_U16BE 1
#define VIS_HP_I32   2
#define VIS_HP_F32   3
int vis_build_coeffs_f64(int in_size, int out_size, int filter, double* k, int32_t* bounds, int* ksize_out);
int vis_resample_hp(const uint8_t* src, int64_t src_pitch, int h, int w, int kind, int vertical,
                    uint8_t* dst, int64_t dst_pitch, int out_size, const double* k, const int32_t* bounds, int ksize,
                    void* stream);

/* Row re-pitch.  The fused kernels stage rows with bulk copies, which need a 16-byte aligned base and pitch; frames that
 * are not (a 502-pixel-wide RGB frame has 1506-byte rows) are copied, a whole batch per launch, into a caller-owned
 * staging buffer with an aligned pitch.  descs: DEVICE array; dst and dst_pitch multiples of 16; bytes of a dst row
 * beyond row_bytes are left untouched.  max_frame_bytes sizes the grid.                                  [device] */
typedef struct VisRepitch {
    const uint8_t* src;
    uint8_t*       dst;
    int64_t        src_pitch, dst_pitch;
    int32_t        rows, row_bytes;
} VisRepitch;
int vis_repitch_u8(const VisRepitch* descs, int n, int64_t max_frame_bytes, void* stream);

/* resized RGB uint8 HWC [h,w,3] (h,w multiples of 28) -> rows [row0, row0 + (h/14)*(w/14)) of
 * pixel_values [*,1176] f32 in Qwen2-VL patch order (tf:models/qwen2_vl/image_processing_pil_qwen2_vl.py:182-214):
 * rescale+normalize through lut768, temporal duplicate (T=2), 14x14 patches, 2x2 merge order.  [device] */
int vis_normalize_patchify(const uint8_t* src, int64_t src_pitch, int h, int w,
                           const float* lut768, float* pixel_values, int64_t row0, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused hot path: frame -> pixel_values in ONE launch for a whole batch.           [device]
 * Replaces, per frame, tf:...image_processing_pil_qwen2_vl.py:164-214 (smart-resized bicubic resample,
 * rescale, normalize, patchify).  Horizontal-then-vertical order only (VIS_E_UNSUPPORTED otherwise).
 * ------------------------------------------------------------------------------------------ */
typedef struct VisFrame {
    const uint8_t* src;        /* device, RGB uint8 HWC, 16-byte aligned                          */
    int64_t        src_pitch;  /* bytes per row, multiple of 16                                  */
    int32_t        src_h, src_w;
    int32_t        dst_h, dst_w;   /* smart_resize result, multiples of 28                       */
    const int32_t* hrec;       /* device, packed horizontal records (vis_pack_records, dst_w+1)  */
    const int32_t* vrec;       /* device, packed vertical records   (vis_pack_records, dst_h+1)  */
    int64_t        row0;       /* first row of this frame in pixel_values                        */
} VisFrame;

typedef struct VisStrip {      /* one unit of work (one CTA): a column strip of one frame        */
    int32_t frame;             /* index into frames[]                                            */
    int32_t x0, x1;            /* output columns [x0,x1), multiples of 28                        */
    int32_t y0, y1;            /* output rows    [y0,y1), multiples of 14                        */
} VisStrip;

/* largest tap count of a bounds table                                                 [host] */
int vis_max_taps(const int32_t* bounds, int out_size);
/* tap class the fused kernel is instantiated for (6, 8, 12, 16) or 0 if kt > 16        [host] */
int vis_fused_kt_class(int kt);
/* int32 slots per record for tap class kt                                              [host] */
int vis_record_stride(int kt);
/* pack a vis_build_coeffs table into push-order records: out_size+1 records (last = sentinel),
 * each vis_record_stride(kt) int32: [0..kt) coefficients newest tap first (zero padded),
 * [stride-2] = first, [stride-1] = last input index of the window.  Both tables of a frame must be
 * packed with the same kt = vis_fused_kt_class(max taps of both).                      [host] */
int vis_pack_records(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int kt,
                     int32_t* rec, int64_t rec_capacity);
/* VIS_OK if the fused kernel can take this geometry (alignment, <= 16 taps, horizontal-first order),
 * else VIS_E_UNSUPPORTED: use the generic passes                                        [host] */
int vis_fused_supported(int64_t src_addr, int64_t src_pitch, int src_h, int src_w,
                        int dst_h, int dst_w, int hkt, int vkt);
/* upper bound of strips vis_plan_strips emits for one frame                            [host] */
int vis_plan_strips_max(int dst_h, int dst_w);
/* split one frame into column strips x `vsplit` row segments; hbounds = HOST horizontal bounds table,
 * kt = max taps of both axes.  Returns the strip count; *span_bytes_out / *strip_w_out receive the widest
 * staged input span and strip width (they size the launch's shared memory).            [host] */
int vis_plan_strips(int frame_index, int dst_h, int dst_w, const int32_t* hbounds, int kt, int vsplit,
                    VisStrip* strips, int capacity, int* span_bytes_out, int* strip_w_out);
/* frames/strips: DEVICE arrays; every frame of one call shares the tap class of max_kt.
 * max_span_bytes / max_strip_w: maxima of vis_plan_strips over the batch.              [device] */
int vis_preprocess_fused(const VisFrame* frames, int n_frames, const VisStrip* strips, int n_strips,
                         int max_kt, int max_span_bytes, int max_strip_w,
                         const float* lut768, float* pixel_values, void* stream);

/* ------------------------------------------------------------------------------------------
 * Statically scheduled fused kernel: same computation as vis_preprocess_fused for ONE geometry per
 * launch whose scale is > 0.5 on both axes (at most two output samples end at any input index; one when
 * nothing is upscaled).
 * The host precomputes, from the bounds tables, which input column of every 8-pixel step / which input
 * row of every 8-row group completes an output sample; the schedule travels as a kernel parameter
 * (constant bank), so every branch of the resampling loops is warp-uniform and no window bookkeeping
 * is left on the device.  Geometries it declines (VIS_E_UNSUPPORTED) go to vis_preprocess_fused.
 * ------------------------------------------------------------------------------------------ */
#define VIS_SCHED_MAX_STRIPS 16      /* column strips per frame (<= 336 output columns each)       */
#define VIS_SCHED_MAX_SEGS   16      /* row segments per frame                                     */
#define VIS_SCHED_SUBS       12      /* column sub-ranges per strip of the 8- and 16-slot kernels (one per horizontal-pass warp) */
#define VIS_SCHED_MAX_SUBS   20      /* ... at most (the packed-byte kernel runs 16 horizontal-pass warps)        */
#define VIS_SCHED_MASK_BYTES 6144

typedef struct VisSchedStrip { int32_t x0, x1, px0, row_bytes; } VisSchedStrip;
typedef struct VisSchedSub   { uint16_t xa, xb, p0, nsteps, mask_off, pad; } VisSchedSub;
typedef struct VisSchedSeg   { int32_t y0, y1, r_first, r_end, mask_off, pad; } VisSchedSeg;

typedef struct VisSched {            /* opaque to callers: filled by vis_sched_build, passed back by pointer */
    int32_t src_h, src_w, dst_h, dst_w;
    int64_t src_pitch;
    int32_t kt;                      /* tap class (6 or 8)                                         */
    int32_t n_strips, n_segs;
    int32_t stage_pitch, max_strip_w;
    int32_t per_index;               /* 1: scale >= 1 on both axes; 2: mild upscale, two mask bytes are live per step */
    int32_t ring;                    /* register window of the kernel: 8 (<= 8 taps) or 16 (<= 16 taps)            */
    int32_t n_subs;                  /* column sub-ranges per strip = horizontal-pass warps of the kernel (12 / 16) */
    int32_t out_mode;                /* VIS_SCHED_OUT_PIXEL_VALUES or VIS_SCHED_OUT_U8                              */
    int32_t h_pull;                  /* 1: 17..32 taps, the horizontal role pulls its window (no step masks)        */
    int32_t n_vwarps;                /* 16-slot kernel: vertical-pass warps of the launch (6 / 4 / 3: fewer for strong downscales) */
    int32_t dp_words;                /* > 0: packed-byte kernels (vis_fused_dp.cu / vis_fused_mma.cu), W words of 4 taps per window (4..9) */
    int32_t mma_ks;                  /* > 0: integer tensor-path kernel (vis_fused_mma.cu); 32-pixel k-steps per tile of 16 output columns */
    int32_t chunk_rows;              /* input rows a chunk advances by (32; 28 / 24 for the tensor-path kernel at vertical scales < 2) */
    VisSchedStrip strip[VIS_SCHED_MAX_STRIPS];
    VisSchedSub   sub[VIS_SCHED_MAX_STRIPS][VIS_SCHED_MAX_SUBS];
    VisSchedSeg   seg[VIS_SCHED_MAX_SEGS];
    uint8_t       mask[VIS_SCHED_MASK_BYTES];
} VisSched;

typedef struct VisFrameRef {         /* per frame of a scheduled launch (device array)             */
    const uint8_t* src;              /* RGB uint8 HWC, 16-byte aligned, row pitch = sched.src_pitch */
    int64_t        row0;             /* first row of this frame in pixel_values                    */
} VisFrameRef;

#define VIS_SCHED_OUT_PIXEL_VALUES 0   /* LUT + Qwen2-VL patch layout, fp32 (dst_h, dst_w multiples of 28)            */
#define VIS_SCHED_OUT_U8           1   /* resized RGB uint8 HWC (dst_w multiple of 4): Image.resize / thumbnails      */
#define VIS_SCHED_FLAG_DP4A    0x100   /* OR-ed into out_mode: serve 9+ tap geometries with the packed-byte (IDP.4A) kernel
                                          instead of the 16-slot IMAD kernel (same results; the faster one where measured) */

#define VIS_SCHED_FLAG_MMA     0x200   /* OR-ed into out_mode: serve 9+ tap geometries with the integer tensor-path kernel
                                          (mma.sync u8 x 8-bit limbs -> s32, same limb arithmetic and results as the packed-byte
                                          kernel; falls back to it when a tile's window exceeds three k-steps)            */

/* sizeof(VisSched), for bindings that treat it as an opaque byte buffer                  [host] */
int vis_sched_sizeof(void);
/* hbounds / vbounds: HOST bounds tables of vis_build_coeffs (dst_w x 2, dst_h x 2); vsplit = row segments.
 * VIS_OK, or VIS_E_UNSUPPORTED when the geometry needs the general kernel.              [host] */
int vis_sched_build(int src_h, int src_w, int dst_h, int dst_w, int64_t src_pitch,
                    const int32_t* hbounds, const int32_t* vbounds, int vsplit, int out_mode, VisSched* out);
/* records for the scheduled kernel: like vis_pack_records, but samples whose window the far border clamps are
 * moved to the virtual end index the schedule gives them (leading zero coefficients).  kt, per_index: the
 * schedule's.                                                                             [host] */
int vis_sched_pack_records(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int kt, int per_index,
                           int32_t* rec, int64_t rec_capacity);
/* records for the packed-byte kernel (sched.dp_words = words > 0): per output sample 3 limb rows of `words` 32-bit
 * words, byte j of word q = that limb of the coefficient of input index 4 * ((end >> 2) - (words - 1) + q) + j, where
 * `end` is the sample's scheduled window end; coefficient k = k0 + 256 k1 + 65536 k2, k0 / k1 unsigned bytes, k2 signed.
 * rec: (out_size + 1) * vis_sched_record_stride_dp(words) int32.                              [host] */
int vis_sched_record_stride_dp(int words);
int vis_sched_pack_records_dp(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int words,
                              int32_t* rec, int64_t rec_capacity);
/* records for the integer tensor-path kernel (sched.mma_ks > 0): the limb bytes of vis_sched_pack_records_dp, limb-minor
 * (32-bit word 3 q + l = limb l of the four samples of window word q), followed by two int32: the absolute 4-sample word index of record byte 0 ((end >> 2) - (words - 1)) and of the window's first tap
 * (first >> 2).  rec: (out_size + 1) * vis_sched_record_stride_mma(words) int32.             [host] */
int vis_sched_record_stride_mma(int words);
int vis_sched_pack_records_mma(int out_size, const int32_t* k, const int32_t* bounds, int ksize, int words,
                               int32_t* rec, int64_t rec_capacity);
/* frames: DEVICE array; hrec / vrec: DEVICE records from vis_sched_pack_records.        [device] */
int vis_preprocess_fused_sched(const VisSched* sched, const VisFrameRef* frames, int n_frames,
                               const int32_t* hrec, const int32_t* vrec,
                               const float* lut768, float* pixel_values, void* stream);
/* The same with an optional second destination per frame.  Dual Inspector + Auditor inputs
 * (src/agents/vlm_inspector.py:59-69 thumbnail 2048, src/agents/vlm_auditor.py:87-96 thumbnail 1024): a frame that
 * neither agent thumbnails (longer side <= 1024) gives both processors the SAME pixel_values rows; dup_rows[i] >= 0
 * (DEVICE int64[n_frames], or NULL) names the first row of frame i's copy in the same pixel_values allocation, written
 * from the same registers as the primary rows — computed once, never copied.  8-slot kernel (<= 8 taps) only:
 * VIS_E_UNSUPPORTED otherwise.                                                            [device] */
int vis_preprocess_fused_sched_dup(const VisSched* sched, const VisFrameRef* frames, int n_frames,
                                   const int32_t* hrec, const int32_t* vrec,
                                   const float* lut768, float* pixel_values, const int64_t* dup_rows, void* stream);

/* Image.resize((dst_w, dst_h), filter) of RGB uint8 HWC frames in ONE fused launch (both passes, uint8 between them,
 * horizontal first): the agents' thumbnails (src/agents/vlm_inspector.py:64, vlm_auditor.py:91) and resize_image
 * (utils/image_utils.py:75).  sched: built with VIS_SCHED_OUT_U8 from the filter's bounds tables (<= 16 taps).
 * frames: DEVICE array of (src, dst) pointers; dst rows are dst_pitch bytes apart (multiple of 4).   [device] */
typedef struct VisResizeRef { const uint8_t* src; uint8_t* dst; } VisResizeRef;
int vis_resize_fused_sched(const VisSched* sched, const VisResizeRef* frames, int n_frames, int64_t dst_pitch,
                           const int32_t* hrec, const int32_t* vrec, void* stream);

/* ------------------------------------------------------------------------------------------
 * Defect overlay rasteriser.
 * Replaces the cv2 drawing calls of draw_bounding_boxes (utils/image_utils.py:259-313):
 * cv2.rectangle/line (LINE_AA, thickness 2), cv2.circle (filled, and thickness 3), cv2.putText
 * (FONT_HERSHEY_SIMPLEX).  The host expands validated pixel boxes into an ORDERED list of leaf
 * primitives (cv: ThickLine/PolyLine/EllipseEx/FillConvexPoly/LineAA/Line2/Circle/putText) and
 * the device applies them per pixel in list order, which reproduces OpenCV's sequential result.
 * ------------------------------------------------------------------------------------------ */
typedef struct VisBox {        /* a box AFTER the reference's validation + percent->pixel conversion (utils/image_utils.py:200-237) */
    int32_t x, y, w, h;        /* pixels                                                         */
    uint8_t b, g, r;           /* BGR colour (utils/image_utils.py:250-252)                       */
    uint8_t dashed;            /* 1 iff confidence == "low" (utils/image_utils.py:257)            */
    const char* label;         /* HOST pointer: NUL-terminated UTF-8 text drawn in the marker ('#' already removed,
                                  utils/image_utils.py:240-242), any length; bytes outside 32..126 draw '?' as cv2 does */
} VisBox;

#define VIS_LEAF_WORDS 12
typedef struct VisLeaf { int32_t w[VIS_LEAF_WORDS]; } VisLeaf;   /* opaque 48-byte leaf primitive */

/* expand the boxes of ONE frame into its leaf array: n_boxes group headers (one per box: leaf range +
 * bounding box, indices relative to the start of this array) followed by the ordered leaves.
 * Returns the leaf count, or VIS_E_CAPACITY with *needed = required count (call again with a larger buffer),
 * Never fails on label content: like cv2.putText, any byte outside printable ASCII is drawn as '?'.   [host] */
int vis_overlay_expand(int img_h, int img_w, const VisBox* boxes, int n_boxes,
                       VisLeaf* leaves, int capacity, int* needed);

typedef struct VisOverlayFrame {
    const uint8_t* src;        /* device BGR uint8 HWC                                           */
    uint8_t*       dst;        /* device BGR uint8 HWC; may equal src (in place)                 */
    int64_t        src_pitch, dst_pitch;
    int32_t        h, w;
    int32_t        group_begin, group_end;   /* this frame's group headers in leaves[]; its leaf array starts at group_begin */
} VisOverlayFrame;

/* bins the leaves of ONE frame (written by vis_overlay_expand; culled per leaf, in leaf order) into the 64x16-pixel tiles
 * of the draw kernel.  tiles_out: 3 int32 per touched tile, row-major: tx | ty << 16, first ref, one past last ref.
 * refs_out: 2 int32 per ref, per tile IN LEAF ORDER: first leaf, one past last leaf (indices inside the frame's leaf
 * array); this function emits single-leaf refs, which the draw kernel fetches 32 at a time (it also accepts longer
 * runs).  Returns the tile count, or VIS_E_CAPACITY with the needed counts.                               [host] */
int vis_overlay_tiles(int img_h, int img_w, const VisLeaf* leaves, int n_boxes,
                      int32_t* tiles_out, int tile_capacity, int32_t* refs_out, int ref_capacity,
                      int* tiles_needed, int* refs_needed);

typedef struct VisOverlayTile {   /* one CTA of the draw kernel */
    int32_t frame;                /* index into frames[]                                         */
    int32_t txy;                  /* tx | ty << 16                                               */
    int32_t ref_begin, ref_end;   /* this tile's refs in the batch-wide refs[] array             */
} VisOverlayTile;
typedef struct VisOverlayRef { int32_t leaf_begin, leaf_end; } VisOverlayRef;   /* relative to the frame's leaf array */

/* Batch form of vis_overlay_expand + vis_overlay_tiles for n_frames independent frames, spread over n_threads host
 * threads (0 = all cores).  hw: (h, w) per frame; boxes of frame i = boxes[box_begin[i] .. box_begin[i+1]).
 * Outputs, ready for vis_overlay_draw after upload: leaves (concatenated per-frame arrays), leaf_begin[n_frames + 1]
 * (each frame's base = VisOverlayFrame.group_begin; group_end = base + its box count), tiles (frame index and
 * batch-wide ref ranges filled in), refs.  needed[3] receives the leaf / tile / ref counts; returns the tile count,
 * or VIS_E_CAPACITY (call again with buffers of the needed sizes).                        [host] */
int vis_overlay_plan_batch(int n_frames, const int32_t* hw, const VisBox* boxes, const int32_t* box_begin,
                           VisLeaf* leaves, int64_t leaf_capacity, int32_t* leaf_begin,
                           VisOverlayTile* tiles, int64_t tile_capacity,
                           VisOverlayRef* refs, int64_t ref_capacity, int64_t* needed, int n_threads);

/* Marker sprites.  The marker of a box — white disc, coloured ring, black label (utils/image_utils.py:290-313) — is made
 * of opaque primitives only, so its pixels do not depend on the frame: it is rasterised ONCE per (radius, colour, label)
 * ON THE DEVICE (vis_overlay_sprite_expand -> vis_overlay_tiles -> vis_overlay_draw_cn on a zeroed BGRA canvas: alpha 255
 * where something was drawn) and every box that shows it whole gets one sprite leaf instead of its ~600 leaves.  Markers
 * that touch the image border are still expanded in place.                                 */
typedef struct VisSprite {
    int32_t  radius;
    uint8_t  b, g, r, pad;
    const char* label;             /* HOST pointer, NUL-terminated, as VisBox.label                                 */
    uint64_t pixels;               /* DEVICE address of the BGRA canvas, rows of 4 * w bytes                        */
    int32_t  w, h, ox, oy;         /* canvas size and the position of the marker centre inside it                   */
} VisSprite;
/* leaves of ONE marker on its own canvas (one group header first, colours with alpha 255); *w / *h / *ox / *oy receive
 * the canvas geometry.  Same return convention as vis_overlay_expand.                     [host] */
int vis_overlay_sprite_expand(int radius, int b, int g, int r, const char* label, VisLeaf* leaves, int capacity,
                              int* needed, int* w, int* h, int* ox, int* oy);
/* Dash stamps.  A dash of a low-confidence box (cv2.line, thickness 2, LINE_AA, utils/image_utils.py:264-283) blends with
 * the frame, so it cannot be a sprite; but WHAT happens to each pixel — an optional opaque write followed by a chain of
 * blends with fixed 8-bit alphas — does not depend on the frame or the colour.  That chain is recorded once per dash
 * geometry ON THE DEVICE (vis_overlay_stamp_expand -> vis_overlay_tiles -> vis_overlay_draw_cn with channels = 8 on a
 * zeroed canvas of 8 bytes per pixel: byte 0 = blend count | 0x80 opaque first | 0x40 overflow, bytes 1..7 = alphas) and
 * replayed by ONE leaf per interior dash instead of its 19.  A table entry with radius == -1 is a stamp: label holds
 * "dx,dy", pixels the record canvas.  Stamps whose record overflowed must not be listed.
 * vis_overlay_stamp_expand: leaves of the dash (0,0)-(dx,dy) on its own canvas, one group header first.   [host] */
int vis_overlay_stamp_expand(int dx, int dy, VisLeaf* leaves, int capacity, int* needed, int* w, int* h, int* ox, int* oy);
/* vis_overlay_plan_batch with a table of rendered sprites / stamps (HOST array; may be empty)      [host] */
int vis_overlay_plan_batch_sprites(int n_frames, const int32_t* hw, const VisBox* boxes, const int32_t* box_begin,
                                   VisLeaf* leaves, int64_t leaf_capacity, int32_t* leaf_begin,
                                   VisOverlayTile* tiles, int64_t tile_capacity,
                                   VisOverlayRef* refs, int64_t ref_capacity, int64_t* needed, int n_threads,
                                   const VisSprite* sprites, int n_sprites);

/* frames / tiles / refs / leaves: DEVICE arrays (<= 65535 frames).  copy_frames != 0: every frame with dst != src
 * is first copied src -> dst (vectorised), then the listed tiles are drawn in place on dst; frames drawn in place
 * (dst == src) are only touched inside listed tiles.                                    [device] */
int vis_overlay_draw(const VisOverlayFrame* frames, int n_frames, int copy_frames,
                     const VisOverlayTile* tiles, int n_tiles, const VisOverlayRef* refs,
                     const VisLeaf* leaves, void* stream);

/* ------------------------------------------------------------------------------------------
 * Image-quality statistics (SURVEY.md 8f "next" row).  Replaces the array work of
 * src/safety/image_quality.py:42-125: cv2.cvtColor(BGR2GRAY), cv2.Laplacian(gray, CV_64F).var(), np.mean(gray).
 * sums[3*i + 0..2] = sum(gray), sum(laplacian), sum(laplacian^2) of frame i, exact int64; the caller finishes
 * variance, mean and the scores on the host.                                            [device] */
typedef struct VisQualityFrame {
    const uint8_t* src;        /* device BGR uint8 HWC */
    int64_t        pitch;
    int32_t        h, w;
} VisQualityFrame;
int vis_quality_stats(const VisQualityFrame* frames, int n_frames, int max_h, int max_w,
                      int64_t* sums, void* stream);

/* ------------------------------------------------------------------------------------------
 * Defect heat-map overlay (SURVEY.md 8f "next" row).  Replaces the array work of create_heatmap_overlay
 * (utils/image_utils.py:364-601): per-defect Gaussian heat with boosts, cv2.GaussianBlur of each defect region and
 * of the whole mask (float32, BORDER_REFLECT_101), max-composite, normalisation by the global maximum,
 * cv2.applyColorMap(COLORMAP_JET) and cv2.addWeighted(img, 0.6, colour, 0.4, 0).  Floating point: specified with a
 * tolerance (+-1 on the 8-bit heat index = <= 2 output levels), not bit-exact.
 * The host mirror of the reference's per-defect scalar code fills VisHeatDefect in list order;
 * kernels: DEVICE float32 Gaussian kernels (cv2.getGaussianKernel values), indexed by koff; jet768: DEVICE BGR table. */
typedef struct VisHeatDefect {
    int32_t kind;                    /* 0: box defect, 1: widespread (whole image, no blur)                      */
    int32_t x, y, w, h;              /* pixel box                                                                 */
    int32_t x1, y1, x2, y2;          /* region [x1,x2) x [y1,y2) the defect is evaluated and blurred on           */
    int32_t ksize, koff, pad;        /* odd blur kernel size (1 = none) and its offset in kernels[]              */
    double  intensity, cx, cy, sigma;
} VisHeatDefect;
typedef struct VisHeatItem {         /* one defect of one frame of a batch                                          */
    VisHeatDefect d;
    int32_t frame;                   /* index into frames[]                                                        */
    int32_t pad;
    int64_t tmp_off;                 /* float offset of its region buffer in tmp[] ((x2-x1)*(y2-y1) floats; box defects
                                        with ksize > 1)                                                            */
    int64_t tab_off;                 /* double offset of its 1-D tables in tabs[]: 3 * ((x2-x1) + (y2-y1)) doubles     */
} VisHeatItem;
typedef struct VisHeatFrame {
    const uint8_t* src;              /* device BGR uint8 HWC                                                        */
    uint8_t*       dst;              /* device BGR uint8 HWC (may not alias src)                                    */
    int64_t        src_pitch, dst_pitch;
    int32_t        h, w;
    int64_t        plane_off;        /* float offset of this frame's h*w plane in heat[] / fa[] / fb[]              */
    int32_t        final_ksize, final_koff;   /* whole-mask blur: odd size (1 = none) and offset in kernels[]        */
} VisHeatFrame;
/* The whole batch in six launches.  frames / items: DEVICE arrays (items in the reference's list order per frame; every
 * region inside its frame, ksize odd and <= 51); max_w / max_h / max_rw / max_rh: maxima over the frames and the item
 * regions (they size the grids); heat / fa / fb: plane_floats floats each; tmp / tabs: as the item offsets say;
 * max_bits: n_frames words.  All scratch is the caller's; heat and max_bits are cleared here.          [device] */
int vis_heatmap_batch(const VisHeatFrame* frames, int n_frames, const VisHeatItem* items, int n_items,
                      int max_w, int max_h, int max_rw, int max_rh, int64_t plane_floats,
                      const float* kernels, const uint8_t* jet768, float* heat, float* fa, float* fb,
                      float* tmp, double* tabs, unsigned int* max_bits, void* stream);

/* ------------------------------------------------------------------------------------------
 * Comparison panel and status stamp (SURVEY.md 8f "next" row 4).
 * create_side_by_side_comparison (utils/image_utils.py:608-686): both frames cv2.resize()d to a height of 800
 * (default interpolation INTER_LINEAR), a 40-row header and a 10-column divider filled with 45, two centred
 * white Hershey labels.  create_status_stamp (utils/image_utils.py:689-739): a 4-channel canvas with a 4-px
 * LINE_8 rectangle and a text.  Integer arithmetic throughout: bit-exact against OpenCV 4.13.
 * ------------------------------------------------------------------------------------------ */
/* what cv::resize does for INTER_LINEAR on 8-bit data (cv: resize.cpp): VIS_RESIZE_COPY when the sizes are equal,
 * VIS_RESIZE_AREA2 when both scale factors are exactly 2 (OpenCV switches to INTER_AREA's (a+b+c+d+2)>>2 fast path),
 * VIS_RESIZE_BILINEAR otherwise.                                                       [host] */
#define VIS_RESIZE_COPY     0
#define VIS_RESIZE_AREA2    1
#define VIS_RESIZE_BILINEAR 2
int vis_resize_linear_mode(int src_h, int src_w, int dst_h, int dst_w);
/* one axis of the fixed-point bilinear resizer (cv: resize.cpp, the xofs/alpha and yofs/beta tables):
 * ofs[d] = first source index, coef[2d], coef[2d+1] = 11-bit weights (saturate_cast<short>(w * 2048)) of ofs[d] and
 * ofs[d] + 1.  is_x != 0 folds the border clamp into the table as OpenCV does for columns; rows keep the raw index
 * (may be -1 or src_size - 1) and the kernel clamps the row it fetches.                 [host] */
int vis_linear_table(int src_size, int dst_size, int is_x, int32_t* ofs, int16_t* coef);

typedef struct VisPanel {        /* cv2.resize(src, (dst_w, dst_h)) placed at (org_x, org_y) of the canvas        */
    const uint8_t* src;          /* device, 3-channel uint8 HWC                                                  */
    int64_t        src_pitch;
    int32_t        src_h, src_w;
    int32_t        dst_h, dst_w;
    int32_t        org_x, org_y;
    int32_t        mode;         /* vis_resize_linear_mode                                                       */
    int32_t        pad;
    const int32_t* xofs;         /* device tables of vis_linear_table (VIS_RESIZE_BILINEAR only)                 */
    const int16_t* alpha;
    const int32_t* yofs;
    const int16_t* beta;
} VisPanel;
/* canvas [h, w, 3] = `fill` everywhere except inside the panels (HOST array, <= 4, non-overlapping); one launch.
 * Replaces utils/image_utils.py:637-679 (two cv2.resize calls, header/divider fill, hstack, vstack).  [device] */
int vis_compose_panels(uint8_t* canvas, int64_t canvas_pitch, int h, int w, int fill,
                       const VisPanel* panels, int n_panels, void* stream);
/* The same for a BATCH of canvases in one launch (a report run composes one panel per inspected image).  canvases is a
 * DEVICE array; the library cannot validate it: every panel's `mode` must be vis_resize_linear_mode of its sizes, its
 * tables those of vis_linear_table, panels of a canvas must not overlap.  max_h / max_w: the largest canvas.  [device] */
typedef struct VisPanelCanvas {
    uint8_t* canvas;             /* device, [h, w, 3] uint8, rows of `pitch` bytes                                */
    int64_t  pitch;
    int32_t  h, w;
    int32_t  fill;               /* 0..255: every byte outside the panels                                         */
    int32_t  n_panels;           /* 0..4                                                                          */
    VisPanel panels[4];
} VisPanelCanvas;
int vis_compose_panels_batch(const VisPanelCanvas* canvases, int n_canvases, int max_h, int max_w, void* stream);

/* generic draw list on top of the overlay rasteriser: each command becomes one group of leaves (like one VisBox),
 * so vis_overlay_tiles / vis_overlay_draw(_cn) take the result unchanged.                */
#define VIS_DRAW_LINE      1     /* cv2.line((x1,y1),(x2,y2), color, thickness, line_type)                        */
#define VIS_DRAW_RECTANGLE 2     /* cv2.rectangle((x1,y1),(x2,y2), color, thickness >= 1, line_type)              */
#define VIS_DRAW_CIRCLE    3     /* cv2.circle((x1,y1), x2 = radius, color, thickness: -1 filled or > 1, LINE_8)   */
#define VIS_DRAW_TEXT      4     /* cv2.putText(text, (x1,y1), FONT_HERSHEY_SIMPLEX, font_scale, color, thickness >= 2) */
typedef struct VisDrawCmd {
    int32_t kind;
    int32_t x1, y1, x2, y2;
    int32_t thickness;
    int32_t line_type;           /* 8 or 16 (LINE / RECTANGLE)                                                    */
    uint8_t color[4];            /* B, G, R, A (A is used by 4-channel canvases only)                             */
    double  font_scale;
    const char* text;            /* HOST pointer, NUL-terminated UTF-8, any length (bytes outside 32..126 draw '?') */
} VisDrawCmd;
/* cv2.getTextSize(text, FONT_HERSHEY_SIMPLEX, font_scale, thickness)[0] (bytes outside 32..126 measure as '?').  [host] */
int vis_text_size(const char* text, double font_scale, int thickness, int* width, int* height);
/* like vis_overlay_expand, for a draw list: n_cmds group headers followed by the ordered leaves.   [host] */
int vis_draw_expand(int img_h, int img_w, const VisDrawCmd* cmds, int n_cmds,
                    VisLeaf* leaves, int capacity, int* needed);
/* vis_overlay_draw for canvases with `channels` = 3 or 4 interleaved bytes per pixel; channels = 8 records blend chains
 * (dash stamps, see VisSprite).                                                       [device] */
int vis_overlay_draw_cn(const VisOverlayFrame* frames, int n_frames, int channels, int copy_frames,
                        const VisOverlayTile* tiles, int n_tiles, const VisOverlayRef* refs,
                        const VisLeaf* leaves, void* stream);

/* ------------------------------------------------------------------------------------------
 * JPEG codec stage on the GPU (SURVEY.md 8f "next" row 1) — binding of NVIDIA's nvJPEG (libnvjpeg.so.12), not a
 * kernel of this library.  Replaces, for JPEG files, Image.open / cv2.imread (utils/image_utils.py:39-41, :170;
 * src/agents/vlm_inspector.py:59) in front of the kernels and cv2.imwrite / img.save(JPEG) (utils/image_utils.py:316;
 * src/agents/vlm_inspector.py:73) behind them.  nvJPEG's IDCT / chroma upsampling differ from libjpeg-turbo's:
 * decoded pixels are specified with a tolerance against the reference's decoders, not bit-exact.
 * The ONE family that owns device memory: nvJPEG allocates work buffers behind the opaque handle.  A handle is not
 * thread-safe: one per thread of use.
 * ------------------------------------------------------------------------------------------ */
typedef struct VisJpeg VisJpeg;
#define VIS_JPEG_BACKEND_DEFAULT    0
#define VIS_JPEG_BACKEND_HYBRID     1   /* Huffman decode on the CPU                                              */
#define VIS_JPEG_BACKEND_GPU_HYBRID 2   /* Huffman decode on the GPU for large batches of baseline streams         */
#define VIS_JPEG_BACKEND_HARDWARE   3   /* the NVJPG engine (baseline, single scan)                                */
#define VIS_JPEG_CSS_444 0
#define VIS_JPEG_CSS_422 1
#define VIS_JPEG_CSS_420 2
/* interpolate_chroma != 0: triangle-filter chroma upsampling (closest to libjpeg-turbo's "fancy upsampling").
 * VIS_E_UNSUPPORTED when the backend does not exist on this GPU.                        [host] */
int  vis_jpeg_create(int backend, int interpolate_chroma, VisJpeg** out);
void vis_jpeg_destroy(VisJpeg* j);
/* size, component count and chroma subsampling (VIS_JPEG_CSS_*, 6 = gray, -1 = other) of a JPEG stream   [host] */
int  vis_jpeg_info(VisJpeg* j, const uint8_t* data, int64_t length, int* width, int* height, int* components,
                   int* subsampling);
/* data: HOST JPEG stream; dst: DEVICE [h, w, 3] uint8 interleaved (RGB, or BGR when bgr != 0)            [device] */
int  vis_jpeg_decode(VisJpeg* j, const uint8_t* data, int64_t length, uint8_t* dst, int64_t dst_pitch, int h, int w,
                     int bgr, void* stream);
/* n streams in one call (nvjpegDecodeBatched); dst[i] must hold the size vis_jpeg_info reports          [device] */
int  vis_jpeg_decode_batch(VisJpeg* j, int n, const uint8_t* const* data, const int64_t* lengths, uint8_t* const* dst,
                           const int64_t* dst_pitch, int bgr, int cpu_threads, void* stream);
/* host buffer size that always holds the encoded stream of an h x w frame                                [host] */
int64_t vis_jpeg_encode_bound(int h, int w);
/* src: DEVICE [h, w, 3] uint8 interleaved; out: HOST buffer.  Synchronises `stream` (the stream of bytes is handed
 * back on the host).  VIS_E_CAPACITY with *length = needed size when `out` is too small.          [device, syncs] */
int  vis_jpeg_encode(VisJpeg* j, const uint8_t* src, int64_t src_pitch, int h, int w, int bgr, int quality,
                     int subsampling, int optimized_huffman, uint8_t* out, int64_t capacity, int64_t* length,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIS_B200_H */
