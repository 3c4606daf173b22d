// vis_jpeg.cpp — JPEG codec stage on the GPU through nvJPEG (SURVEY.md 8f "next" row 1).
//
// Replaces, for JPEG files, the host codecs either side of the kernels: Image.open / cv2.imread in front of the
// preprocessing and overlay kernels (utils/image_utils.py:39-41, :170; src/agents/vlm_inspector.py:59) and
// cv2.imwrite / img.save(JPEG) behind them (utils/image_utils.py:316; src/agents/vlm_inspector.py:73).
// nvJPEG is NVIDIA's library (libnvjpeg.so.12 of the CUDA toolkit): this file is binding code, not a kernel of ours.
// Its IDCT and chroma upsampling are not libjpeg-turbo's, so the decoded pixels are specified with a tolerance
// against the reference's decoders (tests/test_gpu_jpeg.py), never bit-exact; the Python layer keeps the host codecs
// as the default and takes this path only when asked (codec="nvjpeg").
//
// This is the one entry-point family that owns device memory: nvJPEG allocates its own work buffers behind the
// opaque VisJpeg handle (created and destroyed by the caller, one per thread of use).
#include <cstring>
#include <mutex>
#include <vector>

#include <nvjpeg.h>

#include "vis_internal.h"

struct VisJpeg {
    nvjpegHandle_t handle = nullptr;
    nvjpegJpegState_t state = nullptr;            // single-image decodes
    nvjpegJpegState_t batch_state = nullptr;      // batched decodes
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    int backend = 0;
    int batch_size = 0, batch_format = -1, batch_threads = 0;
};

namespace {

const char* status_text(nvjpegStatus_t s) {
    switch (s) {
        case NVJPEG_STATUS_SUCCESS: return "success";
        case NVJPEG_STATUS_NOT_INITIALIZED: return "not initialized";
        case NVJPEG_STATUS_INVALID_PARAMETER: return "invalid parameter";
        case NVJPEG_STATUS_BAD_JPEG: return "bad JPEG";
        case NVJPEG_STATUS_JPEG_NOT_SUPPORTED: return "JPEG not supported";
        case NVJPEG_STATUS_ALLOCATOR_FAILURE: return "allocator failure";
        case NVJPEG_STATUS_EXECUTION_FAILED: return "execution failed";
        case NVJPEG_STATUS_ARCH_MISMATCH: return "architecture mismatch";
        case NVJPEG_STATUS_INTERNAL_ERROR: return "internal error";
        case NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED: return "implementation not supported";
        case NVJPEG_STATUS_INCOMPLETE_BITSTREAM: return "incomplete bitstream";
        default: return "unknown status";
    }
}

int fail(nvjpegStatus_t s, const char* where) {
    vis::set_error("%s: nvJPEG status %d (%s)", where, (int)s, status_text(s));
    if (s == NVJPEG_STATUS_BAD_JPEG || s == NVJPEG_STATUS_INVALID_PARAMETER || s == NVJPEG_STATUS_INCOMPLETE_BITSTREAM)
        return VIS_E_INVALID;
    if (s == NVJPEG_STATUS_JPEG_NOT_SUPPORTED || s == NVJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED ||
        s == NVJPEG_STATUS_ARCH_MISMATCH)
        return VIS_E_UNSUPPORTED;
    return VIS_E_CUDA;
}

#define NVJ(call, where)                                          \
    do {                                                          \
        const nvjpegStatus_t s_ = (call);                         \
        if (s_ != NVJPEG_STATUS_SUCCESS) return fail(s_, where);  \
    } while (0)

}  // namespace

extern "C" {

int vis_jpeg_create(int backend, int interpolate_chroma, VisJpeg** out) {
    if (!out || backend < 0 || backend > 3) {
        vis::set_error("vis_jpeg_create: bad arguments (backend %d)", backend);
        return VIS_E_INVALID;
    }
    *out = nullptr;
    VisJpeg* j = new VisJpeg();
    j->backend = backend;
    const unsigned flags = interpolate_chroma ? NVJPEG_FLAGS_UPSAMPLING_WITH_INTERPOLATION : NVJPEG_FLAGS_DEFAULT;
    nvjpegStatus_t s = nvjpegCreateEx((nvjpegBackend_t)backend, nullptr, nullptr, flags, &j->handle);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegJpegStateCreate(j->handle, &j->state);
    if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegJpegStateCreate(j->handle, &j->batch_state);
    if (s != NVJPEG_STATUS_SUCCESS) {
        const int rc = fail(s, "vis_jpeg_create");
        if (j->state) nvjpegJpegStateDestroy(j->state);
        if (j->handle) nvjpegDestroy(j->handle);
        delete j;
        return rc;
    }
    *out = j;
    return VIS_OK;
}

void vis_jpeg_destroy(VisJpeg* j) {
    if (!j) return;
    if (j->enc_params) nvjpegEncoderParamsDestroy(j->enc_params);
    if (j->enc_state) nvjpegEncoderStateDestroy(j->enc_state);
    if (j->batch_state) nvjpegJpegStateDestroy(j->batch_state);
    if (j->state) nvjpegJpegStateDestroy(j->state);
    if (j->handle) nvjpegDestroy(j->handle);
    delete j;
}

int vis_jpeg_info(VisJpeg* j, const uint8_t* data, int64_t length, int* width, int* height, int* components,
                  int* subsampling) {
    if (!j || !data || length <= 0) {
        vis::set_error("vis_jpeg_info: bad arguments");
        return VIS_E_INVALID;
    }
    int nc = 0, ws[NVJPEG_MAX_COMPONENT] = {0}, hs[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t css = NVJPEG_CSS_UNKNOWN;
    NVJ(nvjpegGetImageInfo(j->handle, data, (size_t)length, &nc, &css, ws, hs), "vis_jpeg_info");
    if (width) *width = ws[0];
    if (height) *height = hs[0];
    if (components) *components = nc;
    if (subsampling) *subsampling = (int)css;
    return VIS_OK;
}

int vis_jpeg_decode(VisJpeg* j, const uint8_t* data, int64_t length, uint8_t* dst, int64_t dst_pitch, int h, int w,
                    int bgr, void* stream) {
    if (!j || !data || length <= 0 || !dst || h <= 0 || w <= 0 || dst_pitch < (int64_t)w * 3) {
        vis::set_error("vis_jpeg_decode: bad arguments");
        return VIS_E_INVALID;
    }
    int iw = 0, ih = 0, nc = 0, css = 0;
    const int rc = vis_jpeg_info(j, data, length, &iw, &ih, &nc, &css);
    if (rc != VIS_OK) return rc;
    if (iw != w || ih != h) {
        vis::set_error("vis_jpeg_decode: stream is %dx%d, destination %dx%d", iw, ih, w, h);
        return VIS_E_INVALID;
    }
    nvjpegImage_t img;
    std::memset(&img, 0, sizeof img);
    img.channel[0] = dst;
    img.pitch[0] = (size_t)dst_pitch;
    NVJ(nvjpegDecode(j->handle, j->state, data, (size_t)length, bgr ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI, &img,
                     (cudaStream_t)stream), "vis_jpeg_decode");
    return VIS_OK;
}

int vis_jpeg_decode_batch(VisJpeg* j, int n, const uint8_t* const* data, const int64_t* lengths, uint8_t* const* dst,
                          const int64_t* dst_pitch, int bgr, int cpu_threads, void* stream) {
    if (!j || n <= 0 || !data || !lengths || !dst || !dst_pitch || cpu_threads < 1) {
        vis::set_error("vis_jpeg_decode_batch: bad arguments");
        return VIS_E_INVALID;
    }
    const nvjpegOutputFormat_t fmt = bgr ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_RGBI;
    if (j->batch_size != n || j->batch_format != (int)fmt || j->batch_threads != cpu_threads) {
        NVJ(nvjpegDecodeBatchedInitialize(j->handle, j->batch_state, n, cpu_threads, fmt), "vis_jpeg_decode_batch(init)");
        j->batch_size = n; j->batch_format = (int)fmt; j->batch_threads = cpu_threads;
    }
    std::vector<nvjpegImage_t> imgs((size_t)n);
    std::vector<size_t> lens((size_t)n);
    std::memset(imgs.data(), 0, sizeof(nvjpegImage_t) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        if (!data[i] || lengths[i] <= 0 || !dst[i]) {
            vis::set_error("vis_jpeg_decode_batch: image %d has no data or no destination", i);
            return VIS_E_INVALID;
        }
        imgs[i].channel[0] = dst[i];
        imgs[i].pitch[0] = (size_t)dst_pitch[i];
        lens[i] = (size_t)lengths[i];
    }
    NVJ(nvjpegDecodeBatched(j->handle, j->batch_state, data, lens.data(), imgs.data(), (cudaStream_t)stream),
        "vis_jpeg_decode_batch");
    return VIS_OK;
}

int64_t vis_jpeg_encode_bound(int h, int w) {
    // generous: raw size plus headers (nvjpegEncodeGetBufferSize needs a handle; callers size host buffers with this)
    return h > 0 && w > 0 ? (int64_t)h * w * 3 + 65536 : VIS_E_INVALID;
}

int vis_jpeg_encode(VisJpeg* j, const uint8_t* src, int64_t src_pitch, int h, int w, int bgr, int quality,
                    int subsampling, int optimized_huffman, uint8_t* out, int64_t capacity, int64_t* length,
                    void* stream) {
    if (!j || !src || h <= 0 || w <= 0 || src_pitch < (int64_t)w * 3 || quality < 1 || quality > 100 || !length ||
        (subsampling != NVJPEG_CSS_444 && subsampling != NVJPEG_CSS_422 && subsampling != NVJPEG_CSS_420)) {
        vis::set_error("vis_jpeg_encode: bad arguments");
        return VIS_E_INVALID;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (!j->enc_state) NVJ(nvjpegEncoderStateCreate(j->handle, &j->enc_state, st), "vis_jpeg_encode(state)");
    if (!j->enc_params) NVJ(nvjpegEncoderParamsCreate(j->handle, &j->enc_params, st), "vis_jpeg_encode(params)");
    NVJ(nvjpegEncoderParamsSetQuality(j->enc_params, quality, st), "vis_jpeg_encode(quality)");
    NVJ(nvjpegEncoderParamsSetSamplingFactors(j->enc_params, (nvjpegChromaSubsampling_t)subsampling, st),
        "vis_jpeg_encode(sampling)");
    NVJ(nvjpegEncoderParamsSetOptimizedHuffman(j->enc_params, optimized_huffman ? 1 : 0, st), "vis_jpeg_encode(huffman)");
    nvjpegImage_t img;
    std::memset(&img, 0, sizeof img);
    img.channel[0] = const_cast<uint8_t*>(src);
    img.pitch[0] = (size_t)src_pitch;
    NVJ(nvjpegEncodeImage(j->handle, j->enc_state, j->enc_params, &img, bgr ? NVJPEG_INPUT_BGRI : NVJPEG_INPUT_RGBI, w, h, st),
        "vis_jpeg_encode");
    size_t len = 0;
    NVJ(nvjpegEncodeRetrieveBitstream(j->handle, j->enc_state, nullptr, &len, st), "vis_jpeg_encode(size)");
    *length = (int64_t)len;
    if (!out || capacity < (int64_t)len) {
        vis::set_error("vis_jpeg_encode: %lld bytes needed, capacity %lld", (long long)len, (long long)capacity);
        return VIS_E_CAPACITY;
    }
    const cudaError_t e = cudaStreamSynchronize(st);          // the bitstream is handed back on the host
    if (e != cudaSuccess) return vis::cuda_fail(e, "vis_jpeg_encode(sync)");
    NVJ(nvjpegEncodeRetrieveBitstream(j->handle, j->enc_state, out, &len, st), "vis_jpeg_encode(retrieve)");
    const cudaError_t e2 = cudaStreamSynchronize(st);
    if (e2 != cudaSuccess) return vis::cuda_fail(e2, "vis_jpeg_encode(sync)");
    *length = (int64_t)len;
    return VIS_OK;
}

}  // extern "C"
