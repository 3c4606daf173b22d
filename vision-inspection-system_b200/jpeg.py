"""JPEG codec stage on the GPU: ctypes wrapper of the ``vis_jpeg_*`` entry points (nvJPEG behind the C ABI).

Replaces, for JPEG files and only when a caller asks for ``codec="nvjpeg"``, the host decoders / encoders around the
kernels (Image.open, cv2.imread, cv2.imwrite, img.save — utils/image_utils.py:39-41, :170, :316 and
src/agents/vlm_inspector.py:59, :73 in the reference).  The decoded pixels differ from libjpeg-turbo's by a few levels
(different IDCT and chroma upsampling), so this stage is tolerance-specified and stays opt-in; every other stage of the
path is bit-exact.  Non-JPEG files keep the host codecs.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _native as N

_SUBSAMPLING = {"4:4:4": N.JPEG_CSS_444, "4:2:2": N.JPEG_CSS_422, "4:2:0": N.JPEG_CSS_420,
                "444": N.JPEG_CSS_444, "422": N.JPEG_CSS_422, "420": N.JPEG_CSS_420}
BACKENDS = {"default": N.JPEG_BACKEND_DEFAULT, "hybrid": N.JPEG_BACKEND_HYBRID, "gpu_hybrid": N.JPEG_BACKEND_GPU_HYBRID,
            "hardware": N.JPEG_BACKEND_HARDWARE}


def is_jpeg(data: bytes) -> bool:
    return len(data) > 3 and data[0] == 0xFF and data[1] == 0xD8 and data[2] == 0xFF


def exif_orientation(data: bytes) -> int:
    """The EXIF Orientation tag (1..8) of a JPEG stream, 1 when absent or unreadable.  ``cv2.imread`` applies it by
    default (utils/image_utils.py:170, :347, :629 in the reference go through cv2.imread); nvJPEG and PIL do not."""
    import struct
    pos, n = 2, len(data)
    while pos + 4 <= n and data[pos] == 0xFF:
        marker = data[pos + 1]
        if marker in (0xD8, 0x01) or 0xD0 <= marker <= 0xD7:          # markers without a length
            pos += 2
            continue
        if marker == 0xDA or marker == 0xD9:                          # start of scan / end of image: no EXIF ahead
            break
        seg_len = struct.unpack(">H", data[pos + 2:pos + 4])[0]
        if marker == 0xE1 and data[pos + 4:pos + 10] == b"Exif\0\0":
            tiff = data[pos + 10:pos + 2 + seg_len]
            if len(tiff) < 8 or tiff[:2] not in (b"II", b"MM"):
                return 1
            e = "<" if tiff[:2] == b"II" else ">"
            ifd = struct.unpack(e + "I", tiff[4:8])[0]
            if ifd + 2 > len(tiff):
                return 1
            count = struct.unpack(e + "H", tiff[ifd:ifd + 2])[0]
            for i in range(count):
                ent = tiff[ifd + 2 + 12 * i:ifd + 14 + 12 * i]
                if len(ent) < 12:
                    return 1
                tag, typ = struct.unpack(e + "HH", ent[:4])
                if tag == 0x0112:
                    v = struct.unpack(e + "H", ent[8:10])[0] if typ == 3 else struct.unpack(e + "I", ent[8:12])[0]
                    return v if 1 <= v <= 8 else 1
            return 1
        pos += 2 + seg_len
    return 1


class JpegCodec:
    """One nvJPEG handle on one device (not thread-safe: one codec per thread of use)."""

    def __init__(self, device, backend: str = "default", interpolate_chroma: bool = True):
        self.L = N.lib()
        self.device = torch.device(device)
        self.backend = backend
        self.interpolate_chroma = bool(interpolate_chroma)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(self.L.vis_jpeg_create(BACKENDS[backend], 1 if interpolate_chroma else 0, C.byref(handle)),
                    f"vis_jpeg_create({backend})")
        self._h = handle

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.L.vis_jpeg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ decode
    def info(self, data: bytes):
        """(width, height, components, subsampling code) of a JPEG stream."""
        w, h, nc, css = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        N.check(self.L.vis_jpeg_info(self._h, data, len(data), C.byref(w), C.byref(h), C.byref(nc), C.byref(css)),
                "vis_jpeg_info")
        return w.value, h.value, nc.value, css.value

    def decode(self, data: bytes, bgr: bool = False, pitch_align: int = 16) -> torch.Tensor:
        """JPEG stream -> [H, W, 3] uint8 CUDA tensor (RGB, or BGR like cv2.imread).  Rows are padded to
        ``pitch_align`` bytes so that the fused preprocessing kernels take the frame without a repack."""
        w, h, _, _ = self.info(data)
        out = self._alloc(h, w, pitch_align)
        N.check(self.L.vis_jpeg_decode(self._h, data, len(data), out.data_ptr(), out.stride(0), h, w, 1 if bgr else 0,
                                       C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)), "vis_jpeg_decode")
        return out

    def decode_batch(self, streams, bgr: bool = False, cpu_threads: int | None = None, pitch_align: int = 16) -> list:
        """Many JPEG streams in one ``nvjpegDecodeBatched`` call -> list of [H, W, 3] uint8 CUDA tensors."""
        n = len(streams)
        if n == 0:
            return []
        cpu_threads = cpu_threads or min(n, os.cpu_count() or 1)
        outs = []
        for s in streams:
            w, h, _, _ = self.info(s)
            outs.append(self._alloc(h, w, pitch_align))
        bufs = [np.frombuffer(s, np.uint8) for s in streams]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
        lens = np.array([len(s) for s in streams], np.int64)
        dsts = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
        pitches = np.array([o.stride(0) for o in outs], np.int64)
        N.check(self.L.vis_jpeg_decode_batch(self._h, n, ptrs, lens.ctypes.data_as(C.c_void_p), dsts,
                                             pitches.ctypes.data_as(C.c_void_p), 1 if bgr else 0, cpu_threads,
                                             C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                "vis_jpeg_decode_batch")
        return outs

    def _alloc(self, h: int, w: int, pitch_align: int) -> torch.Tensor:
        pitch = (w * 3 + pitch_align - 1) // pitch_align * pitch_align
        buf = torch.empty((h, pitch), dtype=torch.uint8, device=self.device)
        return buf[:, :w * 3].unflatten(1, (w, 3))

    # ------------------------------------------------------------------ encode
    def encode(self, frame: torch.Tensor, quality: int = 95, subsampling: str = "4:2:0", bgr: bool = True,
               optimize: bool = False) -> bytes:
        """[H, W, 3] uint8 CUDA frame -> JPEG bytes.  Defaults mirror cv2.imwrite (quality 95, 4:2:0, BGR input);
        the agents' ``img.save(format="JPEG", quality=85, optimize=True)`` is ``quality=85, optimize=True, bgr=False``."""
        if frame.dtype != torch.uint8 or not frame.is_cuda or frame.dim() != 3 or frame.shape[2] != 3 or \
                frame.stride(2) != 1 or frame.stride(1) != 3:
            raise ValueError("expected a [H, W, 3] uint8 CUDA frame with contiguous pixels")
        h, w = int(frame.shape[0]), int(frame.shape[1])
        cap = int(self.L.vis_jpeg_encode_bound(h, w))
        out = np.empty(cap, np.uint8)
        length = C.c_int64(0)
        N.check(self.L.vis_jpeg_encode(self._h, frame.data_ptr(), frame.stride(0), h, w, 1 if bgr else 0, int(quality),
                                       _SUBSAMPLING[subsampling], 1 if optimize else 0, out.ctypes.data_as(C.c_void_p), cap,
                                       C.byref(length), C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                "vis_jpeg_encode")
        return out[:length.value].tobytes()
