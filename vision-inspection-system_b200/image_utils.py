"""Drop-in mirror of the reference's ``utils/image_utils`` functions that sit on the accelerated path.

Same names, arguments, return types and error behaviour as the reference:

    load_image(image_path) -> PIL.Image                         utils/image_utils.py:20-43
    resize_image(img, max_dimension=None) -> PIL.Image          utils/image_utils.py:46-78
    draw_bounding_boxes(image_path, boxes, output_path,
                        confidence_threshold="low",
                        criticality="medium") -> Path           utils/image_utils.py:148-317
    create_heatmap_overlay(image_path, defects, output_path,
                           alpha=0.4, ...) -> Path              utils/image_utils.py:320-604  (tolerance-specified)
    create_side_by_side_comparison(original_path, annotated_path,
                                   output_path, labels) -> Path utils/image_utils.py:608-686
    create_status_stamp(verdict, output_path, size) -> Path     utils/image_utils.py:689-739

Every file-level function takes ``codec="host"`` (default: PIL / cv2 codecs exactly as the reference) or
``codec="nvjpeg"`` (JPEG decode / encode on the GPU through nvJPEG; tolerance-specified, see jpeg.py).

plus the entry points the reference delegates to a remote server today (the Qwen2-VL image processor behind
``_encode_image_optimized``, src/agents/vlm_inspector.py:46-88 / src/agents/vlm_auditor.py:85-108):

    agent_thumbnail(img, role) -> PIL.Image                     geometry half of _encode_image_optimized
    preprocess_for_vlm(images, *, min_pixels, max_pixels, role) -> (pixel_values, image_grid_thw)
    normalize = preprocess_for_vlm                              (the "normalize" entry BASELINE.json names)

File decode/encode stays on the host (PIL / cv2 codecs, as in the reference); every resample, normalisation,
patch layout and overlay pixel is computed by the CUDA library.  Nothing here imports ``utils.config`` (no API-key
check, no directory creation at import).  There is no CPU fallback.
"""
from __future__ import annotations

import logging
import warnings
from pathlib import Path

import numpy as np
from PIL import Image

from . import geometry as G

logger = logging.getLogger("vision_inspection_system_b200.image_utils")

MAX_IMAGE_DIMENSION = G.INSPECTOR_MAX_SIZE     # config.max_image_dimension default (utils/config.py:184)

_ROLE_MAX_SIZE = {"inspector": G.INSPECTOR_MAX_SIZE, "auditor": G.AUDITOR_MAX_SIZE}


def _engine():
    from .engine import get_engine
    return get_engine()


def load_image(image_path: Path) -> Image.Image:
    """Load an image file; ``FileNotFoundError`` if absent, ``ValueError`` if it cannot be decoded."""
    image_path = Path(image_path)
    if not image_path.exists():
        raise FileNotFoundError(f"Image not found: {image_path}")
    try:
        img = Image.open(image_path)
        img.load()
        logger.debug("Loaded image: %s, size: %s, mode: %s", image_path.name, img.size, img.mode)
        return img
    except Exception as e:
        raise ValueError(f"Failed to load image: {e}")


def get_image_info(image_path: Path) -> dict:
    """Image metadata, same keys as the reference (utils/image_utils.py:81-101); host only."""
    image_path = Path(image_path)
    img = load_image(image_path)
    return {"path": str(image_path), "filename": image_path.name, "width": img.size[0], "height": img.size[1],
            "mode": img.mode, "format": img.format, "size_bytes": image_path.stat().st_size}


ALLOWED_EXTENSIONS = ["jpg", "jpeg", "png", "bmp", "tiff", "webp"]     # config.allowed_extensions default (utils/config.py:144)
MAX_FILE_SIZE_MB = 10                                                   # config.max_file_size_mb default (utils/config.py:143)


def validate_image(image_path: Path, allowed_extensions: list = None, max_size_mb: float = None):
    """(is_valid, error_message) with the reference's checks and messages (utils/image_utils.py:104-145); host only.
    Defaults mirror the config defaults instead of importing ``utils.config`` (no API-key check at import)."""
    image_path = Path(image_path)
    allowed_extensions = allowed_extensions or ALLOWED_EXTENSIONS
    max_size_mb = max_size_mb or MAX_FILE_SIZE_MB
    if not image_path.exists():
        return False, "File does not exist"
    ext = image_path.suffix.lower().lstrip(".")
    if ext not in allowed_extensions:
        return False, f"Invalid extension '{ext}'. Allowed: {allowed_extensions}"
    size_mb = image_path.stat().st_size / (1024 * 1024)
    if size_mb > max_size_mb:
        return False, f"File too large: {size_mb:.1f}MB (max: {max_size_mb}MB)"
    try:
        img = load_image(image_path)
        if img.size[0] < 10 or img.size[1] < 10:
            return False, "Image too small (minimum 10x10 pixels)"
    except Exception as e:
        return False, f"Invalid image file: {e}"
    return True, None


_CODECS = ("host", "nvjpeg")
# modes Pillow resamples in double precision: (vis_resample_hp kind, bytes per pixel); tobytes() of "I" / "F" is native order
# Byte order of the 16-bit modes AS PILLOW 12.2 RESAMPLES THEM on a little-endian host (Resample.c keys the order on the
# mode name "I;16N" only): "I;16B" words are read little-endian, "I;16N" words big-endian.  Mirrored, not corrected: the
# reference's resize_image returns exactly those bytes (tests/test_gpu_resize.py pins all four names against Pillow).
_HP_MODES = {"I;16": (0, 2), "I;16L": (0, 2), "I;16B": (0, 2), "I;16N": (1, 2), "I": (2, 4), "F": (3, 4)}
_JPEG_SUFFIXES = (".jpg", ".jpeg", ".jpe")


def _check_codec(codec: str) -> None:
    if codec not in _CODECS:
        raise ValueError(f"codec must be one of {_CODECS}, got {codec!r}")


def _imread_cuda(path, codec: str, bgr: bool = True):
    """``cv2.imread(path)`` as a BGR uint8 CUDA tensor, or None when the file cannot be decoded.  ``codec="nvjpeg"``
    decodes JPEG streams on the GPU (a few levels away from libjpeg-turbo, see jpeg.py); anything else, and every file
    with ``codec="host"``, goes through cv2 exactly like the reference."""
    import cv2
    import torch
    _check_codec(codec)
    if codec == "nvjpeg":
        from .jpeg import exif_orientation, is_jpeg
        try:
            data = Path(path).read_bytes()
        except OSError:
            return None
        # cv2.imread rotates by the EXIF Orientation tag, nvJPEG does not: such files (phone cameras) keep the host
        # decoder so that percent boxes land on the same pixels as in the reference
        if is_jpeg(data) and exif_orientation(data) == 1:
            try:
                return _engine().jpeg_codec().decode(data, bgr=bgr)
            except Exception as e:          # CMYK, arithmetic coding, damaged stream: let the host decoder decide
                logger.warning("nvJPEG declined %s (%s); decoding on the host like the reference", path, e)
    img = cv2.imread(str(path))
    if img is None:
        return None
    return torch.from_numpy(img if bgr else np.ascontiguousarray(img[:, :, ::-1])).cuda()


def _imwrite_cuda(path, frame, codec: str) -> None:
    """``cv2.imwrite(path, frame)`` for a BGR uint8 CUDA tensor; ``codec="nvjpeg"`` encodes .jpg/.jpeg on the GPU with
    cv2's defaults (quality 95, 4:2:0)."""
    import cv2
    _check_codec(codec)
    path = Path(path)
    if codec == "nvjpeg" and path.suffix.lower() in _JPEG_SUFFIXES and frame.shape[2] == 3:
        path.write_bytes(_engine().jpeg_codec().encode(frame, quality=95, subsampling="4:2:0", bgr=True))
        return
    cv2.imwrite(str(path), frame.cpu().numpy())


def decode_image(image_path: Path, bgr: bool = False, codec: str = "nvjpeg"):
    """File -> [H, W, 3] uint8 CUDA tensor (RGB by default): the device-side counterpart of ``load_image`` for callers
    that feed ``preprocess_for_vlm`` / ``Engine.annotate`` directly.  Same errors as ``load_image``."""
    image_path = Path(image_path)
    if not image_path.exists():
        raise FileNotFoundError(f"Image not found: {image_path}")
    t = _imread_cuda(image_path, codec, bgr)
    if t is None:
        raise ValueError(f"Failed to load image: {image_path}")
    return t


def _resample_pil(img: Image.Image, size: tuple[int, int], filt: int, box=None, reducing_gap=None) -> Image.Image:
    """``img.resize(size, filt, box, reducing_gap)`` with the resampling (and the ``reduce`` pre-pass a
    ``reducing_gap`` asks for) done on the GPU.  As in Pillow, alpha modes are resampled premultiplied and WITHOUT the
    pre-pass (``Image.resize`` drops ``reducing_gap`` on that branch, PIL:Image.py:2399-2402)."""
    import torch
    full = (0, 0) + img.size
    box = full if box is None else tuple(box)
    if img.size == tuple(size) and box == full:
        return img.copy()
    mode = img.mode
    if mode in ("P", "1"):                           # Pillow forces NEAREST for these, whatever filter is named
        work = img.convert("L") if mode == "1" else img          # "1" is stored as 0 / 255 bytes
        arr = np.frombuffer(work.tobytes(), np.uint8).reshape(work.size[1], work.size[0])
        out = _engine().resize_nearest_u8(torch.from_numpy(arr.copy()).cuda(), size[1], size[0],
                                          None if box == full else box).cpu().numpy()
        res = Image.frombytes(work.mode, tuple(size), out.tobytes())
        if mode == "P":                              # Image._new: the palette and the info dict travel with the result
            if img.palette is not None:
                raw = img.palette.mode if img.palette.mode in ("RGB", "RGBA") else "RGB"
                res.putpalette(img.getpalette(raw), raw)
            res.info = img.info.copy()
            return res
        return res.convert("1", dither=Image.Dither.NONE)
    premultiply = mode in ("LA", "RGBA")             # Pillow resamples these in premultiplied form ...
    if premultiply:
        work, reducing_gap = img, None               # ... and without the reduce pre-pass
    elif mode in ("L", "RGB", "RGBX", "CMYK", "YCbCr", "HSV", "LAB", "La", "RGBa"):
        work = img
    elif mode in _HP_MODES:
        # "I;16", "I", "F": Pillow's double-precision passes (ImagingResample*_16bpc / _32bpc), one channel
        if box != full or reducing_gap is not None:
            raise NotImplementedError(f"box / reducing_gap resampling of mode {mode!r} images is not implemented")
        kind, bpp = _HP_MODES[mode]
        raw = np.frombuffer(img.tobytes(), np.uint8).reshape(img.size[1], img.size[0] * bpp)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)
            dev = torch.from_numpy(raw).cuda()
        out = _engine().resize_hp(dev, size[1], size[0], int(filt), kind)
        return Image.frombytes(mode, tuple(size), out.cpu().numpy().tobytes())
    else:
        raise NotImplementedError(f"image mode {mode!r} is not resampled by this engine")
    bands = len(work.getbands())
    arr = np.frombuffer(work.tobytes(), np.uint8).reshape(work.size[1], work.size[0], bands)   # one host copy (read-only)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", UserWarning)        # torch notes that the array is read-only; it is only copied
        dev = torch.from_numpy(arr).cuda()
    eng = _engine()
    if premultiply:
        eng.alpha_premultiply_(dev, True)
    if reducing_gap is not None:
        out = eng.resize_reducing_u8(dev, size[1], size[0], int(filt), None if box == full else box, reducing_gap)
    elif box != full:
        out = eng.resize_box_u8(dev, size[1], size[0], int(filt), box)
    else:
        out = eng.resize_u8(dev, size[1], size[0], int(filt))
    if premultiply:
        eng.alpha_premultiply_(out, False)
    return Image.frombytes(work.mode, tuple(size), out.cpu().numpy().tobytes())


def pil_thumbnail(img: Image.Image, max_size: int) -> Image.Image:
    """``img.thumbnail((max_size, max_size), LANCZOS)`` (PIL:Image.py:2831-2915) as a function: the JPEG draft request
    (a decoder-side 1/2, 1/4, 1/8 scale for streams that are still unread and >= 4x larger than the request — host
    codec work, done by PIL itself), the integer ``reduce`` pre-pass from 4x downscales on, and the LANCZOS resample over
    the resulting (possibly fractional) box — the last two on the GPU.  Returns ``img`` itself when it already fits."""
    size = G.thumbnail_size(img.size[0], img.size[1], max_size)
    if size is None:
        return img
    box = None
    res = img.draft(None, (int(max_size * 2.0), int(max_size * 2.0)))      # reducing_gap = 2.0, as thumbnail()
    if res is not None:
        box = res[1]
    if img.size == tuple(size):
        return img
    return _resample_pil(img, size, Image.Resampling.LANCZOS, box, 2.0)


def resize_image(img: Image.Image, max_dimension: int = None) -> Image.Image:
    """Fit within ``max_dimension`` keeping the aspect ratio (LANCZOS); returns ``img`` itself when it already fits."""
    max_dimension = max_dimension or MAX_IMAGE_DIMENSION
    width, height = img.size
    target = G.resize_image_size(width, height, max_dimension)
    if target is None:
        return img
    resized = _resample_pil(img, target, Image.Resampling.LANCZOS)
    logger.debug("Resized image from %s to %s", img.size, resized.size)
    return resized


def agent_thumbnail(img: Image.Image, role: str = "inspector", max_size: int | None = None) -> Image.Image:
    """The geometry of ``_encode_image_optimized``: LANCZOS thumbnail to fit 2048 (Inspector) / 1024 (Auditor),
    then RGB for RGBA/P/LA inputs.  The JPEG q85 round trip that follows in the reference is a codec step and is
    not part of this path."""
    max_size = max_size or _ROLE_MAX_SIZE[role]
    if max(img.size) > max_size:
        img = pil_thumbnail(img, max_size)
    if img.mode in ("RGBA", "P", "LA"):
        img = img.convert("RGB")
    return img


def _to_rgb_array(image) -> np.ndarray:
    """PIL image / path / HWC uint8 array -> RGB uint8 HWC (tf:image_processing_backends.py:437-470 semantics)."""
    if isinstance(image, (str, Path)):
        image = load_image(Path(image))
    if isinstance(image, Image.Image):
        if image.mode != "RGB":
            image = image.convert("RGB")
        return np.asarray(image)
    arr = np.asarray(image)
    if arr.dtype != np.uint8 or arr.ndim != 3 or arr.shape[2] != 3:
        raise ValueError("expected a PIL image, a path or an RGB uint8 HWC array")
    return arr


def preprocess_for_vlm(images, *, min_pixels: int = G.DEFAULT_MIN_PIXELS, max_pixels: int = G.DEFAULT_MAX_PIXELS,
                       role: str | None = None, codec: str = "host"):
    """Frames -> (``pixel_values`` float32 CUDA tensor [sum N_i, 1176], ``image_grid_thw`` int64 tensor [B, 3]).

    ``images``: one or a list of PIL images / paths / RGB uint8 HWC arrays / CUDA uint8 HWC tensors.
    ``role``: None, or "inspector"/"auditor" to first apply that agent's thumbnail limit (2048 / 1024, LANCZOS).
    ``codec``: "host" decodes paths with PIL like the reference; "nvjpeg" decodes the JPEG paths of the list in one
    batched GPU call (tolerance-specified, see jpeg.py).
    """
    import torch
    eng = _engine()
    _check_codec(codec)
    if not isinstance(images, (list, tuple)):
        images = [images]
    images = list(images)
    if codec == "nvjpeg":
        from .jpeg import is_jpeg
        streams = {}
        for i, im in enumerate(images):
            if isinstance(im, (str, Path)):
                if not Path(im).exists():
                    raise FileNotFoundError(f"Image not found: {im}")
                data = Path(im).read_bytes()
                if is_jpeg(data):
                    streams[i] = data
        if streams:
            try:
                decoded = eng.jpeg_codec().decode_batch(list(streams.values()))
                for i, t in zip(streams, decoded):
                    images[i] = t
            except Exception as e:
                logger.warning("nvJPEG declined the batch (%s); decoding on the host like the reference", e)
    frames = []
    for im in images:
        if isinstance(im, torch.Tensor):
            t = im if im.is_cuda else im.cuda()
        else:
            t = torch.from_numpy(np.array(_to_rgb_array(im), dtype=np.uint8, order="C")).cuda()       # copy: PIL arrays are read-only
        frames.append(t)
    if role is not None:
        frames = eng.agent_inputs(frames, role)          # thumbnails batched per source geometry
    return eng.preprocess(frames, min_pixels=min_pixels, max_pixels=max_pixels)


normalize = preprocess_for_vlm


def draw_bounding_boxes(image_path: Path, boxes: list, output_path: Path, confidence_threshold: str = "low",
                        criticality: str = "medium", codec: str = "host") -> Path:
    """Annotated copy of ``image_path`` at ``output_path``: dashed/solid 2-px box, numbered marker, as the reference.

    Raises ``ValueError("Failed to load image: ...")`` when the file cannot be read; invalid boxes are skipped with a
    warning (never raised), exactly like the reference.
    """
    dev = _imread_cuda(image_path, codec)
    if dev is None:
        raise ValueError(f"Failed to load image: {image_path}")
    _engine().annotate([dev], [boxes], confidence_threshold, criticality, inplace=True)
    _imwrite_cuda(output_path, dev, codec)
    return output_path


def create_heatmap_overlay(image_path: Path, defects: list, output_path: Path, alpha: float = 0.4,
                           actual_model_size=None, confidence_threshold: str = "low",
                           criticality: str = "medium", codec: str = "host") -> Path:
    """Semi-transparent JET heat map over the defect regions, saved at ``output_path`` (utils/image_utils.py:320-604).

    As in the reference, ``alpha``, ``actual_model_size``, ``confidence_threshold`` and ``criticality`` are accepted and
    unused: every defect is drawn and the blend is fixed at 60 % image / 40 % heat map.  Raises
    ``ValueError("Failed to load image: ...")`` when the file cannot be read.
    """
    logger.info("Creating heatmap overlay for %s", Path(image_path).name)
    img = _imread_cuda(image_path, codec)
    if img is None:
        raise ValueError(f"Failed to load image: {image_path}")
    out = _engine().heatmap(img.contiguous(), defects)
    _imwrite_cuda(output_path, out, codec)
    return output_path


def create_side_by_side_comparison(original_path: Path, annotated_path: Path, output_path: Path,
                                   labels: tuple = ("Original Input", "AI Analysis Layer"), codec: str = "host") -> Path:
    """Side-by-side comparison image (utils/image_utils.py:608-686): both files resized to a height of 800, a labelled
    header bar and a divider.  Raises ``ValueError("Failed to load images for comparison")`` like the reference."""
    logger.info("Creating side-by-side comparison")
    original = _imread_cuda(original_path, codec)
    annotated = _imread_cuda(annotated_path, codec)
    if original is None or annotated is None:
        raise ValueError("Failed to load images for comparison")
    result = _engine().side_by_side(original, annotated, labels)
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    _imwrite_cuda(output_path, result, codec)
    logger.info("Saved comparison image: %s", output_path)
    return output_path


def create_status_stamp(verdict: str, output_path: Path, size: tuple = (300, 100)) -> Path:
    """Status stamp PNG with a transparent background: "SAFE" -> PASSED, "UNSAFE" -> REJECTED, anything else -> REVIEW
    (utils/image_utils.py:689-739)."""
    import cv2
    stamp = _engine().status_stamp(verdict, size)
    output_path = Path(output_path)
    output_path.parent.mkdir(parents=True, exist_ok=True)
    cv2.imwrite(str(output_path), stamp.cpu().numpy())
    logger.info("Saved status stamp: %s", output_path)
    return output_path
