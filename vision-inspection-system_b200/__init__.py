"""vision-inspection-system_b200 — B200-native image preprocessing engine for the one data-parallel hot path of
Aditya-Somasi/Vision-Inspection-System (inspection frame -> Qwen2-VL pixel_values / image_grid_thw, agent
thumbnails, defect overlay), behind the reference's ``utils/image_utils`` call signatures.

Layout
    csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/vis_b200.h)  -> libvis_b200.so
    _native.py       ctypes binding of the C ABI (fails loudly when the library is missing; no CPU fallback)
    tables.py        host tables (coefficients, normalisation LUT, records, strip plans) via the [host] ABI
    geometry.py      size rules (smart_resize, thumbnail, resize_image)
    engine.py        per-GPU engine: device tables, batch planning, kernel launches
    image_utils.py   drop-in mirror of the reference's utils/image_utils functions on this path
    overlay.py       draw_bounding_boxes box logic -> VisBox -> leaves -> overlay kernel
    compare.py       create_side_by_side_comparison / create_status_stamp host logic (panel list, draw list)
    heatmap.py       create_heatmap_overlay host logic; image_quality.py: assess_image_quality
    jpeg.py          nvJPEG codec stage behind the C ABI (opt-in, tolerance-specified)
    agents.py        the callers either side: encode_image_optimized, build_visual_evidence_images
    sharding.py      image sharding across the GPUs of one box, optional NCCL gather
    synth.py         seeded synthetic workloads of BASELINE.json's configs

The directory name carries a hyphen (it is the repository's name); import it as ``vision_inspection_system_b200``
(the alias package next to it).
"""
from .geometry import (smart_resize, thumbnail_size, resize_image_size, DEFAULT_MIN_PIXELS, DEFAULT_MAX_PIXELS,
                       HUB_MAX_PIXELS, ROW_FLOATS)

__all__ = ["smart_resize", "thumbnail_size", "resize_image_size", "DEFAULT_MIN_PIXELS", "DEFAULT_MAX_PIXELS",
           "HUB_MAX_PIXELS", "ROW_FLOATS"]
__version__ = "0.1.0"
