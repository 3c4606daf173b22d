"""The image-quality oracle (numpy restatement of src/safety/image_quality.py + the two cv2 routines it calls)
against results captured from the reference's own ImageQualityAssessment, and against cv2 itself when importable."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import quality as Q

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
VAR_RTOL = 1e-12          # numpy's two-pass variance vs other float64 evaluation orders of the same integers


def quality_frames(arrays):
    import make_goldens
    cases = dict(make_goldens.quality_cases())
    cases["mouri_bgr"] = np.ascontiguousarray(arrays["mouri_rgb"][:, :, ::-1])
    return cases


def same_result(got: dict, want: dict, rtol: float):
    assert set(got) == set(want)
    for k in ("quality_score", "quality_passed"):
        assert got[k] == want[k], k
    for sec in ("sharpness", "brightness", "resolution"):
        assert set(got[sec]) == set(want[sec]), sec
        for k, v in want[sec].items():
            if k == "laplacian_variance":
                assert got[sec][k] == pytest.approx(v, rel=rtol, abs=0.0), k
            else:
                assert got[sec][k] == v, (sec, k)


def test_oracle_reproduces_the_reference_results(goldens, arrays):
    frames = quality_frames(arrays)
    assert len(goldens["quality"]) == len(frames)
    for rec in goldens["quality"]:
        bgr = frames[rec["name"]]
        assert hashlib.sha256(np.ascontiguousarray(bgr).tobytes()).hexdigest() == rec["input_sha256"], rec["name"]
        same_result(Q.assess(bgr), rec["result"], VAR_RTOL)


def test_oracle_against_cv2_directly():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for shape in [(37, 53), (1, 1), (1, 9), (9, 1), (2, 2), (480, 640), (333, 517)]:
        bgr = rng.integers(0, 256, shape + (3,), dtype=np.uint8)
        gray = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
        assert np.array_equal(Q.gray_from_bgr(bgr), gray), shape
        assert np.array_equal(Q.laplacian(gray).astype(np.float64), cv2.Laplacian(gray, cv2.CV_64F)), shape
        sg, sl, sl2 = Q.stats(bgr)
        assert sg == int(gray.astype(np.int64).sum())
        lap = cv2.Laplacian(gray, cv2.CV_64F)
        assert sl == int(lap.sum()) and sl2 == int((lap * lap).sum())


def test_host_finish_from_exact_sums(goldens, arrays):
    """image_quality.result_from_sums (host half of the product) fed with the oracle's exact sums."""
    from vision_inspection_system_b200 import image_quality as IQ
    frames = quality_frames(arrays)
    for rec in goldens["quality"]:
        bgr = frames[rec["name"]]
        h, w = bgr.shape[:2]
        same_result(IQ.result_from_sums(w, h, *Q.stats(bgr)), rec["result"], 1e-9)
