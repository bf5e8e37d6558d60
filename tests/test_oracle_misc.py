"""Pins the small restatements of oracle/glue.py against cv2."""
import numpy as np

from oracle import glue


def test_bgr_to_gray_restatement_matches_cv2():
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (1500, 1500, 3), dtype=np.uint8)
    assert np.array_equal(glue.bgr_to_gray(img), glue.bgr_to_gray_restated(img))
    # every gray level and the saturated corners
    ramp = np.stack([np.arange(256, dtype=np.uint8)] * 3, -1)[None]
    assert np.array_equal(glue.bgr_to_gray(ramp), glue.bgr_to_gray_restated(ramp))
    corners = np.array([[[0, 0, 0], [255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 255], [255, 255, 0]]], np.uint8)
    assert np.array_equal(glue.bgr_to_gray(corners), glue.bgr_to_gray_restated(corners))
