"""TEST INFRASTRUCTURE ONLY — CPU restatement of ReinhardNormalizer (tiatoolbox/tools/stainnorm.py:222-367).

  lab_split      :286-312   cv2 8-bit RGB->Lab, float32 channels L / 2.55, a - 128, b - 128
  get_mean_std   :337-367   cv2.meanStdDev per channel
  fit            :251-260
  transform      :262-284   (chan - mean) * (target_std / std) + target_mean per channel
  merge_back     :314-335   L * 2.55, a + 128, b + 128 -> clip -> uint8 -> cv2 8-bit Lab->RGB

The colour conversions and the statistics are OpenCV calls in the reference, and OpenCV is present wherever the tests run, so this
restatement calls cv2 exactly like the reference does; what it restates is the numpy arithmetic around them (dtype promotions included).
Parity status: PINNED — equals the unmodified reference class (loaded by oracle/ref_shim.load_stain) bit for bit on the seeded tiles of
oracle/make_golden_reinhard.py, whose outputs are committed as tests/golden/reinhard_golden.npz.
Only tests/, __graft_entry__.smoke() and bench.py may import this module.
"""
from __future__ import annotations

import cv2
import numpy as np


def lab_split(img: np.ndarray):
    lab = cv2.cvtColor(np.asarray(img).astype("uint8"), cv2.COLOR_RGB2LAB).astype(np.float32)
    c1, c2, c3 = cv2.split(lab)
    c1 /= np.asarray(2.55)     # float64 division, stored back as float32
    c2 -= np.asarray(128.0)
    c3 -= np.asarray(128.0)
    return c1, c2, c3


def get_mean_std(img: np.ndarray):
    means, stds = [], []
    for chan in lab_split(img):
        m, s = cv2.meanStdDev(np.asarray(chan))
        means.append(float(m[0][0]))
        stds.append(float(s[0][0]))
    return tuple(means), tuple(stds)


def merge_back(c1, c2, c3) -> np.ndarray:
    c1 *= 2.55
    c2 += 128.0
    c3 += 128.0
    lab = np.clip(cv2.merge((c1, c2, c3)), 0, 255).astype(np.uint8)
    return cv2.cvtColor(lab, cv2.COLOR_LAB2RGB)


def transform(img: np.ndarray, target_means, target_stds) -> np.ndarray:
    chans = lab_split(img)
    means, stds = get_mean_std(img)
    out = [((c - means[i]) * (target_stds[i] / stds[i])) + target_means[i] for i, c in enumerate(chans)]
    return merge_back(*out)


def level_maps(means, stds, target_means, target_stds) -> np.ndarray:
    """The same arithmetic applied to the 256 possible levels of each 8-bit Lab channel: (3, 256) uint8 — what the device tabulates."""
    lv = np.arange(256, dtype=np.float32)
    c1, c2, c3 = lv.copy(), lv.copy(), lv.copy()
    c1 /= np.asarray(2.55)
    c2 -= np.asarray(128.0)
    c3 -= np.asarray(128.0)
    out = [((c - means[i]) * (target_stds[i] / stds[i])) + target_means[i] for i, c in enumerate((c1, c2, c3))]
    out[0] *= 2.55
    out[1] += 128.0
    out[2] += 128.0
    return np.clip(np.stack(out), 0, 255).astype(np.uint8)
