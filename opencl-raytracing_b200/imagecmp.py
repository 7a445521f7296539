"""Image comparison for the acceptance step around the backend boundary.

* `mse_maxdiff` — the metric of the reference's own harness (`rttest`: MSE and max abs difference over all channels,
  visual-testing/src/rttest/diff.py:64-89, pass iff MSE <= tolerance).
* `flip` — LDR-FLIP (Andersson et al. 2020, "FLIP: A Difference Evaluator for Alternating Images"), which the north
  star names but the reference harness does not contain and `flip_evaluator` is not installable offline: restated
  here from the published algorithm (colour pipeline: YCxCz contrast-sensitivity filtering, Hunt-adjusted L*a*b*,
  HyAB, error redistribution; feature pipeline: edge / point detection on the achromatic channel; default
  67 pixels per degree).
* `tonemap` — HDR radiance -> LDR sRGB for FLIP (exposure, clamp, sRGB transfer).
* `mean_luminance_z` — the north star's "mean luminance within 3 sigma" check between two renders at equal spp.

Pure numpy / scipy: this is caller-side tooling, not part of the render path.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage


def mse_maxdiff(a: np.ndarray, b: np.ndarray):
    d = a.astype(np.float64) - b.astype(np.float64)
    return float(np.mean(d * d)), float(np.max(np.abs(d)))


def luminance(rgb: np.ndarray) -> np.ndarray:
    return 0.2126 * rgb[..., 0] + 0.7152 * rgb[..., 1] + 0.0722 * rgb[..., 2]


def mean_luminance_z(a: np.ndarray, b: np.ndarray, spp: int) -> float:
    """z-score of the difference of the two frames' mean luminance; the per-frame standard error is estimated from the
    pixel-wise difference of the two independent renders (each pixel difference has variance 2 sigma_pixel^2 / spp
    already folded in)."""
    la, lb = luminance(a).astype(np.float64), luminance(b).astype(np.float64)
    d = la - lb
    se = d.std() / np.sqrt(d.size)
    return float(abs(d.mean()) / max(se, 1e-30))


def tonemap(rgb: np.ndarray, exposure: float = 1.0) -> np.ndarray:
    x = np.clip(rgb.astype(np.float64) * exposure, 0.0, 1.0)
    return np.where(x <= 0.0031308, 12.92 * x, 1.055 * np.power(x, 1.0 / 2.4) - 0.055)


# ---- FLIP ---------------------------------------------------------------------------------------------------
_RGB2XYZ = np.array([[0.4124564, 0.3575761, 0.1804375], [0.2126729, 0.7151522, 0.0721750], [0.0193339, 0.1191920, 0.9503041]])
_XYZ2RGB = np.linalg.inv(_RGB2XYZ)
_WHITE = _RGB2XYZ @ np.ones(3)


def _srgb_to_linear(s):
    return np.where(s <= 0.04045, s / 12.92, np.power((s + 0.055) / 1.055, 2.4))


def _xyz_to_ycxcz(xyz):
    n = xyz / _WHITE
    return np.stack([116.0 * n[..., 1] - 16.0, 500.0 * (n[..., 0] - n[..., 1]), 200.0 * (n[..., 1] - n[..., 2])], axis=-1)


def _ycxcz_to_xyz(c):
    y = (c[..., 0] + 16.0) / 116.0
    return np.stack([y + c[..., 1] / 500.0, y, y - c[..., 2] / 200.0], axis=-1) * _WHITE


def _xyz_to_lab(xyz):
    n = xyz / _WHITE
    d = 6.0 / 29.0
    f = np.where(n > d ** 3, np.cbrt(np.maximum(n, 1e-30)), n / (3 * d * d) + 4.0 / 29.0)
    return np.stack([116.0 * f[..., 1] - 16.0, 500.0 * (f[..., 0] - f[..., 1]), 200.0 * (f[..., 1] - f[..., 2])], axis=-1)


def _hunt(lab):
    return np.stack([lab[..., 0], 0.01 * lab[..., 0] * lab[..., 1], 0.01 * lab[..., 0] * lab[..., 2]], axis=-1)


def _hyab(a, b):
    d = a - b
    return np.abs(d[..., 0]) + np.sqrt(d[..., 1] ** 2 + d[..., 2] ** 2)


def _csf_kernels(ppd):
    p = {"A": (1.0, 0.0047, 0.0, 1e-5), "RG": (1.0, 0.0053, 0.0, 1e-5), "BY": (34.1, 0.04, 13.5, 0.025)}
    r = int(np.ceil(3.0 * np.sqrt(0.04 / (2.0 * np.pi ** 2)) * ppd))
    x = np.arange(-r, r + 1) / ppd
    xx, yy = np.meshgrid(x, x)
    d2 = xx ** 2 + yy ** 2
    ks = []
    for a1, b1, a2, b2 in (p["A"], p["RG"], p["BY"]):
        g = a1 * np.sqrt(np.pi / b1) * np.exp(-np.pi ** 2 * d2 / b1) + a2 * np.sqrt(np.pi / b2) * np.exp(-np.pi ** 2 * d2 / b2)
        ks.append(g / g.sum())
    return ks


def _filter_colour(img_ycxcz, kernels):
    f = np.stack([ndimage.convolve(img_ycxcz[..., i], kernels[i], mode="nearest") for i in range(3)], axis=-1)
    rgb = np.clip(_ycxcz_to_xyz(f) @ _XYZ2RGB.T, 0.0, 1.0)
    return _hunt(_xyz_to_lab(rgb @ _RGB2XYZ.T))


def _feature_kernels(ppd, w=0.082):
    sd = 0.5 * w * ppd
    r = int(np.ceil(3.0 * sd))
    x = np.arange(-r, r + 1)
    xx, yy = np.meshgrid(x, x)
    g = np.exp(-(xx ** 2 + yy ** 2) / (2.0 * sd * sd))
    edge = -xx * g
    point = (xx ** 2 / (sd * sd) - 1.0) * g

    def norm(k):
        pos, neg = k[k > 0].sum(), -k[k < 0].sum()
        return np.where(k > 0, k / pos, k / neg)
    return norm(edge), norm(point)


def _features(y, edge, point):
    ex, ey = ndimage.convolve(y, edge, mode="nearest"), ndimage.convolve(y, edge.T, mode="nearest")
    px, py = ndimage.convolve(y, point, mode="nearest"), ndimage.convolve(y, point.T, mode="nearest")
    return np.sqrt(ex ** 2 + ey ** 2), np.sqrt(px ** 2 + py ** 2)


def flip_map(reference_srgb: np.ndarray, test_srgb: np.ndarray, ppd: float = 0.7 * 3840 / 0.7 * np.pi / 180.0) -> np.ndarray:
    """Per-pixel LDR-FLIP error in [0, 1]; inputs are [H, W, 3] sRGB images in [0, 1]."""
    qc, pc, pt, qf = 0.7, 0.4, 0.95, 0.5
    ref = _xyz_to_ycxcz(_srgb_to_linear(np.clip(reference_srgb, 0, 1).astype(np.float64)) @ _RGB2XYZ.T)
    tst = _xyz_to_ycxcz(_srgb_to_linear(np.clip(test_srgb, 0, 1).astype(np.float64)) @ _RGB2XYZ.T)
    ks = _csf_kernels(ppd)
    de = _hyab(_filter_colour(ref, ks), _filter_colour(tst, ks)) ** qc
    green = _hunt(_xyz_to_lab(np.array([0.0, 1.0, 0.0]) @ _RGB2XYZ.T))
    blue = _hunt(_xyz_to_lab(np.array([0.0, 0.0, 1.0]) @ _RGB2XYZ.T))
    cmax = _hyab(green, blue) ** qc
    pccmax = pc * cmax
    dec = np.where(de < pccmax, pt / pccmax * de, pt + (de - pccmax) / (cmax - pccmax) * (1.0 - pt))
    edge, point = _feature_kernels(ppd)
    er, pr = _features((ref[..., 0] + 16.0) / 116.0, edge, point)
    et, ptt = _features((tst[..., 0] + 16.0) / 116.0, edge, point)
    def_ = (np.maximum(np.abs(er - et), np.abs(pr - ptt)) / np.sqrt(2.0)) ** qf
    return np.clip(dec, 0.0, 1.0) ** (1.0 - np.clip(def_, 0.0, 1.0))


def flip(reference_rgb: np.ndarray, test_rgb: np.ndarray, exposure: float = 1.0, hdr: bool = True) -> float:
    """Mean FLIP error. With hdr=True the inputs are linear radiance and are tone-mapped (exposure, clamp, sRGB) first."""
    a = tonemap(reference_rgb, exposure) if hdr else reference_rgb
    b = tonemap(test_rgb, exposure) if hdr else test_rgb
    return float(flip_map(a, b).mean())
