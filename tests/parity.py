"""Shared parity predicates (BASELINE.json north_star gates)."""
import numpy as np

ID_AGREEMENT = 0.9999     # primary-hit geom/prim id agreement
AOV_ATOL = 1e-4           # normal / uv / depth absolute tolerance ...
AOV_RTOL = 2e-5           # ... plus a relative term: planes holding values >> 1 (uv = +-500 on the checkered
                          # plane, depth ~ 100 at grazing angles) cannot meet 1e-4 absolute in f32
# albedo / mip level are not gated by the north star; they amplify the (in-tolerance) uv difference by the
# texture gradient (a 2048^2 checker image turns 5e-6 in uv into 1e-3 in albedo) and log2 of a derivative
PLANE_ATOL = {"albedo": 5e-3, "mip_level": 2e-3}
AOV_OUTLIER_FRAC = 2e-3
# (checkered_plane: uv = +-500 and t up to ~1000 at the horizon -> a few 1e-3 of its pixels sit at f32 resolution)   # pixels on silhouettes / sphere poles / uv seams where ulp-level input differences are
                          # amplified (acos near +-1, grazing hits): allowed to exceed the tolerance


def luminance(img):
    return 0.2126 * img[..., 0] + 0.7152 * img[..., 1] + 0.0722 * img[..., 2]


def id_agreement(a, b):
    return float((a == b).all(axis=-1).mean())


def aov_close(a, b, same_hit=None, atol=AOV_ATOL):
    """fraction of samples outside atol + rtol*|b| (NaN == NaN counts as equal)"""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    bad = ~(np.abs(a - b) <= atol + AOV_RTOL * np.abs(b)) & ~both_nan
    if bad.ndim == 3:
        bad = bad.any(axis=-1)
    if same_hit is not None:
        bad &= same_hit
    return float(bad.mean())


def assert_first_hit_parity(out, ref):
    same = None
    if out.debug_ids is not None:
        agree = id_agreement(out.debug_ids, ref.debug_ids)
        assert agree >= ID_AGREEMENT, f"primary-hit id agreement {agree:.6f}"
        same = (out.debug_ids == ref.debug_ids).all(axis=-1)
    for plane in ("normals", "uv", "debug_depth", "albedo", "mip_level"):
        a, b = getattr(out, plane), getattr(ref, plane)
        if a is None:
            continue
        frac = aov_close(a, b, same, PLANE_ATOL.get(plane, AOV_ATOL))
        assert frac <= AOV_OUTLIER_FRAC, f"{plane}: {frac:.2e} of pixels outside {PLANE_ATOL.get(plane, AOV_ATOL)} (+2e-5 rel)"


def mean_luminance_z(a, b, spp_a, spp_b):
    """z-score of the difference of mean luminance of two independent renders. The per-image standard error
    is estimated from 8x8 block means (robust to the heavy per-pixel tails of a path tracer)."""
    la, lb = luminance(a), luminance(b)
    d = la - lb
    h, w = d.shape
    bh, bw = h // 8 * 8, w // 8 * 8
    blocks = d[:bh, :bw].reshape(bh // 8, 8, bw // 8, 8).mean(axis=(1, 3)).ravel()
    se = blocks.std(ddof=1) / np.sqrt(blocks.size)
    return float(d[:bh, :bw].mean() / max(se, 1e-30))


def beauty_close(a, b, rel=1e-3):
    """Same seed => same sampler streams => both renders walk the same paths up to float rounding: the median
    per-pixel relative difference is tiny (a few pixels diverge where a rounding flips a hit / branch)."""
    a = np.nan_to_num(np.asarray(a, dtype=np.float64))
    b = np.nan_to_num(np.asarray(b, dtype=np.float64))
    scale = max(float(np.abs(b).mean()), 1e-12)
    return float(np.median(np.abs(a - b))) <= rel * scale
