"""Parity comparators shared by the tests (TEST INFRASTRUCTURE).

Bar (BASELINE.json north_star): FP32 results within 1e-4 relative of the reference; bit-exact for integer /
index outputs.  Two correct FP32 implementations of the blend can legitimately disagree on a *threshold
decision* (alpha vs 1/255, T vs 1e-4) when the compared quantity is within rounding of the threshold; those
cases are isolated with the oracle's `ambiguous` mask (images) or bounded by an outlier fraction (gradients,
where one flipped pixel moves a whole Gaussian's sum).
"""
from __future__ import annotations

import numpy as np

RTOL = 1e-4


def to_np(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x)


def rel_report(a, b, name="", atol=None):
    a, b = to_np(a).astype(np.float64), to_np(b).astype(np.float64)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if a.size == 0:
        return dict(name=name, max_abs=0.0, norm_rel=0.0, norm_rel_trim=0.0, frac_bad=0.0, scale=0.0)
    scale = float(np.sqrt(np.mean(b * b))) if b.size else 0.0
    atol = (RTOL * scale) if atol is None else atol
    err = np.abs(a - b)
    bad = err > (RTOL * np.abs(b) + atol)
    nb = float(np.linalg.norm(b))
    nb = nb if nb > 0 else 1.0
    # the same norm WITHOUT the elements already counted as outliers (threshold flips; their number is bounded
    # separately): what is left is the arithmetic difference of the two implementations
    trimmed = float(np.sqrt(np.sum(np.where(bad, 0.0, err) ** 2)) / nb)
    return dict(name=name, max_abs=float(err.max()), norm_rel=float(np.linalg.norm(a - b) / nb), norm_rel_trim=trimmed,
                frac_bad=float(bad.mean()), scale=scale)


def assert_float_parity(a, b, name, mask=None, max_norm_rel=RTOL, max_frac_bad=0.0, atol=None, verbose=True,
                        min_bad_count=0, max_norm_rel_trim=None):
    """Elementwise |a-b| <= 1e-4*|b| + atol (atol defaults to 1e-4 * rms(b)) on all but `max_frac_bad` of the
    elements (or `min_bad_count` elements, whichever is larger: on a 3 000-Gaussian scene ONE flipped threshold
    decision already moves ~4 gradient entries = 7e-4 of a 6 000-element tensor), AND normwise relative error
    <= max_norm_rel.  `max_norm_rel_trim` additionally bounds the normwise error of the NON-outlier elements (a single
    flipped threshold decision on a 15 M-element gradient can carry the whole untrimmed norm).  `mask` (bool, True =
    compare) restricts the check."""
    a, b = to_np(a), to_np(b)
    if mask is not None:
        mask = to_np(mask).astype(bool)
        a, b = a[mask], b[mask]
    assert np.all(np.isfinite(a)), f"{name}: non-finite values in result"
    r = rel_report(a, b, name, atol)
    if verbose:
        print(f"[parity] {name:28s} max_abs={r['max_abs']:.3e} norm_rel={r['norm_rel']:.3e} trimmed={r['norm_rel_trim']:.3e} "
              f"frac_bad={r['frac_bad']:.2e} rms_ref={r['scale']:.3e}")
    assert r["norm_rel"] <= max_norm_rel, f"{name}: normwise relative error {r['norm_rel']:.3e} > {max_norm_rel:.1e}"
    if max_norm_rel_trim is not None:
        assert r["norm_rel_trim"] <= max_norm_rel_trim, (
            f"{name}: trimmed normwise relative error {r['norm_rel_trim']:.3e} > {max_norm_rel_trim:.1e}")
    allowed = max(max_frac_bad, (min_bad_count + 0.5) / max(1, a.size))
    assert r["frac_bad"] <= allowed, f"{name}: {r['frac_bad']:.3e} of elements outside 1e-4 (allowed {allowed:.1e})"
    return r


def assert_int_equal(a, b, name, max_frac_bad=0.0, mask=None):
    a, b = to_np(a), to_np(b)
    assert a.shape == b.shape, f"{name}: shape {a.shape} vs {b.shape}"
    if mask is not None:
        mask = to_np(mask).astype(bool)
        a, b = a[mask], b[mask]
    frac = float((a != b).mean()) if a.size else 0.0
    print(f"[parity] {name:28s} int mismatches={int((a != b).sum())} of {a.size}")
    assert frac <= max_frac_bad, f"{name}: {frac:.3e} of integer elements differ (allowed {max_frac_bad:.1e})"
