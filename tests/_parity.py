"""Plan-parity metrics shared by the GPU parity tests.

Two figures are reported for every plan comparison:

* ``max_norm``  max |P - Pref| / max |Pref|  -- the north star's "within fp32 relative tolerance 1e-4"
  as SURVEY 8(c) allows it for entries that carry no mass;
* ``elem``      max over the entries with Pref >= 1e-6 * max(Pref) of |P - Pref| / Pref -- the elementwise
  relative error on every entry that carries transport mass.  A plan entry is exp((f + g - C)/eps), so this
  is the absolute error of the exponent: fp32 rounding of k*C (k = log2(e)/eps, |k C| up to ~120 at
  eps = 0.05) alone is ~4e-6, the potentials add their own fp32 rounding per iteration.

Every comparison is appended to ``REPORT`` and written to ``gpurun_out/parity_report.json`` at session end
(tests/conftest.py) so the measured figures travel back from the GPU box.
"""
import numpy as np

RTOL = 1e-4        # north star: plan / loss / fused embedding, fp32 vs the float64 oracle
ELEM_RTOL = 5e-4   # elementwise bound on entries >= 1e-6 * max; measured values are in the report
REPORT = []


def plan_errors(P, Pref, floor=1e-6):
    P = np.asarray(P, dtype=np.float64)
    Pref = np.asarray(Pref, dtype=np.float64)
    diff = np.abs(P - Pref)
    mx = float(np.abs(Pref).max())
    max_norm = float(diff.max() / mx) if mx > 0 else float(diff.max())
    mask = np.abs(Pref) >= floor * mx
    elem = float((diff[mask] / np.abs(Pref[mask])).max()) if mask.any() else 0.0
    return max_norm, elem, int(mask.sum())


def rel(P, Pref, what="", elem_rtol=ELEM_RTOL):
    """max-normalised error (returned, the caller asserts it against RTOL); the elementwise error on the
    entries that carry mass is asserted here against ``elem_rtol`` (None = report only)."""
    max_norm, elem, cnt = plan_errors(P, Pref)
    REPORT.append({"what": what, "shape": list(np.shape(Pref)), "max_norm": max_norm, "elem_ge_1e-6max": elem,
                   "entries_ge_1e-6max": cnt})
    if elem_rtol is not None:
        assert elem < elem_rtol, (what, "elementwise relative error", elem, "max-normalised", max_norm)
    return max_norm
