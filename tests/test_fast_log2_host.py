"""
CPU: fast_log2 (csrc/bfg_common.cuh -- 128-entry table of rounded reciprocals + degree-5 series; the log2 behind ln r of every
non-lean read-out, and the constants the lean loops share) compiled for the HOST (bfg_test_fast_log2_host) against np.log2:
absolute error below 2e-15 (+ the rounding of the result) over the whole double range, special inputs -> NaN (every caller turns that into
"outside the table").  tools/sass_fingerprint.py shows no kernel changed.
"""
import numpy as np


def fast_log2(x):
    from baryonforge_b200 import _lib
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    _lib.check(_lib.lib().bfg_test_fast_log2_host(x.size, x.ctypes.data, out.ctypes.data))
    return out


def test_fast_log2_source_on_host():
    rng = np.random.default_rng(11)
    x = np.concatenate([10.0 ** rng.uniform(-300, 300, 400000),           # the whole normal range
                        rng.uniform(0.5, 2.0, 400000),                     # around log2 = 0, where the relative error is hardest
                        2.0 ** np.arange(-1000, 1000, 7.0),                # exact powers of two
                        1.0 + (np.arange(129) / 128.0),                    # the table's interval boundaries
                        np.nextafter(1.0 + (np.arange(129) / 128.0), 0)])
    got = fast_log2(x)
    want = np.log2(x)
    # absolute error: below 2e-15 from table + series (mantissa part in [-1, 1)) plus the rounding of the sum with the exponent
    err = np.abs(got - want)
    assert np.all(err <= 2e-15 + 2 * np.spacing(np.abs(want))), (err.max(), x[np.argmax(err)])
    k = np.arange(-1000.0, 1000.0)
    assert np.abs(fast_log2(2.0 ** k) - k).max() < 2e-15 + 2 * np.spacing(1000.0)
    # what this costs a table coordinate: ln r = 0.5 ln2 log2(r^2) over a step of ln(3e5)/499 -> below 1e-13 cells
    near = (x > 1e-8) & (x < 1e8)                                          # r^2 of any table radius
    assert err[near].max() * 0.5 * np.log(2) / (np.log(3e5) / 499) < 1e-12
    bad = fast_log2(np.array([0.0, -0.0, -1.0, np.inf, -np.inf, np.nan, 5e-324, 2.2250738585072009e-308]))
    assert np.all(np.isnan(bad))
    assert np.isfinite(fast_log2(np.array([2.2250738585072014e-308, 1.7976931348623157e308]))).all()
