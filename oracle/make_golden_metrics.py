"""Golden vectors for eta_raw from the REAL reference functions (run in the build container):

    python -m oracle.make_golden_metrics

* ``calculate_metric.calculate_eta_raw`` is imported from /root/reference (pure numpy, importable).
* ``nraw_vs_wer.calculate_eta_raw`` cannot be imported (jiwer / matplotlib missing), so its function source
  (nraw_vs_wer.py:20-62) is extracted with ``ast`` and executed on its own with numpy.
Inputs are seeded; inputs' seeds and the reference outputs go to tests/golden/eta_raw.npz.
"""
import ast
import logging
import os
import sys

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "eta_raw.npz")


def load_variant():
    src = open(os.path.join(REF, "nraw_vs_wer.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "calculate_eta_raw")
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "nraw_vs_wer.py", "exec"), ns)
    return ns["calculate_eta_raw"]


def case(seed, L, T, speech_lo, speech_hi):
    rng = np.random.default_rng(seed)
    clean = np.zeros(L)
    clean[speech_lo:speech_hi] = rng.standard_normal(speech_hi - speech_lo)
    noise = 0.6 * rng.standard_normal(L)
    shap = rng.standard_normal((L, T)) * (0.2 + 2.0 * (np.arange(L)[:, None] % 977 < 40))
    shap[speech_lo:speech_hi] *= (1.5 if seed < 4 else 0.9)
    return clean, noise, shap


def main():
    sys.path.insert(0, REF)
    logging.disable(logging.CRITICAL)
    import calculate_metric as cm

    variant = load_variant()
    cases, out = [(1, 16000, 7, 3000, 9000), (2, 12345, 3, 0, 6000), (3, 32000, 11, 20000, 32000), (4, 24000, 5, 8000, 20000),
                  (5, 48000, 4, 100, 30000)], {}
    params = [(20, 99.0), (20, 90.0), (5, 95.0), (0.0625, 99.0)]
    rows = []
    for seed, L, T, lo, hi in cases:
        clean, noise, shap = case(seed, L, T, lo, hi)
        for seg_ms, pct in params:
            a = cm.calculate_eta_raw(clean, noise, shap, 16000, segment_ms=seg_ms, percentile=pct)
            b = variant(clean, noise, shap, 16000, segment_ms=seg_ms, percentile=pct)
            rows.append([seed, L, T, lo, hi, seg_ms, pct, a, b])
    np.savez_compressed(OUT, rows=np.array(rows, dtype=np.float64))
    print(np.array(rows)[:, -4:])


if __name__ == "__main__":
    main()
