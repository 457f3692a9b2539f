"""Agreement metrics between two batched searches over the same roots (e.g. the bf16 tensor-core path against the
fp32 path): what `bench.py` reports as `parity` and tests/test_gpu_search_baseline.py asserts.

All inputs are [G, A] visit-count arrays / [G] root values (numpy or torch)."""
import numpy as np


def _np(x):
    return x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)


def visit_agreement(visits_a, visits_b, root_value_a=None, root_value_b=None):
    a, b = _np(visits_a).astype(np.float64), _np(visits_b).astype(np.float64)
    pa, pb = a / a.sum(1, keepdims=True), b / b.sum(1, keepdims=True)
    tv = 0.5 * np.abs(pa - pb).sum(1)                       # total-variation distance of the visit distributions
    out = {"games": int(a.shape[0]),
           "visit_agreement": float(1.0 - tv.mean()),       # 1 = identical child_visits targets
           "identical_visit_counts": float((a == b).all(1).mean()),
           "top_action_agreement": float((a.argmax(1) == b.argmax(1)).mean()),
           "visit_tv_p95": float(np.quantile(tv, 0.95))}
    if root_value_a is not None:
        ra, rb = _np(root_value_a).astype(np.float64), _np(root_value_b).astype(np.float64)
        out["root_value_mae"] = float(np.abs(ra - rb).mean())
        out["root_value_max_err"] = float(np.abs(ra - rb).max())
    return out
