"""Drop-in for the evaluation functions DUPLICATED inside the reference's entry script
(`src/benchmark/iris_evaluate.py:11-265`) — the copy `evaluate_model` (:390) actually calls.

The arithmetic is line for line that of `utils/evaluate.py` and runs in the same GPU kernel (`evaluate.py` here);
what differs between the two reference copies is only conventions, and this module mirrors the script's:

  * the report spells "Easy" (`iris_evaluate.py:244`; `utils/evaluate.py:189` prints "Eeay");
  * an unknown dataset name prints `Unknown dataset: <name>` and returns `(None, None, None)` (:263-265) instead of
    falling through and returning `None`;
  * the old-protocol branch would return `(map, None, None)` (:207-210) — but, like the `utils` copy, it first unpacks
    four values from compute_map's 2-tuple and raises `ValueError` (:208).

`evaluate_features` is the script's evaluation tail (:378-398) on librir kernels: normalise -> similarity + ranking ->
revisited mAP.  Patch the script with `from research_image_retrieval_b200.iris_evaluate import *`.
"""
from __future__ import annotations

import numpy as np

from .evaluate import compute_ap, compute_map, revisited_map  # noqa: F401  (same functions in both reference copies)

__all__ = ["compute_ap", "compute_map", "compute_map_and_print", "evaluate_features"]


def compute_map_and_print(dataset, featuretype, mode, ranks, gnd, kappas=[1, 5, 10], verbose=False, li=False):
    """iris_evaluate.py:189-265 — same prints, same return conventions."""
    if dataset.startswith('oxford5k') or dataset.startswith('paris6k'):
        map, aps, _, _ = compute_map(ranks, gnd)  # raises ValueError exactly like the reference (:208)
        print('>> {}: mAP {:.2f}'.format(dataset, np.around(map * 100, decimals=2)))
        return np.around(map * 100, decimals=2), None, None
    if not (dataset.startswith('roxford5k') or dataset.startswith('rparis6k')):
        print(f"Unknown dataset: {dataset}")
        return None, None, None
    res = revisited_map(ranks, gnd, kappas, li=li)          # Easy, Medium, Hard in one launch
    maps = [np.around(r[0] * 100, decimals=2) for r in res]
    print('>> Test Dataset: {} *** Feature Type: {} >>'.format(dataset, featuretype))
    print('>> mAP Easy: {}, Medium: {}, Hard: {}'.format(*maps))
    print('>> mP@k{} Easy: {}, Medium: {}, Hard: {}'.format(kappas, *[np.around(r[2] * 100, decimals=2) for r in res]))
    if verbose:
        print('>> Query aps: >>\nEasy: {}\nMedium: {}\nHard: {}'.format(*[np.around(r[1] * 100, decimals=2) for r in res]))
    return maps[0], maps[1], maps[2]


def evaluate_features(query_features, gallery_features, gnd, dataset_name, featuretype='IRIS', kappas=[1, 5, 10],
                      verbose=True, dtype="fp32", k=None):
    """The evaluation tail of `evaluate_model` (iris_evaluate.py:378-398): F.normalize both sides, Q.X^T, descending
    ranking, revisited mAP report.  Returns (mapE, mapM, mapH).  `ranks` is produced in the [N, nq] layout compute_map
    documents (utils/evaluate.py:49)."""
    from .search import rank
    ranks = rank(query_features, gallery_features, k=k, dtype=dtype, normalize=True)
    return compute_map_and_print(dataset=dataset_name, featuretype=featuretype, mode='global', ranks=ranks, gnd=gnd,
                                 kappas=kappas, verbose=verbose)
