"""CPU restatement of gradient-boosted-tree inference as the reference uses it for its alternative readout head:
``xgb_readout.predict(h_V)`` (rdesign/model/rdesign.py:152-155, rnampnn/model/rnampnn.py:298) on an
``xgboost.XGBClassifier(objective='multi:softmax', num_class=4)`` fitted by ``XGBTrainer``
(rdesign/utils/train.py:51-73).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the checker for ``rmp_forest_predict``; never imported by the product.

Parity pin.  The arithmetic lives in a third-party dependency that is absent from ``/root/reference`` and from this
image: ``xgboost~=2.1.1`` (requirements.txt:7).  PARITY UNPINNED against the library itself; what is pinned is
(a) the model: ``tests/golden/xgb_v0_forest.npz`` holds the trees of the reference's own fitted artefact
``out/checkpoints/RDesign-X/XGB-V0.pkl`` (400 trees = 100 rounds x 4 classes, depth 6, 128 features), decoded by
``tests/golden/make_golden_xgb.py``; (b) the published algorithm of xgboost 2.1 (``include/xgboost/tree_model.h``
``RegTree::GetLeafIndex / GetNext``, ``src/predictor/cpu_predictor.cc`` ``PredValueByOneTree`` /
``PredictBatchByBlockOfRowsKernel``, ``src/objective/multiclass_obj.cu`` ``SoftmaxMultiClassObj::PredTransform``):

* a node with ``left_children == -1`` is a leaf whose value is stored in ``split_conditions``;
* otherwise go left iff ``x[split_indices] < split_conditions`` (fp32 compare); a missing (NaN) feature follows
  ``default_left``;
* tree ``t`` belongs to output group ``tree_info[t]``; the margin of a group is ``base_score`` plus the leaf values of its
  trees, accumulated in fp32 in tree order;
* ``multi:softmax`` predicts the first maximal margin (the softmax is monotone).
"""
from __future__ import annotations

import numpy as np


def leaf_values(forest: dict, X: np.ndarray) -> np.ndarray:
    """[rows, n_trees] fp32: the leaf each tree assigns to each row (``RegTree::GetLeafIndex``)."""
    X = np.asarray(X, dtype=np.float32)
    rows = X.shape[0]
    n_trees = len(forest["tree_root"])
    out = np.empty((rows, n_trees), dtype=np.float32)
    thr, feat = forest["threshold"], forest["feature"]
    left, right, dleft = forest["left"], forest["right"], forest["default_left"]
    ar = np.arange(rows)
    for t in range(n_trees):
        node = np.full(rows, forest["tree_root"][t], dtype=np.int64)
        active = left[node] >= 0
        while active.any():
            n = node[active]
            fv = X[ar[active], feat[n]]
            go_left = np.where(np.isnan(fv), dleft[n] != 0, fv < thr[n])
            node[active] = np.where(go_left, left[n], right[n])
            active = left[node] >= 0
        out[:, t] = thr[node]
    return out


def margins(forest: dict, X: np.ndarray) -> np.ndarray:
    """[rows, n_class] fp32 margins: base_score + leaves of the group's trees, summed in tree order in fp32
    (``PredictBatchByBlockOfRowsKernel``)."""
    lv = leaf_values(forest, X)
    n_class = int(forest["n_class"])
    m = np.full((lv.shape[0], n_class), np.float32(forest["base_score"]), dtype=np.float32)
    for t, c in enumerate(forest["tree_class"]):
        m[:, c] = m[:, c] + lv[:, t]
    return m


def predict(forest: dict, X: np.ndarray) -> np.ndarray:
    """Class per row: first maximal margin (``multi:softmax``)."""
    return margins(forest, X).argmax(axis=1).astype(np.int64)


def random_forest(rng: np.random.Generator, n_trees: int, depth: int, n_feat: int, n_class: int, p_leaf: float = 0.15):
    """A seeded random forest in the flat layout of ``rmpnn_b200.host.xgb_model.Forest`` (ragged trees, random default
    directions): size-independent test cases beside the reference's fitted artefact."""
    thr, feat, left, right, dleft, roots = [], [], [], [], [], []
    for _ in range(n_trees):
        base = len(thr)
        roots.append(base)
        nodes = [(0, 0)]                       # (local id, level)
        local = [[0.0, 0, -1, -1, 0]]
        while nodes:
            i, lvl = nodes.pop()
            if lvl == depth or (lvl > 0 and rng.random() < p_leaf):
                local[i] = [float(rng.normal() * 0.3), 0, -1, -1, 0]
                continue
            l, r = len(local), len(local) + 1
            local += [None, None]
            local[i] = [float(rng.normal()), int(rng.integers(n_feat)), base + l, base + r, int(rng.integers(2))]
            nodes += [(l, lvl + 1), (r, lvl + 1)]
        for t_, f_, l_, r_, d_ in local:
            thr.append(t_); feat.append(f_); left.append(l_); right.append(r_); dleft.append(d_)
    return {
        "threshold": np.asarray(thr, np.float32), "feature": np.asarray(feat, np.int32),
        "left": np.asarray(left, np.int32), "right": np.asarray(right, np.int32),
        "default_left": np.asarray(dleft, np.uint8), "tree_root": np.asarray(roots, np.int32),
        "tree_class": (np.arange(n_trees) % n_class).astype(np.int32), "n_class": n_class, "n_feat": n_feat,
        "base_score": np.float32(0.5),
    }
