"""Fixture of the tree-ensemble readout head: decodes the reference's own fitted artefact
``/root/reference/out/checkpoints/RDesign-X/XGB-V0.pkl`` (a pickled ``xgboost.XGBClassifier``; 100 rounds x 4 classes,
depth 6, 128 features) with a decoder written for this script only (independent of the product loader
``rmpnn_b200/host/xgb_model.py``, which the tests check against this file), flattens the trees and stores them with a
seeded input block and the oracle's margins / classes for it.

    python tests/golden/make_golden_xgb.py        (build container only: needs /root/reference)

xgboost itself is not installed here, so the expected outputs come from ``oracle/forest_oracle.py`` (the published
predictor algorithm), not from the library: see that file's header."""
import os
import pickle
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import forest_oracle as FO          # noqa: E402

SRC = "/root/reference/out/checkpoints/RDesign-X/XGB-V0.pkl"
SIZES = {'i': 1, 'U': 1, 'I': 2, 'l': 4, 'L': 8, 'd': 4, 'D': 8}
FMT = {'i': 'b', 'U': 'B', 'I': 'h', 'l': 'i', 'L': 'q', 'd': 'f', 'D': 'd'}


def parse(b, p=0, t=None):
    """Recursive-descent UBJSON: returns (value, next position)."""
    if t is None:
        t = chr(b[p]); p += 1
    if t in FMT:
        return struct.unpack_from('>' + FMT[t], b, p)[0], p + SIZES[t]
    if t == 'S':
        n, p = parse(b, p)
        return b[p:p + n].decode(), p + n
    if t in 'TFZ':
        return {'T': True, 'F': False, 'Z': None}[t], p
    if t in '{[':
        typ = cnt = None
        if chr(b[p]) == '$':
            typ = chr(b[p + 1]); p += 2
        if chr(b[p]) == '#':
            cnt, p = parse(b, p + 1)
        out = {} if t == '{' else []
        close = '}' if t == '{' else ']'
        i = 0
        while (cnt is None and chr(b[p]) != close) or (cnt is not None and i < cnt):
            if t == '{':
                k, p = parse(b, p, 'S')
                out[k], p = parse(b, p, typ)
            else:
                v, p = parse(b, p, typ)
                out.append(v)
            i += 1
        return out, (p + 1 if cnt is None else p)
    raise ValueError(t)


class Bag:
    def __setstate__(self, st):
        self.__dict__.update(st)


class Unp(pickle.Unpickler):
    def find_class(self, module, name):
        return type(name, (Bag,), {}) if module.startswith("xgboost") else super().find_class(module, name)


def main():
    clf = Unp(open(SRC, "rb")).load()
    doc, _ = parse(bytes(clf._Booster.handle))
    learner = doc["Model"]["learner"]
    assert learner["gradient_booster"]["name"] == "gbtree" and learner["objective"]["name"] == "multi:softmax"
    trees = learner["gradient_booster"]["model"]["trees"]
    info = learner["gradient_booster"]["model"]["tree_info"]
    thr, feat, left, right, dl, roots = [], [], [], [], [], []
    for t in trees:
        off = len(thr)
        roots.append(off)
        assert not any(t["split_type"])
        for c, f, l, r, d in zip(t["split_conditions"], t["split_indices"], t["left_children"], t["right_children"],
                                 t["default_left"]):
            thr.append(c); dl.append(d)
            feat.append(f if l >= 0 else 0)
            left.append(l + off if l >= 0 else -1)
            right.append(r + off if l >= 0 else -1)
    forest = {
        "threshold": np.asarray(thr, np.float32), "feature": np.asarray(feat, np.int32),
        "left": np.asarray(left, np.int32), "right": np.asarray(right, np.int32),
        "default_left": np.asarray(dl, np.uint8), "tree_root": np.asarray(roots, np.int32),
        "tree_class": np.asarray(info, np.int32), "n_class": int(learner["learner_model_param"]["num_class"]),
        "n_feat": int(learner["learner_model_param"]["num_feature"]),
        "base_score": np.float32(float(learner["learner_model_param"]["base_score"])),
    }
    rng = np.random.default_rng(20250)
    X = rng.normal(size=(768, forest["n_feat"])).astype(np.float32)      # LayerNorm-ed embeddings are O(1)
    X[::7, rng.integers(0, forest["n_feat"], 110)] = np.nan               # missing values follow default_left
    X[5] *= 40.0
    m = FO.margins(forest, X)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "xgb_v0_forest.npz")
    np.savez_compressed(out, X=X, margins=m, pred=m.argmax(1).astype(np.int64),
                        sklearn_params=np.asarray([clf.n_estimators, clf.max_depth, clf.n_classes_]), **forest)
    print(out, os.path.getsize(out), "bytes;", len(trees), "trees,", len(thr), "nodes; class histogram",
          np.bincount(m.argmax(1), minlength=4))


if __name__ == "__main__":
    main()
