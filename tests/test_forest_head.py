"""Tree-ensemble readout head (reference call sites rdesign/model/rdesign.py:152-155,
rnampnn/model/rnampnn.py:298, model written by rdesign/utils/train.py:56-73).

CPU: the loader (pickle stub + UBJSON / JSON) against the fixture decoded from the reference's own ``XGB-V0.pkl`` and
against an independent encoder; the oracle on a hand-computed case.  GPU: ``rmp_forest_predict`` bit-exact (margins and
classes) against the oracle on the fixture and on random forests, and the drop-in ``predict`` writing the forest's
sequence."""
import io
import json
import os
import pickle
import struct
import sys
import types

import numpy as np
import pytest
import torch

from oracle import forest_oracle as FO
from rmpnn_b200.host import xgb_model as XM

GOLD = os.path.join(os.path.dirname(__file__), "golden", "xgb_v0_forest.npz")
REF_PKL = "/root/reference/out/checkpoints/RDesign-X/XGB-V0.pkl"
KEYS = ("threshold", "feature", "left", "right", "default_left", "tree_root", "tree_class")


def _gold():
    g = np.load(GOLD)
    forest = {k: g[k] for k in KEYS}
    forest.update(n_class=int(g["n_class"]), n_feat=int(g["n_feat"]), base_score=np.float32(g["base_score"]))
    return g, forest


# ---------------------------------------------------------------- a tiny UBJSON writer (independent of the reader)
def _ubj(v) -> bytes:
    if isinstance(v, bool):
        return b"T" if v else b"F"
    if isinstance(v, str):
        e = v.encode()
        return b"S" + b"L" + struct.pack(">q", len(e)) + e
    if isinstance(v, int):
        return b"l" + struct.pack(">i", v)
    if isinstance(v, float):
        return b"d" + struct.pack(">f", v)
    if isinstance(v, np.ndarray):
        t = {np.dtype("float32"): b"d", np.dtype("int32"): b"l", np.dtype("uint8"): b"U", np.dtype("int64"): b"L"}[v.dtype]
        return b"[$" + t + b"#L" + struct.pack(">q", v.size) + v.astype(v.dtype.newbyteorder(">")).tobytes()
    if isinstance(v, list):
        return b"[" + b"".join(_ubj(x) for x in v) + b"]"
    if isinstance(v, dict):
        out = b"{"
        for k, x in v.items():
            e = k.encode()
            out += b"L" + struct.pack(">q", len(e)) + e + _ubj(x)
        return out + b"}"
    raise TypeError(type(v))


def _model_doc(forest, as_arrays: bool):
    """The forest as an xgboost model document (per-tree local child ids, leaf values in split_conditions)."""
    roots = list(forest["tree_root"]) + [len(forest["threshold"])]
    trees = []
    for t in range(len(roots) - 1):
        a, b = roots[t], roots[t + 1]
        lc = np.where(forest["left"][a:b] >= 0, forest["left"][a:b] - a, -1).astype(np.int32)
        rc = np.where(forest["right"][a:b] >= 0, forest["right"][a:b] - a, -1).astype(np.int32)
        tr = {"split_conditions": forest["threshold"][a:b].astype(np.float32), "split_indices": forest["feature"][a:b].astype(np.int32),
              "left_children": lc, "right_children": rc, "default_left": forest["default_left"][a:b].astype(np.uint8),
              "split_type": np.zeros(b - a, np.uint8), "tree_param": {"num_nodes": str(b - a)}}
        trees.append(tr if as_arrays else {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in tr.items()})
    info = forest["tree_class"].astype(np.int32)
    return {"learner": {"learner_model_param": {"base_score": "5E-1", "num_class": str(forest["n_class"]),
                                                "num_feature": str(forest["n_feat"])},
                        "objective": {"name": "multi:softmax"},
                        "gradient_booster": {"name": "gbtree", "model": {"trees": trees,
                                                                          "tree_info": info if as_arrays else info.tolist()}}},
            "version": [2, 1, 1]}


def _same(f: XM.Forest, forest: dict):
    for k in KEYS:
        assert np.array_equal(getattr(f, k), forest[k]), k
    assert f.n_class == forest["n_class"] and f.n_feat == forest["n_feat"] and f.base_score == float(forest["base_score"])


# ---------------------------------------------------------------- CPU
def test_oracle_known_answer():
    """Hand-computed: two stumps and one depth-2 tree, two groups, a missing value in each direction."""
    forest = {
        #            t0: x0<0.5 ? 1.0 : -1.0     t1: x1<0 ? 0.25 : 0.5 (default left)     t2: x0<2 ? (x1<1 ? 3 : 4) : 5
        "threshold": np.asarray([0.5, 1.0, -1.0, 0.0, 0.25, 0.5, 2.0, 1.0, 5.0, 3.0, 4.0], np.float32),
        "feature": np.asarray([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], np.int32),
        "left": np.asarray([1, -1, -1, 4, -1, -1, 7, 9, -1, -1, -1], np.int32),
        "right": np.asarray([2, -1, -1, 5, -1, -1, 8, 10, -1, -1, -1], np.int32),
        "default_left": np.asarray([0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0], np.uint8),
        "tree_root": np.asarray([0, 3, 6], np.int32), "tree_class": np.asarray([0, 1, 0], np.int32),
        "n_class": 2, "n_feat": 2, "base_score": np.float32(0.5),
    }
    X = np.asarray([[0.0, -1.0], [0.5, 0.0], [3.0, np.nan], [np.nan, 2.0]], np.float32)
    m = FO.margins(forest, X)
    # row 0: t0 1.0, t1 0.25, t2 3      row 1: t0 -1 (0.5 < 0.5 false), t1 0.5, t2 3
    # row 2: t0 -1, t1 0.25 (missing -> left), t2 5      row 3: t0 -1 (missing -> right), t1 0.5, t2 5 (missing -> right)
    assert np.array_equal(m, np.asarray([[4.5, 0.75], [2.5, 1.0], [4.5, 0.75], [4.5, 1.0]], np.float32))
    assert FO.predict(forest, X).tolist() == [0, 0, 0, 0]


def test_fixture_is_consistent_with_the_oracle():
    g, forest = _gold()
    assert forest["tree_root"].shape[0] == 400 and forest["n_class"] == 4 and forest["n_feat"] == 128
    assert g["sklearn_params"].tolist() == [100, 6, 4]                       # rdesign/model/rdesign.py:70-76 defaults
    assert np.array_equal(forest["tree_class"], np.arange(400) % 4)
    m = FO.margins(forest, g["X"][:96])
    assert np.array_equal(m, g["margins"][:96]) and np.array_equal(m.argmax(1), g["pred"][:96])


@pytest.mark.skipif(not os.path.exists(REF_PKL), reason="reference artefact only exists in the build container")
def test_loader_reads_the_reference_pickle():
    _, forest = _gold()
    _same(XM.load_forest(REF_PKL), forest)


def test_loader_pickle_ubjson_and_json_round_trip(tmp_path, monkeypatch):
    rng = np.random.default_rng(4)
    forest = FO.random_forest(rng, n_trees=9, depth=4, n_feat=17, n_class=3)
    # the sklearn wrapper pickle: a module-level class of a module named xgboost.sklearn holding a Booster whose state is
    # one UBJSON buffer {"Config": ..., "Model": ...}
    pkg, sk, core = types.ModuleType("xgboost"), types.ModuleType("xgboost.sklearn"), types.ModuleType("xgboost.core")
    XGBClassifier = type("XGBClassifier", (), {"__module__": "xgboost.sklearn"})
    Booster = type("Booster", (), {"__module__": "xgboost.core"})
    sk.XGBClassifier, core.Booster = XGBClassifier, Booster
    for name, mod in (("xgboost", pkg), ("xgboost.sklearn", sk), ("xgboost.core", core)):
        monkeypatch.setitem(sys.modules, name, mod)
    clf, bst = XGBClassifier(), Booster()
    bst.handle = bytearray(_ubj({"Config": {"learner": {}}, "Model": _model_doc(forest, as_arrays=True)}))
    clf._Booster, clf.n_classes_ = bst, 3
    blob = pickle.dumps(clf, protocol=4)
    for name in ("xgboost", "xgboost.sklearn", "xgboost.core"):
        monkeypatch.delitem(sys.modules, name)
    _same(XM.load_forest(io.BytesIO(blob)), forest)
    p = tmp_path / "model.json"
    p.write_text(json.dumps(_model_doc(forest, as_arrays=False)))
    _same(XM.load_forest(str(p)), forest)
    p = tmp_path / "model.ubj"
    p.write_bytes(_ubj(_model_doc(forest, as_arrays=True)))
    _same(XM.load_forest(str(p)), forest)


def test_loader_rejects_what_it_does_not_support():
    forest = FO.random_forest(np.random.default_rng(1), 2, 2, 4, 2)
    doc = _model_doc(forest, as_arrays=False)
    doc["learner"]["gradient_booster"]["name"] = "gblinear"
    with pytest.raises(ValueError, match="gbtree"):
        XM.forest_from_model(doc)
    doc = _model_doc(forest, as_arrays=False)
    doc["learner"]["gradient_booster"]["model"]["trees"][0]["split_type"][0] = 1
    with pytest.raises(ValueError, match="categorical"):
        XM.forest_from_model(doc)
    unfitted = types.SimpleNamespace(n_estimators=100)
    with pytest.raises(ValueError, match="not fitted"):
        XM.load_forest(io.BytesIO(pickle.dumps(unfitted)))


# ---------------------------------------------------------------- GPU
def _device_forest(forest):
    from rmpnn_b200 import ops
    f = XM.Forest(**{k: forest[k] for k in KEYS}, n_class=forest["n_class"], n_feat=forest["n_feat"],
                  base_score=float(forest["base_score"]))
    return f, ops.pack_forest(f, torch.device("cuda"))


@pytest.mark.gpu
def test_forest_kernel_matches_fixture_bit_for_bit():
    from rmpnn_b200 import ops
    g, forest = _gold()
    _, dev = _device_forest(forest)
    m, pred = ops.forest_predict(dev, torch.from_numpy(g["X"]).cuda())
    assert torch.equal(m.cpu(), torch.from_numpy(g["margins"]))
    assert torch.equal(pred.cpu(), torch.from_numpy(g["pred"]))
    # a strided view (embeddings inside a wider buffer) and the classes-only call
    wide = torch.zeros(g["X"].shape[0], 256, device="cuda")
    wide[:, :128] = torch.from_numpy(g["X"]).cuda()
    m2, pred2 = ops.forest_predict(dev, wide[:, :128], want_margins=False)
    assert m2 is None and torch.equal(pred2, pred)


@pytest.mark.gpu
@pytest.mark.parametrize("n_trees,depth,n_feat,n_class,rows", [(1, 1, 1, 1, 3), (37, 5, 60, 4, 1000), (64, 10, 300, 33, 257),
                                                               (700, 3, 128, 7, 4099), (12, 6, 128, 2, 0)])
def test_forest_kernel_matches_oracle_on_random_forests(n_trees, depth, n_feat, n_class, rows):
    from rmpnn_b200 import ops
    rng = np.random.default_rng(n_trees * 131 + depth)
    forest = FO.random_forest(rng, n_trees, depth, n_feat, n_class)
    X = rng.normal(size=(rows, n_feat)).astype(np.float32)
    if rows:
        X[rng.random(X.shape) < 0.1] = np.nan
        X[0] = forest["threshold"][: n_feat] if forest["threshold"].shape[0] >= n_feat else X[0]     # values equal to thresholds
    _, dev = _device_forest(forest)
    m, pred = ops.forest_predict(dev, torch.from_numpy(X).cuda())
    ref = FO.margins(forest, X)
    assert torch.equal(m.cpu(), torch.from_numpy(ref))
    assert torch.equal(pred.cpu(), torch.from_numpy(ref.argmax(1).astype(np.int64)) if rows else torch.empty(0, dtype=torch.int64))


@pytest.mark.gpu
def test_classifier_interface_and_dropin_predict(tmp_path):
    """``xgb_readout.predict`` as the reference calls it (numpy in / numpy out) and the drop-in ``RNAModel.predict``
    writing the forest's sequences (rdesign/model/rdesign.py:144-172) with ``h_V`` kept on the device."""
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200.synthetic import synthetic_backbones
    g, forest = _gold()
    f, _ = _device_forest(forest)
    clf = XM.GpuForestClassifier(f)
    out = clf.predict(g["X"][:50])
    assert isinstance(out, np.ndarray) and np.array_equal(out, g["pred"][:50])
    assert torch.equal(clf.predict(torch.from_numpy(g["X"][:50]).cuda()).cpu(), torch.from_numpy(g["pred"][:50]))
    torch.manual_seed(3)
    model = RNAModel().cuda().eval()
    model.xgb_readout = clf
    X, S, mask = synthetic_backbones(3, 48, 5, lengths=[48, 20, 9])
    lengths = torch.tensor([48, 20, 9])
    model.predict((X, S, mask, lengths, ["a", "b", "c"]), 0, str(tmp_path), "submit.csv")
    with torch.no_grad():
        h_V, _ = model(X.cuda(), S.cuda(), mask.cuda())
    ref = FO.predict(forest, h_V.cpu().numpy())
    lines = (tmp_path / "submit.csv").read_text().strip().split("\n")
    assert lines[0] == "pdb_id,seq"
    start = 0
    for line, n, name in zip(lines[1:], [48, 20, 9], "abc"):
        assert line == f"{name}," + "".join("AUCG"[i] for i in ref[start:start + n])
        start += n
