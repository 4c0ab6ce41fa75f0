"""Host side of the tree-ensemble readout head: reads the model the reference fits and pickles
(``XGBTrainer.on_fit_end``, rdesign/utils/train.py:56-73 -> ``XGB-V{version}.pkl``; loaded again in
``RNAModel.on_load_checkpoint``, rdesign/model/rdesign.py:196-203) WITHOUT the xgboost package, flattens it, and predicts
on the GPU through ``rmp_forest_predict`` so that ``predict()`` keeps ``h_V`` on the device
(reference: ``xgb_readout.predict(h_V.cpu().numpy())``, rdesign/model/rdesign.py:153).

Formats: the sklearn-wrapper pickle (the Booster's state is one UBJSON buffer, ``Booster.__getstate__`` of xgboost >= 1.6)
and xgboost's JSON / UBJSON model files (``Booster.save_model``).  Only ``gbtree`` with numerical splits is supported
(what ``XGBClassifier(objective='multi:softmax')`` with the reference's arguments produces); anything else raises."""
from __future__ import annotations

import io
import json
import pickle
import struct
from dataclasses import dataclass
from typing import Optional

import numpy as np

_NUM = {'i': '>b', 'U': '>B', 'I': '>h', 'l': '>i', 'L': '>q', 'd': '>f', 'D': '>d'}
_ARR = {'i': '>i1', 'U': '>u1', 'I': '>i2', 'l': '>i4', 'L': '>i8', 'd': '>f4', 'D': '>f8'}


class _UBJ:
    """Minimal UBJSON (draft 12) reader with the optimised ``[$type#count`` containers xgboost writes."""

    def __init__(self, buf):
        self.b = memoryview(bytes(buf))
        self.p = 0

    def _tag(self) -> str:
        c = chr(self.b[self.p])
        self.p += 1
        return c

    def _peek(self) -> str:
        return chr(self.b[self.p])

    def _num(self, t: str):
        fmt = _NUM[t]
        v = struct.unpack_from(fmt, self.b, self.p)[0]
        self.p += struct.calcsize(fmt)
        return v

    def _str(self) -> str:
        n = self._num(self._tag())
        s = bytes(self.b[self.p:self.p + n]).decode()
        self.p += n
        return s

    def _header(self):
        typ = cnt = None
        if self._peek() == '$':
            self.p += 1
            typ = self._tag()
        if self._peek() == '#':
            self.p += 1
            cnt = self._num(self._tag())
        return typ, cnt

    def value(self, t: Optional[str] = None):
        t = t or self._tag()
        if t in _NUM:
            return self._num(t)
        if t == 'S':
            return self._str()
        if t == 'C':
            return self._tag()
        if t in 'TFZ':
            return {'T': True, 'F': False, 'Z': None}[t]
        if t == '{':
            typ, cnt = self._header()
            out = {}
            if cnt is not None:
                for _ in range(cnt):
                    k = self._str()
                    out[k] = self.value(typ)
                return out
            while self._peek() != '}':
                k = self._str()
                out[k] = self.value()
            self.p += 1
            return out
        if t == '[':
            typ, cnt = self._header()
            if cnt is not None:
                if typ in _ARR:
                    a = np.frombuffer(self.b, dtype=_ARR[typ], count=cnt, offset=self.p)
                    self.p += a.nbytes
                    return a
                return [self.value(typ) for _ in range(cnt)]
            out = []
            while self._peek() != ']':
                out.append(self.value())
            self.p += 1
            return out
        raise ValueError(f"UBJSON: unknown type marker {t!r} at byte {self.p - 1}")


def decode_ubjson(buf) -> dict:
    return _UBJ(buf).value()


class _Stub:
    def __setstate__(self, state):
        self.__dict__.update(state)


class _StubUnpickler(pickle.Unpickler):
    """Unpickles ``xgboost.sklearn.XGBClassifier`` / ``xgboost.core.Booster`` as attribute bags (no xgboost needed)."""

    def find_class(self, module, name):
        if module == "xgboost" or module.startswith("xgboost."):
            return type(name, (_Stub,), {})
        return super().find_class(module, name)


@dataclass
class Forest:
    """Flat forest: node arrays over all trees (children are global node ids, -1 = leaf; a leaf's value sits in
    ``threshold`` like in xgboost's ``split_conditions``), per-tree root and output group."""
    threshold: np.ndarray        # [n_nodes] fp32
    feature: np.ndarray          # [n_nodes] int32
    left: np.ndarray             # [n_nodes] int32
    right: np.ndarray            # [n_nodes] int32
    default_left: np.ndarray     # [n_nodes] uint8
    tree_root: np.ndarray        # [n_trees] int32
    tree_class: np.ndarray       # [n_trees] int32
    n_class: int
    n_feat: int
    base_score: float

    def as_dict(self) -> dict:
        return dict(self.__dict__)

    @property
    def n_trees(self) -> int:
        return int(self.tree_root.shape[0])


def forest_from_model(model: dict) -> Forest:
    """``model`` = the decoded JSON / UBJSON document (with or without the outer ``Model`` / ``Config`` wrapper)."""
    if "Model" in model:
        model = model["Model"]
    learner = model["learner"]
    gb = learner["gradient_booster"]
    if gb.get("name") != "gbtree":
        raise ValueError(f"unsupported booster {gb.get('name')!r} (only gbtree)")
    lmp = learner["learner_model_param"]
    n_class = max(int(lmp.get("num_class", 0)), 1)
    n_feat = int(lmp["num_feature"])
    base = float(lmp.get("base_score", 0.5))
    obj = learner.get("objective", {}).get("name", "")
    if obj in ("binary:logistic", "reg:logistic"):
        base = float(np.log(base / (1.0 - base)))       # ProbToMargin of the logistic objectives
    elif not (obj.startswith("multi:") or obj in ("reg:squarederror", "binary:logitraw", "")):
        raise ValueError(f"unsupported objective {obj!r}")
    trees = gb["model"]["trees"]
    info = np.asarray(gb["model"]["tree_info"], dtype=np.int32)
    thr, feat, left, right, dl, roots = [], [], [], [], [], []
    off = 0
    for t in trees:
        st = np.asarray(t.get("split_type", []))
        if st.size and st.any():
            raise ValueError("categorical splits are not supported")
        lc = np.asarray(t["left_children"], dtype=np.int32)
        rc = np.asarray(t["right_children"], dtype=np.int32)
        n = lc.shape[0]
        roots.append(off)
        thr.append(np.asarray(t["split_conditions"], dtype=np.float32))
        feat.append(np.where(lc >= 0, np.asarray(t["split_indices"], dtype=np.int32), 0).astype(np.int32))
        left.append(np.where(lc >= 0, lc + off, -1).astype(np.int32))
        right.append(np.where(lc >= 0, rc + off, -1).astype(np.int32))
        dl.append(np.asarray(t["default_left"], dtype=np.uint8))
        off += n
    f = Forest(np.concatenate(thr), np.concatenate(feat), np.concatenate(left), np.concatenate(right),
               np.concatenate(dl), np.asarray(roots, np.int32), info, n_class, n_feat, base)
    if f.feature.size and int(f.feature.max()) >= n_feat:
        raise ValueError("split index beyond num_feature")
    if int(info.max(initial=0)) >= n_class:
        raise ValueError("tree_info beyond num_class")
    return f


def load_forest(path_or_file) -> Forest:
    """``.pkl`` (pickled ``XGBClassifier`` / ``Booster``), ``.json`` or ``.ubj`` -> :class:`Forest`."""
    if hasattr(path_or_file, "read"):
        data = path_or_file.read()
    else:
        with open(path_or_file, "rb") as fh:
            data = fh.read()
    if data[:1] == b"\x80":                                   # pickle protocol >= 2
        obj = _StubUnpickler(io.BytesIO(data)).load()
        booster = getattr(obj, "_Booster", obj)
        handle = getattr(booster, "handle", None)
        if handle is None:
            raise ValueError("the pickled estimator is not fitted (no Booster state)")
        return forest_from_model(decode_ubjson(handle))
    if data.lstrip()[:1] == b"{" and data.lstrip()[1:2] in (b'"', b"\n", b" ", b"}"):
        return forest_from_model(json.loads(data.decode()))
    return forest_from_model(decode_ubjson(data))


class GpuForestClassifier:
    """``predict`` of the fitted classifier on the GPU.  Accepts what ``XGBClassifier.predict`` is given in the reference
    (a host numpy array: copied in, result copied out) and CUDA tensors (nothing leaves the device).  The flattened forest
    is uploaded once per device, on first use."""

    def __init__(self, forest: Forest):
        self.forest = forest
        self.n_classes_ = forest.n_class
        self._packed = {}

    @classmethod
    def load(cls, path) -> "GpuForestClassifier":
        return cls(load_forest(path))

    def _on(self, device):
        from .. import ops
        key = str(device)
        if key not in self._packed:
            self._packed[key] = ops.pack_forest(self.forest, device)
        return self._packed[key]

    def margins_and_classes(self, X, device=None):
        import torch
        from .. import ops
        x = torch.as_tensor(X)
        if not x.is_cuda:
            x = x.to(torch.device(device if device is not None else "cuda"))
        x = x.to(torch.float32)
        if x.stride(-1) != 1:
            x = x.contiguous()
        return ops.forest_predict(self._on(x.device), x)

    def predict_device(self, X):
        """CUDA tensor in -> int64 CUDA tensor of classes out."""
        return self.margins_and_classes(X)[1]

    def predict(self, X):
        import torch
        pred = self.predict_device(X)
        return pred if isinstance(X, torch.Tensor) and X.is_cuda else pred.cpu().numpy()


def load_readout_head(path):
    """What the drop-in ``on_load_checkpoint`` hooks assign to ``model.xgb_readout`` for ``XGB-V{version}.pkl``: the
    pickled classifier as a GPU forest (xgboost is not needed; raises ``ValueError`` for an unfitted / unsupported model)."""
    return GpuForestClassifier.load(path)
