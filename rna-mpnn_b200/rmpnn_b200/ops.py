"""Tensor-level wrappers over the C ABI (include/rmpnn_b200.h).

Each function checks that its tensors live on a CUDA device, allocates outputs with
torch (the library never allocates), and enqueues the kernels on torch's current
stream.  No function here computes anything in PyTorch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import EPI_NORMALIZE, EPI_RESID_ADD, EPI_RESID_LN, EPI_SEGSUM, EPI_STORE, MlpDesc, check, lib

HIDDEN = 128
NODE_RAW = 101
EDGE_RAW = 115
MAX_K = 128

# torch.linspace(0, 20, 16) bit patterns (rdesign/model/feature.py:54) -- passed to the kernels
_RBF_MU = (C.c_float * 16)(*torch.linspace(0.0, 20.0, 16).tolist())


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("rmpnn_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# --------------------------------------------------------------------------- graph
@dataclass
class Graph:
    """Valid-only flattened kNN graph of a padded batch (device int32 bookkeeping)."""
    B: int
    N: int
    K: int
    Nv: int
    Ne: int
    node_map: torch.Tensor     # [B*N] dense -> flat (-1 masked)
    node_src: torch.Tensor     # [B*N] flat -> dense (first Nv)
    sinfo: torch.Tensor        # [B,4] L, deg, node_off, edge_off
    row_ptr: torch.Tensor      # [B*N+1] CSR offsets (first Nv+1)
    counts: torch.Tensor       # [2] Nv, Ne
    E_dense: torch.Tensor      # [B,N,K] int64
    D_neighbors: torch.Tensor  # [B,N,K] fp32
    ctr: torch.Tensor          # [Ne] int32 centre flat id
    nbr: torch.Tensor          # [Ne] int32 neighbour flat id
    E_flat: torch.Tensor       # [2,Ne] int64


def knn(X: torch.Tensor, mask: torch.Tensor, top_k: int, atom: int = 0):
    """``RNAFeatures._dist`` (rdesign/model/feature.py:42-50) -> (D_neighbors, E_idx)."""
    _need_cuda(X, mask)
    X, mask = _f32c(X), _f32c(mask)
    B, N, A, _ = X.shape
    K = min(int(top_k), N)
    E = torch.empty(B, N, K, dtype=torch.int64, device=X.device)
    D = torch.empty(B, N, K, dtype=torch.float32, device=X.device)
    check(lib().rmp_knn(_ptr(X), _ptr(mask), B, N, A, atom, K, _ptr(E), _ptr(D), _stream()), "rmp_knn")
    return D, E


def build_graph(X: torch.Tensor, mask: torch.Tensor, top_k: int, sizes=None) -> Graph:
    """kNN + valid-only flattening (rdesign/model/feature.py:170-177, 227-237).

    ``sizes=(Nv, Ne)`` lets a caller that already knows the lengths on the host skip the
    one device->host read of the two counters.
    """
    _need_cuda(X, mask)
    X, mask = _f32c(X), _f32c(mask)
    B, N, A, _ = X.shape
    K = min(int(top_k), N)
    dev = X.device
    i32 = dict(dtype=torch.int32, device=dev)
    node_map = torch.empty(B * N, **i32)
    node_src = torch.empty(B * N, **i32)
    sinfo = torch.empty(B, 4, **i32)
    row_ptr = torch.empty(B * N + 1, **i32)
    counts = torch.empty(2, **i32)
    L = lib()
    check(L.rmp_graph_index(_ptr(mask), B, N, K, _ptr(node_map), _ptr(node_src), _ptr(sinfo), _ptr(row_ptr),
                            _ptr(counts), _stream()), "rmp_graph_index")
    D, E = knn(X, mask, K)
    if sizes is None:
        Nv, Ne = (int(v) for v in counts.tolist())        # the one host sync of the featurizer
    else:
        Nv, Ne = int(sizes[0]), int(sizes[1])
    ctr = torch.empty(Ne, **i32)
    nbr = torch.empty(Ne, **i32)
    E_flat = torch.empty(2, Ne, dtype=torch.int64, device=dev)
    check(L.rmp_edge_list(_ptr(E), _ptr(node_map), _ptr(node_src), _ptr(sinfo), _ptr(row_ptr), _ptr(counts),
                          B, N, K, Ne, _ptr(ctr), _ptr(nbr), _ptr(E_flat), _stream()), "rmp_edge_list")
    return Graph(B, N, K, Nv, Ne, node_map, node_src, sinfo, row_ptr, counts, E, D, ctr, nbr, E_flat)


def raw_features(X: torch.Tensor, g: Graph, want_edges: bool = True, return_frames: bool = False, pad_edges: bool = False):
    """node_raw [Nv,101], edge_raw [Ne,115] (rdesign/model/feature.py:179-220); ``pad_edges`` -> edge_raw [Ne,128]
    with zero columns 115.. (rmp_rdesign_raw_features_k128)."""
    _need_cuda(X)
    X = _f32c(X[:, :, :6, :])
    dev = X.device
    frames = torch.empty(g.B * g.N, 9, dtype=torch.float32, device=dev)
    node_raw = torch.empty(g.Nv, NODE_RAW, dtype=torch.float32, device=dev)
    edge_raw = torch.empty(g.Ne, 128 if pad_edges else EDGE_RAW, dtype=torch.float32, device=dev) if want_edges else None
    fn = lib().rmp_rdesign_raw_features_k128 if pad_edges else lib().rmp_rdesign_raw_features
    check(fn(_ptr(X), g.B, g.N, _ptr(g.node_src), _ptr(g.ctr), _ptr(g.nbr),
             _ptr(g.counts), g.Nv, g.Ne, C.cast(_RBF_MU, C.c_void_p),
             _ptr(frames), _ptr(node_raw), _ptr(edge_raw), _stream()),
          "rmp_rdesign_raw_features")
    if return_frames:
        return node_raw, edge_raw, frames
    return node_raw, edge_raw


# kernel feature order of rmp_edge_embed_bf16: rbf (reference columns 4..99) | quaternion (0..3) |
# direction (100..114) | 13 zero columns
_EDGE_PERM = list(range(4, 100)) + list(range(0, 4)) + list(range(100, 115))


def pack_edge_embedding(weight: torch.Tensor) -> torch.Tensor:
    """edge_embedding.weight [128,115] -> bf16 [128,128] with columns in the fused kernel's order."""
    _need_cuda(weight)
    W = torch.zeros(128, 128, dtype=torch.float32, device=weight.device)
    W[:, :EDGE_RAW] = weight.detach().to(torch.float32)[:, _EDGE_PERM]
    return W.to(torch.bfloat16).contiguous()


def edge_embed_bf16(X: torch.Tensor, frames: torch.Tensor, g: Graph, W_perm16, bias, gain, beta) -> torch.Tensor:
    """Fused tcgen05 edge featurizer: coordinates -> h_E [Ne,128] bf16 (rmp_edge_embed_bf16)."""
    _need_cuda(X, frames, W_perm16, bias, gain, beta)
    X = _f32c(X[:, :, :6, :])
    h_E = torch.empty(g.Ne, HIDDEN, dtype=torch.bfloat16, device=X.device)
    check(lib().rmp_edge_embed_bf16(_ptr(X), _ptr(frames), _ptr(g.node_src), _ptr(g.ctr), _ptr(g.nbr), g.Ne,
                                    C.cast(_RBF_MU, C.c_void_p), _ptr(W_perm16), _ptr(bias), _ptr(gain), _ptr(beta),
                                    _ptr(h_E), _stream()), "rmp_edge_embed_bf16")
    return h_E


# --------------------------------------------------------------------------- fused row MLP
@dataclass
class PackedLinear:
    """nn.Linear weights in the kernel's layout: Wt[k_pad][n_out] (transposed, zero padded)."""
    Wt: torch.Tensor
    bias: Optional[torch.Tensor]
    k_in: int
    n_out: int
    W16: Optional[torch.Tensor] = None      # bf16 [n_out, k_in] (nn.Linear layout) for the tensor-core arm
    W16h: Optional[torch.Tensor] = None     # the same as IEEE fp16 (RMP_OPERAND_FP16)
    W16x2: Optional[torch.Tensor] = None    # bf16 [2, n_out, k_in]: high and low terms of the fp32 weight (rmp_row_mlp_x3)


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> PackedLinear:
    _need_cuda(weight, bias)
    n_out, k_in = weight.shape
    if n_out % 128:
        raise ValueError(f"rmpnn_b200 row MLP needs output widths that are multiples of 128, got {n_out}")
    k_pad = (k_in + 31) // 32 * 32
    Wt = torch.zeros(k_pad, n_out, dtype=torch.float32, device=weight.device)
    Wt[:k_in] = weight.detach().to(torch.float32).t()
    return PackedLinear(Wt.contiguous(), None if bias is None else _f32c(bias), k_in, n_out,
                        weight.detach().to(torch.bfloat16).contiguous(),
                        weight.detach().to(torch.float32).clamp(-65504.0, 65504.0).to(torch.float16).contiguous())


def split_weight(pl: PackedLinear) -> torch.Tensor:
    """W = W_hi + W_lo with both terms bf16 (|W - W_hi - W_lo| <= 2^-18 |W|), cached on the PackedLinear."""
    if pl.W16x2 is None:
        W = pl.Wt[:pl.k_in].t().contiguous()
        hi = W.to(torch.bfloat16)
        lo = (W - hi.float()).to(torch.bfloat16)
        pl.W16x2 = torch.stack([hi, lo]).contiguous()
    return pl.W16x2


def x3_eligible(segments: Sequence, stages: Sequence) -> bool:
    """Shapes the split-operand tcgen05 row MLP (rmp_row_mlp_x3) takes."""
    return (all(src.shape[-1] % 128 == 0 and src.dtype == torch.float32 for src, _ in segments)
            and all(pl.k_in % 64 == 0 and pl.k_in <= 384 and pl.n_out in (128, 256) and pl.bias is not None
                    for pl, _ in stages))


def tc_eligible(segments: Sequence, stages: Sequence) -> bool:
    """Shapes the tcgen05 row-MLP kernel (rmp_row_mlp_bf16) takes."""
    return (all(src.shape[-1] % 128 == 0 for src, _ in segments) and len({src.dtype for src, _ in segments}) <= 1
            and all(pl.k_in % 64 == 0 and pl.k_in <= 512 and pl.n_out % 128 == 0 and pl.n_out <= 512
                    and pl.bias is not None and pl.W16 is not None for pl, _ in stages))


def row_mlp(segments: Sequence, stages: Sequence, epilogue: int, out: torch.Tensor, rows: int,
            gain=None, beta=None, resid=None, seg_id=None, alpha: float = 1.0,
            rows_dev: Optional[torch.Tensor] = None, rows_slot: int = 0, row_scale=None, tensor_core: bool = False, fp16: bool = False, out_bf16=None,
            split: bool = False, adds: Sequence = (), add_stride: int = 0):
    """One fused launch: concat(gathered segments) -> chained Linear(+GELU) -> epilogue.

    segments: [(src [*, width] fp32, idx int32 [rows] or None), ...]
    stages:   [(PackedLinear, gelu: bool), ...]
    """
    d = MlpDesc()
    if not 1 <= len(segments) <= _lib.MAX_SEGS or not 1 <= len(stages) <= _lib.MAX_STAGES:
        raise ValueError(f"row_mlp: 1..{_lib.MAX_SEGS} input segments and 1..{_lib.MAX_STAGES} stages per launch")
    keep = []
    for s, (src, idx) in enumerate(segments):
        _need_cuda(src, idx)
        assert src.dtype in (torch.float32, torch.bfloat16) and src.is_contiguous()
        assert idx is None or (idx.dtype == torch.int32 and idx.is_contiguous())
        d.seg_src[s], d.seg_idx[s], d.seg_width[s] = _ptr(src), _ptr(idx), src.shape[-1]
        d.seg_bf16[s] = int(src.dtype == torch.bfloat16)
        keep.append((src, idx))
    d.num_segs = len(segments)
    for s, (pl, gelu) in enumerate(stages):
        d.Wt[s], d.bias[s], d.k_in[s], d.n_out[s], d.gelu[s] = _ptr(pl.Wt), _ptr(pl.bias), pl.k_in, pl.n_out, int(gelu)
    d.num_stages = len(stages)
    d.epilogue = epilogue
    _need_cuda(out, gain, beta, resid, seg_id, rows_dev, row_scale)
    d.gain, d.beta, d.resid, d.seg_id = _ptr(gain), _ptr(beta), _ptr(resid), _ptr(seg_id)
    d.alpha = float(alpha)
    assert row_scale is None or (row_scale.dtype == torch.float32 and row_scale.is_contiguous())
    d.row_scale = _ptr(row_scale)
    _need_cuda(out_bf16)
    assert out_bf16 is None or (out_bf16.dtype == torch.bfloat16 and out_bf16.is_contiguous())
    d.out_bf16 = _ptr(out_bf16)
    d.out = _ptr(out)
    d.rows_dev, d.rows_slot, d.rows_cap = _ptr(rows_dev), rows_slot, int(rows)
    if split:
        # fp32-parity arm on tcgen05 (split bf16 operands); adds = [(fp32 [*, n_out0], idx or None), ...] (<= 2)
        w2 = (C.c_void_p * len(stages))(*[_ptr(split_weight(pl)) for pl, _ in stages])
        a = list(adds) + [(None, None)] * (2 - len(adds))
        _need_cuda(*[t for pair in a for t in pair])
        check(lib().rmp_row_mlp_x3(C.byref(d), C.cast(w2, C.c_void_p), _ptr(a[0][0]), _ptr(a[0][1]), _ptr(a[1][0]),
                                   _ptr(a[1][1]), int(add_stride), _stream()), "rmp_row_mlp_x3")
    elif tensor_core:
        w16 = (C.c_void_p * len(stages))(*[_ptr(pl.W16h if fp16 else pl.W16) for pl, _ in stages])
        check(lib().rmp_row_mlp_tc(C.byref(d), C.cast(w16, C.c_void_p), int(fp16), _stream()), "rmp_row_mlp_tc")
    else:
        check(lib().rmp_row_mlp(C.byref(d), _stream()), "rmp_row_mlp")
    return out


def add_layernorm(a, b, gamma, beta, eps=1e-5, zero_b=False, out=None, rows_dev=None, rows_slot=0,
                  reference_normalize=False):
    """out = LayerNorm(a + b) over 128 channels (norm1, rdesign/model/mpnn.py:35); with
    ``reference_normalize`` the reference's ``Normalize`` (rdesign/model/functional.py:79-96)."""
    _need_cuda(a, b, gamma, beta)
    rows = a.shape[0]
    if out is None:
        out = torch.empty_like(a)
    check(lib().rmp_add_layernorm(_ptr(a), _ptr(b), _ptr(gamma), _ptr(beta), float(eps), _ptr(rows_dev),
                                  rows_slot, rows, int(zero_b), int(reference_normalize), _ptr(out), _stream()),
          "rmp_add_layernorm")
    return out


def readout_head(h, weight, bias, temperature: float = 0.0, seed: int = 0, want_probs=True, want_samples=True):
    """Final Linear(->4) + softmax + argmax / temperature sampling
    (rdesign/model/functional.py:121-127, rdesign/model/rdesign.py:112-115)."""
    _need_cuda(h, weight, bias)
    rows, k_in = h.shape
    dev = h.device
    logits = torch.empty(rows, 4, dtype=torch.float32, device=dev)
    probs = torch.empty(rows, 4, dtype=torch.float32, device=dev) if want_probs else None
    samples = torch.empty(rows, dtype=torch.int64, device=dev) if want_samples else None
    check(lib().rmp_readout_head(_ptr(h), k_in, _ptr(weight), _ptr(bias), float(temperature), int(seed), None, 0,
                                 rows, _ptr(logits), _ptr(probs), _ptr(samples), _stream()), "rmp_readout_head")
    return logits, probs, samples


def edge_mlp_bf16(h_E16, W16, biases, Pc, Pn16, ctr, nbr, row_ptr, agg, alpha: float, debug=None):
    """tcgen05 edge-message kernel (rmp_edge_mlp_bf16).

    h_E16 [Ne,128] bf16; W16 [S*128,128] bf16 (rows of W_0e, W_1, ..); biases: list of S fp32 [128]
    tensors (entry 0 ignored); Pc [Nv,128] fp32 / Pn16 [Nv,128] bf16 node projections; agg [Nv,128] fp32 zeroed.
    """
    _need_cuda(h_E16, W16, Pc, Pn16, ctr, nbr, row_ptr, agg, debug, *[b for b in biases if b is not None])
    S = W16.shape[0] // 128
    assert h_E16.dtype == torch.bfloat16 and W16.dtype == torch.bfloat16 and h_E16.is_contiguous() and W16.is_contiguous()
    assert Pc.dtype == torch.float32 and Pc.is_contiguous() and Pc.shape[1] == 128
    assert Pn16.dtype == torch.bfloat16 and Pn16.is_contiguous() and Pn16.shape[1] == 128
    barr = (C.c_void_p * 4)(*([_ptr(b) if b is not None else None for b in biases] + [None] * (4 - len(biases))))
    check(lib().rmp_edge_mlp_bf16(_ptr(h_E16), _ptr(W16), S, C.cast(barr, C.c_void_p), _ptr(Pc), _ptr(Pn16),
                                  _ptr(ctr), _ptr(nbr), _ptr(row_ptr), h_E16.shape[0], float(alpha), _ptr(agg), _ptr(debug),
                                  _stream()), "rmp_edge_mlp_bf16")
    return agg


def edge_mlp_train_fwd(h_E16, W16, biases, Pc, Pn16, ctr, nbr, row_ptr, agg, alpha: float, p_drop: float, seed: int,
                       stream_id0: int, seed_dev=None):
    """Training forward of the tcgen05 edge-message kernel (rmp_edge_mlp_train_fwd_bf16): agg += alpha * segment sums of the last
    activation, Dropout after every GELU.  Returns (zs, as_): zs[s] bf16 [Ne,128] pre-activations; as_[s] (s < S-1) the
    activations in operand form ``2 GELU(z) mask`` (true activation = as_[s] * keep_scale(p_drop) / 2)."""
    _need_cuda(h_E16, W16, Pc, Pn16, ctr, nbr, row_ptr, agg, seed_dev, *[b for b in biases if b is not None])
    S = W16.shape[0] // 128
    assert h_E16.dtype == torch.bfloat16 and W16.dtype == torch.bfloat16 and h_E16.is_contiguous() and W16.is_contiguous()
    assert Pc.dtype == torch.float32 and Pc.is_contiguous() and Pc.shape[1] == 128
    assert Pn16.dtype == torch.bfloat16 and Pn16.is_contiguous() and Pn16.shape[1] == 128
    zs = [torch.empty_like(h_E16) for _ in range(S)]
    as_ = [torch.empty_like(h_E16) for _ in range(S - 1)]
    barr = (C.c_void_p * 4)(*([_ptr(b) if b is not None else None for b in biases] + [None] * (4 - len(biases))))
    zarr = (C.c_void_p * 4)(*([_ptr(z) for z in zs] + [None] * (4 - S)))
    aarr = (C.c_void_p * 4)(*([_ptr(a) for a in as_] + [None] * (4 - (S - 1))))
    check(lib().rmp_edge_mlp_train_fwd_bf16(_ptr(h_E16), _ptr(W16), S, C.cast(barr, C.c_void_p), _ptr(Pc), _ptr(Pn16), _ptr(ctr),
                                            _ptr(nbr), _ptr(row_ptr), h_E16.shape[0], float(alpha), _ptr(agg),
                                            C.cast(zarr, C.c_void_p), C.cast(aarr, C.c_void_p), float(p_drop), int(seed),
                                            _ptr(seed_dev), int(stream_id0), _stream()), "rmp_edge_mlp_train_fwd_bf16")
    return zs, as_


def keep_scale(p_drop: float) -> float:
    """1 / (1 - p) for the 8-bit quantised drop probability of the training kernels."""
    t = 0 if p_drop <= 0 else int(p_drop * 256.0 + 0.5)
    return 256.0 / (256.0 - t)


def node_mlp_bf16(h_in, stages, n_rows, agg=None, ln1=None, ln2=None, x1=None, h_out=None, p_out=None, pn_out=None):
    """tcgen05 node-update kernel (rmp_node_mlp_bf16).

    stages: [(W bf16 [N,K], bias fp32 [N], kind)], kind 0 = GELU, 1 = residual + LayerNorm2 -> h_out,
    2 = projection store -> p_out [n,128] fp32 (columns 0..127) + pn_out [n,128] bf16 (columns 128..255).
    ln1 = (gamma, beta) enables x1 = LayerNorm1(h_in + agg).
    """
    _need_cuda(h_in, agg, x1, h_out, p_out, pn_out, *(ln1 or ()), *(ln2 or ()))
    n = len(stages)
    Ws = (C.c_void_p * n)(*[_ptr(w) for w, _, _ in stages])
    Bs = (C.c_void_p * n)(*[_ptr(b) for _, b, _ in stages])
    Ks = (C.c_int32 * n)(*[w.shape[1] for w, _, _ in stages])
    Ns = (C.c_int32 * n)(*[w.shape[0] for w, _, _ in stages])
    kinds = (C.c_int32 * n)(*[k for _, _, k in stages])
    for w, b, _ in stages:
        _need_cuda(w, b)
        assert w.dtype == torch.bfloat16 and w.is_contiguous() and b.dtype == torch.float32
    g1, b1 = ln1 if ln1 is not None else (None, None)
    g2, b2 = ln2 if ln2 is not None else (None, None)
    check(lib().rmp_node_mlp_bf16(_ptr(h_in), _ptr(agg), int(ln1 is not None), _ptr(g1), _ptr(b1), n,
                                  C.cast(Ws, C.c_void_p), C.cast(Ks, C.c_void_p), C.cast(Ns, C.c_void_p),
                                  C.cast(kinds, C.c_void_p), C.cast(Bs, C.c_void_p), _ptr(g2), _ptr(b2), _ptr(x1),
                                  _ptr(h_out), _ptr(p_out), _ptr(pn_out), int(n_rows), _stream()), "rmp_node_mlp_bf16")
    return h_out, p_out, pn_out


def cast_bf16(src: torch.Tensor) -> torch.Tensor:
    _need_cuda(src)
    src = _f32c(src)
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    check(lib().rmp_cast_bf16(_ptr(src), _ptr(dst), src.numel(), _stream()), "rmp_cast_bf16")
    return dst


def collate(packed: torch.Tensor, row_start: torch.Tensor, lengths: torch.Tensor, N: int, atoms_out: int = 6):
    """Device-side ``featurize`` (rdesign/utils/data.py:84-119): packed [sum L, A, 3] (fp32 / fp64) + per-structure
    (row_start int64 [B], lengths int32 [B]) -> X [B,N,atoms_out,3] fp32 (NaN -> 0, zero padded), mask [B,N]."""
    _need_cuda(packed, row_start, lengths)
    assert packed.dtype in (torch.float32, torch.float64) and packed.is_contiguous() and packed.dim() == 3
    assert row_start.dtype == torch.int64 and lengths.dtype == torch.int32
    B = lengths.shape[0]
    X = torch.empty(B, N, atoms_out, 3, dtype=torch.float32, device=packed.device)
    mask = torch.empty(B, N, dtype=torch.float32, device=packed.device)
    check(lib().rmp_collate(_ptr(packed), int(packed.dtype == torch.float64), _ptr(row_start), _ptr(lengths), B, int(N),
                            packed.shape[1], int(atoms_out), _ptr(X), _ptr(mask), _stream()), "rmp_collate")
    return X, mask


# =========================================================================== tree-ensemble readout head
@dataclass
class DeviceForest:
    """A flattened gradient-boosted forest on the device (``rmpnn_b200.host.xgb_model.Forest`` packed for
    ``rmp_forest_predict``)."""
    nodes: torch.Tensor        # [n_nodes, 4] int32: threshold bits, feature | default_left << 31, left, right
    tree_root: torch.Tensor    # [n_trees] int32
    class_ptr: torch.Tensor    # [n_class+1] int32
    class_trees: torch.Tensor  # [n_trees] int32, the trees of every output group in tree order
    n_class: int
    n_feat: int
    base_score: float


def pack_forest(forest, device) -> DeviceForest:
    import numpy as np
    thr = np.ascontiguousarray(forest.threshold, dtype=np.float32).view(np.int32)
    feat = forest.feature.astype(np.int64) | (forest.default_left.astype(np.int64) << 31)
    feat = (feat & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    nodes = np.stack([thr, feat, forest.left.astype(np.int32), forest.right.astype(np.int32)], axis=1)
    order = np.argsort(forest.tree_class, kind="stable").astype(np.int32)
    counts = np.bincount(forest.tree_class, minlength=forest.n_class)
    ptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return DeviceForest(t(nodes), t(forest.tree_root.astype(np.int32)), t(ptr), t(order), int(forest.n_class),
                        int(forest.n_feat), float(forest.base_score))


def forest_predict(f: DeviceForest, X: torch.Tensor, want_margins: bool = True):
    """(margins [rows, n_class] fp32 | None, classes [rows] int64) of the tree-ensemble head on ``X [rows, >= n_feat]``
    (rdesign/model/rdesign.py:153; semantics of xgboost's gbtree predictor, see oracle/forest_oracle.py)."""
    _need_cuda(X, f.nodes)
    assert X.dim() == 2 and X.dtype == torch.float32 and X.shape[1] >= f.n_feat
    rows = X.shape[0]
    assert rows == 0 or X.stride(1) == 1
    margins = torch.empty(rows, f.n_class, dtype=torch.float32, device=X.device) if want_margins else None
    pred = torch.empty(rows, dtype=torch.int64, device=X.device)
    if rows == 0:
        return margins, pred
    check(lib().rmp_forest_predict(_ptr(X), rows, int(X.stride(0)), f.n_feat, _ptr(f.nodes), f.nodes.shape[0],
                                   _ptr(f.tree_root), f.tree_root.shape[0], _ptr(f.class_ptr), _ptr(f.class_trees),
                                   f.n_class, f.base_score, _ptr(margins), _ptr(pred), _stream()), "rmp_forest_predict")
    return margins, pred


# =========================================================================== rnampnn (secondary path)
@dataclass
class ResGraph:
    """Dense (padded-layout) residue graph of the rnampnn path and its per-edge bookkeeping."""
    B: int
    N: int
    K: int
    edge_index: torch.Tensor   # [B,N,K] int64, -1 sentinel
    ctr_row: torch.Tensor      # [B*N*K] int32 dense row of the centre
    nbr_row: torch.Tensor      # [B*N*K] int32 dense row of the neighbour (-1 -> row b*N)
    edge_keep: torch.Tensor    # [B*N*K] fp32 (edge_index != -1) * mask_i
    edge_w: torch.Tensor       # [B*N*K] fp32 (edge_index != -1) / max(#valid, 1)


def res_graph(coords: torch.Tensor, mask: torch.Tensor, num_neighbours: int, want_dist: bool = False):
    """``ResFeature._get_res_graph`` (rnampnn/model/feature.py:205-256) -> edge_index [B,N,K] int64."""
    _need_cuda(coords, mask)
    coords, mask = _f32c(coords), _f32c(mask)
    B, N, A, _ = coords.shape
    K = int(num_neighbours)
    ei = torch.empty(B, N, K, dtype=torch.int64, device=coords.device)
    D = torch.empty(B, N, K, dtype=torch.float32, device=coords.device) if want_dist else None
    check(lib().rmp_res_graph(_ptr(coords), _ptr(mask), B, N, A, K, _ptr(ei), _ptr(D), _stream()), "rmp_res_graph")
    return (ei, D) if want_dist else ei


def res_edge_rows(edge_index: torch.Tensor, mask: torch.Tensor) -> ResGraph:
    _need_cuda(edge_index, mask)
    edge_index = edge_index.detach().to(torch.int64).contiguous()
    mask = _f32c(mask)
    B, N, K = edge_index.shape
    dev = edge_index.device
    ctr = torch.empty(B * N * K, dtype=torch.int32, device=dev)
    nbr = torch.empty(B * N * K, dtype=torch.int32, device=dev)
    keep = torch.empty(B * N * K, dtype=torch.float32, device=dev)
    w = torch.empty(B * N * K, dtype=torch.float32, device=dev)
    check(lib().rmp_res_edge_rows(_ptr(edge_index), _ptr(mask), B, N, K, _ptr(ctr), _ptr(nbr), _ptr(keep), _ptr(w),
                                  _stream()), "rmp_res_edge_rows")
    return ResGraph(B, N, K, edge_index, ctr, nbr, keep, w)


def res_raw_features(coords, mask, edge_index, inside=(7, 6, 6), cross=(7, 6, 6), want_raw=True, want_edges=True):
    """raw [B*N, raw_dim], edge_raw [B*N*K, edge_dim]   (rnampnn/model/feature.py:298-517)."""
    _need_cuda(coords, mask, edge_index)
    coords, mask = _f32c(coords), _f32c(mask)
    B, N, A, _ = coords.shape
    dev = coords.device
    nd, na, nh = (int(v) for v in inside)
    cd, ca, ch = (int(v) for v in cross)
    raw = torch.empty(B * N, nd * (nd - 1) // 2 + na - 2 + nh - 3, dtype=torch.float32, device=dev) if want_raw else None
    K = 0
    edge_raw = None
    if want_edges:
        edge_index = edge_index.detach().to(torch.int64).contiguous()
        K = edge_index.shape[-1]
        edge_raw = torch.empty(B * N * K, cd * cd + (ca - 1) ** 2 + (ch - 2) ** 2, dtype=torch.float32, device=dev)
    check(lib().rmp_res_raw_features(_ptr(coords), _ptr(mask), _ptr(edge_index) if want_edges else None, B, N, A, K,
                                     nd, na, nh, cd, ca, ch, _ptr(raw), _ptr(edge_raw), _stream()),
          "rmp_res_raw_features")
    return raw, edge_raw


def graph_norm(a, b, mask, B: int, N: int, scale, shift, n_total: Optional[int] = None, zero_b: bool = False,
               out=None, stats_ws=None, out_bf16=None):
    """``GraphNormalization`` (rnampnn/model/functional.py:18-48) of x = a (+ b) over rows [B*N, 128]."""
    _need_cuda(a, b, mask, scale, shift, out, stats_ws)
    assert a.dtype == torch.float32 and a.is_contiguous() and a.shape[-1] == HIDDEN
    if out is None:
        out = torch.empty_like(a)
    if stats_ws is None:
        stats_ws = torch.empty(B * 257, dtype=torch.float64, device=a.device)
    check(lib().rmp_graph_norm(_ptr(a), _ptr(b), _ptr(mask), B, N, int(N if n_total is None else n_total),
                               _ptr(scale), _ptr(shift), int(zero_b), _ptr(stats_ws), _ptr(out), _ptr(out_bf16), _stream()),
          "rmp_graph_norm")
    return out


def res_attention(qkv, mask, B: int, N: int, heads: int, out=None):
    """softmax(q k^T / sqrt(hd)) v over the valid keys of every structure (rnampnn/model/functional.py:165-168)."""
    _need_cuda(qkv, mask, out)
    E = qkv.shape[-1] // 3
    assert qkv.dtype == torch.float32 and qkv.is_contiguous()
    if out is None:
        out = torch.empty(B * N, E, dtype=torch.float32, device=qkv.device)
    check(lib().rmp_res_attention(_ptr(qkv), _ptr(mask), B, N, E, int(heads), _ptr(out), _stream()), "rmp_res_attention")
    return out


def attention_tc_eligible(E: int, heads: int) -> bool:
    """Shapes ``rmp_res_attention_tc`` takes: 128 channels, head dim 16 or 32, an even number of heads."""
    return E == 128 and heads > 0 and heads % 2 == 0 and E % heads == 0 and E // heads in (16, 32)


def res_attention_tc(qkv16, mask, B: int, N: int, heads: int, out=None):
    """``res_attention`` on the tcgen05 tensor cores: qkv16 [B*N, 384] bf16 (the in-projection's bf16 side copy)."""
    _need_cuda(qkv16, mask, out)
    E = qkv16.shape[-1] // 3
    assert qkv16.dtype == torch.bfloat16 and qkv16.is_contiguous() and qkv16.shape[0] == B * N
    if out is None:
        out = torch.empty(B * N, E, dtype=torch.float32, device=qkv16.device)
    check(lib().rmp_res_attention_tc(_ptr(qkv16), _ptr(mask), B, N, E, int(heads), _ptr(out), _stream()),
          "rmp_res_attention_tc")
    return out


def scale_rows(x, w):
    """x[r,:] *= w[r] in place."""
    _need_cuda(x, w)
    assert x.is_contiguous() and x.dtype == torch.float32 and w.dtype == torch.float32
    check(lib().rmp_scale_rows(_ptr(x), _ptr(w), x.shape[0], x.shape[1], _stream()), "rmp_scale_rows")
    return x


def readout_head_masked(h, weight, bias, row_mask, temperature: float = 0.0, seed: int = 0, want_probs=True,
                        want_samples=True):
    """Final Linear(->4), logits *= mask, softmax, argmax / sampling (rnampnn/model/functional.py:86-88)."""
    _need_cuda(h, weight, bias, row_mask)
    rows, k_in = h.shape
    dev = h.device
    logits = torch.empty(rows, 4, dtype=torch.float32, device=dev)
    probs = torch.empty(rows, 4, dtype=torch.float32, device=dev) if want_probs else None
    samples = torch.empty(rows, dtype=torch.int64, device=dev) if want_samples else None
    check(lib().rmp_readout_head_masked(_ptr(h), k_in, _ptr(weight), _ptr(bias), _ptr(row_mask), float(temperature),
                                        int(seed), rows, _ptr(logits), _ptr(probs), _ptr(samples), _stream()),
          "rmp_readout_head_masked")
    return logits, probs, samples


__all__ = ["collate", "tc_eligible", "ResGraph", "res_graph", "res_edge_rows", "res_raw_features", "graph_norm", "res_attention", "scale_rows",
           "readout_head_masked", "EPI_RESID_ADD", "Graph", "PackedLinear", "edge_embed_bf16", "pack_edge_embedding", "edge_mlp_bf16", "node_mlp_bf16", "knn", "build_graph", "raw_features", "pack_linear", "row_mlp",
           "add_layernorm", "readout_head", "cast_bf16", "EPI_STORE", "EPI_NORMALIZE", "EPI_RESID_LN",
           "EPI_SEGSUM", "HIDDEN", "NODE_RAW", "EDGE_RAW", "MAX_K"]


# --------------------------------------------------------------------------- training path
def act_fwd(z16: torch.Tensor, p_drop: float, seed: int, stream_id: int, seed_dev=None) -> torch.Tensor:
    """a = Dropout(GELU(z)) on bf16 rows (rmp_act_fwd_bf16)."""
    _need_cuda(z16)
    assert z16.dtype == torch.bfloat16 and z16.is_contiguous()
    a = torch.empty_like(z16)
    check(lib().rmp_act_fwd_bf16(_ptr(z16), _ptr(a), z16.shape[0], z16.shape[1], float(p_drop), int(seed), _ptr(seed_dev),
                                 int(stream_id), _stream()), "rmp_act_fwd_bf16")
    return a


def act_bwd(z16, p_drop: float, seed: int, stream_id: int, da16=None, dagg=None, ctr=None, alpha: float = 1.0, db=None,
            seed_dev=None, out=None):
    """dz = da * mask/(1-p) * GELU'(z); da bf16, or gathered from dagg[ctr] * alpha; db [128] += column sums."""
    _need_cuda(z16, da16, dagg, ctr, db, out)
    assert z16.dtype == torch.bfloat16 and z16.is_contiguous() and z16.shape[1] == 128
    dz = torch.empty_like(z16) if out is None else out
    assert dz.dtype == torch.bfloat16 and dz.is_contiguous() and dz.shape == z16.shape
    check(lib().rmp_act_bwd_bf16(_ptr(z16), _ptr(da16), _ptr(dagg), _ptr(ctr), float(alpha), _ptr(dz), _ptr(db),
                                 z16.shape[0], float(p_drop), int(seed), _ptr(seed_dev), int(stream_id), _stream()), "rmp_act_bwd_bf16")
    return dz


def scatter_add_rows(x16, idx, out):
    _need_cuda(x16, idx, out)
    assert x16.dtype == torch.bfloat16 and x16.is_contiguous() and x16.shape[1] == 128 and out.dtype == torch.float32
    check(lib().rmp_scatter_add_rows(_ptr(x16), _ptr(idx), _ptr(out), x16.shape[0], _stream()), "rmp_scatter_add_rows")
    return out


def segment_sum_rows(x16, row_ptr, out=None):
    """out[v] = sum of rows row_ptr[v]..row_ptr[v+1]-1 of x16 (bf16 [rows,128]) -> fp32 [n_seg,128] (overwritten, no atomics)."""
    _need_cuda(x16, row_ptr, out)
    assert x16.dtype == torch.bfloat16 and x16.is_contiguous() and x16.shape[1] == 128 and row_ptr.dtype == torch.int32
    n_seg = row_ptr.shape[0] - 1
    if out is None:
        out = torch.empty(n_seg, 128, dtype=torch.float32, device=x16.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] >= n_seg
    check(lib().rmp_segment_sum_rows(_ptr(x16), _ptr(row_ptr), _ptr(out), n_seg, _stream()), "rmp_segment_sum_rows")
    return out


def segment_sum_rows_perm(x16, perm, row_ptr, out=None):
    """out[v] = sum over r in [row_ptr[v], row_ptr[v+1]) of x16[perm[r]] -> fp32 [n_seg,128] (rmp_segment_sum_rows_perm)."""
    _need_cuda(x16, perm, row_ptr, out)
    assert x16.dtype == torch.bfloat16 and x16.is_contiguous() and x16.shape[1] == 128
    assert perm.dtype == row_ptr.dtype == torch.int32 and perm.is_contiguous() and row_ptr.is_contiguous()
    n_seg = row_ptr.shape[0] - 1
    if out is None:
        out = torch.empty(n_seg, 128, dtype=torch.float32, device=x16.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] >= n_seg
    check(lib().rmp_segment_sum_rows_perm(_ptr(x16), _ptr(perm), _ptr(row_ptr), _ptr(out), n_seg, _stream()),
          "rmp_segment_sum_rows_perm")
    return out


def transposed_csr(idx, n_seg: int):
    """(perm, row_ptr) of the rows grouped by ``idx`` (int32 [rows], values in [0, n_seg)): perm = stable argsort,
    row_ptr [n_seg+1] = offsets of every value's group.  torch ops only (any device), capturable in a CUDA graph (no host
    sync)."""
    il = idx.long()
    perm = torch.argsort(il, stable=True).to(torch.int32)
    cnt = torch.zeros(n_seg + 1, dtype=torch.int32, device=idx.device)
    cnt[1:].index_add_(0, il, torch.ones_like(idx))
    return perm, torch.cumsum(cnt, 0, dtype=torch.int32)


def gemm_dw(dz16, x16, dW):
    """dW [128,128] fp32 += dz16^T @ x16 (both bf16 [rows,128])."""
    _need_cuda(dz16, x16, dW)
    assert dz16.dtype == x16.dtype == torch.bfloat16 and dz16.shape == x16.shape and dz16.shape[1] == 128
    assert dz16.is_contiguous() and x16.is_contiguous() and dW.dtype == torch.float32 and dW.is_contiguous()
    check(lib().rmp_gemm_dw_bf16(_ptr(dz16), _ptr(x16), _ptr(dW), dz16.shape[0], _stream()), "rmp_gemm_dw_bf16")
    return dW


def gemm_rows(a16, w16, bias=None, Pc=None, Pn=None, ctr=None, nbr=None, want_z=True, act=None, acc32=None,
              w_transposed=False, seed_dev=None):
    """Z = a16 @ w16^T (or a16 @ w16 with ``w_transposed``) (+bias) (+Pc[ctr] + Pn[nbr]) on the lean tcgen05 row GEMM.  ``act=(p_drop, seed, stream_id)``
    also returns Dropout(GELU(Z)); ``acc32`` (fp32 [rows,128]) is incremented by Z.  Returns (z16 | None, a16 | None)."""
    _need_cuda(a16, w16, bias, Pc, Pn, ctr, nbr, acc32)
    assert a16.dtype == w16.dtype == torch.bfloat16 and a16.is_contiguous() and w16.stride(1) == 1
    assert a16.shape[1] == 128 and tuple(w16.shape) == (128, 128)      # w16 may be a column block of a wider matrix
    assert Pc is None or (Pc.dtype == Pn.dtype == torch.bfloat16 and Pc.is_contiguous() and Pn.is_contiguous())
    rows = a16.shape[0]
    z = torch.empty_like(a16) if want_z else None
    a = torch.empty_like(a16) if act is not None else None
    p_drop, seed, sid = act if act is not None else (0.0, 0, 0)
    check(lib().rmp_gemm_rows_bf16(_ptr(a16), _ptr(w16), _ptr(bias), _ptr(Pc), _ptr(Pn), _ptr(ctr), _ptr(nbr), _ptr(z), _ptr(a),
                                   _ptr(acc32), rows, int(w16.stride(0)), int(w_transposed), float(p_drop), int(seed), _ptr(seed_dev), int(sid), _stream()),
          "rmp_gemm_rows_bf16")
    return z, a


def add_layernorm_fwd(a, b, gamma, beta, eps: float):
    """x = a + b (b fp32 / bf16 / None), y = LayerNorm(x) over 128 channels -> (x, y, mean, rstd)   (rmp_add_layernorm_fwd)."""
    _need_cuda(a, b, gamma, beta)
    assert a.dtype == torch.float32 and a.is_contiguous() and a.dim() == 2 and a.shape[1] == 128
    assert b is None or (b.shape == a.shape and b.is_contiguous() and b.dtype in (torch.float32, torch.bfloat16))
    rows = a.shape[0]
    x = torch.empty_like(a) if b is not None else a
    y = torch.empty_like(a)
    stats = torch.empty(2, rows, dtype=torch.float32, device=a.device)
    check(lib().rmp_add_layernorm_fwd(_ptr(a), _ptr(b), int(b is not None and b.dtype == torch.bfloat16), _ptr(_f32c(gamma)),
                                      _ptr(_f32c(beta)), float(eps), rows, _ptr(x) if b is not None else None, _ptr(y),
                                      _ptr(stats[0]), _ptr(stats[1]), _stream()), "rmp_add_layernorm_fwd")
    return x, y, stats[0], stats[1]


def layernorm_bwd(dy, x, mean, rstd, gamma, want_bf16: bool = False):
    """-> (dx fp32, dx bf16 | None, dgamma [128], dbeta [128])   (rmp_layernorm_bwd)."""
    _need_cuda(dy, x, mean, rstd, gamma)
    assert dy.dtype == x.dtype == torch.float32 and dy.is_contiguous() and x.is_contiguous() and dy.shape == x.shape and x.shape[1] == 128
    dx = torch.empty_like(x)
    dx16 = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if want_bf16 else None
    sums = torch.zeros(2, 128, dtype=torch.float32, device=x.device)
    check(lib().rmp_layernorm_bwd(_ptr(dy), _ptr(x), _ptr(mean), _ptr(rstd), _ptr(_f32c(gamma)), x.shape[0], _ptr(dx), _ptr(dx16),
                                  _ptr(sums[0]), _ptr(sums[1]), _stream()), "rmp_layernorm_bwd")
    return dx, dx16, sums[0], sums[1]


def colsum(a, x=None, mean=None, rstd=None):
    """Column sums of ``a`` [rows,width] (fp32 or bf16) -> fp32 [width]; with ``x`` (fp32), ``mean``, ``rstd`` [rows] also the
    sums of ``a * (x - mean) * rstd`` (LayerNorm weight gradient).  Returns (s0, s1 | None).  rmp_colsum."""
    _need_cuda(a, x, mean, rstd)
    assert a.dim() == 2 and a.is_contiguous() and a.dtype in (torch.float32, torch.bfloat16)
    rows, width = a.shape
    out = torch.zeros(2 if x is not None else 1, width, dtype=torch.float32, device=a.device)
    if x is not None:
        assert x.dtype == torch.float32 and x.is_contiguous() and x.shape == a.shape
        mean, rstd = mean.reshape(-1), rstd.reshape(-1)
        assert mean.dtype == rstd.dtype == torch.float32 and mean.is_contiguous() and rstd.is_contiguous() and mean.numel() == rows
    check(lib().rmp_colsum(_ptr(a), int(a.dtype == torch.bfloat16), _ptr(x), _ptr(mean), _ptr(rstd), rows, width, _ptr(out[0]),
                           _ptr(out[1]) if x is not None else None, _stream()), "rmp_colsum")
    return out[0], (out[1] if x is not None else None)


def gemm_rows_kcat(a16, w16, w_transposed=False, out=None):
    """out = sum_l a16[l] @ w16[l]^T (or a16[l] @ w16[l] with ``w_transposed``): a16 bf16 [L,rows,128], w16 bf16 [L,128,128]
    -> fp32 [rows,128] in one pass (rmp_gemm_rows_kcat_bf16)."""
    _need_cuda(a16, w16, out)
    assert a16.dtype == w16.dtype == torch.bfloat16 and a16.is_contiguous() and w16.is_contiguous()
    L, rows = a16.shape[0], a16.shape[1]
    assert a16.shape[2] == 128 and tuple(w16.shape) == (L, 128, 128) and L >= 2
    if out is None:
        out = torch.empty(rows, 128, dtype=torch.float32, device=a16.device)
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == (rows, 128)
    check(lib().rmp_gemm_rows_kcat_bf16(_ptr(a16), _ptr(w16), L, _ptr(out), rows, int(w_transposed), _stream()),
          "rmp_gemm_rows_kcat_bf16")
    return out


def normalize_fwd(x, gain, bias, eps: float = 1e-6):
    """Reference Normalize over 128 channels -> bf16 rows (rmp_normalize_fwd_bf16)."""
    _need_cuda(x, gain, bias)
    assert x.dtype == torch.float32 and x.is_contiguous() and x.shape[1] == 128
    y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib().rmp_normalize_fwd_bf16(_ptr(x), _ptr(gain), _ptr(bias), _ptr(y), x.shape[0], float(eps), _stream()),
          "rmp_normalize_fwd_bf16")
    return y


def normalize_bwd(x, dy, gain, eps: float = 1e-6):
    """-> (dx bf16 [rows,128], dgain [128], dbias [128], column sums of dx [128])."""
    _need_cuda(x, dy, gain)
    assert x.dtype == dy.dtype == torch.float32 and x.is_contiguous() and dy.is_contiguous() and x.shape == dy.shape
    dx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    sums = torch.zeros(3, 128, dtype=torch.float32, device=x.device)
    check(lib().rmp_normalize_bwd_bf16(_ptr(x), _ptr(dy), _ptr(gain), _ptr(dx), _ptr(sums[0]), _ptr(sums[1]), _ptr(sums[2]),
                                       x.shape[0], float(eps), _stream()), "rmp_normalize_bwd_bf16")
    return dx, sums[0], sums[1], sums[2]


def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, grad_scale: float = 1.0, betas=(0.9, 0.999),
              eps: float = 1e-8, step_dev=None):
    """``step_dev`` (int32 device scalar) overrides ``step`` so that a captured graph can advance it on the device."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq, step_dev)
    check(lib().rmp_adam_step(_ptr(param), _ptr(grad), _ptr(exp_avg), _ptr(exp_avg_sq), param.numel(), float(grad_scale),
                              float(lr), float(betas[0]), float(betas[1]), float(eps), int(step), _ptr(step_dev), _stream()),
          "rmp_adam_step")


# --------------------------------------------------------------------------- device guard
def _cuda_device_of(objs, depth=0):
    """Device of the first CUDA tensor found in ``objs`` (tensors, Graph / ResGraph / PackedLinear, short sequences)."""
    for o in objs:
        if torch.is_tensor(o):
            if o.is_cuda:
                return o.device
        elif isinstance(o, (Graph, ResGraph, PackedLinear, DeviceForest)):
            d = _cuda_device_of(vars(o).values(), depth + 1)
            if d is not None:
                return d
        elif isinstance(o, (list, tuple)) and depth < 3:
            d = _cuda_device_of(o, depth + 1)
            if d is not None:
                return d
    return None


def _on_tensor_device(fn):
    """Run ``fn`` with the CUDA device of its tensor arguments current: kernels, function attributes and
    ``torch.cuda.current_stream()`` then all refer to the GPU that owns the memory (a model on cuda:1 while
    cuda:0 is current would otherwise enqueue device-1 pointers on device 0)."""
    import functools

    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        dev = _cuda_device_of(list(args) + list(kwargs.values()))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


for _name, _fn in list(globals().items()):
    if callable(_fn) and not _name.startswith("_") and getattr(_fn, "__module__", None) == __name__ \
            and not isinstance(_fn, type):
        globals()[_name] = _on_tensor_device(_fn)
del _name, _fn
