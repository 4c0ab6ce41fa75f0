"""CPU restatement of the reference's ``rdesign`` structure-to-sequence forward path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- the checker for the CUDA
path and the timed CPU arm of ``bench.py``; never imported by the product.

Parity pin: ``tests/golden/*.npz`` were produced by importing the *unmodified*
reference (``/root/reference/rdesign``) in the build container through
``tests/golden/make_golden.py``; ``tests/test_oracle_golden.py`` checks every
function here against those vectors (neighbour indices exact, floats 1e-5/1e-6).

Written as stateless functions over a flat parameter dict (the reference's
``state_dict`` keys) instead of the reference's ``nn.Module`` tree.  All
arithmetic is fp32 torch-CPU, like the reference's CPU path.  Citations are
``file:line`` under ``/root/reference``.

Conventions: B structures, N padded length, K' = min(top_k, N), H hidden.
Valid-only flattening is (b, i, k) row major; ``E_idx[0]`` = centre node,
``E_idx[1]`` = neighbour node, both global flat ids.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

NUM_RBF = 16
ATOM_P, ATOM_O5, ATOM_C5, ATOM_C4, ATOM_C3, ATOM_O3 = range(6)


# --------------------------------------------------------------------------- a-1
def sqrt_rn(x: torch.Tensor) -> torch.Tensor:
    """Correctly rounded fp32 square root (via fp64).  torch's CPU ``sqrt`` on this build
    (MKL VML, high-accuracy mode) is only faithful to < 1 ulp, so the reference's own
    distances can sit 1 ulp off the IEEE value; the canonical kNN is defined on the IEEE
    value, which is what ``sqrt.rn.f32`` gives on the GPU."""
    return torch.sqrt(x.double()).float()


def pairwise_p_distance(P: torch.Tensor, mask: torch.Tensor, exact_sqrt: bool = True) -> torch.Tensor:
    """Adjusted P-P distance matrix [B,N,N]   (rdesign/model/feature.py:42-48).

    valid pair   -> sqrt(((dx*dx + dy*dy) + dz*dz) + 1e-6)
    invalid pair -> 10000 + (rowmax + 1)
    ``exact_sqrt=False`` uses torch's CPU sqrt exactly as the reference does.
    """
    pair_ok = mask[:, None, :] * mask[:, :, None]
    diff = P[:, None, :, :] - P[:, :, None, :]
    root = sqrt_rn if exact_sqrt else torch.sqrt
    D = (1.0 - pair_ok) * 10000 + pair_ok * root((diff ** 2).sum(3) + 1e-6)
    row_max = D.max(-1, keepdim=True).values
    return D + (1.0 - pair_ok) * (row_max + 1)


def knn_reference_order(P, mask, top_k):
    """``torch.topk`` selection exactly as the reference calls it
    (rdesign/model/feature.py:49).  Tie order is whatever ATen gives."""
    D = pairwise_p_distance(P, mask, exact_sqrt=False)
    return torch.topk(D, min(top_k, D.shape[-1]), dim=-1, largest=False)


def knn(P, mask, top_k):
    """Canonical kNN: ascending adjusted distance (IEEE-rounded sqrt), ties broken by
    LOWER index.

    This is the documented rule the CUDA kernel implements (the reference's ``topk``
    order among exact ties is unspecified, and its CPU sqrt is 1-ulp faithful only); a
    stable ascending sort realises it.  Returns (D_neighbors [B,N,K'], E_idx [B,N,K'] int64).
    """
    D = pairwise_p_distance(P, mask)
    Ds, idx = torch.sort(D, dim=-1, stable=True)
    k = min(top_k, D.shape[-1])
    return Ds[..., :k].contiguous(), idx[..., :k].contiguous()


def canonicalize_knn(D_neighbors, E_idx):
    """Re-order a (D, idx) top-k result into the canonical (distance, index) order."""
    key_idx = torch.argsort(E_idx, dim=-1, stable=True)
    D1 = torch.gather(D_neighbors, -1, key_idx)
    I1 = torch.gather(E_idx, -1, key_idx)
    key_d = torch.argsort(D1, dim=-1, stable=True)
    return torch.gather(D1, -1, key_d), torch.gather(I1, -1, key_d)


# --------------------------------------------------------------------------- helpers
def _unit_nan0(v):
    """v / |v| with 0/0 -> 0   (rdesign/utils/data.py:164-171)."""
    out = v / torch.norm(v, dim=-1, keepdim=True)
    return torch.where(torch.isnan(out), torch.zeros_like(out), out)


def _take_nodes(x, E_idx):
    """x [B,N,C], E_idx [B,N,K] -> x[b, E_idx[b,i,k]]  [B,N,K,C]."""
    B, N, K = E_idx.shape
    flat = E_idx.reshape(B, N * K, 1).expand(-1, -1, x.shape[-1])
    return torch.gather(x, 1, flat).reshape(B, N, K, x.shape[-1])


def rbf_expand(D):
    """16 Gaussian bumps, centres linspace(0,20,16), width 1.25
    (rdesign/model/feature.py:52-58)."""
    mu = torch.linspace(0.0, 20.0, NUM_RBF)
    sigma = (20.0 - 0.0) / NUM_RBF
    return torch.exp(-(((D[..., None] - mu) / sigma) ** 2))


def rbf_centres():
    return torch.linspace(0.0, 20.0, NUM_RBF)


# --------------------------------------------------------------------------- a-3
def backbone_dihedral_features(X):
    """[B,N,12] = cos|sin of six pseudo-dihedrals per residue
    (rdesign/model/feature.py:134-154; stride-5 atom difference as written,
    zero-padded residues take part)."""
    B, N = X.shape[:2]
    flat = X[:, :, :6, :].reshape(B, 6 * N, 3)
    U = F.normalize(flat[:, 5:, :] - flat[:, :-5, :], dim=-1)
    ua, ub, uc = U[:, :-2, :], U[:, 1:-1, :], U[:, 2:, :]
    na = F.normalize(torch.linalg.cross(ua, ub), dim=-1)
    nb = F.normalize(torch.linalg.cross(ub, uc), dim=-1)
    cosD = torch.clamp((na * nb).sum(-1), -1 + 1e-7, 1 - 1e-7)
    ang = torch.sign((ua * nb).sum(-1)) * torch.acos(cosD)
    ang = F.pad(ang, (3, 4), 'constant', 0).reshape(B, N, 6)
    return torch.cat((torch.cos(ang), torch.sin(ang)), 2)


# --------------------------------------------------------------------------- a-5
def local_frames(X):
    """Per-residue frame at C3' [B,N,9] (rows b1, n0, b1 x n0); last residue = 0
    (rdesign/model/feature.py:90-107)."""
    B, N = X.shape[:2]
    flat = X[:, :, :6, :].reshape(B, 6 * N, 3)
    U = _unit_nan0(flat[:, 1:, :] - flat[:, :-1, :])
    u0, u1 = U[:, :-2, :], U[:, 1:-1, :]
    n0 = _unit_nan0(torch.linalg.cross(u0, u1))[:, 4::6, :]
    b1 = _unit_nan0(u0 - u1)[:, 4::6, :]
    Q = torch.stack((b1, n0, torch.linalg.cross(b1, n0)), 2).reshape(B, -1, 9)
    return F.pad(Q, (0, 0, 0, 1), 'constant', 0)


def rotation_quaternion(R):
    """[...,3,3] -> unit quaternion (x,y,z,w)  (rdesign/model/feature.py:70-88)."""
    Rxx, Ryy, Rzz = R[..., 0, 0], R[..., 1, 1], R[..., 2, 2]
    mag = 0.5 * torch.sqrt(torch.abs(1 + torch.stack(
        [Rxx - Ryy - Rzz, -Rxx + Ryy - Rzz, -Rxx - Ryy + Rzz], -1)))
    sgn = torch.sign(torch.stack([R[..., 2, 1] - R[..., 1, 2],
                                  R[..., 0, 2] - R[..., 2, 0],
                                  R[..., 1, 0] - R[..., 0, 1]], -1))
    w = torch.sqrt(F.relu(1 + (Rxx + Ryy + Rzz)[..., None])) / 2.0
    return F.normalize(torch.cat((sgn * mag, w), -1), dim=-1)


def orientation_features(X, E_idx):
    """(node_dir [B,N,9], edge_dir [B,N,K,15], edge_quat [B,N,K,4])
    (rdesign/model/feature.py:90-132)."""
    B, N, K = E_idx.shape
    Q = local_frames(X)
    C3 = X[:, :, ATOM_C3, :]
    Qi = Q.reshape(B, N, 1, 3, 3)
    Qj = _take_nodes(Q, E_idx).reshape(B, N, K, 3, 3)
    nb_atoms = torch.stack([_take_nodes(X[:, :, a, :], E_idx)
                            for a in (ATOM_P, ATOM_O5, ATOM_C5, ATOM_C4, ATOM_O3)], dim=3)
    rel = nb_atoms - C3[:, :, None, None, :]
    edge_dir = _unit_nan0(torch.matmul(Qi[:, :, :, None, :, :], rel[..., None]).squeeze(-1))
    edge_dir = edge_dir.reshape(B, N, K, 15)
    edge_quat = rotation_quaternion(torch.matmul(Qi.transpose(-1, -2), Qj))
    own = X[:, :, [ATOM_P, ATOM_C5, ATOM_C4], :] - C3[:, :, None, :]
    node_dir = _unit_nan0(torch.matmul(Qi, own[..., None]).squeeze(-1)).reshape(B, N, 9)
    return node_dir, edge_dir, edge_quat


# --------------------------------------------------------------------------- a-4
def _eps_dist(a, b):
    return torch.sqrt(((a - b) ** 2).sum(-1) + 1e-6)


def node_rbf(X):
    """[B,N,80]: RBF of |atom_a(i) - P(i)| for a in O5',C5',C4',C3',O3'
    (rdesign/model/feature.py:184-191, :65-67)."""
    P = X[:, :, ATOM_P, :]
    return torch.cat([rbf_expand(_eps_dist(X[:, :, a, :], P))
                      for a in (ATOM_O5, ATOM_C5, ATOM_C4, ATOM_C3, ATOM_O3)], -1)


def edge_rbf(X, E_idx):
    """[B,N,K,96]: RBF of |atom_a(centre i) - P(neighbour j)|, a = P..O3'
    (rdesign/model/feature.py:200-206, :61-64)."""
    Pj = _take_nodes(X[:, :, ATOM_P, :], E_idx)
    return torch.cat([rbf_expand(_eps_dist(X[:, :, a, None, :], Pj)) for a in range(6)], -1)


# --------------------------------------------------------------------------- a-2
def attend_mask(mask, E_idx):
    """mask_i * mask_j(E_idx) == 1   (rdesign/model/feature.py:173-175)."""
    mj = _take_nodes(mask[..., None], E_idx)[..., 0]
    return (mask[..., None] * mj) == 1


def flatten_graph(mask, E_idx):
    """Global flat edge list [2,Ne] (centre, neighbour) + batch_id [Nv]
    (rdesign/model/feature.py:227-237)."""
    B, N, K = E_idx.shape
    ok = attend_mask(mask, E_idx)
    counts = mask.sum(1)
    shift = (counts.cumsum(0) - counts).reshape(B, 1, 1)
    nbr = (shift + E_idx)[ok]
    ctr = (shift + torch.arange(N).reshape(1, N, 1).expand(B, N, K))[ok]
    batch_id = mask.nonzero()[:, 0]
    return torch.stack((ctr, nbr), 0).long(), batch_id


# --------------------------------------------------------------------------- a-6
def normalize_rows(x, gain, bias, eps=1e-6):
    """Reference ``Normalize``: unbiased variance, eps added inside and outside
    the sqrt (rdesign/model/functional.py:79-96)."""
    mu = x.mean(-1, keepdim=True)
    sigma = torch.sqrt(x.var(-1, keepdim=True) + eps)
    return gain * (x - mu) / (sigma + eps) + bias


def raw_features(X, mask, top_k, E_idx=None):
    """Valid-only raw features before the embeddings:
    node [Nv,101] = angle12 | rbf80 | dir9, edge [Ne,115] = quat4 | rbf96 | dir15
    (rdesign/model/feature.py:179-220).  Returns (node, edge, E_idx_dense)."""
    if E_idx is None:
        _, E_idx = knn(X[:, :, ATOM_P, :], mask, top_k)
    ok_node = mask == 1
    ok_edge = attend_mask(mask, E_idx)
    node_dir, edge_dir, edge_quat = orientation_features(X, E_idx)
    node = torch.cat((backbone_dihedral_features(X), node_rbf(X), node_dir), -1)[ok_node]
    edge = torch.cat((edge_quat, edge_rbf(X, E_idx), edge_dir), -1)[ok_edge]
    return node, edge, E_idx


def features(params, X, S, mask, top_k, E_idx=None):
    """``RNAFeatures.forward`` (rdesign/model/feature.py:156-238).
    Returns (X_valid, S_valid, h_V [Nv,H], h_E [Ne,H], E_idx [2,Ne], batch_id)."""
    node, edge, E_dense = raw_features(X, mask, top_k, E_idx)
    p = lambda k: params["features." + k]
    h_V = normalize_rows(F.linear(node, p("node_embedding.weight"), p("node_embedding.bias")),
                         p("norm_nodes.gain"), p("norm_nodes.bias"))
    h_E = normalize_rows(F.linear(edge, p("edge_embedding.weight"), p("edge_embedding.bias")),
                         p("norm_edges.gain"), p("norm_edges.bias"))
    E_flat, batch_id = flatten_graph(mask, E_dense)
    ok_node = mask == 1
    return X[ok_node], S[ok_node], h_V, h_E, E_flat, batch_id


# --------------------------------------------------------------------------- a-7 / a-8
def _layer_linears(params, prefix):
    idx = sorted({int(k[len(prefix):].split(".")[0]) for k in params if k.startswith(prefix)})
    return [(params[f"{prefix}{i}.weight"], params[f"{prefix}{i}.bias"]) for i in idx]


def mpnn_layer(params, l, h_V, h_E, E_idx, scale=30.0):
    """One message-passing layer incl. the caller's gather+concat
    (rdesign/model/rdesign.py:85, rdesign/model/mpnn.py:31-38)."""
    pre = f"mpnn_layers.{l}."
    ctr, nbr = E_idx[0], E_idx[1]
    m = torch.cat([h_E, h_V[ctr], h_V[nbr]], -1)
    for W, b in _layer_linears(params, pre + "message_layers."):
        m = F.gelu(F.linear(m, W, b))            # GELU after every message Linear
    agg = torch.zeros_like(h_V).index_add_(0, ctr, m) / scale
    H = h_V.shape[-1]
    h = F.layer_norm(h_V + agg, (H,), params[pre + "norm1.weight"], params[pre + "norm1.bias"], 1e-5)
    dense = _layer_linears(params, pre + "dense.")
    d = h
    for W, b in dense[:-1]:
        d = F.gelu(F.linear(d, W, b))
    d = F.linear(d, *dense[-1])                  # dense ends with a bare Linear
    return F.layer_norm(h + d, (H,), params[pre + "norm2.weight"], params[pre + "norm2.bias"], 1e-5)


# --------------------------------------------------------------------------- a-9
def readout(params, h_V):
    """``Readout.forward`` (rdesign/model/functional.py:123-127)."""
    lin = _layer_linears(params, "readout.readout_layers.")
    x = h_V
    for W, b in lin[:-1]:
        x = F.gelu(F.linear(x, W, b))
    return F.linear(x, *lin[-1])


def num_layers(params):
    return 1 + max(int(k.split(".")[1]) for k in params if k.startswith("mpnn_layers."))


@torch.no_grad()
def forward(params, X, S, mask, top_k, scale=30.0, E_idx=None, return_layers=False):
    """``RNAModel.forward`` + ``Readout`` + softmax/argmax
    (rdesign/model/rdesign.py:82-88, :101, :112-115).
    Returns dict(h_V, S, logits, probs, samples, E_idx, h_E, h_V0[, layers])."""
    _, S_v, h_V, h_E, E_flat, batch_id = features(params, X, S, mask, top_k, E_idx)
    out = {"h_V0": h_V, "h_E": h_E, "E_idx": E_flat, "S": S_v, "batch_id": batch_id}
    layers = []
    for l in range(num_layers(params)):
        h_V = mpnn_layer(params, l, h_V, h_E, E_flat, scale)
        if return_layers:
            layers.append(h_V)
    logits = readout(params, h_V)
    probs = F.softmax(logits, -1)
    out.update(h_V=h_V, logits=logits, probs=probs, samples=probs.argmax(-1))
    if return_layers:
        out["layers"] = layers
    return out
