"""CPU restatement of the reference's ``rnampnn`` forward path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- the checker for the CUDA
path; never imported by the product.

Parity pin: ``tests/golden/rnampnn_*.npz`` were produced by importing the *unmodified*
reference (``/root/reference/rnampnn``) in the build container through
``tests/golden/make_golden_rnampnn.py``; ``tests/test_rnampnn_oracle_golden.py`` checks every
function here against those vectors.

Stateless functions over a flat parameter dict (the reference ``RNAMPNN.state_dict()`` keys).
fp32 torch-CPU arithmetic like the reference's CPU path.  Citations are ``file:line`` under
``/root/reference``.  The reference keeps the padded [B,N,...] layout end to end (no
valid-only flattening), so does this file.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LEPS = 1e6          # rnampnn/config/glob.py:16
SEPS = 1e-6         # rnampnn/config/glob.py:17


def sqrt_rn(x: torch.Tensor) -> torch.Tensor:
    """Correctly rounded fp32 sqrt (torch's CPU sqrt is 1-ulp faithful only, see rdesign_oracle)."""
    return torch.sqrt(x.double()).float()


# --------------------------------------------------------------------------- a-10
def centroids(coords: torch.Tensor) -> torch.Tensor:
    """Mean of the atoms of every residue [B,N,3] (rnampnn/model/feature.py:218).

    ``coords.mean(dim=2)`` on torch-CPU adds the seven atoms in the order 0,4,5,6,1,2,3 and then
    divides by 7 (checked bit for bit in the build container); the canonical definition spells
    that order out so that it does not depend on the host's vector ISA.  Other atom counts are
    summed in index order (unpinned)."""
    A = coords.shape[2]
    order = (0, 4, 5, 6, 1, 2, 3) if A == 7 else tuple(range(A))
    s = coords[:, :, order[0], :].clone()
    for a in order[1:]:
        s = s + coords[:, :, a, :]
    return s / float(A)


def centroid_distance(coords, mask, exact_sqrt: bool = True):
    """Adjusted centroid distance matrix [B,N,N] (rnampnn/model/feature.py:218-227): valid pairs
    sqrt(((dx^2+dy^2)+dz^2)+1e-6), pairs with a padded end 1e6, diagonal + 1e6."""
    c = centroids(coords)
    d = c[:, None, :, :] - c[:, :, None, :]
    root = sqrt_rn if exact_sqrt else torch.sqrt
    D = root((d ** 2).sum(-1) + SEPS)
    m2 = mask[:, None, :] * mask[:, :, None]
    D = D * m2 + (1.0 - m2) * LEPS
    return D + torch.eye(D.shape[-1])[None] * LEPS


def res_graph(coords, mask, num_neighbours: int):
    """``ResFeature._get_res_graph`` (rnampnn/model/feature.py:205-256) with the canonical order:
    ascending adjusted distance (IEEE sqrt), ties -> lower index.  Self excluded, -1 sentinels.
    Returns edge_index [B,N,K] int64."""
    B, N = mask.shape
    K = num_neighbours
    D = centroid_distance(coords, mask)
    _, idx = torch.sort(D, dim=-1, stable=True)
    idx = idx[..., :min(K, N)]
    if K > N:
        idx = torch.cat([idx, torch.full((B, N, K - N), -1, dtype=idx.dtype)], -1)
    own = torch.arange(N).view(1, N, 1)
    idx = torch.where(idx == own, torch.full_like(idx, -1), idx)
    m2 = mask[:, None, :] * mask[:, :, None]
    valid_nb = (m2.sum(-1) - 1).clamp(min=0)
    # NB (reference quirk, feature.py:248-251): slot k is cleared only when valid_nb < k, so slot
    # k == valid_nb survives; under the canonical tie-break with prefix masks it holds "self" -> -1.
    idx = idx.masked_fill(valid_nb[..., None] < torch.arange(K), -1)
    return idx.masked_fill((mask == 0)[..., None], -1)


# --------------------------------------------------------------------------- a-11
def inside_features(coords, mask, n_dist=7, n_angle=6, n_dihedral=6):
    """raw [B,N,raw_dim] = intra-residue distances | bond-angle cosines | bond-plane cosines
    (rnampnn/model/feature.py:298-384, :531-535)."""
    c = coords[:, :, :n_dist, :]
    diff = c[:, :, :, None, :] - c[:, :, None, :, :]
    pd = torch.sqrt((diff ** 2).sum(-1) + SEPS)
    iu = torch.triu_indices(n_dist, n_dist, offset=1)
    dists = pd[:, :, iu[0], iu[1]].masked_fill((mask == 0)[..., None], 1e6)

    c = coords[:, :, :n_angle, :]
    v = c[:, :, 1:] - c[:, :, :-1]
    nrm = torch.norm(v, dim=-1)
    ang = (v[:, :, :-1] * v[:, :, 1:]).sum(-1) / (nrm[:, :, :-1] * nrm[:, :, 1:] + SEPS)
    ang = ang * mask[..., None]

    c = coords[:, :, :n_dihedral, :]
    u = F.normalize(c[:, :, 1:] - c[:, :, :-1], dim=-1, eps=SEPS)
    nor = F.normalize(torch.linalg.cross(u[:, :, :-1], u[:, :, 1:]), dim=-1, eps=SEPS)
    dih = (nor[:, :, 1:] * nor[:, :, :-1]).sum(-1) * mask[..., None]
    return torch.cat([dists, ang, dih], -1)


# --------------------------------------------------------------------------- a-12
def _gather_res(x, safe_idx):
    """x [B,N,...], safe_idx [B,N,K] -> x[b, safe_idx[b,i,k]]  [B,N,K,...]."""
    B, N, K = safe_idx.shape
    flat = x.reshape(B, N, -1)
    out = torch.gather(flat, 1, safe_idx.reshape(B, N * K, 1).expand(-1, -1, flat.shape[-1]))
    return out.reshape(B, N, K, *x.shape[2:])


def cross_features(coords, mask, edge_index, n_dist=7, n_angle=6, n_dihedral=6):
    """raw edge features [B,N,K,n_dist^2 + (n_angle-1)^2 + (n_dihedral-2)^2]
    (rnampnn/model/feature.py:386-517, :553-558).  Invalid edges: distances 1e6, cosines 0."""
    B, N, K = edge_index.shape
    invalid = edge_index == -1
    safe = edge_index.clamp(min=0)
    valid = (mask[:, :, None] * (~invalid)).unsqueeze(-1)                  # [B,N,K,1]
    fill = (invalid | (mask == 0)[:, :, None])[..., None, None]

    def nb(n_atoms):
        return _gather_res(coords[:, :, :n_atoms, :], safe).masked_fill(fill, LEPS)

    c = coords[:, :, :n_dist, :]
    diff = c[:, :, None, :, None, :] - nb(n_dist)[:, :, :, None, :, :]
    dists = torch.sqrt((diff ** 2).sum(-1) + SEPS).reshape(B, N, K, -1)
    dists = dists * valid + (1 - valid) * LEPS

    c = coords[:, :, :n_angle, :]
    cu = F.normalize(c[:, :, 1:] - c[:, :, :-1], dim=-1)                   # eps 1e-12
    q = nb(n_angle)
    nu = F.normalize(q[:, :, :, 1:] - q[:, :, :, :-1], dim=-1)
    ang = (cu[:, :, None, :, None, :] * nu[:, :, :, None, :, :]).sum(-1).reshape(B, N, K, -1) * valid

    c = coords[:, :, :n_dihedral, :]
    cv = c[:, :, 1:] - c[:, :, :-1]
    cn = F.normalize(torch.linalg.cross(cv[:, :, :-1], cv[:, :, 1:]), dim=-1, eps=SEPS)
    q = nb(n_dihedral)
    nv = q[:, :, :, 1:] - q[:, :, :, :-1]
    nn_ = F.normalize(torch.linalg.cross(nv[:, :, :, :-1], nv[:, :, :, 1:]), dim=-1, eps=SEPS)
    dih = (cn[:, :, None, :, None, :] * nn_[:, :, :, None, :, :]).sum(-1).reshape(B, N, K, -1) * valid
    return torch.cat([dists, ang, dih], -1)


# --------------------------------------------------------------------------- a-13
def graph_norm(x, mask, scale, shift, n_total=None):
    """``GraphNormalization`` (rnampnn/model/functional.py:18-48) on the first N rows of a graph that
    the reference holds padded to ``n_total`` rows (N itself, or 4500 inside ``RNABert``): the
    padded rows contribute ``mean^2`` each to the variance (:33-38)."""
    B, N, _ = x.shape
    n_total = N if n_total is None else n_total
    m = mask[..., None]
    n = m.sum(1, keepdim=True)
    n_valid = torch.where(n == 0, torch.ones_like(n), n)
    xm = x * m
    mean = xm.sum(1, keepdim=True) / n_valid
    var = (((xm - mean) ** 2).sum(1, keepdim=True) + (n_total - N) * mean ** 2) / n_valid
    y = (x - mean) / torch.sqrt(var + SEPS)
    return (y * scale.reshape(1, 1, -1) + shift.reshape(1, 1, -1)) * m


# --------------------------------------------------------------------------- helpers
def _linears(params, prefix):
    ids = sorted({int(k[len(prefix):].split(".")[0]) for k in params if k.startswith(prefix)
                  and k[len(prefix):].split(".")[0].isdigit()})
    return [(params[f"{prefix}{i}.weight"], params[f"{prefix}{i}.bias"]) for i in ids]


def _gelu_chain(x, lins, last_bare: bool):
    for n, (W, b) in enumerate(lins):
        x = F.linear(x, W, b)
        if not (last_bare and n == len(lins) - 1):
            x = F.gelu(x)
    return x


# --------------------------------------------------------------------------- a-15
def attention(x, mask, in_w, in_b, out_w, out_b, heads: int):
    """``nn.MultiheadAttention(batch_first=True)`` self-attention with ``key_padding_mask = ~mask``
    (rnampnn/model/functional.py:165-168), evaluated on the N real rows only: padded keys get zero
    weight, padded queries are discarded by the caller's mask."""
    B, N, E = x.shape
    hd = E // heads
    qkv = F.linear(x, in_w, in_b)
    q, k, v = (t.reshape(B, N, heads, hd).transpose(1, 2) for t in qkv.split(E, -1))
    s = (q / (hd ** 0.5)) @ k.transpose(-1, -2)
    s = s.masked_fill((mask == 0)[:, None, None, :], float("-inf"))
    p = torch.softmax(s, -1)
    p = torch.nan_to_num(p, nan=0.0)                 # a structure with no valid key
    o = (p @ v).transpose(1, 2).reshape(B, N, E)
    return F.linear(o, out_w, out_b)


def rna_bert(params, prefix, x, mask, heads: int, padding_len: int):
    """``RNABert.forward`` (rnampnn/model/functional.py:161-172) without materialising the
    4500-row padding."""
    n_attn = len({k.split(".")[-2] for k in params if k.startswith(prefix + "bi_attention_layers.")
                  and k.endswith("in_proj_weight")})
    for l in range(n_attn):
        a = prefix + f"bi_attention_layers.{l}."
        x = x + attention(x, mask, params[a + "in_proj_weight"], params[a + "in_proj_bias"],
                          params[a + "out_proj.weight"], params[a + "out_proj.bias"], heads)
        g = prefix + f"graph_norm_layers.{l}."
        x = graph_norm(x, mask, params[g + "scale"], params[g + "shift"], n_total=padding_len)
    x = _gelu_chain(x, _linears(params, prefix + "ffn_layers."), last_bare=True)
    return x * mask[..., None]


# --------------------------------------------------------------------------- a-11/a-12 embeddings
def res_feature(params, coords, mask, cfg, edge_index=None):
    """``ResFeature.forward`` (rnampnn/model/feature.py:573-592)
    -> (raw [B,N,28], res_embedding [B,N,H], res_edge_embedding [B,N,K,H], edge_index [B,N,K])."""
    if edge_index is None:
        edge_index = res_graph(coords, mask, cfg.num_res_neighbours)
    er = cross_features(coords, mask, edge_index, cfg.num_cross_dist_atoms, cfg.num_cross_angle_atoms,
                        cfg.num_cross_dihedral_atoms)
    e = _gelu_chain(er, _linears(params, "res_feature.res_edge_embedding_layers."), last_bare=False)
    dead = ((edge_index == -1) | (mask == 0)[:, :, None])[..., None]
    e = e.masked_fill(dead, 0.0)
    raw = inside_features(coords, mask, cfg.num_inside_dist_atoms, cfg.num_inside_angle_atoms,
                          cfg.num_inside_dihedral_atoms)
    x = F.linear(raw, params["res_feature.raw_project.weight"], params["res_feature.raw_project.bias"])
    x = rna_bert(params, "res_feature.res_embedding.", x, mask, cfg.num_embedding_heads, cfg.padding_len)
    x = graph_norm(x, mask, params["res_feature.graph_norm.scale"], params["res_feature.graph_norm.shift"])
    return raw, x, e, edge_index


# --------------------------------------------------------------------------- a-14
def res_mpnn_layer(params, l, h, e, edge_index, mask):
    """``ResMPNN.forward`` (rnampnn/model/mpnn.py:267-294): masked-mean message update, GraphNorm,
    then the in-place edge update with the UPDATED node features (also on -1 edges, which gather
    row 0)."""
    pre = f"res_mpnn_layers.{l}."
    B, N, K = edge_index.shape
    safe = edge_index.clamp(min=0)
    valid = (edge_index != -1)

    def cat_inputs(hh):
        ctr = hh[:, :, None, :].expand(-1, -1, K, -1)
        return torch.cat([ctr, _gather_res(hh, safe), e], -1)

    hm = h * mask[..., None]
    msg = _gelu_chain(cat_inputs(hm), _linears(params, pre + "message_layers."), last_bare=False)
    msg = msg * valid[..., None].float()
    cnt = valid.sum(-1, keepdim=True).float()
    cnt = torch.where(cnt == 0, torch.ones_like(cnt), cnt)
    h = (h + msg.sum(2) / cnt) * mask[..., None]
    h = graph_norm(h, mask, params[pre + "graph_norm.scale"], params[pre + "graph_norm.shift"])
    e = e + _gelu_chain(cat_inputs(h), _linears(params, pre + "edge_layers."), last_bare=False)
    return h, e


# --------------------------------------------------------------------------- a-16
def raw_ffn(params, raw, mask):
    """``RawFFN.forward`` (rnampnn/model/functional.py:191-202)."""
    x = _gelu_chain(raw, _linears(params, "raw_embedding.raw_ffn."), last_bare=True)
    return graph_norm(x, mask, params["raw_embedding.graph_norm.scale"], params["raw_embedding.graph_norm.shift"])


def readout(params, emb, mask):
    """``Readout.forward`` (rnampnn/model/functional.py:76-90): logits * mask."""
    return _gelu_chain(emb, _linears(params, "readout.readout_layers."), last_bare=True) * mask[..., None]


def num_mpnn_layers(params):
    return 1 + max(int(k.split(".")[1]) for k in params if k.startswith("res_mpnn_layers."))


@torch.no_grad()
def forward(params, coords, mask, cfg, edge_index=None, return_layers=False):
    """``RNAMPNN.forward`` (rnampnn/model/rnampnn.py:161-185) + softmax/argmax (:201-207).
    Returns dict(raw, h0, e0, edge_index, h, e, fused, raw_embedding, embedding, logits, probs, samples)."""
    raw, h, e, edge_index = res_feature(params, coords, mask, cfg, edge_index)
    out = {"raw": raw, "h0": h, "e0": e, "edge_index": edge_index}
    layers = []
    for l in range(num_mpnn_layers(params)):
        h, e = res_mpnn_layer(params, l, h, e, edge_index, mask)
        if return_layers:
            layers.append((h, e))
    fused = rna_bert(params, "post_fusion.", h, mask, cfg.num_post_fusion_heads, cfg.padding_len)
    remb = raw_ffn(params, raw, mask)
    emb = torch.cat([fused, remb], -1)
    logits = readout(params, emb, mask)
    probs = F.softmax(logits, -1)
    out.update(h=h, e=e, fused=fused, raw_embedding=remb, embedding=emb, logits=logits, probs=probs,
               samples=probs.argmax(-1))
    if return_layers:
        out["layers"] = layers
    return out
