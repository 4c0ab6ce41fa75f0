"""Forward engine of the secondary (rnampnn) path: sequences the C-ABI kernels over a padded batch.

``RNAMPNNEngine`` is what the drop-in ``rnampnn.model.rnampnn.RNAMPNN`` and ``bench.py --model rnampnn``
call.  It owns the packed (kernel-layout) parameters; activations are torch-allocated per call.  The
reference keeps the padded [B,N,..] layout end to end, so rows here are dense residues b*N+i and edges are
(row, slot k); padded rows are carried as zeros exactly like the reference carries them.

Reference call sites: ``RNAMPNN.forward`` / ``embedding`` rnampnn/model/rnampnn.py:161-185,269-278,
``ResFeature.forward`` rnampnn/model/feature.py:573-592, ``ResMPNN.forward`` rnampnn/model/mpnn.py:267-294,
``RNABert.forward`` rnampnn/model/functional.py:161-172, ``RawFFN`` :191-202, ``Readout`` :76-90.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Tuple

import torch

from . import ops
from ._lib import PREC_BF16, PREC_FP32


def _indexed(p: Mapping[str, torch.Tensor], prefix: str):
    ids = sorted({int(k[len(prefix):].split(".")[0]) for k in p if k.startswith(prefix)
                  and k[len(prefix):].split(".")[0].isdigit()})
    return [(p[f"{prefix}{i}.weight"], p[f"{prefix}{i}.bias"]) for i in ids]


def _stack(p, prefix, last_bare: bool):
    """[(PackedLinear, gelu)] of an nn.Sequential of Linear/GELU/Dropout under ``prefix``."""
    lin = _indexed(p, prefix)
    return [(ops.pack_linear(w, b), not (last_bare and n == len(lin) - 1)) for n, (w, b) in enumerate(lin)]


@dataclass
class BertWeights:
    """``RNABert`` (rnampnn/model/functional.py:93-127)."""
    attn: List[Tuple[ops.PackedLinear, ops.PackedLinear]] = field(default_factory=list)   # (in_proj, out_proj)
    norms: List[Tuple[torch.Tensor, torch.Tensor]] = field(default_factory=list)
    ffn: List[Tuple[ops.PackedLinear, bool]] = field(default_factory=list)

    @staticmethod
    def load(p, prefix):
        bw = BertWeights()
        n_attn = len([k for k in p if k.startswith(prefix + "bi_attention_layers.") and k.endswith("in_proj_weight")])
        for l in range(n_attn):
            a = f"{prefix}bi_attention_layers.{l}."
            bw.attn.append((ops.pack_linear(p[a + "in_proj_weight"], p[a + "in_proj_bias"]),
                            ops.pack_linear(p[a + "out_proj.weight"], p[a + "out_proj.bias"])))
            g = f"{prefix}graph_norm_layers.{l}."
            bw.norms.append((p[g + "scale"].reshape(-1).contiguous(), p[g + "shift"].reshape(-1).contiguous()))
        bw.ffn = _stack(p, prefix + "ffn_layers.", last_bare=True)
        return bw


@dataclass
class ResLayerWeights:
    message: List[Tuple[ops.PackedLinear, bool]]
    edge: List[Tuple[ops.PackedLinear, bool]]
    gn: Tuple[torch.Tensor, torch.Tensor]


class RNAMPNNWeights:
    """Kernel-layout copy of an ``RNAMPNN`` state_dict (same key names as the reference)."""

    def __init__(self, params: Mapping[str, torch.Tensor], device):
        p = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in params.items()
             if torch.is_tensor(v)}
        self.device = torch.device(device)
        H = p["res_feature.raw_project.weight"].shape[0]
        if H != ops.HIDDEN:
            raise ValueError("rmpnn_b200 kernels are specialised for res_embedding_dim = 128")
        gn = lambda k: (p[k + ".scale"].reshape(-1).contiguous(), p[k + ".shift"].reshape(-1).contiguous())
        self.raw_project = ops.pack_linear(p["res_feature.raw_project.weight"], p["res_feature.raw_project.bias"])
        self.embed_bert = BertWeights.load(p, "res_feature.res_embedding.")
        self.embed_gn = gn("res_feature.graph_norm")
        self.edge_embed = _stack(p, "res_feature.res_edge_embedding_layers.", last_bare=False)
        n_layers = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("res_mpnn_layers."))
        self.layers = [ResLayerWeights(_stack(p, f"res_mpnn_layers.{l}.message_layers.", False),
                                       _stack(p, f"res_mpnn_layers.{l}.edge_layers.", False),
                                       gn(f"res_mpnn_layers.{l}.graph_norm")) for l in range(n_layers)]
        self.post_fusion = BertWeights.load(p, "post_fusion.")
        self.raw_ffn = _stack(p, "raw_embedding.raw_ffn.", last_bare=True)
        self.raw_gn = gn("raw_embedding.graph_norm")
        ro = _indexed(p, "readout.readout_layers.")
        self.readout_hidden = [(ops.pack_linear(w, b), True) for w, b in ro[:-1]]
        self.readout_w, self.readout_b = ro[-1]


class RNAMPNNEngine:
    """graph + raw features -> embeddings -> ResMPNN layers -> attention fusion -> readout, on the C-ABI kernels."""

    def __init__(self, params: Mapping[str, torch.Tensor], num_neighbours: int = 3, inside=(7, 6, 6), cross=(7, 6, 6),
                 embed_heads: int = 8, fusion_heads: int = 8, padding_len: int = 4500, device="cuda",
                 precision: int = PREC_FP32):
        if not torch.cuda.is_available():
            raise RuntimeError("rmpnn_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.w = RNAMPNNWeights(params, device)
        self.K = int(num_neighbours)
        self.inside, self.cross = tuple(inside), tuple(cross)
        self.embed_heads, self.fusion_heads = int(embed_heads), int(fusion_heads)
        self.padding_len = int(padding_len)
        self.kernel_timer = None
        if precision not in (PREC_FP32, PREC_BF16):
            raise ValueError("precision must be RMP_PREC_FP32 or RMP_PREC_BF16")
        # RMP_PREC_BF16: every Linear stack runs on the tcgen05 row-MLP kernel (bf16 operands, fp32 accumulate),
        # except a first stage fed by raw geometry (28 / 90 unbounded distances), which stays in fp32
        self.precision = precision
        # the two stacks that consume raw geometry (embedding FFN, RawFFN) feed a GraphNorm directly; on badly scaled
        # inputs (atoms parked at the origin) their operand rounding is what the GraphNorm amplifies.  They run with
        # fp16 operands (11 significant bits, saturating conversions) instead of bf16; "fp32" keeps them on the
        # CUDA-core kernel, "bf16" treats them like every other stack.
        self.raw_stacks = "fp16"

    @classmethod
    def from_config(cls, params, cfg, device="cuda", precision: int = PREC_FP32):
        """``cfg``: any object with the ``RNAMPNN`` constructor's field names (rnampnn/model/rnampnn.py:19-54)."""
        return cls(params, cfg.num_res_neighbours,
                   (cfg.num_inside_dist_atoms, cfg.num_inside_angle_atoms, cfg.num_inside_dihedral_atoms),
                   (cfg.num_cross_dist_atoms, cfg.num_cross_angle_atoms, cfg.num_cross_dihedral_atoms),
                   cfg.num_embedding_heads, cfg.num_post_fusion_heads, cfg.padding_len, device, precision)

    # ---- building blocks ----------------------------------------------------------
    @staticmethod
    def _mlp(segments, stages, rows, row_scale=None, tensor_core: bool = False, fp16: bool = False, out_bf16=None):
        """Chained Linear(+GELU) stages -> fp32 [rows, n_out], result rows times ``row_scale``.  With ``tensor_core``
        the stages the tcgen05 kernel can take run on it; a leading stage it cannot take (raw geometry, K = 28 / 90)
        runs on the fp32 kernel first."""
        dev = segments[0][0].device
        groups = [(stages, False)]                    # [(stages, tensor_core)]
        if tensor_core:
            if ops.tc_eligible(segments, stages):
                groups = [(stages, True)]
            elif len(stages) > 1 and stages[0][0].n_out % 128 == 0 and ops.tc_eligible([], stages[1:]):
                groups = [(stages[:1], False), (stages[1:], True)]
        out = None
        for gi, (group, tc) in enumerate(groups):
            for s0 in range(0, len(group), ops._lib.MAX_STAGES):
                chunk = group[s0:s0 + ops._lib.MAX_STAGES]
                last = gi == len(groups) - 1 and s0 + ops._lib.MAX_STAGES >= len(group)
                out = torch.empty(rows, chunk[-1][0].n_out, dtype=torch.float32, device=dev)
                ops.row_mlp(segments, chunk, ops.EPI_STORE, out, rows, row_scale=row_scale if last else None,
                            tensor_core=tc, fp16=fp16, out_bf16=out_bf16 if (last and tc) else None)
                segments = [(out, None)]
        return out

    @property
    def tc(self) -> bool:
        return getattr(self, "precision", PREC_FP32) == PREC_BF16

    def bert(self, bw: BertWeights, x, mask_flat, B, N, heads, pre: Optional[list] = None, tc: Optional[bool] = None,
             fp16: bool = False):
        """``RNABert.forward`` on the N real rows (the 4500-row padding only enters GraphNorm's variance).
        ``pre``: Linear stages applied before the FFN when there is no attention layer (raw_project)."""
        rows = B * N
        tc = self.tc if tc is None else tc
        stages = list(bw.ffn)
        if pre:
            if bw.attn:
                x = self._mlp([(x, None)], pre, rows, tensor_core=tc, fp16=fp16)
            else:
                stages = pre + stages
        if len(bw.attn) and self.padding_len < N:
            raise ValueError(f"padding_len={self.padding_len} is shorter than the batch (N={N})")
        for (in_proj, out_proj), (scale, shift) in zip(bw.attn, bw.norms):
            seg, st = [(x, None)], [(in_proj, False)]
            if tc and not fp16 and ops.attention_tc_eligible(x.shape[1], heads) and ops.tc_eligible(seg, st):
                # bf16 arm: the in-projection writes only its bf16 side copy, attention runs on tcgen05 (QK^T / PV tiles)
                qkv16 = torch.empty(rows, 3 * x.shape[1], dtype=torch.bfloat16, device=x.device)
                ops.row_mlp(seg, st, ops.EPI_STORE, None, rows, tensor_core=True, out_bf16=qkv16)
                att = ops.res_attention_tc(qkv16, mask_flat, B, N, heads)
            else:
                qkv = self._mlp(seg, st, rows, tensor_core=tc, fp16=fp16)
                att = ops.res_attention(qkv, mask_flat, B, N, heads)
            y = self._mlp([(att, None)], [(out_proj, False)], rows, tensor_core=tc, fp16=fp16)
            x = ops.graph_norm(x, y, mask_flat, B, N, scale, shift, n_total=self.padding_len)
        return self._mlp([(x, None)], stages, rows, row_scale=mask_flat, tensor_core=tc, fp16=fp16)

    def features(self, coords, mask, edge_index=None):
        """``ResFeature.forward`` -> (raw [B*N,28], h [B*N,128], e [B*N*K,128], ResGraph)."""
        B, N = mask.shape
        mask_flat = mask.reshape(-1)
        if edge_index is None:
            edge_index = ops.res_graph(coords, mask, self.K)
        g = ops.res_edge_rows(edge_index, mask)
        raw, edge_raw = ops.res_raw_features(coords, mask, edge_index, self.inside, self.cross)
        # bf16 arm: h and e carry bf16 side copies (written by the producing kernels' epilogues) that the message /
        # edge kernels gather with 256-bit loads; the fp32 tensors stay the masters (residuals, GraphNorm)
        self._e16 = torch.empty(B * N * g.K, ops.HIDDEN, dtype=torch.bfloat16, device=raw.device) if self.tc else None
        e = self._mlp([(edge_raw, None)], self.w.edge_embed, B * N * g.K, row_scale=g.edge_keep, tensor_core=self.tc,
                      out_bf16=self._e16)
        x = self.bert(self.w.embed_bert, raw, mask_flat, B, N, self.embed_heads, pre=[(self.w.raw_project, False)],
                      tc=self.tc and self.raw_stacks != "fp32", fp16=self.raw_stacks == "fp16")
        self._h16 = torch.empty(B * N, ops.HIDDEN, dtype=torch.bfloat16, device=raw.device) if self.tc else None
        h = ops.graph_norm(x, None, mask_flat, B, N, *self.w.embed_gn, out_bf16=self._h16)
        return raw, h, e, g

    def layer(self, l: int, h, e, g: ops.ResGraph, mask_flat, agg=None):
        """``ResMPNN.forward`` (rnampnn/model/mpnn.py:267-294); ``e`` is updated in place like the reference."""
        lw = self.w.layers[l]
        rows, edges = g.B * g.N, g.B * g.N * g.K
        if agg is None:
            agg = torch.zeros_like(h)
        if self.kernel_timer is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream())
        seg = [(h, g.ctr_row), (h, g.nbr_row), (e, None)]
        tc = self.tc and ops.tc_eligible(seg, lw.message) and ops.tc_eligible(seg, lw.edge)
        h16 = getattr(self, "_h16", None) if tc else None
        e16 = getattr(self, "_e16", None) if tc else None
        side = h16 is not None and e16 is not None and h16.shape[0] == h.shape[0] and e16.shape[0] == e.shape[0]
        if side:
            seg = [(h16, g.ctr_row), (h16, g.nbr_row), (e16, None)]
        # masked mean of the messages: sum_k valid_k * m_k / max(#valid, 1)  (mpnn.py:191-192, :212-219)
        ops.row_mlp(seg, lw.message, ops.EPI_SEGSUM, agg, edges, seg_id=g.ctr_row, alpha=1.0, row_scale=g.edge_w,
                    tensor_core=tc)
        h = ops.graph_norm(h, agg, mask_flat, g.B, g.N, *lw.gn, zero_b=True, out_bf16=h16 if side else None)
        # e += MLP([h_i | h_j | e]) with the UPDATED h, on every slot (mpnn.py:229-265)
        seg = [(h16, g.ctr_row), (h16, g.nbr_row), (e16, None)] if side else [(h, g.ctr_row), (h, g.nbr_row), (e, None)]
        ops.row_mlp(seg, lw.edge, ops.EPI_RESID_ADD, e, edges, resid=e, tensor_core=tc, out_bf16=e16 if side else None)
        if self.kernel_timer is not None:
            ev[1].record(torch.cuda.current_stream())
            self.kernel_timer.append(ev)
        return h, e

    def encode(self, coords, mask, edge_index=None, return_layers=False):
        B, N = mask.shape
        mask_flat = mask.reshape(-1)
        raw, h, e, g = self.features(coords, mask, edge_index)
        out = {"raw": raw, "graph": g}
        if return_layers:
            out["h0"], out["e0"], out["layers"] = h, e.clone(), []
        agg = torch.zeros_like(h)
        for l in range(len(self.w.layers)):
            h, e = self.layer(l, h, e, g, mask_flat, agg)
            if return_layers:
                out["layers"].append(h)
        fused = self.bert(self.w.post_fusion, h, mask_flat, B, N, self.fusion_heads)
        remb = self._mlp([(raw, None)], self.w.raw_ffn, B * N, tensor_core=self.tc and self.raw_stacks != "fp32",
                         fp16=self.raw_stacks == "fp16")
        remb = ops.graph_norm(remb, None, mask_flat, B, N, *self.w.raw_gn)
        out.update(h=h, e=e, fused=fused, raw_embedding=remb)
        return out

    def readout(self, fused, remb, mask_flat, temperature: float = 0.0, seed: int = 0):
        """``Readout(cat[res | raw])`` + softmax/argmax (rnampnn/model/functional.py:76-90, rnampnn.py:201-207)."""
        rows = fused.shape[0]
        if self.w.readout_hidden:
            x = self._mlp([(fused, None), (remb, None)], self.w.readout_hidden, rows, tensor_core=self.tc)
        else:
            x = torch.cat([fused, remb], -1)
        return ops.readout_head_masked(x, self.w.readout_w, self.w.readout_b, mask_flat, temperature, seed)

    @torch.no_grad()
    def forward(self, coords, mask, edge_index=None, temperature: float = 0.0, seed: int = 0,
                return_layers: bool = False) -> Dict[str, torch.Tensor]:
        """All tensors keep the reference's padded layout, flattened over (b, i): logits [B*N,4] etc."""
        ops._need_cuda(coords, mask)
        coords, mask = ops._f32c(coords), ops._f32c(mask)
        out = self.encode(coords, mask, edge_index, return_layers)
        logits, probs, samples = self.readout(out["fused"], out["raw_embedding"], mask.reshape(-1), temperature, seed)
        out.update(logits=logits, probs=probs, samples=samples)
        return out

    @torch.no_grad()
    def capture(self, coords, mask, temperature: float = 0.0, seed: int = 0):
        """Capture one forward for fixed shapes into a CUDA graph: ``(replay, static)`` like ``RDesignEngine.capture``
        (no host read-back exists on this path, so no sizes are needed)."""
        static = {"coords": coords.clone(), "mask": mask.clone()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.forward(static["coords"], static["mask"], None, temperature, seed)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward(static["coords"], static["mask"], None, temperature, seed)

        def replay():
            graph.replay()
            return out
        return replay, static
