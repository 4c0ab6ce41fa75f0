"""Forward engine of the rdesign path: sequences the C-ABI kernels over a padded batch.

``RDesignEngine`` is what both the drop-in ``rdesign.model.rdesign.RNAModel`` and
``bench.py`` call.  It owns the packed (kernel-layout) copies of the parameters and
nothing else; activations are torch-allocated per call.

Reference call sites: ``RNAModel.forward`` rdesign/model/rdesign.py:82-88,
``RNAFeatures.forward`` rdesign/model/feature.py:156-238, ``MPNNLayer.forward``
rdesign/model/mpnn.py:31-38, ``Readout`` rdesign/model/functional.py:123-127.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Mapping, Optional

import torch

from . import ops
from ._lib import PREC_BF16, PREC_FP32, PREC_FP32X3


@dataclass
class LayerWeights:
    message: List[ops.PackedLinear]
    dense: List[ops.PackedLinear]
    norm1_w: torch.Tensor
    norm1_b: torch.Tensor
    norm2_w: torch.Tensor
    norm2_b: torch.Tensor
    # bf16 tensor-core arm: W_0 split by column block (edge | centre | neighbour)
    msg_w16: Optional[torch.Tensor] = None        # [S*128,128] bf16: W_0e, W_1, ..
    msg_bias: Optional[List[Optional[torch.Tensor]]] = None
    proj: Optional[ops.PackedLinear] = None       # h_V -> [Pc | Pn]  (128 -> 256, bias [b_0 | 0])
    proj16: Optional[tuple] = None                # (W bf16 [256,128], bias fp32 [256]) for the tcgen05 node kernel
    dense16: Optional[list] = None                # [(W bf16 [N,K], bias fp32 [N])] dense stages, or None if unsupported
    msg_x3: Optional[list] = None                 # message stages of the fp32-parity tensor-core arm (W_0e alone, no b_0)


def _indexed(params: Mapping[str, torch.Tensor], prefix: str):
    ids = sorted({int(k[len(prefix):].split(".")[0]) for k in params if k.startswith(prefix)})
    return [(params[f"{prefix}{i}.weight"], params[f"{prefix}{i}.bias"]) for i in ids]


class RDesignWeights:
    """Kernel-layout copy of an ``RNAModel`` state_dict (same key names as the reference)."""

    def __init__(self, params: Mapping[str, torch.Tensor], device):
        p = {k: v.detach().to(device=device, dtype=torch.float32).contiguous() for k, v in params.items()
             if torch.is_tensor(v)}
        self.device = torch.device(device)
        if p["features.node_embedding.weight"].shape[0] != ops.HIDDEN:
            raise ValueError("rmpnn_b200 kernels are specialised for hidden_dim = 128")
        self.node_emb = ops.pack_linear(p["features.node_embedding.weight"], p["features.node_embedding.bias"])
        self.edge_emb = ops.pack_linear(p["features.edge_embedding.weight"], p["features.edge_embedding.bias"])
        self.norm_nodes = (p["features.norm_nodes.gain"], p["features.norm_nodes.bias"])
        self.norm_edges = (p["features.norm_edges.gain"], p["features.norm_edges.bias"])
        self.edge_emb16 = ops.pack_edge_embedding(p["features.edge_embedding.weight"])
        self.edge_emb128 = ops.pack_linear(torch.nn.functional.pad(p["features.edge_embedding.weight"],
                                                                   (0, 128 - ops.EDGE_RAW)),
                                           p["features.edge_embedding.bias"])
        self.edge_emb_bias = p["features.edge_embedding.bias"]
        n_layers = 1 + max(int(k.split(".")[1]) for k in p if k.startswith("mpnn_layers."))
        self.layers: List[LayerWeights] = []
        for l in range(n_layers):
            pre = f"mpnn_layers.{l}."
            msg = [ops.pack_linear(w, b) for w, b in _indexed(p, pre + "message_layers.")]
            den = [ops.pack_linear(w, b) for w, b in _indexed(p, pre + "dense.")]
            if len(msg) > 4 or len(den) > 4:
                raise ValueError("rmpnn_b200 fuses at most 4 Linear stages per launch")
            lw = LayerWeights(msg, den, p[pre + "norm1.weight"], p[pre + "norm1.bias"],
                              p[pre + "norm2.weight"], p[pre + "norm2.bias"])
            mw = _indexed(p, pre + "message_layers.")
            H = ops.HIDDEN
            if mw[0][0].shape == (H, 3 * H) and all(w.shape == (H, H) for w, _ in mw[1:]):
                W0, b0 = mw[0]
                lw.msg_w16 = torch.cat([W0[:, :H]] + [w for w, _ in mw[1:]], 0).to(torch.bfloat16).contiguous()
                lw.msg_bias = [None] + [b for _, b in mw[1:]]
                Wp = torch.cat([W0[:, H:2 * H], W0[:, 2 * H:]], 0)
                bp = torch.cat([b0, torch.zeros_like(b0)], 0).contiguous()
                lw.proj = ops.pack_linear(Wp, bp)
                lw.proj16 = (Wp.to(torch.bfloat16).contiguous(), bp)
                # fp32-parity tensor-core arm: W_0's edge columns alone (its bias rides on the centre projection)
                lw.msg_x3 = [ops.pack_linear(W0[:, :H].contiguous(), torch.zeros_like(b0))] + msg[1:]
            dw = _indexed(p, pre + "dense.")
            widths_ok = all(w.shape[0] in (128, 256) and w.shape[1] in (128, 256) for w, _ in dw)
            if widths_ok and dw[0][0].shape[1] == H and dw[-1][0].shape[0] == H and len(dw) <= 4:
                lw.dense16 = [(w.to(torch.bfloat16).contiguous(), b.contiguous()) for w, b in dw]
            self.layers.append(lw)
        ro = _indexed(p, "readout.readout_layers.")
        self.readout_hidden = [ops.pack_linear(w, b) for w, b in ro[:-1]]
        self.readout_w, self.readout_b = ro[-1]


class RDesignEngine:
    """featurizer -> N message-passing layers -> readout head, all on the C-ABI kernels."""

    def __init__(self, params: Mapping[str, torch.Tensor], top_k: int, scale: float = 30.0,
                 device="cuda", precision: int = PREC_FP32):
        if not torch.cuda.is_available():
            raise RuntimeError("rmpnn_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.w = RDesignWeights(params, device)
        self.top_k = int(top_k)
        self.scale = float(scale)
        self.precision = precision
        self.kernel_timer = None      # bench.py: list collecting (start, end) events of the message kernel
        if precision not in (PREC_FP32, PREC_BF16, PREC_FP32X3):
            raise ValueError("precision must be RMP_PREC_FP32, RMP_PREC_BF16 or RMP_PREC_FP32X3")

    def dominant_kernel_name(self) -> str:
        if self.precision == PREC_BF16:
            return "edge_mlp_tc_kernel (edge message MLP + neighbour sum, bf16 tcgen05/TMEM/TMA)"
        if self.precision == PREC_FP32X3:
            return "row_mlp_x3_kernel (edge message MLP + segmented sum, split-bf16 fp32 parity on tcgen05)"
        return "row_mlp_kernel<64> (edge message MLP + segmented sum, fp32 CUDA cores)"

    # ---- stages -----------------------------------------------------------------
    def graph(self, X, mask, sizes=None) -> ops.Graph:
        return ops.build_graph(X, mask, self.top_k, sizes)

    def embed(self, X, g: ops.Graph):
        """h_V [Nv,128], h_E [Ne,128]   (rdesign/model/feature.py:179-224)."""
        bf16 = self.precision == PREC_BF16
        x3 = self.precision == PREC_FP32X3
        node_raw, edge_raw, frames = ops.raw_features(X, g, want_edges=not bf16, return_frames=True, pad_edges=x3)
        dev = node_raw.device
        h_V = torch.empty(g.Nv, ops.HIDDEN, dtype=torch.float32, device=dev)
        ops.row_mlp([(node_raw, None)], [(self.w.node_emb, False)], ops.EPI_NORMALIZE, h_V, g.Nv,
                    gain=self.w.norm_nodes[0], beta=self.w.norm_nodes[1])
        if bf16:
            # fused tcgen05 edge featurizer: h_E is produced directly as bf16 (it is layer-invariant)
            h_E = ops.edge_embed_bf16(X, frames, g, self.w.edge_emb16, self.w.edge_emb_bias,
                                      self.w.norm_edges[0], self.w.norm_edges[1])
        else:
            h_E = torch.empty(g.Ne, ops.HIDDEN, dtype=torch.float32, device=dev)
            # fp32x3: the 115 features zero-padded to K = 128, Linear + Normalize in one split-operand tcgen05 launch
            ops.row_mlp([(edge_raw, None)], [(self.w.edge_emb128 if x3 else self.w.edge_emb, False)], ops.EPI_NORMALIZE, h_E,
                        g.Ne, gain=self.w.norm_edges[0], beta=self.w.norm_edges[1], split=x3)
        return h_V, h_E

    def layer(self, l: int, h_V, h_E, g: ops.Graph, agg: Optional[torch.Tensor] = None):
        """One MPNN layer (rdesign/model/rdesign.py:85 + rdesign/model/mpnn.py:31-38)."""
        lw = self.w.layers[l]
        if agg is None:
            agg = torch.zeros_like(h_V)
        if self.kernel_timer is not None:
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream())
        if self.precision == PREC_BF16:
            if lw.msg_w16 is None:
                raise ValueError("bf16 layer kernels need message layers of shape 384->128->128..")
            # per-node projections of W_0's centre / neighbour column blocks: emitted by the previous
            # layer's node kernel, or computed here for the first layer
            tagged = getattr(self, "_next_P", None)
            self._next_P = None
            P = None
            if tagged is not None and tagged[0] == (l, h_V.data_ptr(), g.Nv):
                P = tagged[1]              # emitted by the previous layer's node kernel for exactly this h_V
            if P is None:
                P = self._proj_buffers(g.Nv, h_V.device)
                ops.node_mlp_bf16(h_V, [(*lw.proj16, 2)], g.Nv, p_out=P[0], pn_out=P[1])
            if self.kernel_timer is not None:
                ev[0].record(torch.cuda.current_stream())
            ops.edge_mlp_bf16(h_E, lw.msg_w16, lw.msg_bias, P[0], P[1], g.ctr, g.nbr, g.row_ptr, agg, 1.0 / self.scale)
        elif self.precision == PREC_FP32X3 and lw.msg_x3 is not None:
            # [Pc + b_0 | Pn] = h_V [W_0c | W_0n]^T per node, then the message MLP over h_E alone with the two
            # projections gathered into stage 0's pre-activation
            P = torch.empty(g.Nv, 2 * ops.HIDDEN, dtype=torch.float32, device=h_V.device)
            ops.row_mlp([(h_V, None)], [(lw.proj, False)], ops.EPI_STORE, P, g.Nv, split=True)
            if self.kernel_timer is not None:
                ev[0].record(torch.cuda.current_stream())
            ops.row_mlp([(h_E, None)], [(m, True) for m in lw.msg_x3], ops.EPI_SEGSUM, agg, g.Ne, seg_id=g.ctr,
                        alpha=1.0 / self.scale, split=True, adds=[(P, g.ctr), (P[:, ops.HIDDEN:], g.nbr)],
                        add_stride=2 * ops.HIDDEN)
        else:
            ops.row_mlp([(h_E, None), (h_V, g.ctr), (h_V, g.nbr)], [(m, True) for m in lw.message],
                        ops.EPI_SEGSUM, agg, g.Ne, seg_id=g.ctr, alpha=1.0 / self.scale)
        if self.kernel_timer is not None:
            ev[1].record(torch.cuda.current_stream())
            self.kernel_timer.append(ev)
        if self.precision == PREC_BF16 and lw.dense16 is not None:
            return self._node_update_tc(l, h_V, agg, g)
        x1 = ops.add_layernorm(h_V, agg, lw.norm1_w, lw.norm1_b, 1e-5, zero_b=True)
        out = torch.empty_like(h_V)
        stages = [(d, True) for d in lw.dense[:-1]] + [(lw.dense[-1], False)]
        ops.row_mlp([(x1, None)], stages, ops.EPI_RESID_LN, out, g.Nv, gain=lw.norm2_w, beta=lw.norm2_b, resid=x1,
                    split=self.precision == PREC_FP32X3 and ops.x3_eligible([(x1, None)], stages))
        return out

    def _node_update_tc(self, l, h_V, agg, g):
        """norm1 + dense FFN + norm2 on tcgen05, plus the next layer's [Pc | Pn] projections."""
        lw = self.w.layers[l]
        stages = [(w, b, 0) for w, b in lw.dense16[:-1]] + [(*lw.dense16[-1], 1)]
        nxt = self.w.layers[l + 1] if l + 1 < len(self.w.layers) else None
        pc = pn = None
        if nxt is not None and nxt.proj16 is not None:
            stages.append((*nxt.proj16, 2))
            pc, pn = self._proj_buffers(g.Nv, h_V.device)
        out = torch.empty_like(h_V)
        x1 = torch.empty_like(h_V)
        ops.node_mlp_bf16(h_V, stages, g.Nv, agg=agg, ln1=(lw.norm1_w, lw.norm1_b), ln2=(lw.norm2_w, lw.norm2_b),
                          x1=x1, h_out=out, p_out=pc, pn_out=pn)
        self._next_P = ((l + 1, out.data_ptr(), g.Nv), (pc, pn)) if pc is not None else None
        return out

    @staticmethod
    def _proj_buffers(n, device):
        return (torch.empty(n, ops.HIDDEN, dtype=torch.float32, device=device),
                torch.empty(n, ops.HIDDEN, dtype=torch.bfloat16, device=device))

    def encode(self, X, mask, sizes=None, return_layers=False):
        g = self.graph(X, mask, sizes)
        h_V, h_E = self.embed(X, g)
        agg = torch.zeros_like(h_V)
        layers = []
        self._next_P = None
        for l in range(len(self.w.layers)):
            h_V = self.layer(l, h_V, h_E, g, agg)
            if return_layers:
                layers.append(h_V)
        return (h_V, h_E, g, layers) if return_layers else (h_V, h_E, g)

    def readout(self, h_V, temperature: float = 0.0, seed: int = 0):
        """logits, probs, samples   (rdesign/model/functional.py:123-127, rdesign.py:112-115)."""
        x = h_V
        for pl in self.w.readout_hidden:
            y = torch.empty(x.shape[0], pl.n_out, dtype=torch.float32, device=x.device)
            ops.row_mlp([(x, None)], [(pl, True)], ops.EPI_STORE, y, x.shape[0],
                        split=self.precision == PREC_FP32X3 and ops.x3_eligible([(x, None)], [(pl, True)]))
            x = y
        return ops.readout_head(x, self.w.readout_w, self.w.readout_b, temperature, seed)

    @torch.no_grad()
    def capture(self, X, mask, sizes, temperature: float = 0.0, seed: int = 0):
        """Capture one forward for fixed shapes into a CUDA graph.

        Returns ``(replay, static)``: ``static["X"]`` / ``static["mask"]`` are the input buffers to
        overwrite, ``replay()`` launches the whole forward (about 30 kernels) with one graph launch and
        returns the same dict as :meth:`forward` (tensors are reused between replays).  ``sizes=(Nv, Ne)``
        must be given: a captured forward cannot read the two counters back to the host.
        """
        static = {"X": X.clone(), "mask": mask.clone()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up outside the capture (lazy init, attributes)
            for _ in range(2):
                self.forward(static["X"], static["mask"], sizes, temperature, seed)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self.forward(static["X"], static["mask"], sizes, temperature, seed)

        def replay():
            graph.replay()
            return out
        return replay, static

    @torch.no_grad()
    def forward(self, X, mask, sizes=None, temperature: float = 0.0, seed: int = 0) -> Dict[str, torch.Tensor]:
        h_V, h_E, g = self.encode(X, mask, sizes)
        logits, probs, samples = self.readout(h_V, temperature, seed)
        return {"h_V": h_V, "logits": logits, "probs": probs, "samples": samples, "graph": g, "h_E": h_E}
