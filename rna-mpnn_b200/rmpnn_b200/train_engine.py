"""Training step of the rdesign path on the sm_100a kernels (BASELINE configs[3]).

``training_step`` of the reference (rdesign/model/rdesign.py:95-104) = forward with dropout, cross entropy, backward,
gradient all-reduce (Lightning DDP), Adam (rdesign.py:90-93).  Here:

* everything per EDGE -- the three Linear + GELU + Dropout message stages on ``[h_E | h_V[ctr] | h_V[nbr]]``, the
  ``scatter_sum / scale`` and their backward (96 % of the step's flops and bytes) -- runs on the C-ABI kernels:
  tcgen05 row GEMMs (``rmp_row_mlp_tc``), the weight-gradient GEMM ``rmp_gemm_dw_bf16``, and the HBM-bound
  activation / scatter passes of ``csrc/train.cu``; dropout masks are regenerated from Philox, never stored;
  ``d h_E`` is accumulated over the layers in one fp32 buffer by the GEMM epilogue;
* the per-NODE remainder (LayerNorms, the dense FFN on [Nv,128] rows, readout, loss) is composed from torch-CUDA ops
  under bf16 autocast -- 4 % of the work, Nv rows instead of Ne;
* gradients live in ONE flat fp32 buffer; ``step`` all-reduces it with a single NCCL call on a side stream and applies
  the fused Adam kernel ``rmp_adam_step`` with the 1/world scale folded in.

``EdgeMessageFn`` / ``EdgeEmbedFn`` are ``torch.autograd.Function`` wrappers so that autograd only orders the blocks.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from . import ops


def _pl(weight: torch.Tensor, bias: Optional[torch.Tensor]) -> ops.PackedLinear:
    """bf16 nn.Linear-layout weight + fp32 bias for the tcgen05 row GEMM (the fp32 transposed copy is not needed)."""
    n_out, k_in = weight.shape
    b = torch.zeros(n_out, dtype=torch.float32, device=weight.device) if bias is None else bias.detach().float().contiguous()
    return ops.PackedLinear(None, b, k_in, n_out, weight.detach().to(torch.bfloat16).contiguous(), None)


def _gemm16(segments, weight, bias, rows, resid_acc=None):
    """bf16 rows x W^T (+ bias) on tcgen05: -> bf16 [rows, N], or accumulated into the fp32 buffer ``resid_acc``."""
    pl = _pl(weight, bias)
    if resid_acc is not None:
        ops.row_mlp(segments, [(pl, False)], ops.EPI_RESID_ADD, resid_acc, rows, resid=resid_acc, tensor_core=True)
        return resid_acc
    out16 = torch.empty(rows, pl.n_out, dtype=torch.bfloat16, device=weight.device)
    ops.row_mlp(segments, [(pl, False)], ops.EPI_STORE, None, rows, tensor_core=True, out_bf16=out16)
    return out16


class EdgeMessageFn(torch.autograd.Function):
    """agg[v] = sum_{e: ctr[e]=v} Dropout(GELU(W3 Dropout(GELU(W2 Dropout(GELU(W1 [h_E|h_V[ctr]|h_V[nbr]]+b1))+b2))+b3)) / scale
    (rdesign/model/rdesign.py:85 + rdesign/model/mpnn.py:33-34).  ``token`` orders the edge-embedding backward behind
    every layer; d h_E goes to ``ctx_state['dhE']`` (fp32, accumulated over the layers)."""

    @staticmethod
    def forward(ctx, h_V, token, W1, b1, W2, b2, W3, b3, state):
        g, hE16 = state["graph"], state["hE16"]
        p, seed, sid, scale = state["p_drop"], state["seed"], state["stream_id"], state["scale"]
        Ne, Nv = g.Ne, g.Nv
        hV16 = h_V.detach().to(torch.bfloat16).contiguous()
        z0 = _gemm16([(hE16, None), (hV16, g.ctr), (hV16, g.nbr)], W1, b1, Ne)
        a0 = ops.act_fwd(z0, p, seed, sid)
        z1 = _gemm16([(a0, None)], W2, b2, Ne)
        a1 = ops.act_fwd(z1, p, seed, sid + 1)
        z2 = _gemm16([(a1, None)], W3, b3, Ne)
        a2 = ops.act_fwd(z2, p, seed, sid + 2)
        agg = torch.zeros(Nv, 128, dtype=torch.float32, device=h_V.device)
        ops.scatter_add_rows(a2, g.ctr, agg)
        ctx.save_for_backward(h_V, W1, W2, W3, z0, z1, z2, a0, a1)
        ctx.state = (state, p, seed, sid, scale)
        return agg * (1.0 / scale)

    @staticmethod
    def backward(ctx, dagg):
        h_V, W1, W2, W3, z0, z1, z2, a0, a1 = ctx.saved_tensors
        state, p, seed, sid, scale = ctx.state
        g, hE16 = state["graph"], state["hE16"]
        Ne, dev = g.Ne, h_V.device
        dagg = dagg.detach().float().contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        db1, db2, db3 = (torch.zeros(128, **f32) for _ in range(3))
        dW2, dW3, dW1e = (torch.zeros(128, 128, **f32) for _ in range(3))
        dz2 = ops.act_bwd(z2, p, seed, sid + 2, dagg=dagg, ctr=g.ctr, alpha=1.0 / scale, db=db3)
        ops.gemm_dw(dz2, a1, dW3)
        da1 = _gemm16([(dz2, None)], W3.detach().t().contiguous(), None, Ne)
        dz1 = ops.act_bwd(z1, p, seed, sid + 1, da16=da1, db=db2)
        ops.gemm_dw(dz1, a0, dW2)
        da0 = _gemm16([(dz1, None)], W2.detach().t().contiguous(), None, Ne)
        dz0 = ops.act_bwd(z0, p, seed, sid, da16=da0, db=db1)
        ops.gemm_dw(dz0, hE16, dW1e)
        W1f = W1.detach().float()
        _gemm16([(dz0, None)], W1f[:, :128].t().contiguous(), None, Ne, resid_acc=state["dhE"])     # d h_E += dz0 W1e
        dPc = torch.zeros(g.Nv, 128, **f32)
        dPn = torch.zeros(g.Nv, 128, **f32)
        ops.scatter_add_rows(dz0, g.ctr, dPc)
        ops.scatter_add_rows(dz0, g.nbr, dPn)
        hV = h_V.detach().float()
        dW1 = torch.cat([dW1e, dPc.t() @ hV, dPn.t() @ hV], dim=1)
        dhV = dPc @ W1f[:, 128:256] + dPn @ W1f[:, 256:]
        return dhV, torch.zeros_like(state["token"]), dW1, db1, dW2, db2, dW3, db3, None


class EdgeEmbedFn(torch.autograd.Function):
    """h_E = Normalize(edge_raw W^T + b) (rdesign/model/feature.py:224) -> bf16 rows for the layers; its backward
    consumes the d h_E accumulated by the layers (not a tensor gradient: the bf16 copy is never differentiated)."""

    @staticmethod
    def forward(ctx, W, b, gain, beta, state):
        raw16 = state["edge_raw16"]                                # [Ne,128] bf16, columns >= 115 zero
        Wp = F.pad(W.detach(), (0, 128 - W.shape[1]))
        z = torch.empty(raw16.shape[0], 128, dtype=torch.float32, device=W.device)
        ops.row_mlp([(raw16, None)], [(_pl(Wp, b), False)], ops.EPI_STORE, z, raw16.shape[0], tensor_core=True)
        ctx.save_for_backward(z, W, gain, beta)
        ctx.state = state
        with torch.no_grad():
            state["hE16"] = _normalize(z, gain, beta).to(torch.bfloat16)
        return torch.zeros(1, device=W.device)                     # the ordering token

    @staticmethod
    def backward(ctx, _dtoken):
        z, W, gain, beta = ctx.saved_tensors
        state = ctx.state
        with torch.enable_grad():
            zz = z.detach().requires_grad_(True)
            gg, bb = gain.detach().requires_grad_(True), beta.detach().requires_grad_(True)
            h = _normalize(zz, gg, bb)
            dz, dgain, dbeta = torch.autograd.grad(h, (zz, gg, bb), state["dhE"])
        dW = torch.zeros(128, 128, dtype=torch.float32, device=W.device)
        ops.gemm_dw(dz.to(torch.bfloat16).contiguous(), state["edge_raw16"], dW)
        return dW[:, :W.shape[1]].contiguous(), dz.sum(0), dgain, dbeta, None


def _normalize(x, gain, bias, eps=1e-6):
    """Reference Normalize (rdesign/model/functional.py:79-96): unbiased variance, eps inside and outside the sqrt."""
    mu = x.mean(-1, keepdim=True)
    sigma = torch.sqrt(x.var(-1, keepdim=True) + eps)
    return gain * (x - mu) / (sigma + eps) + bias


class RDesignTrainer:
    """One process per GPU; parameters, gradients and Adam moments are flat fp32 buffers."""

    def __init__(self, model, lr: Optional[float] = None, world: int = 1, p_drop: Optional[float] = None,
                 autocast: bool = True):
        self.model = model
        self.world = world
        self.autocast = autocast
        self.lr = float(model.hparams.lr if lr is None else lr)
        self.p_drop = float(model.hparams.dropout if p_drop is None else p_drop)
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.n_params = n
        n_pad = (n + 3) // 4 * 4
        self.flat_p = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n_pad, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        at = 0
        for p in self.params:                                      # parameters and gradients become views of the flat buffers
            k = p.numel()
            self.flat_p[at:at + k].copy_(p.data.reshape(-1))
            p.data = self.flat_p[at:at + k].view_as(p)
            p.grad = self.flat_g[at:at + k].view_as(p)
            at += k
        self.step_no = 0
        self.comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None

    # ------------------------------------------------------------------ forward + backward
    def loss_and_grads(self, X, S, mask, seed: int = 0):
        m = self.model
        f = m.features
        B, N = X.shape[:2]
        g = ops.build_graph(X, mask, f.top_k)
        node_raw, edge_raw = ops.raw_features(X, g)
        state: Dict = {"graph": g, "p_drop": self.p_drop, "seed": seed, "scale": float(m.mpnn_layers[0].scale)}
        state["edge_raw16"] = F.pad(edge_raw, (0, 128 - edge_raw.shape[1])).to(torch.bfloat16).contiguous()
        state["dhE"] = torch.zeros(g.Ne, 128, dtype=torch.float32, device=X.device)
        del edge_raw
        S_valid = S.reshape(B * N)[g.node_src[:g.Nv].long()]
        self.flat_g.zero_()
        token = EdgeEmbedFn.apply(f.edge_embedding.weight, f.edge_embedding.bias, f.norm_edges.gain, f.norm_edges.bias, state)
        state["token"] = token
        h_V = f.norm_nodes(f.node_embedding(node_raw))             # [Nv,128] fp32, torch (node level)
        for l, layer in enumerate(m.mpnn_layers):
            lin = [x for x in layer.message_layers if isinstance(x, torch.nn.Linear)]
            st = dict(state, stream_id=16 * l)
            st["dhE"], st["token"] = state["dhE"], token
            agg = EdgeMessageFn.apply(h_V, token, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                      lin[2].bias, st)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
                x1 = layer.norm1(h_V + agg)
                h_V = layer.norm2(x1 + layer.dense(x1)).float()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            logits = m.readout(h_V)
        loss = F.cross_entropy(logits.float(), S_valid)
        loss.backward()
        return loss.detach()

    # ------------------------------------------------------------------ all-reduce + Adam
    def apply_gradients(self):
        """ONE all-reduce of the flat gradient buffer (the only data-path collective of training), then fused Adam."""
        import torch.distributed as dist
        self.step_no += 1
        if self.world > 1:
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.flat_g)
            cur.wait_stream(self.comm_stream)
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_no, self.lr,
                      grad_scale=1.0 / self.world)

    def step(self, X, S, mask, seed: Optional[int] = None):
        loss = self.loss_and_grads(X, S, mask, self.step_no if seed is None else seed)
        self.apply_gradients()
        return loss
