"""Training step of the rdesign path on the sm_100a kernels (BASELINE configs[3]).

``training_step`` of the reference (rdesign/model/rdesign.py:95-104) = forward with dropout, cross entropy, backward,
gradient all-reduce (Lightning DDP), Adam (rdesign.py:90-93).  Here:

* everything per EDGE -- the three Linear + GELU + Dropout message stages on ``[h_E | h_V[ctr] | h_V[nbr]]``, the
  ``scatter_sum / scale`` and their backward (96 % of the step's flops and bytes) -- runs on the C-ABI kernels:
  the lean tcgen05 row GEMM ``rmp_gemm_rows_bf16`` (GELU + Dropout fused into its forward epilogue), the
  weight-gradient GEMM ``rmp_gemm_dw_bf16``, and the HBM-bound activation-backward / scatter passes of ``csrc/train.cu``; dropout masks are regenerated from Philox, never stored;
  ``d h_E`` = sum over the layers of ``dz0_l W1e_l`` is ONE K-concatenated GEMM (``rmp_gemm_rows_kcat_bf16``) over the saved ``dz0_l``;
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


def _mm16(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """Small node-level products ([Nv,128] operands) on the bf16 tensor cores, fp32 result."""
    return (a.to(torch.bfloat16) @ b.to(torch.bfloat16)).float()


class EdgeMessageFn(torch.autograd.Function):
    """agg[v] = sum_{e: ctr[e]=v} Dropout(GELU(W3 Dropout(GELU(W2 Dropout(GELU(W1 [h_E|h_V[ctr]|h_V[nbr]]+b1))+b2))+b3)) / scale
    (rdesign/model/rdesign.py:85 + rdesign/model/mpnn.py:33-34).  W1 = [W1e | W1c | W1n]: the h_V part is applied per
    node (Pc = h_V W1c^T + b1, Pn = h_V W1n^T) and gathered in the GEMM epilogue.  ``token`` orders the edge-embedding
    backward behind every layer; the layer's ``dz0`` goes to ``state['dz0_all'][layer]`` (d h_E is formed from them, once).
    ``state['w16'](param)`` is the parameter's view in the step's bf16 copy of the flat parameter buffer."""

    @staticmethod
    def forward(ctx, h_V, token, W1, b1, W2, b2, W3, b3, state):
        g, hE16, w16 = state["graph"], state["hE16"], state["w16"]
        p, seed, sid, scale, sd = state["p_drop"], state["seed"], state["stream_id"], state["scale"], state["seed_dev"]
        hV16 = h_V.detach().to(torch.bfloat16)
        W1h = w16(W1)                                              # [128,384] bf16 view, columns e | c | n
        Pc = torch.addmm(w16(b1), hV16, W1h[:, 128:256].t())      # [Nv,128] bf16 (bias from the step's bf16 parameter copy)
        Pn = hV16 @ W1h[:, 256:].t()
        ctx.fused = state.get("wmsg_all") is not None
        if ctx.fused:
            # ONE pass over h_E: the inference edge kernel's three chained tcgen05 stages with Dropout in the epilogue, the
            # neighbour sums by the selector MMA, and z0..z2 / a0, a1 written out for the backward (a in operand form
            # 2 GELU(z) mask: the weight gradients below are rescaled by keep_scale / 2)
            agg = torch.zeros(g.Nv, 128, dtype=torch.float32, device=h_V.device)
            (z0, z1, z2), (a0, a1) = ops.edge_mlp_train_fwd(
                hE16, state["wmsg_all"][state["layer"]], [None, b2.detach().float(), b3.detach().float()], Pc.float(), Pn,
                g.ctr, g.nbr, g.row_ptr, agg, 1.0 / scale, p, seed, sid, seed_dev=sd)
            ctx.save_for_backward(hV16, W1, W2, W3, z0, z1, z2, a0, a1)
            ctx.state = (state, p, seed, sid, scale)
            return agg
        z0, a0 = ops.gemm_rows(hE16, W1h[:, :128], Pc=Pc, Pn=Pn, ctr=g.ctr, nbr=g.nbr, act=(p, seed, sid), seed_dev=sd)
        z1, a1 = ops.gemm_rows(a0, w16(W2), bias=b2.detach(), act=(p, seed, sid + 1), seed_dev=sd)
        z2, a2 = ops.gemm_rows(a1, w16(W3), bias=b3.detach(), act=(p, seed, sid + 2), seed_dev=sd)
        agg = ops.segment_sum_rows(a2, g.row_ptr[:g.Nv + 1])        # edges of a node are contiguous: no atomics
        ctx.save_for_backward(hV16, W1, W2, W3, z0, z1, z2, a0, a1)
        ctx.state = (state, p, seed, sid, scale)
        return agg * (1.0 / scale)

    @staticmethod
    def backward(ctx, dagg):
        hV16, W1, W2, W3, z0, z1, z2, a0, a1 = ctx.saved_tensors
        state, p, seed, sid, scale = ctx.state
        g, hE16, w16, sd = state["graph"], state["hE16"], state["w16"], state["seed_dev"]
        dev = hV16.device
        dagg = dagg.detach().float().contiguous()
        db = torch.zeros(3, 128, dtype=torch.float32, device=dev)
        dW = torch.zeros(3, 128, 128, dtype=torch.float32, device=dev)           # dW3, dW2, dW1e
        dP = torch.empty(2, g.Nv, 128, dtype=torch.float32, device=dev)           # d Pc, d Pn (both overwritten)
        W1h = w16(W1)
        dz2 = ops.act_bwd(z2, p, seed, sid + 2, dagg=dagg, ctr=g.ctr, alpha=1.0 / scale, db=db[2], seed_dev=sd)
        ops.gemm_dw(dz2, a1, dW[0])
        da1, _ = ops.gemm_rows(dz2, w16(W3), w_transposed=True)                  # da1 = dz2 W3
        dz1 = ops.act_bwd(z1, p, seed, sid + 1, da16=da1, db=db[1], seed_dev=sd)
        ops.gemm_dw(dz1, a0, dW[1])
        da0, _ = ops.gemm_rows(dz1, w16(W2), w_transposed=True)
        # dz0 of every layer is kept: d h_E = sum_l dz0_l W1e_l is ONE K-concatenated GEMM in the edge embedding's backward
        dz0 = ops.act_bwd(z0, p, seed, sid, da16=da0, db=db[0], seed_dev=sd, out=state["dz0_all"][state["layer"]])
        ops.gemm_dw(dz0, hE16, dW[2])
        if ctx.fused:
            dW[:2] *= 0.5 * ops.keep_scale(p)                      # a0, a1 were saved as 2 GELU(z) mask
        ops.segment_sum_rows(dz0, g.row_ptr[:g.Nv + 1], dP[0])
        # d Pn: rows of dz0 summed by NEIGHBOUR node through the edges sorted by neighbour (built once per step)
        ops.segment_sum_rows_perm(dz0, state["nbr_perm"], state["nbr_row_ptr"], dP[1])
        dP16 = dP.to(torch.bfloat16)
        dW1 = torch.cat([dW[2], torch.mm(dP16[0].t(), hV16, out_dtype=torch.float32),
                         torch.mm(dP16[1].t(), hV16, out_dtype=torch.float32)], dim=1)
        dhV = (torch.mm(dP16[0], W1h[:, 128:256], out_dtype=torch.float32) +
               torch.mm(dP16[1], W1h[:, 256:], out_dtype=torch.float32))
        return dhV, torch.zeros(1, device=dev), dW1, db[0], dW[1], db[1], dW[0], db[2], None


class EdgeEmbedFn(torch.autograd.Function):
    """h_E = Normalize(edge_raw W^T + b) (rdesign/model/feature.py:224) -> bf16 rows for the layers; its backward
    forms d h_E from the layers' saved ``dz0`` (not a tensor gradient: the bf16 copy is never differentiated)."""

    @staticmethod
    def forward(ctx, W, b, gain, beta, state):
        raw16 = state["edge_raw16"]                                # [Ne,128] bf16, columns >= 115 zero
        Wp = F.pad(state["w16"](W), (0, 128 - W.shape[1]))
        z = torch.zeros(raw16.shape[0], 128, dtype=torch.float32, device=W.device)       # fp32 result through the += output
        ops.gemm_rows(raw16, Wp, bias=b.detach(), want_z=False, acc32=z)
        ctx.save_for_backward(z, W, gain, beta)
        ctx.state = state
        state["hE16"] = ops.normalize_fwd(z, gain.detach(), beta.detach())
        return torch.zeros(1, device=W.device)                     # the ordering token

    @staticmethod
    def backward(ctx, _dtoken):
        z, W, gain, beta = ctx.saved_tensors
        state = ctx.state
        dz0_all, w1e_all = state["dz0_all"], state["w1e_all"]
        if dz0_all.shape[0] >= 2:                                  # d h_E = sum_l dz0_l W1e_l: one pass, K = 128 L
            dhE = ops.gemm_rows_kcat(dz0_all, w1e_all, w_transposed=True)
        else:
            dhE = torch.zeros(dz0_all.shape[1], 128, dtype=torch.float32, device=W.device)
            ops.gemm_rows(dz0_all[0], w1e_all[0], want_z=False, acc32=dhE, w_transposed=True)
        dz16, dgain, dbeta, db = ops.normalize_bwd(z, dhE, gain.detach())
        dW = torch.zeros(128, 128, dtype=torch.float32, device=W.device)
        ops.gemm_dw(dz16, state["edge_raw16"], dW)
        return dW[:, :W.shape[1]].contiguous(), db, dgain, dbeta, None


class LayerNormFn(torch.autograd.Function):
    """``nn.LayerNorm`` over the last dimension (rdesign/model/mpnn.py:11-12, :35, :37) with the weight / bias gradients as one
    column-sum pass (``rmp_colsum``): torch's ``GammaBetaBackward`` kernel takes 75 us per call on [16384,128] (1.35 ms of the
    C4 step), two ``[1,Nv] x [Nv,128]`` split-K products plus the ``xhat`` elementwise kernels 40 us, the pass 4 us."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        x = x.float().contiguous()
        y, mean, rstd = torch.native_layer_norm(x, (x.shape[-1],), weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight = ctx.saved_tensors
        dy = dy.float().contiguous()
        with torch.autocast("cuda", enabled=False):
            dx = torch.ops.aten.native_layer_norm_backward(dy, x, [x.shape[-1]], mean, rstd, weight, None,
                                                           [True, False, False])[0]
            db, dw = ops.colsum(dy, x, mean, rstd)                 # one pass: sum dy, sum dy * xhat
        return dx, dw, db, None


class AddLayerNormFn(torch.autograd.Function):
    """``norm(a + b)`` (rdesign/model/mpnn.py:35, :37) as one forward and one backward kernel (``rmp_add_layernorm_fwd`` /
    ``rmp_layernorm_bwd``): the residual add, the statistics, ``dx`` and the weight / bias gradients; ``b`` fp32 or bf16."""

    @staticmethod
    def forward(ctx, a, b, weight, bias, eps):
        x, y, mean, rstd = ops.add_layernorm_fwd(a.detach().float().contiguous(), b.detach().contiguous(), weight.detach(),
                                                 bias.detach(), eps)
        ctx.save_for_backward(x, mean, rstd, weight)
        ctx.b_bf16 = b.dtype == torch.bfloat16
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd, weight = ctx.saved_tensors
        dx, dx16, dw, db = ops.layernorm_bwd(dy.detach().float().contiguous(), x, mean, rstd, weight.detach(), want_bf16=ctx.b_bf16)
        return dx, (dx16 if ctx.b_bf16 else dx), dw, db, None


class NodeActFn(torch.autograd.Function):
    """``Dropout(GELU(z))`` on the per-node FFN rows (rdesign/model/mpnn.py:20-28) with the edge path's kernels: tanh-form GELU
    and its exact derivative, the keep mask regenerated from (seed, stream id, element) in the backward -- one kernel per
    direction instead of GELU + Dropout (+ their two backward kernels).  ``z`` bf16 [rows, 128 k]."""

    @staticmethod
    def forward(ctx, z, p, seed, stream_id, seed_dev):
        z = z.detach().contiguous()
        ctx.save_for_backward(z)
        ctx.args = (p, seed, stream_id, seed_dev)
        return ops.act_fwd(z, p, seed, stream_id, seed_dev=seed_dev)

    @staticmethod
    def backward(ctx, da):
        (z,) = ctx.saved_tensors
        p, seed, stream_id, seed_dev = ctx.args
        da = da.detach().to(torch.bfloat16).contiguous()
        # the kernel indexes groups of 8 elements globally, so a [rows, 256] matrix is its own [2 rows, 128] view
        dz = ops.act_bwd(z.view(-1, 128), p, seed, stream_id, da16=da.view(-1, 128), seed_dev=seed_dev)
        return dz.view(z.shape), None, None, None, None


class LinearFn(torch.autograd.Function):
    """``nn.Linear`` of the per-node FFN on the bf16 tensor cores (what bf16 autocast computes), the bias gradient as one
    ``rmp_colsum`` pass (torch's column reduction kernel: 25 us per call, a split-K ``[1,Nv] x [Nv,N]`` product: 14 us)."""

    @staticmethod
    def forward(ctx, x, weight, bias, w16, b16=None):
        x16 = x.to(torch.bfloat16)
        ctx.x_fp32 = x.dtype == torch.float32
        ctx.save_for_backward(x16, w16)
        return torch.addmm(bias.to(torch.bfloat16) if b16 is None else b16, x16, w16.t())

    @staticmethod
    def backward(ctx, dy):
        x16, w16 = ctx.saved_tensors
        dy = dy.to(torch.bfloat16).contiguous()
        with torch.autocast("cuda", enabled=False):
            dx = torch.mm(dy, w16, out_dtype=torch.float32) if ctx.x_fp32 else dy @ w16   # the gradient in the input's dtype
            dw = torch.mm(dy.t(), x16, out_dtype=torch.float32)
            db = ops.colsum(dy)[0]
        return dx, dw, db, None, None


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
        # forward of the per-edge message MLP: one fused launch per layer (rmp_edge_mlp_train_fwd_bf16) or, with
        # RMP_TRAIN_FUSED_FWD=0, one row GEMM per stage + segmented sum
        import os
        self.fused_forward = os.environ.get("RMP_TRAIN_FUSED_FWD", "1") != "0"
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
        self._slot = {}
        at = 0
        for p in self.params:
            self._slot[id(p)] = (at, tuple(p.shape))
            at += p.numel()
        self.flat_p16 = torch.zeros(n_pad, dtype=torch.bfloat16, device=dev)     # the step's bf16 copy of every parameter
        self.step_no = 0
        # device-side step number and dropout seed offset: a captured CUDA graph advances them itself
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.comm_stream = torch.cuda.Stream(device=dev) if world > 1 else None
        self._graph = None

    def _w16(self, param: torch.Tensor) -> torch.Tensor:
        at, shape = self._slot[id(param)]
        n = 1
        for d in shape:
            n *= d
        return self.flat_p16[at:at + n].view(shape)

    # ------------------------------------------------------------------ forward + backward
    def loss_and_grads(self, X, S, mask, seed: int = 0, sizes=None):
        """Forward + backward into the flat gradient buffer; returns the detached loss."""
        loss = self.forward_loss(X, S, mask, seed, sizes)
        # torch.autograd.grad instead of backward(): no AccumulateGrad nodes, whose stream (the one the parameters were
        # created on) would have to be synchronised with the capturing stream and invalidates a CUDA-graph capture
        grads = torch.autograd.grad(loss, self.params)
        torch.cat([g.reshape(-1) for g in grads], out=self.flat_g[:self.n_params])
        return loss.detach()

    def forward_loss(self, X, S, mask, seed: int = 0, sizes=None):
        """Cross-entropy loss of one batch with its autograd graph (the per-edge and per-node blocks are autograd Functions
        over the C-ABI kernels): ``loss.backward()`` fills ``p.grad`` like any torch loss, which is how the drop-in
        ``RNAModel.training_step`` hands the native path to Lightning's optimizer loop."""
        m = self.model
        f = m.features
        B, N = X.shape[:2]
        self.flat_p16.copy_(self.flat_p)                           # one cast for all weights of the step
        g = ops.build_graph(X, mask, f.top_k, sizes)
        node_raw, edge_raw = ops.raw_features(X, g)
        state: Dict = {"graph": g, "p_drop": self.p_drop, "seed": seed, "scale": float(m.mpnn_layers[0].scale),
                       "w16": self._w16, "seed_dev": self.seed_dev}
        state["edge_raw16"] = F.pad(edge_raw, (0, 128 - edge_raw.shape[1])).to(torch.bfloat16).contiguous()
        state["nbr_perm"], state["nbr_row_ptr"] = ops.transposed_csr(g.nbr[:g.Ne], g.Nv)
        n_layers = len(m.mpnn_layers)
        state["dz0_all"] = torch.empty(n_layers, g.Ne, 128, dtype=torch.bfloat16, device=X.device)
        lins = [[x for x in layer.message_layers if isinstance(x, torch.nn.Linear)] for layer in m.mpnn_layers]
        state["w1e_all"] = torch.stack([self._w16(ll[0].weight)[:, :128] for ll in lins]).contiguous()
        state["wmsg_all"] = None
        if self.fused_forward and all(len(ll) == 3 and tuple(ll[0].weight.shape) == (128, 384) and tuple(ll[1].weight.shape) ==
                                      tuple(ll[2].weight.shape) == (128, 128) for ll in lins):
            # per layer [W1e ; W2 ; W3] as the edge kernel's resident [3 x 128, 128] bf16 operand
            state["wmsg_all"] = torch.stack([torch.cat([self._w16(ll[0].weight)[:, :128], self._w16(ll[1].weight),
                                                        self._w16(ll[2].weight)]) for ll in lins]).contiguous()
        del edge_raw
        S_valid = S.reshape(B * N)[g.node_src[:g.Nv].long()]
        token = EdgeEmbedFn.apply(f.edge_embedding.weight, f.edge_embedding.bias, f.norm_edges.gain, f.norm_edges.bias, state)
        # (the token is passed to the layers as an argument and never stored in ``state``: state is referenced by the
        # autograd contexts, and a reference cycle would keep this step's graph -- and its AccumulateGrad nodes with
        # the stream they were created on -- alive into the next step, which breaks CUDA-graph capture)
        h_V = f.norm_nodes(f.node_embedding(node_raw))             # [Nv,128] fp32, torch (node level)
        for l, layer in enumerate(m.mpnn_layers):
            lin = [x for x in layer.message_layers if isinstance(x, torch.nn.Linear)]
            st = dict(state, stream_id=16 * l, layer=l)
            agg = EdgeMessageFn.apply(h_V, token, lin[0].weight, lin[0].bias, lin[1].weight, lin[1].bias, lin[2].weight,
                                      lin[2].bias, st)
            if self.autocast:
                # per-node remainder: LayerNorm and Linear through the two functions above (same math as the modules under
                # bf16 autocast), GELU / Dropout / residual adds as torch ops
                x1 = AddLayerNormFn.apply(h_V, agg, layer.norm1.weight, layer.norm1.bias, layer.norm1.eps)
                d = x1
                mods = list(layer.dense)
                k = 0
                while k < len(mods):
                    mod = mods[k]
                    if isinstance(mod, torch.nn.Linear):
                        d = LinearFn.apply(d, mod.weight, mod.bias, self._w16(mod.weight), self._w16(mod.bias))
                    elif (isinstance(mod, torch.nn.GELU) and k + 1 < len(mods) and isinstance(mods[k + 1], torch.nn.Dropout)
                          and d.dtype == torch.bfloat16 and d.shape[1] % 128 == 0):
                        # GELU + Dropout in one kernel per direction (stream ids 8.. of the layer: the edge stages use 0..2)
                        d = NodeActFn.apply(d, self.p_drop, seed, 16 * l + 8 + k, self.seed_dev)
                        k += 1
                    else:
                        d = mod(d)
                    k += 1
                h_V = AddLayerNormFn.apply(x1, d, layer.norm2.weight, layer.norm2.bias, layer.norm2.eps)
            else:
                x1 = layer.norm1(h_V + agg)
                h_V = layer.norm2(x1 + layer.dense(x1)).float()
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast):
            logits = m.readout(h_V)
        return F.cross_entropy(logits.float(), S_valid)

    # ------------------------------------------------------------------ all-reduce + Adam
    def apply_gradients(self):
        """ONE all-reduce of the flat gradient buffer (the only data-path collective of training), then fused Adam."""
        import torch.distributed as dist
        self.step_no += 1
        self.step_dev += 1
        self.seed_dev += 0x5851F42D4C957F2D
        if self.world > 1:
            cur = torch.cuda.current_stream()
            self.comm_stream.wait_stream(cur)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.flat_g)
            cur.wait_stream(self.comm_stream)
        ops.adam_step(self.flat_p, self.flat_g, self.exp_avg, self.exp_avg_sq, self.step_no, self.lr,
                      grad_scale=1.0 / self.world, step_dev=self.step_dev)

    def step(self, X, S, mask, seed: int = 0, sizes=None):
        """One eager training step (the dropout seed is ``seed`` plus the device-side offset that advances every step)."""
        loss = self.loss_and_grads(X, S, mask, seed, sizes)
        self.apply_gradients()
        return loss

    def capture(self, X, S, mask, sizes):
        """Capture forward + backward of one training step for fixed shapes into a CUDA graph (about 1,100 launches);
        the gradient all-reduce (one NCCL call) and the fused Adam kernel follow the replay eagerly, so the collective
        is an ordinary NCCL launch on the side stream rather than a captured node.
        Returns ``(replay, static)``: overwrite ``static['X'|'S'|'mask']`` and call ``replay()`` -> loss tensor.
        ``sizes = (Nv, Ne)`` must be given (a captured step cannot read the two counters back to the host)."""
        static = {"X": X.clone(), "S": S.clone(), "mask": mask.clone()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.autograd.set_multithreading_enabled(False):   # warm-up outside the capture
            for _ in range(3):
                self.step(static["X"], static["S"], static["mask"], sizes=sizes)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        # the backward must run on the capturing thread / stream: the ctypes launches take torch's current stream, and a
        # launch from an autograd worker thread that sees the legacy stream invalidates the capture
        with torch.autograd.set_multithreading_enabled(False), torch.cuda.graph(graph, capture_error_mode="thread_local"):
            loss = self.loss_and_grads(static["X"], static["S"], static["mask"], 0, sizes)
        self._graph = graph

        def replay():
            graph.replay()
            self.apply_gradients()
            return loss
        return replay, static
