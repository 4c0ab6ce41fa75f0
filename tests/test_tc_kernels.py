"""bf16 tensor-core arm (tcgen05 / TMEM / TMA) against plain fp32 PyTorch references of
the same ops, and the whole bf16 forward against the oracle (2e-2 logit budget)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rdesign_oracle as O
from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones

pytestmark = pytest.mark.gpu

LOGIT_TOL_BF16 = 2e-2


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def _edge_reference(hE, W, biases, P, ctr, nbr, alpha, Nv):
    """fp32 reference of rmp_edge_mlp_bf16 with the kernel's rounding points: bf16 h_E / weights /
    inter-stage activations, fp32 accumulation and fp32 gathered projections."""
    S = W.shape[0] // 128
    acts = []
    x = _bf(hE)
    for s in range(S):
        pre = x @ _bf(W[s * 128:(s + 1) * 128]).t()
        pre = pre + (P[ctr.long(), :128] + _bf(P[nbr.long(), 128:]) if s == 0 else biases[s])
        a = F.gelu(pre, approximate="tanh")     # the bf16 arm evaluates the tanh form on MUFU.TANH
        acts.append(a)
        x = _bf(a)
    agg = torch.zeros(Nv, 128).index_add_(0, ctr.long(), x) * alpha
    return acts, agg


def _segments(Nv, degs):
    ctr = torch.repeat_interleave(torch.arange(Nv), torch.as_tensor(degs))
    return ctr.to(torch.int32)


def _row_ptr(ctr, Nv):
    counts = torch.bincount(ctr.long(), minlength=Nv)
    return torch.cat([torch.zeros(1, dtype=torch.long), torch.cumsum(counts, 0)]).to(torch.int32)


@pytest.mark.parametrize("Nv,deg,S", [(4, 32, 3), (129, 30, 3), (1600, 25, 3), (300, 2, 3), (500, 1, 2), (37, 48, 2),
                                      (3, 128, 3), (16384, 30, 3), (2000, 7, 3), (700, 9, 2)])
def test_edge_mlp_tc_matches_fp32_reference(Nv, deg, S):
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(Nv * 7 + deg)
    degs = [deg] * Nv
    if Nv > 10:
        degs[3] = max(1, deg // 2)                     # ragged: one short segment
    ctr = _segments(Nv, degs)
    Ne = ctr.numel()
    nbr = torch.randint(0, Nv, (Ne,), generator=g, dtype=torch.int32)
    hE = torch.randn(Ne, 128, generator=g)
    W = torch.randn(S * 128, 128, generator=g) / 128 ** 0.5
    biases = [None] + [0.1 * torch.randn(128, generator=g) for _ in range(S - 1)]
    P = torch.randn(Nv, 256, generator=g) * 0.5
    alpha = 1.0 / 30
    acts_ref, agg_ref = _edge_reference(hE, W, biases, P, ctr, nbr, alpha, Nv)

    dbg = torch.zeros(S, Ne, 128, device="cuda")
    agg = torch.zeros(Nv, 128, device="cuda")
    ops.edge_mlp_bf16(hE.cuda().to(torch.bfloat16), W.cuda().to(torch.bfloat16),
                      [None] + [b.cuda() for b in biases[1:]], P[:, :128].contiguous().cuda(),
                      P[:, 128:].contiguous().cuda().to(torch.bfloat16), ctr.cuda(), nbr.cuda(),
                      _row_ptr(ctr, Nv).cuda(), agg, alpha, debug=dbg)
    torch.cuda.synchronize()
    for s in range(S):
        tol = 2e-3 if s == 0 else 6e-3      # tanh.approx.f32 is 2^-11 accurate; later stages see bf16 re-rounding
        np.testing.assert_allclose(dbg[s].cpu().numpy(), acts_ref[s].numpy(), atol=tol, rtol=4e-3, err_msg=f"stage {s}")
    np.testing.assert_allclose(agg.cpu().numpy(), agg_ref.numpy(), atol=6e-3 * max(1.0, deg / 30), rtol=0)
    # production path (no dump): packed bf16x2 GELU written straight into the next operand
    agg2 = torch.zeros(Nv, 128, device="cuda")
    ops.edge_mlp_bf16(hE.cuda().to(torch.bfloat16), W.cuda().to(torch.bfloat16),
                      [None] + [b.cuda() for b in biases[1:]], P[:, :128].contiguous().cuda(),
                      P[:, 128:].contiguous().cuda().to(torch.bfloat16), ctr.cuda(), nbr.cuda(),
                      _row_ptr(ctr, Nv).cuda(), agg2, alpha)
    err = (agg2.cpu() - agg_ref).abs()
    assert err.max().item() < 2.5e-2 * max(1.0, deg / 30) and err.mean().item() < 2e-3 * max(1.0, deg / 30), \
        (err.max().item(), err.mean().item())


def test_edge_mlp_tc_accumulates_into_agg_and_is_repeatable():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(1)
    Nv, deg = 700, 30
    ctr = _segments(Nv, [deg] * Nv).cuda()
    rp = _row_ptr(ctr.cpu(), Nv).cuda()
    Ne = ctr.numel()
    nbr = torch.randint(0, Nv, (Ne,), generator=g, dtype=torch.int32).cuda()
    hE = torch.randn(Ne, 128, generator=g).cuda().to(torch.bfloat16)
    W = (torch.randn(384, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
    b = [None, torch.zeros(128).cuda(), torch.zeros(128).cuda()]
    P = torch.randn(Nv, 256, generator=g).cuda()
    Pc, Pn = P[:, :128].contiguous(), P[:, 128:].contiguous().to(torch.bfloat16)
    outs = []
    for _ in range(3):
        agg = torch.zeros(Nv, 128, device="cuda")
        ops.edge_mlp_bf16(hE, W, b, Pc, Pn, ctr, nbr, rp, agg, 1.0)
        outs.append(agg.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])   # <= 2 partial sums per node
    agg2 = outs[0].clone()
    ops.edge_mlp_bf16(hE, W, b, Pc, Pn, ctr, nbr, rp, agg2, 1.0)
    np.testing.assert_allclose(agg2.cpu().numpy(), 2 * outs[0].cpu().numpy(), rtol=1e-5, atol=1e-5)


def _ln(x, g, b):
    return F.layer_norm(x, (128,), g, b, 1e-5)


@pytest.mark.parametrize("n_rows,widths,with_proj", [(128, [256, 256, 256], True), (300, [256, 256, 256], True),
                                                     (77, [256], False), (20000, [256, 256], True),
                                                     (130, [128, 128], False)])
def test_node_mlp_tc_matches_fp32_reference(n_rows, widths, with_proj):
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(n_rows)
    h = torch.randn(n_rows, 128, generator=g)
    agg = 0.3 * torch.randn(n_rows, 128, generator=g)
    ln1 = (1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g))
    ln2 = (1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g))
    dims = [128] + widths + [128]
    lin = [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5, 0.1 * torch.randn(dims[i + 1], generator=g))
           for i in range(len(dims) - 1)]
    proj = (torch.randn(256, 128, generator=g) / 128 ** 0.5, 0.1 * torch.randn(256, generator=g))
    # fp32 reference with the kernel's rounding points
    x1 = _ln(h + agg, *ln1)
    x = _bf(x1)
    for W, b in lin[:-1]:
        x = _bf(F.gelu(x @ _bf(W).t() + b, approximate="tanh"))
    h_ref = _ln(x1 + x @ _bf(lin[-1][0]).t() + lin[-1][1], *ln2)
    p_ref = _bf(h_ref) @ _bf(proj[0]).t() + proj[1]

    dev = "cuda"
    stages = [(W.to(dev).to(torch.bfloat16), b.to(dev), 0) for W, b in lin[:-1]]
    stages.append((lin[-1][0].to(dev).to(torch.bfloat16), lin[-1][1].to(dev), 1))
    if with_proj:
        stages.append((proj[0].to(dev).to(torch.bfloat16), proj[1].to(dev), 2))
    agg_d = agg.to(dev)
    h_out = torch.empty(n_rows, 128, device=dev)
    x1_d = torch.empty(n_rows, 128, device=dev)
    p_out = torch.empty(n_rows, 128, device=dev) if with_proj else None
    pn_out = torch.empty(n_rows, 128, device=dev, dtype=torch.bfloat16) if with_proj else None
    ops.node_mlp_bf16(h.to(dev), stages, n_rows, agg=agg_d, ln1=tuple(t.to(dev) for t in ln1),
                      ln2=tuple(t.to(dev) for t in ln2), x1=x1_d, h_out=h_out, p_out=p_out, pn_out=pn_out)
    torch.cuda.synchronize()
    np.testing.assert_allclose(x1_d.cpu().numpy(), x1.numpy(), atol=2e-5, rtol=0)
    np.testing.assert_allclose(h_out.cpu().numpy(), h_ref.numpy(), atol=2e-2, rtol=0)
    assert (h_out.cpu() - h_ref).abs().mean().item() < 2e-3
    assert float(agg_d.abs().max()) == 0.0                     # neighbour-sum buffer re-armed for the next layer
    if with_proj:
        got = torch.cat([p_out.cpu(), pn_out.float().cpu()], 1)
        np.testing.assert_allclose(got.numpy(), p_ref.numpy(), atol=4e-2, rtol=0)


def test_node_mlp_tc_projection_only():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(5)
    h = torch.randn(1000, 128, generator=g)
    W, b = torch.randn(256, 128, generator=g) / 11.3, 0.1 * torch.randn(256, generator=g)
    Pc = torch.empty(1000, 128, device="cuda")
    Pn = torch.empty(1000, 128, device="cuda", dtype=torch.bfloat16)
    ops.node_mlp_bf16(h.cuda(), [(W.cuda().to(torch.bfloat16), b.cuda(), 2)], 1000, p_out=Pc, pn_out=Pn)
    ref = _bf(h) @ _bf(W).t() + b
    np.testing.assert_allclose(Pc.cpu().numpy(), ref[:, :128].numpy(), atol=1e-4, rtol=0)
    np.testing.assert_allclose(Pn.float().cpu().numpy(), _bf(ref[:, 128:]).numpy(), atol=2e-2, rtol=0)


@pytest.mark.parametrize("B,L,K,lengths", [(2, 64, 25, [64, 50]), (3, 100, 48, [100, 61, 7]), (1, 300, 30, None)])
def test_edge_embed_tc_matches_oracle(B, L, K, lengths):
    from rmpnn_b200._lib import PREC_BF16
    from rmpnn_b200.engine import RDesignEngine
    cfg = RDesignConfig(k_neighbors=K, num_mpnn_layers=1)
    params = make_state_dict(cfg, 42)
    X, S, mask = synthetic_backbones(B, L, 77, lengths=lengths)
    eng = RDesignEngine(params, K, cfg.scale, device="cuda", precision=PREC_BF16)
    g = eng.graph(X.cuda(), mask.cuda())
    h_V, h_E = eng.embed(X.cuda(), g)
    assert h_E.dtype == torch.bfloat16
    ref = O.features(params, X, S, mask, K)
    assert torch.equal(g.E_flat.cpu(), ref[4])
    np.testing.assert_allclose(h_V.cpu().numpy(), ref[2].numpy(), atol=3e-4, rtol=0)   # fp32 node path (acos near +-1)
    err = (h_E.float().cpu() - ref[3]).abs()
    # bf16 raw features x bf16 weights, bf16 output: a few bf16 ulps of an O(1..4) normalised value
    assert err.max().item() < 8e-2 and err.mean().item() < 6e-3, (err.max().item(), err.mean().item())


@pytest.mark.parametrize("B,L,K,lengths", [(2, 64, 25, [64, 50]), (4, 256, 30, None), (3, 100, 48, [100, 61, 7])])
def test_forward_bf16_within_budget(B, L, K, lengths):
    from rmpnn_b200._lib import PREC_BF16
    from rmpnn_b200.engine import RDesignEngine
    cfg = RDesignConfig(k_neighbors=K)
    params = make_state_dict(cfg, 42)
    X, S, mask = synthetic_backbones(B, L, 1002, lengths=lengths)
    ref = O.forward(params, X, S, mask, K, cfg.scale)
    eng = RDesignEngine(params, K, cfg.scale, device="cuda", precision=PREC_BF16)
    out = eng.forward(X.cuda(), mask.cuda())
    assert torch.equal(out["graph"].E_flat.cpu(), ref["E_idx"])
    err = (out["logits"].cpu() - ref["logits"]).abs().max().item()
    print(f"bf16 forward max |dlogit| = {err:.4f}")
    assert err < LOGIT_TOL_BF16, err
    agree = (out["samples"].cpu() == ref["samples"]).float().mean().item()
    assert agree > 0.97, agree          # argmax may flip only where two logits are within the bf16 budget
