"""Training-path kernels (csrc/train.cu) and the native training step (rmpnn_b200/train_engine.py) against PyTorch:
BASELINE configs[3] / SURVEY.md section 8f rank 1.  Reference: training_step rdesign/model/rdesign.py:95-104,
optimizer :90-93, message MLP with Dropout after every GELU rdesign/model/mpnn.py:15-19."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("rows", [1, 127, 128, 1000, 40001])
def test_gemm_dw_matches_torch(rows):
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows)
    dz = torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16)
    x = torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16)
    dW = torch.full((128, 128), 0.5, device="cuda")
    ops.gemm_dw(dz, x, dW)
    ref = 0.5 + dz.float().t() @ x.float()
    err = (dW - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, rows ** 0.5), err          # fp32 accumulation of bf16 products, order differs


def test_act_fwd_bwd_no_dropout_match_torch_gelu():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(0)
    z = (2 * torch.randn(3000, 128, generator=g)).cuda().to(torch.bfloat16)
    da = torch.randn(3000, 128, generator=g).cuda().to(torch.bfloat16)
    a = ops.act_fwd(z, 0.0, 7, 3)
    zf = z.float().requires_grad_(True)
    ref = F.gelu(zf)
    np.testing.assert_allclose(a.float().cpu().numpy(), _bf(ref.detach()).cpu().numpy(), atol=1e-2, rtol=1e-2)
    db = torch.zeros(128, device="cuda")
    dz = ops.act_bwd(z, 0.0, 7, 3, da16=da, db=db)
    (ref * da.float()).sum().backward()
    np.testing.assert_allclose(dz.float().cpu().numpy(), _bf(zf.grad).cpu().numpy(), atol=2e-2, rtol=2e-2)
    np.testing.assert_allclose(db.cpu().numpy(), dz.float().sum(0).cpu().numpy(), atol=2e-2, rtol=1e-3)


def test_dropout_mask_is_regenerated_identically_and_has_the_right_rate():
    from rmpnn_b200 import ops
    z = torch.ones(20000, 128, device="cuda", dtype=torch.bfloat16)
    a = ops.act_fwd(z, 0.1, 11, 5)
    keep = a != 0
    p_eff = 26 / 256                                                   # 0.1 quantised to 8 bits
    assert abs(keep.float().mean().item() - (1 - p_eff)) < 2e-3
    kept_val = a[keep].float().unique()
    assert kept_val.numel() == 1 and abs(kept_val.item() - F.gelu(torch.tensor(1.0)).item() / (1 - p_eff)) < 1e-2
    dz = ops.act_bwd(z, 0.1, 11, 5, da16=torch.ones_like(z))
    assert torch.equal(dz != 0, keep)                                   # same mask in both directions
    assert not torch.equal(ops.act_fwd(z, 0.1, 11, 6) != 0, keep)       # another stream id, another mask
    assert not torch.equal(ops.act_fwd(z, 0.1, 12, 5) != 0, keep)       # another seed, another mask


def test_act_bwd_gathers_the_scatter_sum_gradient():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(2)
    Nv, deg = 300, 7
    ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
    z = torch.randn(Nv * deg, 128, generator=g).cuda().to(torch.bfloat16)
    dagg = torch.randn(Nv, 128, generator=g).cuda()
    dz = ops.act_bwd(z, 0.0, 1, 0, dagg=dagg, ctr=ctr, alpha=1 / 30)
    zf = z.float().requires_grad_(True)
    agg = torch.zeros(Nv, 128, device="cuda").index_add_(0, ctr.long(), F.gelu(zf)) / 30
    (agg * dagg).sum().backward()
    np.testing.assert_allclose(dz.float().cpu().numpy(), _bf(zf.grad).cpu().numpy(), atol=2e-3, rtol=2e-2)


def test_segment_sum_rows_perm_is_the_scatter_by_neighbour():
    """Transposed-CSR gather (edges sorted by neighbour) = index_add by neighbour, without atomics."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(8)
    rows, Nv = 9000, 211
    x = torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16)
    idx = torch.randint(0, Nv - 3, (rows,), generator=g, dtype=torch.int32).cuda()      # the last nodes receive nothing
    perm, row_ptr = ops.transposed_csr(idx, Nv)
    assert row_ptr.shape[0] == Nv + 1 and int(row_ptr[-1]) == rows
    out = ops.segment_sum_rows_perm(x, perm, row_ptr)
    ref = torch.zeros(Nv, 128, device="cuda").index_add_(0, idx.long(), x.float())
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=1e-3, rtol=1e-4)
    assert torch.equal(out, ops.segment_sum_rows_perm(x, perm, row_ptr))               # deterministic


def test_scatter_add_and_adam():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(5000, 128, generator=g).cuda().to(torch.bfloat16)
    idx = torch.randint(0, 97, (5000,), generator=g, dtype=torch.int32).cuda()
    out = torch.zeros(97, 128, device="cuda")
    ops.scatter_add_rows(x, idx, out)
    ref = torch.zeros(97, 128, device="cuda").index_add_(0, idx.long(), x.float())
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=1e-3, rtol=1e-4)
    p = torch.randn(4096, generator=g).cuda()
    p_ref = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([p_ref], lr=2e-3)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        grad = torch.randn(4096, generator=g).cuda()
        p_ref.grad = grad.clone() / 2
        opt.step()
        ops.adam_step(p, grad, m, v, step, 2e-3, grad_scale=0.5)
        np.testing.assert_allclose(p.cpu().numpy(), p_ref.detach().cpu().numpy(), atol=2e-6, rtol=1e-5)


def test_native_training_step_gradients_match_autograd():
    """Whole training step with dropout off: every parameter gradient of the native path (bf16 tensor-core GEMMs)
    against fp32 autograd through the drop-in's differentiable composition, which itself matches the oracle."""
    from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200.train_engine import RDesignTrainer
    cfg = RDesignConfig(k_neighbors=12, num_mpnn_layers=3)
    X, S, mask = synthetic_backbones(3, 60, 9, lengths=[60, 41, 23])
    X, S, mask = X.cuda(), S.cuda(), mask.cuda()

    ref = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    ref.load_state_dict(make_state_dict(cfg, 42), strict=True)
    h_V, S_v = ref(X, S, mask)
    loss_ref = F.cross_entropy(ref.readout(h_V), S_v)
    loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}

    model = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    tr = RDesignTrainer(model, autocast=False)
    loss = tr.loss_and_grads(X, S, mask, seed=1)
    assert abs(loss.item() - loss_ref.item()) < 2e-2, (loss.item(), loss_ref.item())
    worst = 0.0
    for n, p in model.named_parameters():
        a, b = p.grad.float(), ref_grads[n].float()
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        worst = max(worst, rel)
        assert rel < 6e-2 or (a - b).abs().max().item() < 1e-5, (n, rel, b.norm().item())
    print(f"native training step: loss {loss.item():.5f} vs {loss_ref.item():.5f}, worst relative gradient error {worst:.3e}")
    # one optimizer step moves the parameters like torch.optim.Adam on the reference gradients would
    before = tr.flat_p.clone()
    tr.apply_gradients()
    delta = (tr.flat_p - before)[:tr.n_params]
    assert torch.isfinite(delta).all() and (delta.abs() <= model.hparams.lr * 1.001).all() and delta.abs().max() > 0


def test_dropin_training_step_runs_on_the_native_kernels(monkeypatch):
    """``RNAModel.training_step`` (the reference-facing API, rdesign/model/rdesign.py:95-104) + ``loss.backward()`` + the
    optimizer of ``configure_optimizers``: the per-edge / per-node blocks run on the C-ABI kernels (launch counter), the
    gradients match the fp32 torch composition (``RMPNN_NATIVE_TRAIN=0``) and the loss goes down."""
    from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200 import _lib
    cfg = RDesignConfig(k_neighbors=12, num_mpnn_layers=3)
    X, S, mask = synthetic_backbones(3, 60, 9, lengths=[60, 41, 23])
    batch = (X, S, mask, mask.sum(1).long(), ["a", "b", "c"])

    monkeypatch.setenv("RMPNN_NATIVE_TRAIN", "0")
    ref = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    ref.load_state_dict(make_state_dict(cfg, 42), strict=True)
    loss_ref = ref.training_step(batch)
    loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}

    monkeypatch.setenv("RMPNN_NATIVE_TRAIN", "1")
    model = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    (opt,), _ = model.configure_optimizers()
    n0 = _lib.lib().rmp_launch_count()
    loss = model.training_step(batch)
    opt.zero_grad()
    loss.backward()
    assert _lib.lib().rmp_launch_count() - n0 > 20 * len(model.mpnn_layers)          # the step ran on the library's kernels
    assert abs(loss.item() - loss_ref.item()) < 3e-2
    for n, p in model.named_parameters():
        a, b = p.grad.float(), ref_grads[n].float()
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        assert rel < 1e-1 or (a - b).abs().max().item() < 1e-5, (n, rel)
    opt.step()
    sd = model.state_dict()                                                        # still a loadable reference state_dict
    RNAModel(**cfg.as_kwargs(), dropout=0.0).load_state_dict(sd, strict=True)

    model = RNAModel(**cfg.as_kwargs()).cuda().train()                              # dropout 0.1: the loop trains
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    (opt,), _ = model.configure_optimizers()
    losses = []
    for _ in range(25):
        loss = model.training_step(batch)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.05, losses
    model.eval()                                                                   # and the eval path sees the trained weights
    with torch.no_grad():
        h_V, S_v = model(X.cuda(), S.cuda(), mask.cuda())
    assert torch.isfinite(h_V).all() and h_V.shape[0] == int(mask.sum())


def test_training_loss_decreases_with_dropout_on():
    from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200.train_engine import RDesignTrainer
    cfg = RDesignConfig(k_neighbors=10, num_mpnn_layers=2)
    X, S, mask = synthetic_backbones(4, 50, 5)
    X, S, mask = X.cuda(), S.cuda(), mask.cuda()
    model = RNAModel(**cfg.as_kwargs()).cuda().train()
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    tr = RDesignTrainer(model)
    losses = [tr.step(X, S, mask).item() for _ in range(25)]
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5]) - 0.05, losses


@pytest.mark.parametrize("rows", [1, 130, 5000])
def test_gemm_rows_outputs(rows):
    """Lean row GEMM: z, fused Dropout(GELU(z)) with the same mask as rmp_act_fwd_bf16, fp32 accumulation, gathers."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows)
    A = torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16)
    W = (torch.randn(128, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
    b = 0.1 * torch.randn(128, generator=g).cuda()
    Nv = 37
    Pc, Pn = torch.randn(Nv, 128, generator=g).cuda(), torch.randn(Nv, 128, generator=g).cuda()
    ctr = torch.randint(0, Nv, (rows,), generator=g, dtype=torch.int32).cuda()
    nbr = torch.randint(0, Nv, (rows,), generator=g, dtype=torch.int32).cuda()
    acc = torch.full((rows, 128), 2.0, device="cuda")
    Pc, Pn = Pc.to(torch.bfloat16), Pn.to(torch.bfloat16)
    z, a = ops.gemm_rows(A, W, bias=b, Pc=Pc, Pn=Pn, ctr=ctr, nbr=nbr, act=(0.1, 5, 9), acc32=acc)
    ref = A.float() @ W.float().t() + b + Pc.float()[ctr.long()] + Pn.float()[nbr.long()]
    np.testing.assert_allclose(z.float().cpu().numpy(), _bf(ref).cpu().numpy(), atol=3e-2, rtol=1e-2)
    np.testing.assert_allclose(acc.cpu().numpy(), (ref + 2.0).cpu().numpy(), atol=2e-3, rtol=1e-4)
    assert torch.equal(a, ops.act_fwd(z, 0.1, 5, 9))                   # the fused activation IS the stand-alone pass
    z2, a2 = ops.gemm_rows(A, W)
    assert a2 is None
    np.testing.assert_allclose(z2.float().cpu().numpy(), _bf(A.float() @ W.float().t()).cpu().numpy(), atol=3e-2, rtol=1e-2)


@pytest.mark.parametrize("rows,L,transposed", [(1, 2, True), (130, 3, False), (5000, 9, True), (20000, 9, True)])
def test_gemm_rows_kcat_matches_sum_of_products(rows, L, transposed):
    """K-concatenated row GEMM (d h_E = sum over the layers of dz0_l W1e_l): one pass, fp32 output overwritten; a tile of
    block l that runs past its rows reads block l + 1's first rows, which must only reach output rows that are not stored."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows + L)
    A = torch.randn(L, rows, 128, generator=g).cuda().to(torch.bfloat16)
    W = (torch.randn(L, 128, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
    out = torch.full((rows, 128), 7.0, device="cuda")
    got = ops.gemm_rows_kcat(A, W, w_transposed=transposed, out=out)
    assert got is out
    ref = sum(A[l].float() @ (W[l].float() if transposed else W[l].float().t()) for l in range(L))
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), atol=2e-3 * L ** 0.5, rtol=1e-4)


@pytest.mark.parametrize("rows,width,bf16", [(1, 128, False), (1000, 128, True), (16384, 256, True), (777, 256, False)])
def test_colsum_matches_torch(rows, width, bf16):
    """Column sums (bias gradients) and the LayerNorm weight-gradient sums of the per-node modules in one pass."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows + width)
    a = torch.randn(rows, width, generator=g).cuda()
    if bf16:
        a = a.to(torch.bfloat16)
    s0, s1 = ops.colsum(a)
    assert s1 is None
    np.testing.assert_allclose(s0.cpu().numpy(), a.double().sum(0).cpu().numpy(), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-4)
    if not bf16:
        x = (2 * torch.randn(rows, width, generator=g) + 1).cuda()
        y, mean, rstd = torch.native_layer_norm(x, (width,), None, None, 1e-5)
        s0, s1 = ops.colsum(a, x, mean, rstd)
        ref = (a.double() * ((x.double() - mean.double()) * rstd.double())).sum(0)
        np.testing.assert_allclose(s1.cpu().numpy(), ref.cpu().numpy(), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-4)
        np.testing.assert_allclose(s0.cpu().numpy(), a.double().sum(0).cpu().numpy(), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-4)


@pytest.mark.parametrize("rows,b_kind", [(1, "none"), (1000, "fp32"), (16384, "bf16"), (77, "bf16")])
def test_add_layernorm_fwd_bwd_match_autograd(rows, b_kind):
    """norm(a + b) of the MPNN layer (rdesign/model/mpnn.py:35,37): fused forward (sum, statistics) and one-pass backward
    (dx, bf16 copy for a bf16 residual branch, weight / bias gradients) against torch autograd."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows)
    a = (2 * torch.randn(rows, 128, generator=g) + 0.3).cuda()
    b = None if b_kind == "none" else torch.randn(rows, 128, generator=g).cuda().to(torch.bfloat16 if b_kind == "bf16" else torch.float32)
    gamma = (1 + 0.1 * torch.randn(128, generator=g)).cuda()
    beta = (0.1 * torch.randn(128, generator=g)).cuda()
    dy = torch.randn(rows, 128, generator=g).cuda()
    x, y, mean, rstd = ops.add_layernorm_fwd(a, b, gamma, beta, 1e-5)
    xr = (a if b is None else a + b.float()).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (128,), gr, br, 1e-5)
    np.testing.assert_allclose(x.cpu().numpy(), xr.detach().cpu().numpy(), atol=1e-6, rtol=0)
    np.testing.assert_allclose(y.cpu().numpy(), ref.detach().cpu().numpy(), atol=2e-5, rtol=1e-5)
    ref.backward(dy)
    dx, dx16, dgamma, dbeta = ops.layernorm_bwd(dy, x, mean, rstd, gamma, want_bf16=b_kind == "bf16")
    np.testing.assert_allclose(dx.cpu().numpy(), xr.grad.cpu().numpy(), atol=2e-5, rtol=1e-4)
    if b_kind == "bf16":
        assert torch.equal(dx16, dx.to(torch.bfloat16))
    else:
        assert dx16 is None
    np.testing.assert_allclose(dgamma.cpu().numpy(), gr.grad.cpu().numpy(), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-4)
    np.testing.assert_allclose(dbeta.cpu().numpy(), br.grad.cpu().numpy(), atol=2e-3 * max(1.0, rows ** 0.5), rtol=1e-4)


def test_node_act_fn_is_gelu_dropout_with_regenerated_mask():
    """Dropout(GELU(z)) of the per-node FFN rows through the edge path's kernels: [rows,256] rows as their own [2 rows,128] view."""
    from rmpnn_b200.train_engine import NodeActFn
    g = torch.Generator().manual_seed(1)
    z = (2 * torch.randn(500, 256, generator=g)).cuda().to(torch.bfloat16).requires_grad_(True)
    a = NodeActFn.apply(z, 0.0, 3, 40, None)
    zf = z.detach().float().requires_grad_(True)
    ref = F.gelu(zf)
    np.testing.assert_allclose(a.detach().float().cpu().numpy(), _bf(ref.detach()).cpu().numpy(), atol=1e-2, rtol=1e-2)
    da = torch.randn(500, 256, generator=g).cuda().to(torch.bfloat16)
    a.backward(da)
    (ref * da.float()).sum().backward()
    np.testing.assert_allclose(z.grad.float().cpu().numpy(), _bf(zf.grad).cpu().numpy(), atol=2e-2, rtol=2e-2)
    z2 = torch.ones(4000, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
    a2 = NodeActFn.apply(z2, 0.1, 3, 41, None)
    a2.backward(torch.ones_like(a2))
    assert torch.equal(z2.grad != 0, a2.detach() != 0)                      # the backward regenerates the forward's mask
    assert abs((a2.detach() != 0).float().mean().item() - (1 - 26 / 256)) < 3e-3


def test_native_training_step_autocast_gradients_match_autograd():
    """The bench path (bf16 node FFN, fused add + LayerNorm, fused GELU + Dropout, column-sum bias gradients) with dropout
    off against fp32 autograd through the drop-in's composition: looser than the fp32 node path (bf16 activations)."""
    from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200.train_engine import RDesignTrainer
    cfg = RDesignConfig(k_neighbors=12, num_mpnn_layers=3)
    X, S, mask = synthetic_backbones(3, 60, 9, lengths=[60, 41, 23])
    X, S, mask = X.cuda(), S.cuda(), mask.cuda()
    ref = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    ref.load_state_dict(make_state_dict(cfg, 42), strict=True)
    h_V, S_v = ref(X, S, mask)
    loss_ref = F.cross_entropy(ref.readout(h_V), S_v)
    loss_ref.backward()
    ref_grads = {n: p.grad.clone() for n, p in ref.named_parameters()}
    model = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
    model.load_state_dict(make_state_dict(cfg, 42), strict=True)
    tr = RDesignTrainer(model, autocast=True)
    loss = tr.loss_and_grads(X, S, mask, seed=1)
    assert abs(loss.item() - loss_ref.item()) < 3e-2, (loss.item(), loss_ref.item())
    worst = 0.0
    for n, p in model.named_parameters():
        a, b = p.grad.float(), ref_grads[n].float()
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        worst = max(worst, rel)
        assert rel < 1e-1 or (a - b).abs().max().item() < 1e-5, (n, rel, b.norm().item())
    print(f"native training step (autocast): loss {loss.item():.5f} vs {loss_ref.item():.5f}, worst relative gradient error {worst:.3e}")


@pytest.mark.parametrize("Nv,deg,p_drop", [(300, 7, 0.0), (2000, 25, 0.1), (64, 3, 0.1)])
def test_edge_mlp_train_fwd_matches_the_stage_by_stage_path(Nv, deg, p_drop):
    """Fused training forward of the edge kernel against the three row GEMMs + segmented sum it replaces: same bf16 z within
    rounding, the SAME dropout masks (operand-form activations are zero exactly where the stand-alone pass drops), activations
    and neighbour sums within bf16 tolerance.  deg 3 puts more than 16 nodes into a tile (vector-reduction bypass)."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(Nv)
    ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32).cuda()
    Ne = ctr.numel()
    rp = (torch.arange(Nv + 1) * deg).to(torch.int32).cuda()
    nbr = torch.randint(0, Nv, (Ne,), generator=g, dtype=torch.int32).cuda()
    hE = torch.randn(Ne, 128, generator=g).cuda().to(torch.bfloat16)
    W = (torch.randn(384, 128, generator=g) / 11.3).cuda().to(torch.bfloat16)
    b = [None, (0.1 * torch.randn(128, generator=g)).cuda(), (0.1 * torch.randn(128, generator=g)).cuda()]
    Pc = torch.randn(Nv, 128, generator=g).cuda().to(torch.bfloat16)
    Pn = torch.randn(Nv, 128, generator=g).cuda().to(torch.bfloat16)
    seed, sid, alpha = 17, 32, 1 / 30
    agg = torch.zeros(Nv, 128, device="cuda")
    zs, as_ = ops.edge_mlp_train_fwd(hE, W, b, Pc.float(), Pn, ctr, nbr, rp, agg, alpha, p_drop, seed, sid)
    z0, a0 = ops.gemm_rows(hE, W[:128], Pc=Pc, Pn=Pn, ctr=ctr, nbr=nbr, act=(p_drop, seed, sid))
    z1, a1 = ops.gemm_rows(a0, W[128:256], bias=b[1], act=(p_drop, seed, sid + 1))
    z2, a2 = ops.gemm_rows(a1, W[256:], bias=b[2], act=(p_drop, seed, sid + 2))
    agg_ref = ops.segment_sum_rows(a2, rp) * alpha
    ks = ops.keep_scale(p_drop)
    tol = dict(atol=6e-2, rtol=3e-2)                       # bf16 activations through three stages, two evaluation orders
    np.testing.assert_allclose(zs[0].float().cpu().numpy(), z0.float().cpu().numpy(), **tol)
    assert torch.equal(as_[0] != 0, a0 != 0) or ((as_[0] != 0) != (a0 != 0)).float().mean().item() < 1e-3   # GELU(z) == 0 ties
    np.testing.assert_allclose((as_[0].float() * (0.5 * ks)).cpu().numpy(), a0.float().cpu().numpy(), **tol)
    np.testing.assert_allclose(zs[1].float().cpu().numpy(), z1.float().cpu().numpy(), **tol)
    np.testing.assert_allclose((as_[1].float() * (0.5 * ks)).cpu().numpy(), a1.float().cpu().numpy(), **tol)
    np.testing.assert_allclose(zs[2].float().cpu().numpy(), z2.float().cpu().numpy(), **tol)
    np.testing.assert_allclose(agg.cpu().numpy(), agg_ref.cpu().numpy(), atol=2e-2, rtol=3e-2)
    if p_drop > 0:
        assert abs((as_[1] == 0).float().mean().item() - 26 / 256) < 2e-2


def test_normalize_fwd_bwd_match_autograd():
    from rmpnn_b200 import ops
    from rmpnn_b200.train_engine import _normalize
    g = torch.Generator().manual_seed(4)
    x = (3 * torch.randn(2000, 128, generator=g) + 0.5).cuda()
    dy = torch.randn(2000, 128, generator=g).cuda()
    gain = (1 + 0.1 * torch.randn(128, generator=g)).cuda()
    bias = (0.1 * torch.randn(128, generator=g)).cuda()
    y = ops.normalize_fwd(x, gain, bias)
    xr, gr, br = x.clone().requires_grad_(True), gain.clone().requires_grad_(True), bias.clone().requires_grad_(True)
    ref = _normalize(xr, gr, br)
    np.testing.assert_allclose(y.float().cpu().numpy(), _bf(ref.detach()).cpu().numpy(), atol=2e-2, rtol=1e-2)
    ref.backward(dy)
    dx, dgain, dbias, dbsum = ops.normalize_bwd(x, dy, gain)
    np.testing.assert_allclose(dx.float().cpu().numpy(), _bf(xr.grad).cpu().numpy(), atol=2e-3, rtol=2e-2)
    np.testing.assert_allclose(dgain.cpu().numpy(), gr.grad.cpu().numpy(), atol=2e-2, rtol=1e-3)
    np.testing.assert_allclose(dbias.cpu().numpy(), br.grad.cpu().numpy(), atol=2e-2, rtol=1e-3)
    np.testing.assert_allclose(dbsum.cpu().numpy(), dx.float().sum(0).cpu().numpy(), atol=2e-2, rtol=1e-3)


def test_node_side_functions_match_the_modules():
    """LayerNormFn / LinearFn (weight and bias gradients as small products) against nn.LayerNorm / nn.Linear."""
    from rmpnn_b200.train_engine import LayerNormFn, LinearFn
    torch.manual_seed(0)
    x = torch.randn(3000, 128, device="cuda", requires_grad=True)
    ln = torch.nn.LayerNorm(128).cuda()
    ln.weight.data.normal_(); ln.bias.data.normal_()
    g = torch.randn(3000, 128, device="cuda")
    got = torch.autograd.grad(LayerNormFn.apply(x, ln.weight, ln.bias, ln.eps), (x, ln.weight, ln.bias), g)
    want = torch.autograd.grad(ln(x), (x, ln.weight, ln.bias), g)
    for a, b in zip(got, want):
        assert (a - b).abs().max().item() <= 2e-4 * max(1.0, b.abs().max().item())
    lin = torch.nn.Linear(128, 256).cuda()
    g = torch.randn(3000, 256, device="cuda")
    y = LinearFn.apply(x, lin.weight, lin.bias, lin.weight.detach().to(torch.bfloat16))
    assert y.dtype == torch.bfloat16 and (y.float() - lin(x)).abs().max().item() < 3e-2
    got = torch.autograd.grad(y, (x, lin.weight, lin.bias), g.to(torch.bfloat16))
    want = torch.autograd.grad(lin(x), (x, lin.weight, lin.bias), g)
    for a, b in zip(got, want):
        assert (a.float() - b).abs().max().item() <= 2e-2 * b.abs().max().item()


def test_autocast_step_gradients_are_close_to_the_fp32_step():
    """The bf16 node-side path (LayerNormFn / LinearFn) against the same step with the fp32 modules, dropout off."""
    from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones
    from rdesign.model.rdesign import RNAModel
    from rmpnn_b200.train_engine import RDesignTrainer
    cfg = RDesignConfig(k_neighbors=12, num_mpnn_layers=3)
    X, S, mask = (t.cuda() for t in synthetic_backbones(3, 60, 9, lengths=[60, 41, 23]))
    grads = []
    for autocast in (False, True):
        model = RNAModel(**cfg.as_kwargs(), dropout=0.0).cuda().train()
        model.load_state_dict(make_state_dict(cfg, 42), strict=True)
        tr = RDesignTrainer(model, autocast=autocast)
        loss = tr.loss_and_grads(X, S, mask, seed=1)
        grads.append((loss.item(), {n: p.grad.clone().float() for n, p in model.named_parameters()}))
    assert abs(grads[0][0] - grads[1][0]) < 2e-2
    for n, b in grads[0][1].items():
        a = grads[1][1][n]
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        assert rel < 8e-2 or (a - b).abs().max().item() < 1e-5, (n, rel)


def test_segment_sum_equals_scatter_add_by_centre():
    """rmp_segment_sum_rows (no atomics; a node's edges are contiguous) against rmp_scatter_add_rows and torch, on a real
    ragged graph (degrees min(K, L_b)) and with empty segments."""
    from rmpnn_b200 import ops
    from rmpnn_b200.synthetic import synthetic_backbones
    X, S, mask = synthetic_backbones(4, 70, 3, lengths=[70, 33, 9, 1])
    g = ops.build_graph(X.cuda(), mask.cuda(), 25)
    x = torch.randn(g.Ne, 128, device="cuda").to(torch.bfloat16)
    got = ops.segment_sum_rows(x, g.row_ptr[:g.Nv + 1])
    ref = torch.zeros(g.Nv, 128, device="cuda")
    ops.scatter_add_rows(x, g.ctr, ref)
    want = torch.zeros(g.Nv, 128, device="cuda").index_add_(0, g.ctr[:g.Ne].long(), x.float())
    assert (got - want).abs().max().item() < 1e-4 and (got - ref).abs().max().item() < 1e-4
    rp = torch.tensor([0, 0, 3, 3, 10, 10], dtype=torch.int32, device="cuda")          # empty segments -> zeros
    got = ops.segment_sum_rows(x[:10].contiguous(), rp)
    assert float(got[0].abs().max()) == 0.0 and float(got[2].abs().max()) == 0.0 and float(got[4].abs().max()) == 0.0
    assert (got[1] - x[:3].float().sum(0)).abs().max().item() < 1e-5 and (got[3] - x[3:10].float().sum(0)).abs().max().item() < 1e-5
