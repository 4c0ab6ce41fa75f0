"""fp32-parity tensor-core arm (rmp_row_mlp_x3: split-bf16 operands, 3 tcgen05 MMAs per fp32 product) against float64
PyTorch references of the same row MLPs, and the whole RMP_PREC_FP32X3 forward against the oracle / reference goldens at
the fp32 arm's own bar (1e-4 on the logits, identical argmax).  Reference: the Linear/GELU stacks of
rdesign/model/mpnn.py:15-38 and rdesign/model/functional.py:99-127."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import rdesign_oracle as O
from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones

pytestmark = pytest.mark.gpu


def _linears(widths, g, device="cuda"):
    from rmpnn_b200 import ops
    Ws = [torch.randn(n, k, generator=g) / k ** 0.5 for k, n in zip(widths[:-1], widths[1:])]
    bs = [0.1 * torch.randn(n, generator=g) for n in widths[1:]]
    return Ws, bs, [ops.pack_linear(W.to(device), b.to(device)) for W, b in zip(Ws, bs)]


def _ref_mlp(x, Ws, bs, gelu, adds=None):
    y = x.double()
    for s, (W, b) in enumerate(zip(Ws, bs)):
        y = y @ W.double().t() + b.double()
        if s == 0 and adds is not None:
            y = y + adds.double()
        if gelu[s]:
            y = F.gelu(y)
    return y


@pytest.mark.parametrize("rows,widths", [(1, [128, 128]), (127, [128, 128, 128, 128]), (300, [256, 256, 128]),
                                         (5000, [128, 256, 256, 256, 128]), (777, [384, 128, 128]),
                                         (40000, [128, 128, 128, 128])])
def test_x3_store_matches_float64(rows, widths):
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows + len(widths))
    x = torch.randn(rows, widths[0], generator=g)
    Ws, bs, packed = _linears(widths, g)
    gelu = [True] * (len(Ws) - 1) + [False]
    out = torch.empty(rows, widths[-1], device="cuda")
    ops.row_mlp([(x.cuda(), None)], list(zip(packed, gelu)), ops.EPI_STORE, out, rows, split=True)
    ref = _ref_mlp(x, Ws, bs, gelu)
    err = (out.cpu().double() - ref).abs().max().item()
    assert err < 3e-5 * max(1.0, ref.abs().max().item()), err
    # the same call on the bf16 arm is two orders of magnitude further away: the split terms are really used
    out16 = torch.empty_like(out)
    ops.row_mlp([(x.cuda(), None)], list(zip(packed, gelu)), ops.EPI_STORE, out16, rows, tensor_core=True)
    assert (out16.cpu().double() - ref).abs().max().item() > 20 * err


def test_x3_gathered_segments_adds_and_segsum():
    """The message MLP in its decomposed form: h_E rows + gathered per-node projections added to stage 0, GELU after
    every stage, neighbour sum (x alpha) into the centre rows."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(11)
    Nv, deg = 700, 9
    ctr = torch.repeat_interleave(torch.arange(Nv), deg).to(torch.int32)
    Ne = ctr.numel()
    nbr = torch.randint(0, Nv, (Ne,), generator=g, dtype=torch.int32)
    hE = torch.randn(Ne, 128, generator=g)
    P = 0.5 * torch.randn(Nv, 256, generator=g)
    Ws, bs, packed = _linears([128, 128, 128, 128], g)
    adds = P[ctr.long(), :128] + P[nbr.long(), 128:]
    ref = _ref_mlp(hE, Ws, bs, [True] * 3, adds)
    ref_agg = torch.zeros(Nv, 128, dtype=torch.float64).index_add_(0, ctr.long(), ref) / 30.0
    agg = torch.zeros(Nv, 128, device="cuda")
    Pd = P.cuda()
    ops.row_mlp([(hE.cuda(), None)], [(pl, True) for pl in packed], ops.EPI_SEGSUM, agg, Ne, seg_id=ctr.cuda(),
                alpha=1 / 30.0, split=True, adds=[(Pd, ctr.cuda()), (Pd[:, 128:], nbr.cuda())], add_stride=256)
    assert (agg.cpu().double() - ref_agg).abs().max().item() < 2e-5
    # gathered input segments (K = 384, the undecomposed form) and a residual epilogue
    hV = torch.randn(Nv, 128, generator=g)
    Ws2, bs2, packed2 = _linears([384, 128, 128], g)
    xin = torch.cat([hE, hV[ctr.long()], hV[nbr.long()]], 1)
    resid = torch.randn(Ne, 128, generator=g)
    ref2 = _ref_mlp(xin, Ws2, bs2, [True, False]) + resid.double()
    out = torch.empty(Ne, 128, device="cuda")
    hVd = hV.cuda()
    ops.row_mlp([(hE.cuda(), None), (hVd, ctr.cuda()), (hVd, nbr.cuda())], [(packed2[0], True), (packed2[1], False)],
                ops.EPI_RESID_ADD, out, Ne, resid=resid.cuda(), split=True)
    assert (out.cpu().double() - ref2).abs().max().item() < 3e-5


@pytest.mark.parametrize("rows", [5, 128, 1000, 20001])
def test_x3_normalize_and_residual_layernorm_epilogues(rows):
    """Linear + reference Normalize (the edge embedding, rdesign/model/feature.py:224) and FFN + LayerNorm(resid + y)
    (rdesign/model/mpnn.py:37-38) with the row statistics computed from the rows staged in shared memory."""
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, 128, generator=g)
    x[:, 115:] = 0
    Ws, bs, packed = _linears([128, 128], g)
    gain, beta = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    y = _ref_mlp(x, Ws, bs, [False])
    mu = y.mean(-1, keepdim=True)
    sigma = torch.sqrt(y.var(-1, keepdim=True, unbiased=True) + 1e-6)
    ref = gain.double() * (y - mu) / (sigma + 1e-6) + beta.double()
    out = torch.empty(rows, 128, device="cuda")
    ops.row_mlp([(x.cuda(), None)], [(packed[0], False)], ops.EPI_NORMALIZE, out, rows, gain=gain.cuda(), beta=beta.cuda(),
                split=True)
    assert (out.cpu().double() - ref).abs().max().item() < 5e-5     # one split-operand stage, then / sigma (O(1) outputs)
    Ws, bs, packed = _linears([128, 256, 256, 256, 128], g)
    gelu = [True, True, True, False]
    resid = torch.randn(rows, 128, generator=g)
    ref = F.layer_norm(resid.double() + _ref_mlp(resid, Ws, bs, gelu), (128,), gain.double(), beta.double(), 1e-5)
    ops.row_mlp([(resid.cuda(), None)], list(zip(packed, gelu)), ops.EPI_RESID_LN, out, rows, gain=gain.cuda(),
                beta=beta.cuda(), resid=resid.cuda(), split=True)
    assert (out.cpu().double() - ref).abs().max().item() < 6e-5     # four split-operand stages, then / sigma


def test_padded_edge_features_equal_packed():
    from rmpnn_b200 import ops
    X, S, mask = synthetic_backbones(3, 90, 9, lengths=[90, 40, 7])
    g = ops.build_graph(X.cuda(), mask.cuda(), 30)
    _, e115 = ops.raw_features(X.cuda(), g)
    _, e128 = ops.raw_features(X.cuda(), g, pad_edges=True)
    assert torch.equal(e128[:, :115], e115) and float(e128[:, 115:].abs().max()) == 0.0


def test_x3_rejects_unsupported_shapes():
    from rmpnn_b200 import ops
    g = torch.Generator().manual_seed(3)
    Ws, bs, packed = _linears([128, 512], g)
    out = torch.empty(8, 512, device="cuda")
    with pytest.raises(RuntimeError, match="rmp_row_mlp_x3"):
        ops.row_mlp([(torch.randn(8, 128).cuda(), None)], [(packed[0], False)], ops.EPI_STORE, out, 8, split=True)


@pytest.mark.parametrize("B,L,K,lengths", [(4, 256, 30, None), (6, 190, 25, [190, 150, 101, 64, 17, 3]),
                                           (3, 100, 48, [100, 37, 5])])
def test_forward_fp32x3_matches_oracle(B, L, K, lengths):
    from rmpnn_b200 import _lib
    from rmpnn_b200.engine import RDesignEngine
    cfg = RDesignConfig(k_neighbors=K)
    params = make_state_dict(cfg, 42)
    X, S, mask = synthetic_backbones(B, L, 77, lengths=lengths)
    with torch.no_grad():
        ref = O.forward(params, X, S, mask, K, cfg.scale)
    eng = RDesignEngine(params, K, cfg.scale, device="cuda", precision=_lib.PREC_FP32X3)
    out = eng.forward(X.cuda(), mask.cuda())
    err = (out["logits"].cpu() - ref["logits"]).abs().max().item()
    print(f"fp32x3 B={B} L={L} K={K}: max|dlogit| = {err:.3e}")
    assert err < 1e-4, err
    assert torch.equal(out["samples"].cpu(), ref["samples"])


def test_module_api_fp32x3_mode():
    """RNAModel with precision_mode = "fp32x3" (RMPNN_PRECISION=fp32x3) through the reference's module call."""
    from rdesign.model.rdesign import RNAModel
    cfg = RDesignConfig()
    params = make_state_dict(cfg, 42)
    X, S, mask = synthetic_backbones(3, 120, 5, lengths=[120, 80, 33])
    with torch.no_grad():
        ref = O.forward(params, X, S, mask, cfg.k_neighbors, cfg.scale)
    m = RNAModel(**cfg.as_kwargs())
    m.load_state_dict(params, strict=True)
    m = m.cuda().eval()
    m.precision_mode = "fp32x3"
    with torch.no_grad():
        h_V, S_v = m(X.cuda(), S.cuda(), mask.cuda())
        logits = m.readout(h_V)
    assert (logits.cpu() - ref["logits"]).abs().max().item() < 1e-4
