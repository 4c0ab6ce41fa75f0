"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the
golden vectors of the unmodified reference.  Run on the B200 box: pytest -m gpu.

Tolerances (BASELINE.json north_star): neighbour indices bit-exact under the documented
(distance, index) tie-break; fp32 logits within 1e-4 abs; identical argmax sequences.
"""
import numpy as np
import pytest
import torch

from oracle import rdesign_oracle as O
from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones

pytestmark = pytest.mark.gpu

LOGIT_TOL_FP32 = 1e-4
RAW_TOL = 1e-5
EMBED_TOL = 5e-5          # h_V after the node embedding + Normalize
EMBED_TOL_EDGE = 2e-4     # h_E: the quaternion's sqrt(|1 + Rxx - Ryy - Rzz|) amplifies 1e-7 noise to 1e-4

# toy_degenerate sits on an integer lattice with repeated atoms: several dihedral signs come from
# sign() of an *exactly cancelling* dot product, so they flip with any change of rounding order
# (they flip between fp32 CPU runs with and without FMA as well).  Raw features are compared
# entry-wise with those entries detected and skipped; stages downstream of them are only checked
# for graph equality and finiteness on that one case.
FRAGILE_CASE = "toy_degenerate"


def is_fragile(golden):
    return golden["X"].shape[1] == 6


def assert_close_conditioned(actual, f32, f64, atol, what, jitter=None, min_fraction=None):
    """Compare against the fp32 oracle, widening the tolerance where the quantity is
    ill-conditioned in fp32 itself (measured as |oracle_fp32 - oracle_fp64|, and optionally
    as the change under a 1e-6 A jitter of the coordinates: acos near +-1, sqrt of a
    cancelling sum in the quaternion, sign() of a rounding-noise value on degenerate
    geometry).  Entries that move by > 1e-2 are indeterminate (their sign depends on
    rounding order) and are skipped."""
    a = actual.double().cpu()
    cond = (f32.double() - f64).abs()
    if jitter is not None:
        cond = torch.maximum(cond, (jitter.double() - f64).abs())
    tol = atol + 8.0 * cond
    ok = cond <= 1e-2
    bad = ((a - f32.double()).abs() > tol) & ok
    if min_fraction is None:
        assert not bad.any(), (what, int(bad.sum()), float(((a - f32.double()).abs() * ok).max()))
    else:       # fragile geometry: a bounded share of entries may sit on a rounding-order discontinuity
        assert 1.0 - float(bad.double().mean()) >= min_fraction, (what, int(bad.sum()))
    return float(ok.double().mean())


@pytest.fixture(scope="module")
def ops():
    from rmpnn_b200 import ops as _ops
    return _ops


def _engine(params, cfg, **kw):
    from rmpnn_b200.engine import RDesignEngine
    return RDesignEngine(params, cfg.k_neighbors, cfg.scale, device="cuda", **kw)


# --------------------------------------------------------------------------- kNN (a-1)
def test_knn_bit_exact_on_golden_inputs(golden, ops):
    X, mask, cfg = golden["X"], golden["mask"], golden["config"]
    D_o, I_o = O.knn(X[:, :, 0, :], mask, cfg.k_neighbors)
    D_c, I_c = ops.knn(X.cuda(), mask.cuda(), cfg.k_neighbors)
    assert torch.equal(I_c.cpu(), I_o)
    assert torch.equal(D_c.cpu(), D_o)                      # distances bit for bit (IEEE sqrt)
    # the reference's own topk distances come from a 1-ulp-faithful CPU sqrt
    D_r = torch.sort(torch.from_numpy(golden["D_neighbors"]), -1).values
    assert ((D_c.cpu() - D_r).abs() <= 1.2e-7 * D_r.abs()).all()


@pytest.mark.parametrize("B,N,K,lengths", [
    (4, 256, 30, None), (3, 300, 48, [300, 77, 20]), (2, 1000, 25, [1000, 999]), (5, 33, 64, [33, 32, 31, 2, 1]),
    (1, 1, 5, None), (2, 130, 128, [130, 129]),
])
def test_knn_bit_exact_random(ops, B, N, K, lengths):
    X, S, mask = synthetic_backbones(B, N, 900 + N + K, lengths=lengths)
    D_o, I_o = O.knn(X[:, :, 0, :], mask, K)
    D_c, I_c = ops.knn(X.cuda(), mask.cuda(), K)
    assert torch.equal(I_c.cpu(), I_o)
    assert torch.equal(D_c.cpu(), D_o)


def test_knn_ties_lower_index_first(ops):
    # all atoms coincide -> every distance ties; canonical order is 0..K-1
    X = torch.zeros(1, 40, 6, 3)
    mask = torch.ones(1, 40)
    D, I = ops.knn(X.cuda(), mask.cuda(), 7)
    assert torch.equal(I.cpu(), torch.arange(7).expand(1, 40, 7))
    # integer lattice: many exact ties, oracle (stable sort) must agree
    g = torch.Generator().manual_seed(3)
    X = torch.randint(0, 3, (2, 64, 6, 3), generator=g).float()
    D_o, I_o = O.knn(X[:, :, 0, :], torch.ones(2, 64), 9)
    D_c, I_c = ops.knn(X.cuda(), torch.ones(2, 64).cuda(), 9)
    assert torch.equal(I_c.cpu(), I_o) and torch.equal(D_c.cpu(), D_o)


def test_knn_non_prefix_mask(ops):
    X, S, mask = synthetic_backbones(2, 50, 5)
    mask[0, 3:9] = 0
    mask[1, ::4] = 0
    D_o, I_o = O.knn(X[:, :, 0, :], mask, 12)
    D_c, I_c = ops.knn(X.cuda(), mask.cuda(), 12)
    assert torch.equal(I_c.cpu(), I_o) and torch.equal(D_c.cpu(), D_o)


def test_knn_long_chain_global_memory_path(ops):
    # N > 8192 takes the non-staged path
    X, S, mask = synthetic_backbones(1, 8200, 31)
    K = 16
    D_c, I_c = ops.knn(X.cuda(), mask.cuda(), K)
    rows = torch.tensor([0, 1, 4099, 8199])
    P = X[0, :, 0, :]
    for r in rows.tolist():
        d = P - P[r]
        dist = O.sqrt_rn((d * d)[:, 0] + (d * d)[:, 1] + (d * d)[:, 2] + 1e-6)
        ds, idx = torch.sort(dist, stable=True)
        assert torch.equal(I_c[0, r].cpu(), idx[:K]) and torch.equal(D_c[0, r].cpu(), ds[:K])


# --------------------------------------------------------------------------- graph (a-2)
def test_flat_graph_matches_oracle(golden, ops):
    X, mask, cfg = golden["X"], golden["mask"], golden["config"]
    g = ops.build_graph(X.cuda(), mask.cuda(), cfg.k_neighbors)
    _, I_o = O.knn(X[:, :, 0, :], mask, cfg.k_neighbors)
    E_o, batch_id = O.flatten_graph(mask, I_o)
    assert (g.Nv, g.Ne) == (int(mask.sum()), E_o.shape[1])
    assert torch.equal(g.E_flat.cpu(), E_o)
    assert torch.equal(g.ctr.cpu().long(), E_o[0]) and torch.equal(g.nbr.cpu().long(), E_o[1])
    # centres identical to the reference's own flattening; whole list where it is tie free
    assert np.array_equal(g.E_flat[0].cpu().numpy(), golden["E_idx"][0])
    src = g.node_src[: g.Nv].cpu().long()
    assert torch.equal(src // mask.shape[1], batch_id)
    rp = g.row_ptr[: g.Nv + 1].cpu().long()
    assert torch.equal(rp[1:] - rp[:-1], torch.bincount(E_o[0], minlength=g.Nv))


def test_flat_graph_non_prefix_mask(ops):
    """Masks with holes: the reference's own flattening adds dense residue indices to
    compacted offsets (rdesign/model/feature.py:229-232), which only makes sense for prefix
    masks (all its collate produces).  The kernels use compacted ids on both ends."""
    X, S, mask = synthetic_backbones(3, 37, 8, lengths=[37, 20, 3])
    mask[0, 5:11] = 0
    mask[1, 0] = 0
    K = 6
    g = ops.build_graph(X.cuda(), mask.cuda(), K)
    _, I_o = O.knn(X[:, :, 0, :], mask, K)
    B, N = mask.shape
    rank = (torch.cumsum(mask, 1) - 1).long()
    off = torch.cumsum(mask.sum(1), 0) - mask.sum(1)
    want_c, want_n = [], []
    for b in range(B):
        deg = min(K, int(mask[b].sum()))
        for i in range(N):
            if mask[b, i] == 1:
                for k in range(deg):
                    want_c.append(int(off[b]) + int(rank[b, i]))
                    want_n.append(int(off[b]) + int(rank[b, int(I_o[b, i, k])]))
    assert g.E_flat.cpu().tolist() == [want_c, want_n]


# --------------------------------------------------------------------------- features (a-3..a-6)
def test_raw_features_match_oracle_and_reference(golden, ops):
    X, mask, cfg = golden["X"], golden["mask"], golden["config"]
    g = ops.build_graph(X.cuda(), mask.cuda(), cfg.k_neighbors)
    node_c, edge_c = ops.raw_features(X.cuda(), g)
    E = g.E_dense.cpu()
    node_o, edge_o, _ = O.raw_features(X, mask, cfg.k_neighbors, E_idx=E)
    node_64, edge_64, _ = O.raw_features(X.double(), mask.double(), cfg.k_neighbors, E_idx=E)
    gen = torch.Generator().manual_seed(1)
    # jitter only real coordinates: atoms parked at the origin (NaN -> 0, padding) stay exact zeros
    Xj = X.double() + 1e-6 * torch.randn(X.shape, generator=gen, dtype=torch.float64) * (X != 0)
    node_j, edge_j, _ = O.raw_features(Xj, mask.double(), cfg.k_neighbors, E_idx=E)
    frac = 0.95 if is_fragile(golden) else None
    cov_n = assert_close_conditioned(node_c, node_o, node_64, RAW_TOL, "node_raw vs oracle", node_j, frac)
    cov_e = assert_close_conditioned(edge_c, edge_o, edge_64, RAW_TOL, "edge_raw vs oracle", edge_j, frac)
    if not is_fragile(golden):
        assert cov_n > 0.97 and cov_e > 0.995, (cov_n, cov_e)
    # node features do not depend on neighbour order: compare with the reference's own tensors
    assert_close_conditioned(node_c, torch.from_numpy(golden["node_raw"]), node_64, RAW_TOL, "node_raw vs reference",
                             node_j, frac)
    if np.array_equal(g.E_flat.cpu().numpy(), golden["E_idx"]):
        st = int(golden["edge_stride"])
        assert_close_conditioned(edge_c[::st], torch.from_numpy(golden["edge_raw"]), edge_64[::st], RAW_TOL,
                                 "edge_raw vs reference", edge_j[::st], frac)


def test_embeddings_match_reference(golden):
    X, mask, cfg = golden["X"], golden["mask"], golden["config"]
    eng = _engine(golden["params"], cfg)
    g = eng.graph(X.cuda(), mask.cuda())
    h_V, h_E = eng.embed(X.cuda(), g)
    assert torch.isfinite(h_V).all() and torch.isfinite(h_E).all()
    if is_fragile(golden):
        return
    np.testing.assert_allclose(h_V.cpu().numpy(), golden["h_V0"], atol=EMBED_TOL, rtol=0)
    if np.array_equal(g.E_flat.cpu().numpy(), golden["E_idx"]):
        st = int(golden["edge_stride"])
        np.testing.assert_allclose(h_E.cpu().numpy()[::st], golden["h_E"], atol=EMBED_TOL_EDGE, rtol=0)


# --------------------------------------------------------------------------- layers + head (a-7..a-9)
def test_forward_fp32_matches_oracle(golden):
    X, S, mask, cfg = golden["X"], golden["S"], golden["mask"], golden["config"]
    eng = _engine(golden["params"], cfg)
    h_V, h_E, g, layers = eng.encode(X.cuda(), mask.cuda(), return_layers=True)
    logits, probs, samples = eng.readout(h_V)
    ref = O.forward(golden["params"], X, S, mask, cfg.k_neighbors, cfg.scale, E_idx=g.E_dense.cpu(),
                    return_layers=True)
    assert torch.equal(g.E_flat.cpu(), ref["E_idx"])
    assert torch.isfinite(logits).all()
    if is_fragile(golden):
        return
    for l, (a, b) in enumerate(zip(layers, ref["layers"])):
        np.testing.assert_allclose(a.cpu().numpy(), b.numpy(), atol=1e-4, rtol=0, err_msg=f"layer {l}")
    np.testing.assert_allclose(logits.cpu().numpy(), ref["logits"].numpy(), atol=LOGIT_TOL_FP32, rtol=0)
    np.testing.assert_allclose(probs.cpu().numpy(), ref["probs"].numpy(), atol=LOGIT_TOL_FP32, rtol=0)
    assert torch.equal(samples.cpu(), ref["samples"])


def test_forward_fp32_matches_reference_golden(golden):
    """Directly against the unmodified reference's stored outputs (graphs without ties)."""
    X, mask, cfg = golden["X"], golden["mask"], golden["config"]
    eng = _engine(golden["params"], cfg)
    out = eng.forward(X.cuda(), mask.cuda())
    if not np.array_equal(out["graph"].E_flat.cpu().numpy(), golden["E_idx"]):
        pytest.skip("reference topk broke ties differently on this case; covered via the oracle")
    np.testing.assert_allclose(out["logits"].cpu().numpy(), golden["logits"], atol=LOGIT_TOL_FP32, rtol=0)
    assert np.array_equal(out["samples"].cpu().numpy(), golden["samples"])


@pytest.mark.parametrize("B,L,K,lengths", [(4, 256, 30, None), (6, 190, 25, [190, 150, 101, 64, 17, 3])])
def test_forward_fp32_default_model_medium_batch(B, L, K, lengths):
    cfg = RDesignConfig(k_neighbors=K)
    params = make_state_dict(cfg, 42)
    X, S, mask = synthetic_backbones(B, L, 1002, lengths=lengths)
    ref = O.forward(params, X, S, mask, K, cfg.scale)
    out = _engine(params, cfg).forward(X.cuda(), mask.cuda())
    assert torch.equal(out["graph"].E_flat.cpu(), ref["E_idx"])
    np.testing.assert_allclose(out["logits"].cpu().numpy(), ref["logits"].numpy(), atol=LOGIT_TOL_FP32, rtol=0)
    assert torch.equal(out["samples"].cpu(), ref["samples"])


def test_sizes_hint_skips_host_sync_and_matches():
    cfg = RDesignConfig(k_neighbors=10, num_mpnn_layers=1)
    params = make_state_dict(cfg, 3)
    X, S, mask = synthetic_backbones(2, 30, 4, lengths=[30, 12])
    eng = _engine(params, cfg)
    a = eng.forward(X.cuda(), mask.cuda())
    b = eng.forward(X.cuda(), mask.cuda(), sizes=(42, 30 * 10 + 12 * 10))
    assert torch.equal(a["logits"], b["logits"])


# --------------------------------------------------------------------------- sampling
def test_sampling_temperature(ops):
    g = torch.Generator().manual_seed(0)
    h = torch.randn(20000, 128, generator=g).cuda()
    W = (torch.randn(4, 128, generator=g) * 0.05).cuda()
    b = torch.zeros(4).cuda()
    logits, probs, greedy = ops.readout_head(h, W, b, 0.0)
    ref_logits = h.cpu() @ W.cpu().t()
    np.testing.assert_allclose(logits.cpu().numpy(), ref_logits.numpy(), atol=1e-5, rtol=0)
    assert torch.equal(greedy.cpu(), torch.softmax(ref_logits, -1).argmax(-1))
    # T -> 0 reproduces argmax
    _, _, cold = ops.readout_head(h, W, b, 1e-4, seed=7)
    assert (cold == greedy).float().mean().item() > 0.999
    # T = 1: empirical class frequencies follow the mean softmax; draws depend on the seed
    _, _, s1 = ops.readout_head(h, W, b, 1.0, seed=1)
    _, _, s2 = ops.readout_head(h, W, b, 1.0, seed=2)
    _, _, s1b = ops.readout_head(h, W, b, 1.0, seed=1)
    assert torch.equal(s1, s1b) and not torch.equal(s1, s2)
    freq = torch.bincount(s1.cpu(), minlength=4).float() / s1.numel()
    np.testing.assert_allclose(freq.numpy(), probs.cpu().mean(0).numpy(), atol=0.02)


# --------------------------------------------------------------------------- errors
def test_error_paths(ops):
    X, S, mask = synthetic_backbones(1, 200, 1)
    with pytest.raises(RuntimeError, match="RMP_MAX_K|exceeds"):
        ops.knn(X.cuda(), mask.cuda(), 150)


# --------------------------------------------------------------------------- device-side collate (section 8f)
@pytest.mark.parametrize("dtype,atoms_out", [(torch.float64, 6), (torch.float32, 7)])
def test_device_collate_matches_host_featurize(ops, dtype, atoms_out):
    """rmp_collate against the host collate of the reference (rdesign/utils/data.py:84-119): bit-exact."""
    from rdesign.utils.data import ATOMS, featurize
    rng = np.random.default_rng(3)
    lengths = [17, 1, 40, 23, 40]
    structs = []
    for L in lengths:
        c = rng.normal(size=(L, 7, 3)) * 30
        c[rng.random((L, 7)) < 0.05] = np.nan                      # missing atoms
        structs.append(c.astype(np.float64 if dtype == torch.float64 else np.float32))
    packed = torch.from_numpy(np.concatenate(structs, 0)).cuda()
    starts = np.concatenate([[0], np.cumsum(lengths)[:-1]])
    order = [3, 0, 4, 1]                                            # a batch is any subset, in any order
    X, mask = ops.collate(packed, torch.tensor(starts[order], dtype=torch.int64).cuda(),
                          torch.tensor([lengths[i] for i in order], dtype=torch.int32).cuda(), 40, atoms_out)
    if atoms_out == 6:
        batch = [{"name": str(i), "seq": "A" * lengths[i],
                  "coords": {a: structs[i][:, k, :] for k, a in enumerate(ATOMS)}} for i in order]
        Xh, _, mh, _, _ = featurize(batch)
    else:
        Xh = torch.zeros(len(order), 40, 7, 3)
        mh = torch.zeros(len(order), 40)
        for j, i in enumerate(order):
            Xh[j, :lengths[i]] = torch.from_numpy(np.nan_to_num(structs[i], nan=0.0)).float()
            mh[j, :lengths[i]] = 1
    assert torch.equal(X.cpu(), Xh) and torch.equal(mask.cpu(), mh)


def test_knn_shared_memory_opt_in_boundary(ops):
    """N just below 3072: 48 KB of staged points plus the kernel's static rings crosses the 48 KB default limit
    (found on the 3065-nt chain of the reference's validation set)."""
    X, _, mask = synthetic_backbones(1, 3065, 77)
    D_o, I_o = O.knn(X[:, :, 0, :], mask, 25)
    D_c, I_c = ops.knn(X.cuda(), mask.cuda(), 25)
    assert torch.equal(I_c.cpu(), I_o) and torch.equal(D_c.cpu(), D_o)
