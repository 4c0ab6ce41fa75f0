"""Parity on BASELINE.json's own configurations at their own sizes ("what's weak" 1):

* C2: the full 64 x 256 batch, K = 30, seed 1002 (the bench workload itself), both precision arms;
* C3: one ribosome-scale structure, 1 x 4096, K = 48, seed 1003, both arms (the 8-structure batch is eight
  independent copies of this shape; the CPU oracle needs ~4 s for one of them).

fp32 arm and fp32x3 arm (split-bf16 operands on tcgen05, rmp_row_mlp_x3): logits within 1e-4 of the CPU oracle and
the identical argmax sequence; bf16 arm: logits within 2e-2,
agreement of the argmax sequence reported and held above 99 % (flips only where two logits are within the budget).
Reference: RNAModel.forward rdesign/model/rdesign.py:82-88, validation_step :106-123.
"""
import numpy as np
import pytest
import torch

from oracle import rdesign_oracle as O
from oracle.weights import RDesignConfig, make_state_dict, synthetic_backbones

pytestmark = pytest.mark.gpu

CASES = {"c2": (64, 256, 30, 1002), "c3_one": (1, 4096, 48, 1003)}
_ref_cache = {}


def _reference(case):
    if case not in _ref_cache:
        B, L, K, seed = CASES[case]
        cfg = RDesignConfig(k_neighbors=K)
        params = make_state_dict(cfg, 42)
        X, S, mask = synthetic_backbones(B, L, seed)
        with torch.no_grad():
            ref = O.forward(params, X, S, mask, K, cfg.scale)
        _ref_cache[case] = (cfg, params, X, mask, ref)
    return _ref_cache[case]


@pytest.mark.parametrize("case", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "fp32x3", "bf16"])
def test_full_size_forward_matches_oracle(case, precision):
    from rmpnn_b200 import _lib
    from rmpnn_b200.engine import RDesignEngine
    cfg, params, X, mask, ref = _reference(case)
    prec = {"bf16": _lib.PREC_BF16, "fp32x3": _lib.PREC_FP32X3}.get(precision, _lib.PREC_FP32)
    eng = RDesignEngine(params, cfg.k_neighbors, cfg.scale, device="cuda", precision=prec)
    out = eng.forward(X.cuda(), mask.cuda())
    assert torch.equal(out["graph"].E_flat.cpu(), ref["E_idx"]), "E_idx differs from the oracle"
    err = (out["logits"].cpu() - ref["logits"]).abs().max().item()
    agree = (out["samples"].cpu() == ref["samples"]).float().mean().item()
    print(f"{case} {precision}: Nv={ref['logits'].shape[0]} max|dlogit|={err:.3e} argmax agreement={agree:.5f}")
    if precision != "bf16":
        assert err < 1e-4, err
        assert torch.equal(out["samples"].cpu(), ref["samples"])
    else:
        assert err < 2e-2, err
        assert agree > 0.99, agree
        # every flipped position must be a near-tie of the oracle's own top-2 logits (within twice the budget)
        flipped = (out["samples"].cpu() != ref["samples"]).nonzero().flatten()
        if flipped.numel():
            top2 = ref["logits"][flipped].topk(2, dim=-1).values
            assert float((top2[:, 0] - top2[:, 1]).max()) < 4e-2


def test_c2_graph_replay_equals_eager():
    """The CUDA-graph replay that bench.py times returns the same samples/logits as the eager forward."""
    from rmpnn_b200 import _lib
    from rmpnn_b200.engine import RDesignEngine
    cfg, params, X, mask, ref = _reference("c2")
    eng = RDesignEngine(params, cfg.k_neighbors, cfg.scale, device="cuda", precision=_lib.PREC_BF16)
    Xd, md = X.cuda(), mask.cuda()
    eager = eng.forward(Xd, md)
    sizes = (int(ref["logits"].shape[0]), int(ref["E_idx"].shape[1]))
    replay, static = eng.capture(Xd, md, sizes)
    static["X"].copy_(Xd); static["mask"].copy_(md)
    out = replay()
    torch.cuda.synchronize()
    assert (out["logits"] - eager["logits"]).abs().max().item() < 2e-3      # red.global order differs run to run
    assert (out["samples"] == eager["samples"]).float().mean().item() > 0.999
