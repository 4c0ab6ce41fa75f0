import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rna-mpnn_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["toy_degenerate", "ragged_k8", "default_k25", "real_k25", "single_k48"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_global_torch_state():
    """The reference's ``seeding()`` (mirrored by the drop-in) sets ``float32_matmul_precision('medium')`` and cuDNN
    flags process-wide; on hosts with bf16 matmul units that changes the CPU oracle's fp32 results.  Tests that call
    it must not leak the setting into the parity tests that follow."""
    prec = torch.get_float32_matmul_precision()
    det, bench = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    yield
    torch.set_float32_matmul_precision(prec)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = det, bench


def load_golden(name):
    """Golden vectors produced by the unmodified reference (tests/golden/make_golden.py)."""
    from oracle.weights import RDesignConfig, make_state_dict
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    c = [int(v) for v in g["cfg"]]
    cfg = RDesignConfig(hidden_dim=c[0], k_neighbors=c[1], num_message_layers=c[2], num_dense_layers=c[3],
                        dim_dense_layers=c[4], num_mpnn_layers=c[5], readout_hidden_dim=c[6],
                        num_readout_layers=c[7])
    g["config"] = cfg
    g["params"] = make_state_dict(cfg, int(g["seed"]))
    for k in ("X", "S", "mask"):
        g[k] = torch.from_numpy(g[k])
    return g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


RNAMPNN_GOLDEN_CASES = ["rnampnn_toy", "rnampnn_ragged", "rnampnn_default", "rnampnn_real", "rnampnn_alt",
                        "rnampnn_tiny"]


def load_rnampnn_golden(name):
    """Golden vectors of the secondary (rnampnn) path (tests/golden/make_golden_rnampnn.py)."""
    from oracle.weights import RNAMPNNConfig, make_rnampnn_state_dict
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    g = {k: z[k] for k in z.files}
    fields = list(RNAMPNNConfig.__dataclass_fields__)
    cfg = RNAMPNNConfig(**{f: int(v) for f, v in zip(fields, g["cfg"])})
    g["name"] = name
    g["config"] = cfg
    g["params"] = make_rnampnn_state_dict(cfg, int(g["seed"]))
    for k in ("coords", "mask"):
        g[k] = torch.from_numpy(g[k])
    return g


@pytest.fixture(params=RNAMPNN_GOLDEN_CASES)
def rgolden(request):
    return load_rnampnn_golden(request.param)


def assert_close_fp32(actual, o32, o64, atol, what=""):
    """|actual - exact| <= atol + 4 * max|oracle_fp32 - exact| with ``exact`` the oracle evaluated in fp64: the
    CUDA path has to be as close to the exact result as the reference's own fp32 arithmetic is.  Used where a
    GraphNorm over badly scaled channels (real structures with atoms parked at the origin) amplifies the fp32
    rounding noise of the Linear stack in front of it."""
    a = torch.as_tensor(actual).double().cpu().reshape(o64.shape)
    noise = float((o32.double() - o64).abs().max())
    err = float((a - o64).abs().max())
    assert err <= atol + 4.0 * noise, (what, err, noise)


def to_fp64(params):
    return {k: v.double() for k, v in params.items()}
