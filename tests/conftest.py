import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _reset_process_switches():
    """A failing test must not leave the process-wide switches of snipper_b200.ops flipped for the next one."""
    yield
    ops = sys.modules.get("snipper_b200.ops")
    if ops is not None:
        ops.set_deterministic(False)
        ops.set_planar_slots(False)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def rel_err(a, b):
    """max |a-b| / max |b| -- the tolerance measure of BASELINE.json (per output tensor)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def level_start_index(shapes):
    return torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))


def make_case(N, M, D, shapes, P, Lq=None, regime="local", sigma_px=3.0, seed=0, dtype=torch.float32):
    """Seeded synthetic op inputs.

    regime 'uniform': loc ~ U[0,1) (the reference test.py regime).
    regime 'local'  : loc = reference point + N(0, sigma px)/size, a slice lands outside [0,1].
    """
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(shapes, dtype=torch.long)
    L = shapes.shape[0]
    S = int(shapes.prod(1).sum())
    if Lq is None:
        Lq = S
    value = torch.randn(N, S, M, D, generator=g)
    if regime == "uniform":
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    else:
        ref = torch.rand(N, Lq, 1, 1, 1, 2, generator=g) * 1.1 - 0.05
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        off = torch.randn(N, Lq, M, L, P, 2, generator=g) * sigma_px
        loc = ref + off / wh[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    grad_out = torch.randn(N, Lq, M * D, generator=g)
    return dict(value=value.to(dtype), shapes=shapes, lsi=level_start_index(shapes),
                loc=loc.to(dtype), attn=attn.to(dtype), grad_out=grad_out.to(dtype))
