"""Op-level timing: B200 kernels vs the reference's vendored CUDA op (oracle/_ref, recompiled
for sm_100a) on Snipper shapes.  CUDA events over back-to-back launches; L2 is flushed between
launches when --flush is given.  Prints one JSON line per (shape, pass).

    python tools/opbench.py [--iters 50] [--flush] [--ref]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LEVELS = [(75, 100), (38, 50), (19, 25)]
PEAK = 6457.1
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def algorithmic_bytes(N, S, M, D, L, P, Lq, e=4):
    """SURVEY.md section 8d (value terms capped at the bytes actually touched for small Lq)."""
    v = min(N * S * M * D, 4 * N * Lq * M * L * P * D)
    fwd = e * (v + 3 * N * Lq * M * L * P + N * Lq * M * D)
    bwd = e * (2 * v + N * Lq * M * D + 6 * N * Lq * M * L * P)
    return fwd, bwd


def make(N, Lq, M=8, D=48, P=4, levels=LEVELS, regime="local", device="cuda", seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = torch.as_tensor(levels, dtype=torch.long)
    L = len(levels)
    S = int(shapes.prod(1).sum())
    value = torch.randn(N, S, M, D, generator=g)
    if regime == "uniform" or Lq != S:
        loc = torch.rand(N, Lq, M, L, P, 2, generator=g)
    else:  # encoder-like: pixel centres + N(0, 3px)
        refs = []
        for H, W in levels:
            ys, xs = torch.meshgrid(torch.arange(H) + 0.5, torch.arange(W) + 0.5, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1) / W, ys.reshape(-1) / H], -1))
        ref = torch.cat(refs, 0)
        wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float()
        loc = ref[None, :, None, None, None, :] + torch.randn(N, Lq, M, L, P, 2, generator=g) * 3.0 / wh[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn(N, Lq, M, L * P, generator=g), -1).view(N, Lq, M, L, P)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    go = torch.randn(N, Lq, M * D, generator=g)
    return [t.to(device) for t in (value, shapes, lsi, loc, attn, go)]


def time_fn(fn, iters, flush):
    buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if buf is not None:
            buf.zero_()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        times.append(s.elapsed_time(e) * 1e3)
    times.sort()
    return times[len(times) // 2], times[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--flush", action="store_true")
    ap.add_argument("--ref", action="store_true", help="also time the vendored reference op")
    ap.add_argument("--regime", default="local")
    args = ap.parse_args()
    import snipper_b200  # noqa: F401
    ref = None
    if args.ref:
        from oracle.build_ref import load_ref
        ref = load_ref()
    S = sum(h * w for h, w in LEVELS)
    cases = [("enc_N1", 1, S), ("enc_N2", 2, S), ("enc_N8", 8, S), ("dec_N1", 1, 60), ("dec_N2", 2, 60)]
    for name, N, Lq in cases:
        value, shapes, lsi, loc, attn, go = make(N, Lq, regime=args.regime)
        fb, bb = algorithmic_bytes(N, S, 8, 48, 3, 4, Lq)
        fwd = lambda: torch.ops.snipper_b200.msda_forward(value, shapes, lsi, loc, attn, 64)
        bwd = lambda: torch.ops.snipper_b200.msda_backward(value, shapes, lsi, loc, attn, go, 64, False)
        rows = [("ours", "fwd", fwd, fb), ("ours", "bwd", bwd, bb)]
        if ref is not None:
            rows += [("vendored", "fwd", lambda: ref.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), fb),
                     ("vendored", "bwd", lambda: ref.ms_deform_attn_backward(value, shapes, lsi, loc, attn, go, 64), bb)]
        for impl, which, fn, nbytes in rows:
            med, best = time_fn(fn, args.iters, args.flush)
            print(json.dumps({"case": name, "impl": impl, "pass": which, "us_median": round(med, 2),
                              "us_best": round(best, 2), "alg_MB": round(nbytes / 1e6, 2),
                              "GBps": round(nbytes / med / 1e3, 1), "frac_of_measured_hbm": round(nbytes / med / 1e3 / PEAK, 4),
                              "l2_flush": args.flush, "regime": args.regime}))


if __name__ == "__main__":
    main()
