"""BASELINE config 5: MSDeformAttn op sweep -- Len_q, T x L levels (neighbour frames presented as extra
levels), points 4-8, batch 1-64, fp32 / bf16 -- against the reference's vendored CUDA op (oracle/_ref,
recompiled for sm_100a) and the reference's CPU path (grid_sample formulation, oracle port).

    python tools/opsweep.py [--iters 20] [--ref] [--cpu] > sweep.jsonl

One JSON line per (configuration, implementation, pass).  Direct C-ABI calls, CUDA events, `local`
sampling regime (pixel centre / random reference + N(0, 3 px) offsets), inputs larger than L2 from
N = 8 upwards; smaller cases are timed back to back (L2-warm, as they run inside a network).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import opbench  # noqa: E402

BASE = [(75, 100), (38, 50), (19, 25)]
S0 = sum(h * w for h, w in BASE)

# (name, N, Lq (None = S), frames-as-levels k, P)
CONFIGS = [
    ("enc_N1", 1, None, 1, 4), ("enc_N2", 2, None, 1, 4), ("enc_N8", 8, None, 1, 4), ("enc_N64", 64, None, 1, 4),
    ("dec_N1", 1, 60, 1, 4), ("dec_N64", 64, 60, 1, 4), ("q300_N1", 1, 300, 1, 4), ("q300_N8", 8, 300, 1, 4),
    ("enc_N1_k2", 1, S0, 2, 4), ("enc_N1_k3", 1, S0, 3, 4), ("enc_N1_k4", 1, S0, 4, 4),
    ("enc_N1_P8", 1, None, 1, 8), ("enc_N8_P8", 8, None, 1, 8), ("dec_N1_k4_P8", 1, 60, 4, 8),
]


def bytes_for(N, S, M, D, L, P, Lq, e):
    v = min(N * S * M * D, 4 * N * Lq * M * L * P * D)
    samples = N * Lq * M * L * P
    fwd = e * (v + N * Lq * M * D) + 4 * 3 * samples
    bwd = e * (v + N * Lq * M * D) + 4 * (v + 6 * samples)
    return fwd, bwd


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ref", action="store_true", help="time the vendored CUDA op too (oracle/_ref)")
    ap.add_argument("--cpu", action="store_true", help="time the CPU grid_sample path too (N <= 2 only)")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import snipper_b200  # noqa: F401
    peak = opbench.PEAK
    ref = None
    if args.ref:
        from oracle.build_ref import load_ref
        ref = load_ref()
    if args.cpu:
        from oracle import torch_ref
        torch.set_num_threads(os.cpu_count() or 1)
    for name, N, Lq, k, P in CONFIGS:
        if args.only and name not in args.only.split(","):
            continue
        levels = BASE * k
        L = len(levels)
        S = S0 * k
        Lq_ = Lq or S
        value, shapes, lsi, loc, attn, go = opbench.make(N, Lq_, P=P, levels=levels, regime="local")
        # (k > 1: the queries are the pixels of ONE frame sampling k frames' worth of levels; opbench.make
        #  draws uniform locations whenever Lq != S)
        rows = []
        fwd, bwd, _, _, keep = opbench.direct_calls(value, shapes, lsi, loc, attn, go)
        fb, bb = bytes_for(N, S, 8, 48, L, P, Lq_, 4)
        rows += [("ours_fp32", "fwd", fwd, fb), ("ours_fp32", "bwd", bwd, bb)]
        bfwd, bbwd, _, _, bkeep = opbench.direct_calls(value, shapes, lsi, loc, attn, go, bf16=True)
        fb16, bb16 = bytes_for(N, S, 8, 48, L, P, Lq_, 2)
        rows += [("ours_bf16", "fwd", bfwd, fb16), ("ours_bf16", "bwd", bbwd, bb16)]
        if ref is not None:
            rows += [("vendored_fp32", "fwd", lambda: ref.ms_deform_attn_forward(value, shapes, lsi, loc, attn, 64), fb),
                     ("vendored_fp32", "bwd", lambda: ref.ms_deform_attn_backward(value, shapes, lsi, loc, attn, go, 64), bb)]
        flush = N >= 8
        for impl, which, fn, nbytes in rows:
            med, best = opbench.time_fn(fn, args.iters, flush, inner=5)
            print(json.dumps({"config": name, "N": N, "Lq": Lq_, "levels": L, "P": P, "impl": impl, "pass": which,
                              "us": round(med, 2), "alg_MB": round(nbytes / 1e6, 2), "GBps": round(nbytes / med / 1e3, 1),
                              "frac_of_measured_hbm": round(nbytes / med / 1e3 / peak, 4), "l2_flush": flush}), flush=True)
        if args.cpu and N <= 2:
            cv, cl, ca = value.cpu().requires_grad_(True), loc.cpu().requires_grad_(True), attn.cpu().requires_grad_(True)
            cs = shapes.cpu()
            best_f, best_b = 1e9, 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                out = torch_ref.msda_core_torch(cv, cs, cl, ca)
                t1 = time.perf_counter()
                out.backward(go.cpu())
                t2 = time.perf_counter()
                cv.grad = cl.grad = ca.grad = None
                best_f, best_b = min(best_f, t1 - t0), min(best_b, t2 - t1)
            for which, t, nbytes in (("fwd", best_f, fb), ("bwd", best_b, bb)):
                print(json.dumps({"config": name, "N": N, "Lq": Lq_, "levels": L, "P": P, "impl": "cpu_grid_sample_fp32",
                                  "pass": which, "us": round(t * 1e6, 1), "alg_MB": round(nbytes / 1e6, 2),
                                  "GBps": round(nbytes / t / 1e9, 2), "threads": torch.get_num_threads()}), flush=True)
        del keep, bkeep, value, loc, attn, go
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
