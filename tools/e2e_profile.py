"""Where does one end-to-end snippet go?  torch.profiler kernel table of an eager bench step
(the same network / inputs as bench.py), grouped into: MSDA kernels (ours), GEMM, conv, other.

    python tools/e2e_profile.py [--steps 3] [--train]

Prints one JSON line with the per-group device time per step and the top kernels.  Not a bench
value (profiler overhead) -- it only says which part of the step the hot path is.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def group_of(name):
    n = name.lower()
    if "msda_" in n:
        return "msda(ours)"
    if "gemm" in n or "cutlass" in n or "sgemm" in n or "gemv" in n or "xmma" in n and "conv" not in n:
        return "gemm"
    if "conv" in n or "cudnn" in n or "wgrad" in n or "dgrad" in n or "fprop" in n:
        return "conv"
    if "nccl" in n:
        return "nccl"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--no-fused-tails", action="store_true")
    args = ap.parse_args()
    import snipper_b200
    from snipper_b200.harness.snipper_net import build_snipper
    from torch.profiler import ProfilerActivity, profile
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = build_snipper(snipper_b200.MSDeformAttn).to(dev)
    model.train(args.train)
    if not args.no_fused_tails:
        snipper_b200.enable_fused_layer_tails(model)
    x = torch.rand(args.batch, 12, 600, 800, device=dev)

    def step():
        if args.train:
            out, _ = model(x)
            loss = sum(v.float().pow(2).mean() for k, v in out.items() if torch.is_tensor(v))
            loss.backward()
            model.zero_grad(set_to_none=True)
        else:
            with torch.no_grad():
                model(x)

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
    groups, kernels = {}, {}
    for ev in prof.key_averages():
        t = getattr(ev, "device_time_total", None)
        if t is None:
            t = getattr(ev, "cuda_time_total", 0.0)
        if not t:
            continue
        g = group_of(ev.key)
        groups[g] = groups.get(g, 0.0) + t / args.steps / 1e3
        kernels[ev.key] = (t / args.steps / 1e3, ev.count // args.steps)
    top = sorted(kernels.items(), key=lambda kv: -kv[1][0])[:25]
    print(json.dumps({"mode": "train" if args.train else "infer", "batch": args.batch,
                      "ms_per_step_by_group": {k: round(v, 3) for k, v in sorted(groups.items(), key=lambda kv: -kv[1])},
                      "total_ms": round(sum(groups.values()), 3),
                      "top_kernels": [{"name": k[:110], "ms": round(v[0], 3), "calls": v[1]} for k, v in top]}))


if __name__ == "__main__":
    main()
