"""BASELINE config 4: Snipper T=4 training step (forward + backward + clip + AdamW), batch 2 per GPU,
stock DDP over NCCL, synthetic 600x800 snippets and a synthetic loss over every output head.

    python tools/trainbench.py [--steps 10] [--warmup 3] [--batch 2] [--bf16]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
           --master-port P tools/trainbench.py --steps 10

One JSON line from rank 0: training snippets/s (whole job, max-over-ranks device time) and the share of
the step spent in the MSDA kernels (CUDA events around every fused-layer launch in extra eager steps).
The loop mirrors the reference's train_one_epoch (engine.py:36-79: forward, loss, backward,
clip_grad_norm_(0.1), AdamW.step); the Hungarian-matching criterion is replaced by a dense synthetic
loss (it is CPU scipy code outside the hot path).  Not the headline bench (that is bench.py).
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_loss(out):
    loss = out["pred_kpts2d"].square().mean() + out["pred_depth"].square().mean() + out["pred_logits"].square().mean()
    for h in out["heatmaps"]:
        loss = loss + h.square().mean()
    for aux in out.get("aux_outputs", []):
        loss = loss + aux["pred_kpts2d"].square().mean() + aux["pred_depth"].square().mean() + aux["pred_logits"].square().mean()
    return loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--bf16", action="store_true", help="torch.autocast(bfloat16): bf16 GEMMs and bf16 MSDA gathers")
    args = ap.parse_args()
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import snipper_b200
    from snipper_b200 import ops
    from snipper_b200.harness.snipper_net import build_snipper
    torch.manual_seed(42)
    model = build_snipper(snipper_b200.MSDeformAttn).to(dev).train()
    net = model
    if world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])  # reference main.py:184
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)  # reference main.py:206-207
    g = torch.Generator().manual_seed(1000 + rank)
    xs = [torch.rand(args.batch, 12, 600, 800, generator=g).to(dev) for _ in range(2)]

    def step(i):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.bf16):
            out, _ = net(xs[i % len(xs)])
            loss = synthetic_loss(out)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 0.1)  # reference engine.py:75
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(args.steps):
        loss = step(i)
    e.record()
    barrier()
    from snipper_b200 import sharding
    ms = sharding.max_over_ranks(s.elapsed_time(e), dev)
    # MSDA share: events around our launches, two extra steps
    ops.STATS.reset()
    ops.STATS.timing = True
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
    ops.STATS.timing = False
    per = ops.STATS.kernel_ms()
    msda_ms = sum(sum(v) for v in per.values()) / 2
    if rank == 0:
        print(json.dumps({
            "metric": "train_snippets_per_sec_T4_600x800", "value": world * args.batch * args.steps / (ms * 1e-3),
            "unit": "snippets/s", "n_gpus": world, "steps": args.steps, "ms_per_step": ms / args.steps,
            "batch_per_gpu": args.batch, "dtype": "bf16-autocast" if args.bf16 else "f32",
            "parallelism": "DDP x%d (NCCL gradient all-reduce, 171 MB fp32 per step)" % world if world > 1 else "single GPU",
            "msda_ms_per_step": msda_ms, "msda_share": msda_ms / (ms / args.steps),
            "msda_kernels_ms_per_step": {("%s Lq=%d T1=%d" % (k[0], k[1][7], k[1][2])) if len(k[1]) >= 8 else k[0]: round(sum(v) / 2, 3)
                                         for k, v in per.items()},
            "final_loss": float(loss.detach()), "data": "synthetic", "loss": "synthetic dense loss over all heads"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
