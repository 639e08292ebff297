"""Per-layer GPU time of one training iteration (C3 workload, B images in `split` micro-batches): every launch through the C ABI
is bracketed with CUDA events (abi.TIMING) and the convolution entry points are tagged with their layer shape.
    python tools/train_breakdown.py [--batch 16] [--split 2] > breakdown.json"""
import argparse
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--split", type=int, default=2)
    ap.add_argument("--r1", action="store_true", help="time a do_r1 iteration instead of a plain one")
    ap.add_argument("--native", action="store_true", help="instead: torch.profiler over one iteration, kernels that are NOT this library's")
    args = ap.parse_args()
    pkg = importlib.import_module("3dhumangan_b200")
    abi = importlib.import_module("3dhumangan_b200.abi")
    gen = importlib.import_module("3dhumangan_b200.modules.generator")
    disc = importlib.import_module("3dhumangan_b200.modules.discriminator")
    ts = importlib.import_module("3dhumangan_b200.train_step")
    dev = torch.device("cuda", 0)
    cfg = bench.workload_cfg(pkg, "C2")
    cfg["nerf_noise"] = 0.5
    cfg["batch_split"] = args.split
    torch.manual_seed(0)
    G = gen.Map3DGenerator(**cfg).to(dev).train()
    G.set_device(dev)
    D = disc.UNetDiscriminator(**cfg).to(dev).train()
    trainer = ts.Trainer(G, D, cfg, amp=False)
    B, Hg, Wg = args.batch, cfg["gen_height"], cfg["gen_width"]
    g = torch.Generator().manual_seed(5)
    batch = dict(z_d=torch.randn(B, cfg["latent_dim"], generator=g).to(dev), z_g=torch.randn(B, cfg["latent_dim"], generator=g).to(dev),
                 images=torch.randn(B, 3, Hg, Wg, generator=g).clamp_(-1, 1).to(dev),
                 labels=torch.randint(1, cfg["label_dim"], (B, Hg, Wg), generator=g).to(dev),
                 cond={k: v.to(dev) for k, v in pkg.synthetic.make_conditions(B, seed=1).items()})
    for _ in range(2):
        trainer.iteration(batch)
    while bool(cfg["phases"][D.step % len(cfg["phases"])]["do_r1"]) != args.r1:
        trainer.iteration(batch)
    if args.native:
        from torch.profiler import ProfilerActivity, profile
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            trainer.iteration(batch)
            torch.cuda.synchronize()
        rows = []
        for ev in prof.key_averages():
            t = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
            if t > 0:
                rows.append({"name": ev.key[:120], "ms": round(t / 1e3, 3), "calls": ev.count})
        rows.sort(key=lambda r: -r["ms"])
        ours = [r for r in rows if "hg::" in r["name"]]
        other = [r for r in rows if "hg::" not in r["name"]]
        print(json.dumps({"library_ms": sum(r["ms"] for r in ours), "other_ms": sum(r["ms"] for r in other), "other": other[:40],
                          "library": ours[:12]}))
        return
    abi.TIMING_TAGS = True
    abi.TIMING = []
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    trainer.iteration(batch)
    e.record()
    torch.cuda.synchronize()
    per = {}
    for name, a, b in abi.TIMING:
        d = per.setdefault(name, [0.0, 0])
        d[0] += a.elapsed_time(b)
        d[1] += 1
    rows = sorted(per.items(), key=lambda kv: -kv[1][0])
    print(json.dumps({"batch": B, "split": args.split, "do_r1": args.r1, "iteration_ms_with_events": s.elapsed_time(e),
                      "abi_kernels_ms": sum(v[0] for v in per.values()),
                      "rows": [{"kernel": k, "ms": round(v[0], 3), "launches": v[1]} for k, v in rows]}))


if __name__ == "__main__":
    main()
