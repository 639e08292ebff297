"""world_size-2 gloo tests of the N>1 host logic (no GPU): SyncBatchNorm statistics exchange and the
rank protocol of `bench.py --impl reference` under torch.distributed.run."""
import json
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import importlib, os, sys, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
syn = importlib.import_module("3dhumangan_b200.modules.synthesis_ops")
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
g = torch.Generator().manual_seed(7)
x = torch.randn(6, 256, 5, 7, generator=g, dtype=torch.float64) * 2 + 0.3      # the GLOBAL batch, same on both ranks
mine = x[:2] if rank == 0 else x[2:]                                             # uneven shards: 2 and 4 images
row = torch.zeros(syn.STAT_STRIDE, dtype=torch.float64)
row[:256] = mine.sum(dim=(0, 2, 3))
row[256:512] = mine.square().sum(dim=(0, 2, 3))
row[512] = mine.numel() / 256
syn.all_reduce_stats(row)
count = row[512]
mean = row[:256] / count
var = row[256:512] / count - mean * mean
ref_mean = x.mean(dim=(0, 2, 3)); ref_var = x.var(dim=(0, 2, 3), unbiased=False)
assert count == 6 * 35, count
assert torch.allclose(mean, ref_mean, atol=1e-12) and torch.allclose(var, ref_var, atol=1e-10)
# the running estimate uses the unbiased variance of the GLOBAL batch (nn.SyncBatchNorm semantics)
unb = var * count / (count - 1)
assert torch.allclose(unb, x.var(dim=(0, 2, 3), unbiased=True), atol=1e-10)
dist.barrier(); dist.destroy_process_group(); print("rank", rank, "ok")
'''


def test_syncbn_statistics_all_reduce_gloo():
    port = 29500 + os.getpid() % 2000
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(WORKER % {"root": ROOT, "port": port})
        path = f.name
    try:
        procs = [subprocess.Popen([sys.executable, path, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
                 for r in range(2)]
        outs = [p.communicate(timeout=240) for p in procs]
        for p, (o, e) in zip(procs, outs):
            assert p.returncode == 0, e[-2000:]
            assert "ok" in o
    finally:
        os.unlink(path)


def test_bench_reference_arm_under_torchrun_two_ranks():
    """rank 0 alone runs and prints the CPU reference line; the other rank exits 0 without work."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29700 + os.getpid() % 200), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "0", "--workload", "tiny"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "images/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["higher_is_better"] is True and d["metric"] == "images_per_sec_G_fwd_512x512"


def _avg_worker(rank, world, port_, q):
    import importlib, os, torch, torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts = importlib.import_module("3dhumangan_b200.train_step")
    m = torch.nn.Linear(3, 2)
    for i, p in enumerate(m.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    ts.average_gradients(m)
    q.put((rank, [float(p.grad.flatten()[0]) for p in m.parameters()]))
    dist.destroy_process_group()


def test_average_gradients_gloo():
    """The explicit gradient averaging that stands in for DDP's reducer hooks (train_step.average_gradients)."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port_ = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_avg_worker, args=(r, 2, port_, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res[0] == res[1] == [1.5, 3.0]
