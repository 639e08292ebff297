"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the CPU oracle port on the
host cores) prints ONE JSON line with the keys the driver parses, under a plain launch and under a 2-rank torchrun launch
(rank 0 prints, the other rank exits 0 without work); the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
            "config", "cpu_baseline", "e2e")


def _run(cmd, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def _check_line(out, n_gpus):
    lines = [l for l in out.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, (k, sorted(d))
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["higher_is_better"] is True
    assert d["metric"] == "images_per_sec_G_fwd_512x512" and d["unit"] == "images/s" and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_prints_the_contract_line():
    r = _run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 1)


def test_reference_arm_under_torchrun_prints_one_line_from_rank_0():
    r = _run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
              "--master-port", str(29600 + os.getpid() % 300), "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1",
              "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 2)


def test_gpu_arm_refuses_to_run_without_a_device():
    r = _run([sys.executable, "bench.py", "--steps", "1", "--warmup", "0", "--no-cpu", "--no-train"], timeout=300)
    assert r.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in r.stdout.splitlines()), r.stdout[-500:]
