"""Multi-GPU paths (skipped on a single-GPU box): the fused frame-sharded resample -> row-sharded stack (N4) and the
fused reassembly of the stacked image, each checked bit for bit against its plain NCCL form by the tool it wraps."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script_args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port)] + script_args
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:]
    return json.loads(lines[-1])


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_scatter_resample_into_peer_stack_jobs():
    d = _torchrun([os.path.join("tools", "check_scatter.py"), "--frames", "8", "--width", "1500", "--height", "1001", "--reps", "1"], 29571)
    assert d["bit_identical_job_buffers"] and d["stack_identical"] and d["n_gpus"] == 2


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_two_gpu_bench_fused_gather_verified_against_nccl():
    d = _torchrun(["bench.py", "--gpus", "2", "--rows", "256", "--steps", "1", "--warmup", "3", "--no-e2e", "--no-cpu"], 29572)
    assert d["n_gpus"] == 2 and "verified against NCCL" in d["config"]["gather"]


@pytest.mark.skipif(_gpus() < 2, reason="needs two GPUs")
def test_two_gpu_star_detect_resample_stack_pipeline():
    """bench.py --config c3 under torchrun: detection and resample sharded over frames, fused scatter into the row-stripe
    stack jobs (ragged stripes: 201 + 200 rows), stack per stripe, all-gather; the run itself compares stars, the stripe
    rows of an own and of a peer's frame and the stacked rows with the CPU restatement and aborts on a difference"""
    d = _torchrun(["bench.py", "--gpus", "2", "--config", "c3", "--rows", "401", "--steps", "1", "--warmup", "1"], 29573)
    p = d["config"]["parity"]
    assert d["n_gpus"] == 2 and p["stars_bit_exact"] and p["stack_bit_exact"] and p["stripes_in_gathered_image"]
    assert p["resample_bit_exact_own_and_peer_frame"][2] is True
