"""GPU, >= 2 devices: the data-parallel path on hardware (skipped on a single-GPU box; the same host logic runs at
world_size 2 over gloo in tests/test_dp_gloo.py).  Each test launches one process per GPU with torchrun:
  * tests/_dp_peer_check.py  - the peer-memory SUM all-reduce kernel (csrc/dp_allreduce.cu) against NCCL, five steps;
  * tests/_dp_train_check.py - N-rank-sharded TrainWindow step == single-process global-batch step (gradients, clipped
    norm, updated parameters, replicas identical), for the peer kernel and for the NCCL all-reduce.
Logs of a 2-GPU and an 8-GPU run are kept under profiles/ (r2_multi_gpu_tests.md)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(script, nproc):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=os.path.dirname(HERE))
    log = os.path.join(os.path.dirname(HERE), "gpurun_out")
    try:
        os.makedirs(log, exist_ok=True)
        with open(os.path.join(log, f"multi_{script[:-3]}_{nproc}gpu.log"), "w") as f:
            f.write(out.stdout + "\n--- stderr ---\n" + out.stderr[-4000:])
    except OSError:
        pass
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def _world():
    n = torch.cuda.device_count()
    return 8 if n >= 8 else 4 if n >= 4 else 2


def test_peer_allreduce_kernel_matches_nccl():
    out = _torchrun("_dp_peer_check.py", _world())
    assert out.count("peer all-reduce OK") == _world()


def test_sharded_step_equals_global_batch_step():
    out = _torchrun("_dp_train_check.py", _world())
    assert "FAILED" not in out and out.count("PeerGradAllReduce: sharded step == global-batch step: OK") == _world()
    assert out.count("one-launch DP update: sharded step == global-batch step: OK") == _world()
