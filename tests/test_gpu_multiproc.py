"""The real multi-process, multi-GPU step (one process per GPU, CUDA IPC peer mapping over NVLink, torch.distributed for
the plumbing) against the whole-buffer oracle.  Needs at least two devices: skipped on a one-GPU box (the driver's scaling
run exercises the same code: bench.py verifies every N > 1 run against the oracle before it times it)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    if os.environ.get("ADSB_EMU") == "1":
        return 0
    from gr_adsb_b200 import _native
    return int(_native.lib().adsb_device_count())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_bench_multi_gpu_equals_whole_buffer_oracle(world):
    if _n_devices() < world:
        pytest.skip("needs %d CUDA devices" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"),
           "--gpus", str(world), "--log2n", "23", "--steps", "3", "--warmup", "3", "--no-e2e", "--transport", "peer"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["n_gpus"] == world and line["transport"] == "peer"
    assert line["verify"]["ok"] and line["verify"]["ranks"] == world and line["verify"]["samples"] == world << 23
    assert line["carry"]["unsettled_steps"] == 0
