"""N > 1 path: host-side logic on CPU with world_size-2 gloo; full parity on 2 GPUs when the box has them."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(script, nproc, port, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_slab_host_logic_gloo_world2(oracle_built):
    r = _torchrun("_gloo_worker.py", 2, 29541, 300)
    assert r.returncode == 0 and "GLOO_WORKER_PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_slab_host_logic_gloo_world4(oracle_built):
    """four ranks: every rank has two DIFFERENT neighbours (with two, the upper and the lower neighbour are one rank)"""
    r = _torchrun("_gloo_worker.py", 4, 29543, 300)
    assert r.returncode == 0 and "GLOO_WORKER_PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_two_gpu_parity_vs_oracle(oracle_built):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    r = _torchrun("mgpu_check.py", 2, 29542, 600)
    assert r.returncode == 0 and "MGPU_CHECK_PASSED" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_numa_binding_is_best_effort():
    """bench.py's N > 1 runs bind every rank to the NUMA node of its GPU (parm_b200/sharded.py): the cpulist parser,
    and that without a GPU (or without sysfs entries) the call reports why and leaves the affinity alone."""
    from parm_b200 import sharded
    assert sharded._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharded._parse_cpulist("\n") == set()
    before = os.sched_getaffinity(0)
    out = sharded.bind_to_gpu_numa_node(0, sysfs="/nonexistent")
    assert out["bound"] is False and "reason" in out
    assert os.sched_getaffinity(0) == before
