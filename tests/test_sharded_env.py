"""The environment switches of the multi-GPU host path (csrc/sharded.inl) are read once per process, so they are exercised
in a child process: SQY_NO_NCCL=1 sums the quantiser's histograms on the host instead of ncclAllReduce (the blob must be the
same), SQY_CUDA_DEVICES picks the device set. Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, hashlib
sys.path.insert(0, %r)
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
shape = (64, 2048, 2048)
h = torch.empty(shape, dtype=torch.int16).pin_memory()
h.copy_(torch_volume(shape, "scmos", index=22))
vol = h.numpy().view(np.uint16)
blob = sq.encode("quantiser->lz4", vol, nthreads=4)
info = sq.last_shard_info()
back = sq.decode(blob, nthreads=4)
print("RESULT", info["gpus"], int(info["nccl"]), sq.nccl_allreduces(), hashlib.sha256(blob.tobytes()).hexdigest(), hashlib.sha256(back.tobytes()).hexdigest())
""" % ROOT


def _run(env_extra):
    env = dict(os.environ)
    env.pop("SQY_CUDA_DEVICE", None)
    env.update(env_extra)
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    return {"gpus": int(line[1]), "nccl": int(line[2]), "allreduces": int(line[3]), "blob": line[4], "voxels": line[5]}


def test_host_sum_and_nccl_sum_give_the_same_blob(cuda):
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    with_nccl = _run({"SQY_CUDA_DEVICES": "0,1"})
    host_sum = _run({"SQY_CUDA_DEVICES": "0,1", "SQY_NO_NCCL": "1"})
    one_gpu = _run({"SQY_CUDA_DEVICES": "0"})
    assert with_nccl["gpus"] == 2 and with_nccl["nccl"] == 1 and with_nccl["allreduces"] == 1
    assert host_sum["gpus"] == 2 and host_sum["nccl"] == 0 and host_sum["allreduces"] == 0
    assert one_gpu["nccl"] == 0
    assert with_nccl["blob"] == host_sum["blob"] == one_gpu["blob"]
    assert with_nccl["voxels"] == host_sum["voxels"] == one_gpu["voxels"]
