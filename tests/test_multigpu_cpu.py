"""world_size-2 gloo tests of the multi-GPU host logic (partitioning, histogram all-reduce -> identical LUT, distributed
rmestbkrd threshold). CPU only: the per-rank histograms stand in for what the histogram kernel produces on each GPU;
the LUT / support numerics are the product's own host code called through the C ABI."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from sqeazy_b200 import dist as sqdist
from sqeazy_b200.synth import numpy_volume


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _np_hist_into(sub, row):
    row += torch.from_numpy(np.bincount(sub.numpy().view(np.uint16).astype(np.int64), minlength=65536).astype(np.int32))


def _worker(rank, world, port, shape, q):
    import sqeazy_b200 as sq

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vol = numpy_volume(shape, "ref", index=21)
    z0, z1 = sqdist.zslab_for_rank(shape[0], rank, world)
    slab = torch.from_numpy(vol[z0:z1].view(np.int16))
    # quantiser: local histogram -> all-reduce -> LUT
    hist = torch.zeros(65536, dtype=torch.int32)
    _np_hist_into(slab.reshape(-1), hist)
    sqdist.allreduce_histogram(hist)
    enc, dec = sq.quantiser_luts(hist.numpy().view(np.uint32))
    # rmestbkrd: partial face histograms -> all-reduce -> support
    thr, sup = sqdist.global_background_threshold(slab, z0, shape, l2_bytes=1 << 16, histogram_fn=_np_hist_into,
                                                 support_fn=lambda h: sq.histogram_support(h, 0.99))
    q.put((rank, enc.tobytes(), dec.tobytes(), thr, sup.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


def test_partitions():
    assert sqdist.stacks_for_rank(10, 1, 4) == [1, 5, 9]
    covered = []
    for r in range(3):
        z0, z1 = sqdist.zslab_for_rank(10, r, 3)
        covered += list(range(z0, z1))
    assert covered == list(range(10))
    assert sqdist.zslab_for_rank(512, 7, 8) == (448, 512)


@pytest.mark.timeout(120)
def test_two_rank_histogram_allreduce_and_threshold(sq, port):
    shape = (12, 64, 96)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    prt = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, prt, shape, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    vol = numpy_volume(shape, "ref", index=21)
    enc, dec = port.quantiser_luts(port.histogram(vol))
    sup = port.darkest_face_supports(vol, 1 << 16)
    for rank, e, d, thr, s in res:
        assert e == enc.tobytes() and d == dec.tobytes(), "every rank must derive the single-GPU LUT"
        assert s == sup.tobytes() and thr == int(np.uint16(sup.min()))
