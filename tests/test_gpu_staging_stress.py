"""Regression test for the round-1 pinned-ring race (csrc/staging.cu): the slot bookkeeping of the page-locked ring restarted
with every staged_h2d call, so the first chunks of z-slab k+1 were copied into slots whose DMAs of slab k could still be
running — wrong voxels from SQY_PipelineEncode_UI16 on >= 512 MiB pageable stacks with nthreads > 1, depending on timing.
The streamed host encode (256 MiB slabs = 8 ring chunks each, three slots) is repeated with the thread counts that make
the host copies fastest; every blob must decode to the voxels of the device path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_streamed_pageable_encode_is_repeatable(sq, cuda):
    from sqeazy_b200.synth import torch_volume

    shape = (72, 2048, 2048)     # 576 MiB: two full slabs and a ragged third
    d_vol = torch_volume(shape, "scmos", index=11)
    h_vol = cuda.empty(shape, dtype=cuda.int16)          # pageable
    h_vol.copy_(d_vol)
    vol = h_vol.numpy().view(np.uint16)
    d_out = cuda.empty(shape, dtype=cuda.int16, device="cuda")
    out = np.empty(vol.size, dtype=np.uint16)            # pageable
    for rep in range(4):
        for t in (16, 8, 3):
            blob = sq.encode("bitswap1->lz4", vol, nthreads=t)
            sq.decode_device(cuda.from_numpy(blob).cuda(), d_out)
            assert cuda.equal(d_out, d_vol), f"encode rep={rep} nthreads={t}"
            out[...] = 0x5A5A
            sq.decode(blob, nthreads=t, out=out)
            assert np.array_equal(out, vol.reshape(-1)), f"decode rep={rep} nthreads={t}"
    # the ring is shared by the two directions and by consecutive calls on different streams
    blob = sq.encode("rmestbkrd->bitswap1->lz4", vol, nthreads=16)
    want = sq.decode(blob, nthreads=1)
    for t in (16, 5):
        assert np.array_equal(sq.decode(sq.encode("rmestbkrd->bitswap1->lz4", vol, nthreads=t), nthreads=t), want)
    sq.release_scratch()
