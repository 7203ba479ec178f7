"""cfg5: time-lapse stream of 1024x1024x128 uint16 stacks through remove_background(threshold=110)->bitswap4->lz4, the
stacks of a rank (stack v -> rank v mod G, sqeazy_b200/dist.py) as one batch (sqyx_*_batch_device_UI16) vs one call per
stack. usage: bench_cfg5_batch.py [n_stacks]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

n = 32 if len(sys.argv) < 2 else int(sys.argv[1])
shape = (128, 1024, 1024)
pipeline = "remove_background(threshold=110)->bitswap4->lz4"
torch.cuda.set_device(0); sq.set_device(0)
vols = [torch_volume(shape, "scmos", index=i) for i in range(n)]
raw = vols[0].numel() * 2
cap = sq.max_compressed_length(pipeline, raw)
bufs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(n)]
outs = [torch.empty_like(v) for v in vols]


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, r


ms_e1, blobs = timed(lambda: [sq.encode_device(pipeline, v, out=b) for v, b in zip(vols, bufs)])
ms_d1, _ = timed(lambda: [sq.decode_device(b, o) for b, o in zip(blobs, outs)])
ms_eb, blobs_b = timed(lambda: sq.encode_batch_device(pipeline, vols, outs=bufs))
ms_db, _ = timed(lambda: sq.decode_batch_device(blobs_b, outs))
ref = [torch.clamp(v.to(torch.int32) & 0xffff, min=110) - 110 for v in vols[:4]]
ok = all(torch.equal(o.to(torch.int32) & 0xffff, r) for o, r in zip(outs, ref))
tot = n * raw
print(f"cfg5 {n} stacks of 1024x1024x128 (ratio {tot / sum(b.numel() for b in blobs_b):.1f}): one call per stack: encode {ms_e1:.1f} ms = {tot / ms_e1 / 1e6:.0f} GB/s, "
      f"decode {ms_d1:.1f} ms = {tot / ms_d1 / 1e6:.0f} GB/s, pair {tot / (ms_e1 + ms_d1) / 1e6:.0f} voxel-GB/s | batch: encode {ms_eb:.1f} ms = {tot / ms_eb / 1e6:.0f} GB/s, "
      f"decode {ms_db:.1f} ms = {tot / ms_db / 1e6:.0f} GB/s, pair {tot / (ms_eb + ms_db) / 1e6:.0f} voxel-GB/s | bit-exact {ok}")
