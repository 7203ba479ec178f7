"""ncu driver: the LZ4 stage on the three lowest bit planes (13-15: the sparse, noisy ones) of a background-removed slab."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
shape = (128, 2048, 2048)
torch.cuda.set_device(0); sq.set_device(0)
vol = torch_volume(shape, "scmos")
_, thr = sq.estimate_background_device(vol)
planes = torch.empty_like(vol)
sq.bitswap_encode_device(1, vol.view(-1), planes.view(-1), threshold=thr)
n = planes.numel()
low = planes.view(-1)[n // 16 * 13:].contiguous().view(torch.uint8)
out = torch.empty_like(low)
for _ in range(3):
    payload = sq.lz4_encode_device(low, pitch=256)
    sq.lz4_decode_device(payload, out)
torch.cuda.synchronize()
assert torch.equal(out, low)
print("payload", payload.numel(), "of", low.numel(), sq.last_lz4_stats())
