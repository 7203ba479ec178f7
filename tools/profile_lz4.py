"""Small driver for ncu: bitswap a cfg1-sized synthetic stack, then LZ4-encode and -decode it a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

shape = (256, 512, 512) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
reps = 3 if len(sys.argv) < 3 else int(sys.argv[2])
torch.cuda.set_device(0)
sq.set_device(0)
vol = torch_volume(shape, "scmos")
planes = torch.empty_like(vol)
thr = 0
if len(sys.argv) > 3 and sys.argv[3] == "rmest":
    _, thr = sq.estimate_background_device(vol)
if len(sys.argv) > 3 and sys.argv[3] == "quant":
    import numpy as np
    hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
    sq.histogram_device(vol, hist)
    torch.cuda.synchronize()
    enc, dec = sq.quantiser_luts(hist.cpu().numpy().view(np.uint32))
    planes = torch.empty(vol.numel(), dtype=torch.uint8, device="cuda")
    sq.lut_apply_device(vol, planes, enc)
    vol = planes
else:
    sq.bitswap_encode_device(1, vol.view(-1), planes.view(-1), threshold=thr)
pitch = 0
if len(sys.argv) > 4:      # one bit plane only (0 = lowest), encoded the way the pipeline does it: row pitch + no-noise hint
    k = int(sys.argv[4])
    flat = planes.view(torch.uint8).view(-1)
    seg = flat.numel() // 16
    planes = flat[(15 - k) * seg:(16 - k) * seg].contiguous()
    vol = planes
    pitch = shape[2] // 8 | (0x80000000 if thr else 0)
out = torch.empty_like(vol)
payload = None
for _ in range(reps):
    payload = sq.lz4_encode_device(planes, pitch=pitch)
    sq.lz4_decode_device(payload, out)
torch.cuda.synchronize()
assert torch.equal(out, planes)
print("payload", payload.numel(), sq.last_lz4_stats())
