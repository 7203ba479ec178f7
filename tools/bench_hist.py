"""histogram kernel alone: 2 GiB of scmos-like, uniform-random and constant uint16 voxels (CUDA events)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
torch.cuda.set_device(0); sq.set_device(0)
for kind in ("scmos", "random", "zeros"):
    vol = torch_volume((256, 2048, 2048), kind)
    hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
    for _ in range(2):
        sq.histogram_device(vol, hist)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hist.zero_()
    e0.record()
    for _ in range(5):
        sq.histogram_device(vol, hist)
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ref = torch.zeros(65536, dtype=torch.int64, device="cuda")
    flat = vol.view(-1)
    for lo in range(0, flat.numel(), 1 << 28):
        ref += torch.bincount(flat[lo: lo + (1 << 28)].to(torch.int32) & 0xFFFF, minlength=65536)
    ok = torch.equal((ref * 5) & 0xFFFFFFFF, hist.to(torch.int64) & 0xFFFFFFFF)
    print(f"{kind}: {ms:.3f} ms = {vol.numel() * 2 / ms / 1e6:.0f} GB/s, exact {ok}")
