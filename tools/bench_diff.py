"""diff3x3x1 kernels alone (4 B/voxel algorithmic: 2 read + 2 written; the previous plane comes from L2) and pipelines with
the filter in front of bitswap1->lz4, device-resident, CUDA events.  python tools/bench_diff.py [ZxYxX]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

torch.cuda.set_device(0); sq.set_device(0)
shape = (512, 2048, 2048) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
vol = torch_volume(shape, "scmos")
raw = vol.numel() * 2
enc, back = torch.empty_like(vol), torch.empty_like(vol)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, r


ms_e, _ = timed(lambda: sq.diff_device(vol, enc))
n0 = sq.kernel_launches()
ms_d, _ = timed(lambda: sq.diff_device(enc, back, decode=True), 3)
launches = (sq.kernel_launches() - n0) // 4
print(f"diff3x3x1 {shape}: encode {ms_e:.3f} ms = {2 * raw / ms_e / 1e6:.0f} GB/s, decode {ms_d:.3f} ms = {2 * raw / ms_d / 1e6:.0f} GB/s "
      f"({launches} launches, {1e3 * ms_d / launches:.2f} us each), round trip {'ok' if torch.equal(back, vol) else 'WRONG'}")
del enc, back
out = torch.empty_like(vol)
for p in ("diff3x3x1->bitswap1->lz4", "bitswap1->lz4", "rmestbkrd->diff3x3x1->bitswap1->lz4", "rmestbkrd->bitswap1->lz4"):
    buf = torch.empty(sq.max_compressed_length(p, raw), dtype=torch.uint8, device="cuda")
    ms_e, blob = timed(lambda: sq.encode_device(p, vol, out=buf), 3)
    ms_d, _ = timed(lambda: sq.decode_device(blob, out), 3)
    print(f"{p}: encode {ms_e:.2f} ms = {raw / ms_e / 1e6:.0f} GB/s, decode {ms_d:.2f} ms = {raw / ms_d / 1e6:.0f} GB/s, pair {raw / (ms_e + ms_d) / 1e6:.0f} voxel-GB/s, ratio {raw / blob.numel():.2f}, blocks {sq.last_lz4_stats()}")
    del buf
