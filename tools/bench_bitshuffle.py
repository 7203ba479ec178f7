"""bitshuffle kernels alone (4 B/voxel algorithmic: 2 read + 2 written) and the reference's own full-pipeline benchmark
pipelines (bench/benchmark_full_pipeline_impl.cpp:11-12) beside the bitswap1 ones, device-resident, CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

torch.cuda.set_device(0); sq.set_device(0)
shape = (512, 2048, 2048) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
vol = torch_volume(shape, "scmos")
raw = vol.numel() * 2
a, b = vol.view(-1), torch.empty(vol.numel(), dtype=torch.int16, device="cuda")


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps, r


for bs in (0, 512, 1000):
    ms_e, _ = timed(lambda: sq.bitshuffle_encode_device(a, b, bs))
    ms_d, _ = timed(lambda: sq.bitshuffle_decode_device(b, a, bs))
    print(f"bitshuffle block_size={bs}: encode {ms_e:.3f} ms = {2 * raw / ms_e / 1e6:.0f} GB/s, decode {ms_d:.3f} ms = {2 * raw / ms_d / 1e6:.0f} GB/s")
sq.bitshuffle_decode_device(b, a, 1000)   # (a was overwritten by the decodes: restore the stack)
vol = torch_volume(shape, "scmos")
out = torch.empty_like(vol)
for p in ("bitshuffle->lz4", "bitswap1->lz4", "rmestbkrd->bitshuffle->lz4", "rmestbkrd->bitswap1->lz4"):
    buf = torch.empty(sq.max_compressed_length(p, raw), dtype=torch.uint8, device="cuda")
    ms_e, blob = timed(lambda: sq.encode_device(p, vol, out=buf), 3)
    ms_d, _ = timed(lambda: sq.decode_device(blob, out), 3)
    print(f"{p}: encode {ms_e:.2f} ms = {raw / ms_e / 1e6:.0f} GB/s, decode {ms_d:.2f} ms = {raw / ms_d / 1e6:.0f} GB/s, pair {raw / (ms_e + ms_d) / 1e6:.0f} voxel-GB/s, ratio {raw / blob.numel():.2f}, blocks {sq.last_lz4_stats()}")
