"""Times the LZ4 stage alone (CUDA events, device-resident) for both block decoders on bit planes / quantiser codes of
several sizes. usage: bench_decoders.py [ZxYxX ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

torch.cuda.set_device(0); sq.set_device(0)
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(256, 512, 512), (128, 1024, 1024), (512, 2048, 2048)]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps


for shape in shapes:
    vol = torch_volume(shape, "scmos")
    for kind in ("rmest_planes", "planes", "codes"):
        if kind == "codes":
            hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
            sq.histogram_device(vol, hist)
            torch.cuda.synchronize()
            enc, dec = sq.quantiser_luts(hist.cpu().numpy().view(np.uint32))
            data = torch.empty(vol.numel(), dtype=torch.uint8, device="cuda")
            sq.lut_apply_device(vol, data, enc)
        else:
            thr = 0
            if kind == "rmest_planes":
                _, thr = sq.estimate_background_device(vol)
            data = torch.empty_like(vol)
            sq.bitswap_encode_device(1, vol.view(-1), data.view(-1), threshold=thr)
        nbytes = data.numel() * data.element_size()
        payload = sq.lz4_encode_device(data)
        out = torch.empty_like(data)
        line = f"{'x'.join(map(str, shape))} {kind}: {nbytes >> 20} MiB ratio {nbytes / payload.numel():.2f} {sq.last_lz4_stats()}"
        for label, lane_max in (("warp", 0), ("lanes", 65536)):
            prev = sq.set_lz4_lane_max(lane_max)
            ms = timed(lambda: sq.lz4_decode_device(payload, out))
            sq.set_lz4_lane_max(prev)
            ok = torch.equal(out, data)
            line += f" | {label} {ms:.3f} ms {nbytes / ms / 1e6:.0f} GB/s ok={ok}"
        print(line, flush=True)
        del data, out, payload
    del vol
