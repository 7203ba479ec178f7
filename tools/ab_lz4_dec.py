"""A/B of the LZ4 DECODE stage between library builds on the same GPU in the same process (development tool): the streams are
made by the library the package loads (sqeazy_b200/libsqeazy.so). usage: ab_lz4_dec.py libA.so libB.so ..."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
from ctypes import c_long, c_void_p

torch.cuda.set_device(0); sq.set_device(0)
inputs = {}
vol = torch_volume((512, 2048, 2048), "scmos")
_, thr = sq.estimate_background_device(vol)
planes = torch.empty_like(vol)
sq.bitswap_encode_device(1, vol.view(-1), planes.view(-1), threshold=thr)
inputs["cfg2 planes"] = planes.view(torch.uint8).view(-1)
plain = torch.empty_like(vol)
sq.bitswap_encode_device(1, vol.view(-1), plain.view(-1))
inputs["plain planes 4 GiB"] = plain.view(torch.uint8).view(-1)
small = torch_volume((256, 512, 512), "scmos")
sp = torch.empty_like(small)
sq.bitswap_encode_device(1, small.view(-1), sp.view(-1))
inputs["cfg1 planes"] = sp.view(torch.uint8).view(-1)
inputs["noise 2 GiB (all stored)"] = torch.randint(0, 256, (1 << 31,), dtype=torch.uint8, device="cuda")
del vol
only = os.environ.get("AB_ONLY")
if only:
    inputs = {k: v for k, v in inputs.items() if any(o in k for o in only.split(","))}
libs = [(os.path.basename(p), ctypes.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
for name, x in inputs.items():
    payload = sq.lz4_encode_device(x, pitch=256 if "cfg1" not in name else 64).clone()
    out = torch.empty_like(x)
    for rnd in range(2):
        for lname, L in libs:
            n = c_long(0)
            def run():
                rc = L.sqyx_lz4_decode(c_void_p(payload.data_ptr()), c_long(payload.numel()), c_void_p(out.data_ptr()), c_long(out.numel()), ctypes.byref(n), st)
                assert rc == 0 and n.value == x.numel()
            run(); torch.cuda.synchronize()
            assert torch.equal(out, x)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{name:20s} {lname:24s} {ms:7.3f} ms  {x.numel() / ms / 1e6:7.1f} GB/s  payload {payload.numel()}", flush=True)
