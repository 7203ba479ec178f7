"""A/B of the LZ4 encode stage between library builds on the same GPU in the same process (development tool).
usage: ab_lz4.py libA.so libB.so ...   (cfg2 bit planes after rmestbkrd, all-zero buffer, 8-bit quantiser codes)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
from ctypes import c_long, c_void_p

torch.cuda.set_device(0); sq.set_device(0)
shape = (512, 2048, 2048)
vol = torch_volume(shape, "scmos")
_, thr = sq.estimate_background_device(vol)
planes = torch.empty_like(vol)
sq.bitswap_encode_device(1, vol.view(-1), planes.view(-1), threshold=thr)
hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
sq.histogram_device(vol, hist); torch.cuda.synchronize()
enc, dec = sq.quantiser_luts(hist.cpu().numpy().view(np.uint32))
codes = torch.empty(vol.numel(), dtype=torch.uint8, device="cuda")
sq.lut_apply_device(vol, codes, enc)
plain = torch.empty_like(vol)
sq.bitswap_encode_device(1, vol.view(-1), plain.view(-1))
only = os.environ.get("AB_ONLY")
inputs = {"cfg2 planes": planes.view(torch.uint8).view(-1), "zeros": torch.zeros(planes.numel() * 2, dtype=torch.uint8, device="cuda"),
          "quantiser codes": codes, "plain planes": plain.view(torch.uint8).view(-1)}
del vol
libs = [(os.path.basename(p), ctypes.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
outbuf = torch.empty(sq.lz4_bound(planes.numel() * 2), dtype=torch.uint8, device="cuda")
st = c_void_p(torch.cuda.current_stream().cuda_stream)
def checksum(buf, n):
    """position-dependent checksum of buf[:n] (equal between two builds = the same compressed bytes)"""
    w = buf[: n & ~7].view(torch.int64)
    acc = 0
    for a in range(0, w.numel(), 1 << 26):
        c = w[a : a + (1 << 26)]
        acc += int((c * (torch.arange(a, a + c.numel(), device=c.device, dtype=torch.int64) * 2 + 1)).sum().item())
    acc += int(buf[n & ~7 : n].to(torch.int64).sum().item())
    return acc & 0xFFFFFFFFFFFFFFFF
for name, x in inputs.items():
    if only and name not in only.split(","): continue
    for rnd in range(2):
        for lname, L in libs:
            n = c_long(0)
            def run():
                if hasattr(L, "sqyx_lz4_encode_ex"):
                    pitch = 2048 if name == "quantiser codes" else 256
                    if name in ("cfg2 planes", "zeros") and not os.environ.get("AB_NO_HINT"): pitch |= 0x80000000   # what rmestbkrd->bitswap1->lz4 passes
                    rc = L.sqyx_lz4_encode_ex(c_void_p(x.data_ptr()), c_long(x.numel()), c_void_p(outbuf.data_ptr()), c_long(outbuf.numel()), ctypes.byref(n), c_long(pitch), st)
                else:
                    rc = L.sqyx_lz4_encode(c_void_p(x.data_ptr()), c_long(x.numel()), c_void_p(outbuf.data_ptr()), c_long(outbuf.numel()), ctypes.byref(n), st)
                assert rc == 0
            run(); torch.cuda.synchronize()
            ck = checksum(outbuf, n.value)
            if rnd == 0:                # the stream decodes back to the input (decoder of the in-tree library)
                back = torch.empty_like(x)
                assert sq.lz4_decode_device(outbuf[: n.value], back) == x.numel() and torch.equal(back, x), "round trip failed"
                del back
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{name:16s} {lname:24s} {ms:7.3f} ms  {x.numel() / ms / 1e6:7.1f} GB/s  payload {n.value}  bytes {ck:016x}", flush=True)
