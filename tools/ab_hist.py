"""A/B of the histogram kernel between library builds (development tool). usage: ab_hist.py libA.so libB.so ..."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
from ctypes import c_long, c_void_p
torch.cuda.set_device(0); sq.set_device(0)
libs = [(os.path.basename(p), ctypes.CDLL(os.path.abspath(p))) for p in sys.argv[1:]]
st = c_void_p(torch.cuda.current_stream().cuda_stream)
for kind in ("scmos", "zeros", "random"):
    vol = torch_volume((512, 2048, 2048), kind)
    ref = torch.zeros(65536, dtype=torch.int64, device="cuda")
    flat = vol.view(-1)
    for lo in range(0, flat.numel(), 1 << 28):
        ref += torch.bincount(flat[lo: lo + (1 << 28)].to(torch.int32) & 0xFFFF, minlength=65536)
    for rnd in range(2):
        for lname, L in libs:
            hist = torch.zeros(65536, dtype=torch.int32, device="cuda")
            def run():
                assert L.sqyx_histogram_UI16(c_void_p(vol.data_ptr()), c_long(vol.numel()), c_void_p(hist.data_ptr()), st) == 0
            run(); torch.cuda.synchronize()
            ok = torch.equal(ref & 0xFFFFFFFF, hist.to(torch.int64) & 0xFFFFFFFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"{kind:8s} {lname:24s} {ms:7.3f} ms  {vol.numel() * 2 / ms / 1e6:7.1f} GB/s  exact {ok}", flush=True)
