"""host-pointer API with PAGEABLE host buffers (what a plain C / Java / CLI caller passes) vs pinned ones"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume
torch.cuda.set_device(0); sq.set_device(0)
shape = (512, 2048, 2048) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
pipeline = "rmestbkrd->bitswap1->lz4"
vol = torch_volume(shape, "scmos")
raw = vol.numel() * 2
cap = sq.max_compressed_length(pipeline, raw)
for kind, nt in (("pageable", 1), ("pageable", 4), ("pageable", 8), ("pageable", 16), ("pinned", 1)):
    if kind == "pinned":
        h_vol = torch.empty(vol.shape, dtype=torch.int16).pin_memory(); h_blob = torch.empty(cap, dtype=torch.uint8).pin_memory(); h_out = torch.empty(vol.shape, dtype=torch.int16).pin_memory()
    else:
        h_vol = torch.empty(vol.shape, dtype=torch.int16); h_blob = torch.empty(cap, dtype=torch.uint8); h_out = torch.empty(vol.shape, dtype=torch.int16)
    h_vol.copy_(vol)
    np_vol = h_vol.numpy().view(np.uint16); np_blob = h_blob.numpy(); np_out = h_out.numpy().view(np.uint16).reshape(-1)
    b = sq.encode(pipeline, np_vol, nthreads=nt, out=np_blob); sq.decode(b, nthreads=nt, out=np_out)
    t0 = time.perf_counter(); b = sq.encode(pipeline, np_vol, nthreads=nt, out=np_blob); t1 = time.perf_counter(); sq.decode(b, nthreads=nt, out=np_out); t2 = time.perf_counter()
    print(f"{kind} nthreads={nt}: encode {t1 - t0:.3f} s = {raw / (t1 - t0) / 1e9:.1f} GB/s, decode {t2 - t1:.3f} s = {raw / (t2 - t1) / 1e9:.1f} GB/s, pair {raw / (t2 - t0) / 1e9:.1f} voxel-GB/s")
    del h_vol, h_blob, h_out
