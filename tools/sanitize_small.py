"""Small end-to-end run for compute-sanitizer: every kernel on ragged shapes, own and foreign (golden) LZ4 frames."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sqeazy_b200 as sq
from sqeazy_b200.synth import numpy_volume

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "golden_v1.npz"))
for shape, preset in (((7, 33, 47), "ref"), ((16, 64, 128), "scmos"), ((3, 40, 128), "random"), ((4, 32, 128), "zeros")):
    vol = numpy_volume(shape, preset, index=1)
    for p in ("bitswap1->lz4", "rmestbkrd->bitswap1->lz4", "quantiser->lz4", "remove_background(threshold=110)->bitswap4->lz4", "lz4", "bitswap2"):
        blob = sq.encode(p, vol)
        back = sq.decode(blob)
        assert back.shape == vol.shape
import torch
torch.cuda.set_device(0)
for key, src in (("lz4_serial", "lz4_vol"), ("lz4_parallel", "lz4_vol"), ("lz4_linked", "lz4_linked_in")):
    pay = torch.from_numpy(g[key]).cuda()
    out = torch.zeros(g[src].size, dtype=torch.int16, device="cuda")
    assert sq.lz4_decode_device(pay, out) == g[src].nbytes
print("sanitize run ok")
