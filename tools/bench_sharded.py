"""One stack through SQY_PipelineEncode_UI16 + SQY_Decode_UI16 (host buffers) on 1, 2, 4, ... GPUs of one box, sharded
inside the library (csrc/sharded.inl). usage: bench_sharded.py [cfg2|cfg3|cfg1] [pinned|pageable] [steps]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from sqeazy_b200.synth import torch_volume

WORK = {"cfg1": ((256, 512, 512), "bitswap1->lz4"), "cfg2": ((512, 2048, 2048), "rmestbkrd->bitswap1->lz4"),
        "cfg3": ((1024, 2048, 2048), "quantiser->lz4")}
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
kind = sys.argv[2] if len(sys.argv) > 2 else "pinned"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
shape, pipeline = WORK[name]
ngpu = torch.cuda.device_count()
torch.cuda.set_device(0)
h_vol = torch.empty(shape, dtype=torch.int16)
h_out = torch.empty(shape, dtype=torch.int16)
if kind == "pinned":
    h_vol, h_out = h_vol.pin_memory(), h_out.pin_memory()
# the stack is generated slab-wise on GPU 0 (8 GiB stacks do not need a second device copy)
zs = max(1, (1 << 28) // (shape[1] * shape[2]))
for z in range(0, shape[0], zs):
    part = torch_volume((min(zs, shape[0] - z), shape[1], shape[2]), "scmos", index=1000 + z)
    h_vol[z:z + part.shape[0]].copy_(part)
    del part
torch.cuda.empty_cache()
vol = h_vol.numpy().view(np.uint16)
out = h_out.numpy().view(np.uint16).reshape(-1)
cap = sq.max_compressed_length(pipeline, vol.nbytes)
h_blob = torch.empty(cap, dtype=torch.uint8)
if kind == "pinned":
    h_blob = h_blob.pin_memory()
blob_buf = h_blob.numpy()
ref_sum = None
sets = [list(range(g)) for g in (1, 2, 4, 8) if g <= ngpu]
for devs in sets:
    sq.set_devices(devs) if len(devs) > 1 else sq.set_device(0)
    b = sq.encode(pipeline, vol, nthreads=16, out=blob_buf)
    sq.decode(b, nthreads=16, out=out)           # warm-up: arenas, NCCL communicators
    te = td = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        b = sq.encode(pipeline, vol, nthreads=16, out=blob_buf)
        t1 = time.perf_counter()
        sq.decode(b, nthreads=16, out=out)
        t2 = time.perf_counter()
        te += t1 - t0
        td += t2 - t1
    info = sq.last_shard_info()
    s = int(out[:: 4097].astype(np.uint64).sum())
    if ref_sum is None:
        ref_sum = s
        if "quantiser" not in pipeline and "bkrd" not in pipeline and "background" not in pipeline:
            assert np.array_equal(out, vol.reshape(-1))
    print(json.dumps({"workload": name, "buffers": kind, "gpus": len(devs), "shard_info": info, "encode_gbs": vol.nbytes * steps / te / 1e9,
                      "decode_gbs": vol.nbytes * steps / td / 1e9, "e2e_voxel_gbs": vol.nbytes * steps / (te + td) / 1e9,
                      "encode_ms": te / steps * 1e3, "decode_ms": td / steps * 1e3, "blob_bytes": int(b.size), "ratio": vol.nbytes / b.size,
                      "same_voxels_as_1gpu": s == ref_sum}), flush=True)
    for d in devs:
        sq.set_device(d)
        sq.release_scratch()
sq.set_device(0)
