"""cfg4: decode-only of REFERENCE-produced bitswap1->lz4 blobs (oracle/_ref = the reference's stage code + liblz4,
multi-threaded framing: one LZ4 frame per 256 KiB chunk), a batch of B stacks per GPU through
sqyx_decode_batch_device_UI16 vs one stack at a time. usage: bench_cfg4_batch.py [ZxYxX] [B] [serial]
(serial: blobs of the reference's one-thread mode = one block-LINKED frame, the sqy CLI default)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import sqeazy_b200 as sq
from oracle import oracle as orc
from sqeazy_b200.synth import numpy_volume

shape = (256, 2048, 2048) if len(sys.argv) < 2 else tuple(int(v) for v in sys.argv[1].split("x"))
B = 8 if len(sys.argv) < 3 else int(sys.argv[2])
serial = len(sys.argv) > 3 and sys.argv[3] == "serial"
ref = orc.ref()
assert ref.available
torch.cuda.set_device(0); sq.set_device(0)
name = "bitswap1(num_bits_per_plane=1)->lz4(accel=1,blocksize_kb=256,framestep_kb=256,n_chunks_of_input=0)"
distinct = min(B, 2)                      # reference-side encoding is slow: two distinct stacks, repeated
vols, blobs = [], []
for i in range(distinct):
    vol = numpy_volume(shape, "scmos", index=i)
    payload, _ = ref.pipeline_encode_stages(0, vol, 1 if serial else os.cpu_count())
    h = orc.pack_header(vol.shape, name, payload.size, version="0.5.2", headref="4c45a9b")
    vols.append(vol)
    blobs.append(torch.from_numpy(np.concatenate([np.frombuffer(h.encode(), dtype=np.uint8), payload])).cuda())
blobs = [blobs[i % distinct].clone() for i in range(B)]
outs = [torch.empty(shape, dtype=torch.int16, device="cuda") for _ in range(B)]
raw = vols[0].nbytes


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        fn()
    ev1.record(); ev1.synchronize()
    return ev0.elapsed_time(ev1) / reps


ms_one = timed(lambda: [sq.decode_device(b, o) for b, o in zip(blobs, outs)])
ms_batch = timed(lambda: sq.decode_batch_device(blobs, outs))
ok = all(np.array_equal(outs[i].cpu().numpy().view(np.uint16), vols[i % distinct]) for i in range(B))
print(f"cfg4 {shape[2]}x{shape[1]}x{shape[0]} x {B} reference-made blobs, {'one linked frame each' if serial else 'one frame per 256 KiB chunk'} (ratio {raw / blobs[0].numel():.3f}): one at a time {ms_one:.1f} ms = "
      f"{B * raw / ms_one / 1e6:.1f} GB/s; batch {ms_batch:.1f} ms = {B * raw / ms_batch / 1e6:.1f} GB/s; bit-exact {ok}")
