"""Generates tests/golden/golden_v1.npz from the REFERENCE's own code (oracle/_ref/libsqyref.so, built from
/root/reference by `make oracle_ref`) run in the build container. Inputs are seeded; outputs are what
the reference's stage classes produce. Re-run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from sqeazy_b200.synth import numpy_volume  # noqa: E402

ref = orc.ref()
assert ref.available, "build oracle/_ref first (make oracle_ref)"
out = {}

# bitswap: KAT input 0..15 (tests/test_bitswap_scheme_impl.cpp:286-329) and a seeded ragged buffer
kat = np.arange(16, dtype=np.uint16)
rng = np.random.Generator(np.random.Philox(7))
ragged = rng.integers(0, 65536, size=1000 + 128 * 3 + 5, dtype=np.uint16)
aligned = rng.integers(0, 65536, size=128 * 40, dtype=np.uint16)
out["bitswap_kat_in"] = kat
out["bitswap_ragged_in"] = ragged
out["bitswap_aligned_in"] = aligned
for w in (1, 2, 4, 8):
    out[f"bitswap_kat_w{w}"] = ref.bitswap_encode(w, kat)
    out[f"bitswap_ragged_w{w}"] = ref.bitswap_encode(w, ragged, scalar=True)
    out[f"bitswap_aligned_w{w}"] = ref.bitswap_encode(w, aligned)  # SSE path for w == 1

# background: scmos volume whose frame exceeds the L2 portion rule and one that does not
vol = numpy_volume((12, 96, 128), "scmos", index=1)
out["bg_vol"] = vol
out["bg_l2_bytes"] = np.array([ref.l2_cache_bytes()], dtype=np.int64)
out["bg_supports"] = ref.darkest_face_supports(vol)
out["bg_rmest"] = ref.rmestbkrd(vol)
out["bg_rm110"] = ref.remove_background(vol, 110)
big = numpy_volume((4, 1536, 2048), "ref", index=2)  # frame 3.1 M elements > L2 bytes on the build host (2 MiB)
out["bg_big_seed_shape"] = np.array(big.shape, dtype=np.int64)
out["bg_big_supports"] = ref.darkest_face_supports(big)

# quantiser: lossless (<= 256 levels) and lloyd (> 256 levels) cases
q_small = (rng.integers(0, 200, size=1 << 16) * 3).astype(np.uint16)
q_big = np.clip(rng.normal(3000, 700, size=1 << 18), 0, 65535).astype(np.uint16)
ramp = np.arange(4096, dtype=np.uint16)  # tests/test_quantiser_impl.cpp:862-990 uses a 0..4095 ramp
for name, a in (("q_small", q_small), ("q_big", q_big), ("q_ramp", ramp)):
    hist, enc, dec = ref.quantiser_setup(a)
    out[name + "_in"] = a
    out[name + "_enc"] = enc
    out[name + "_dec"] = dec
    out[name + "_hist_nonzero_idx"] = np.flatnonzero(hist).astype(np.uint16)
    out[name + "_hist_nonzero_val"] = hist[hist != 0]
out["q_big_lutstring"] = np.frombuffer(ref.quantiser_lut_string(out["q_big_dec"]).encode(), dtype=np.uint8)

# lz4: reference-produced payloads (liblz4 1.9.4): serial = one block-linked frame, parallel = a frame per chunk
lz_vol = numpy_volume((6, 256, 512), "scmos", index=3)  # 1.5 MiB -> 6 chunks of 256 KiB
planes = ref.bitswap_encode(1, lz_vol)
out["lz4_vol"] = lz_vol
out["lz4_serial"] = ref.lz4_encode(planes, nthreads=1)
out["lz4_parallel"] = ref.lz4_encode(planes, nthreads=4)
lin = np.tile(rng.integers(0, 65536, size=60000, dtype=np.uint16), 6)[: 3 * 131072]  # cross-block matches in a linked frame
out["lz4_linked_in"] = lin
out["lz4_linked"] = ref.lz4_encode(lin, nthreads=1)
out["lz4_max_encoded_size_1MiB_1t"] = np.array([ref.lz4_max_encoded_size(1 << 20, 1)], dtype=np.int64)
out["lz4_max_encoded_size_256KiB_1t"] = np.array([ref.lz4_max_encoded_size(1 << 18, 1)], dtype=np.int64)
out["lz4_default_config"] = np.frombuffer(ref.lz4_config(b"").encode(), dtype=np.uint8)

path = os.path.join(ROOT, "tests", "golden", "golden_v1.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
