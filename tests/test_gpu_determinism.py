"""The LZ4 encoder is a pure function of its input (the round-1 hash rounds read and wrote the table without
ordering, so blobs changed from run to run; liblz4, which the reference calls — encoders/lz4_utils.hpp:99-173 — is
deterministic). Hash inserts now happen behind a barrier with atomicMax, lookups see only the sub-blocks in front."""
import numpy as np
import pytest

from sqeazy_b200.synth import numpy_volume
from test_gpu_parity import dev

pytestmark = pytest.mark.gpu


def _inputs(port):
    rng = np.random.default_rng(77)
    vol = numpy_volume((24, 512, 512), "scmos", index=4)
    filt, _ = port.rmestbkrd(vol, 2 << 20)
    text = (rng.integers(0, 6, 3_000_000, dtype=np.uint8) * 37)
    rep = np.tile(rng.integers(0, 256, 5000, dtype=np.uint8), 500)
    rep[rng.integers(0, rep.size, 20000)] ^= 1
    return {
        "bit planes (scmos)": (port.bitswap_encode(1, vol).view(np.uint8), 64),
        "bit planes after rmestbkrd": (port.bitswap_encode(1, filt).view(np.uint8), 64),
        "nibble planes": (port.bitswap_encode(4, filt).view(np.uint8), 256),
        "low-entropy bytes": (text, 0),
        "far repeats with flips": (rep, 0),
        "reference preset": (port.bitswap_encode(1, numpy_volume((12, 256, 512), "ref", index=1)).view(np.uint8), 64),
    }


def test_lz4_stage_bytes_do_not_depend_on_the_run(sq, cuda, port):
    for name, (a, pitch) in _inputs(port).items():
        d = dev(cuda, a)
        first = sq.lz4_encode_device(d, pitch=pitch).cpu().numpy().copy()
        for rep in range(4):
            if rep == 2:   # something else in between: other blocks on the SMs, other scratch contents
                sq.lz4_encode_device(dev(cuda, np.random.default_rng(rep).integers(0, 256, 1 << 20, dtype=np.uint8)))
            again = sq.lz4_encode_device(d, pitch=pitch).cpu().numpy()
            assert again.size == first.size and np.array_equal(again, first), f"{name}: run {rep} differs"
        out = cuda.zeros(a.size, dtype=cuda.uint8, device="cuda")
        assert sq.lz4_decode_device(dev(cuda, first), out) == a.size
        assert np.array_equal(out.cpu().numpy(), a), name


@pytest.mark.parametrize("pipeline", ["bitswap1->lz4", "rmestbkrd->bitswap1->lz4", "quantiser->lz4",
                                      "remove_background(threshold=110)->bitswap4->lz4", "lz4"])
def test_all_routes_write_the_same_blob(sq, cuda, pipeline):
    """device-pointer API, host API (pageable and pinned), every nthreads: one blob"""
    vol = numpy_volume((48, 512, 1024), "scmos", index=6)
    d_vol = cuda.from_numpy(vol.view(np.int16)).cuda()
    want = sq.encode_device(pipeline, d_vol).cpu().numpy().copy()
    for t in (1, 7):
        assert np.array_equal(sq.encode(pipeline, vol, nthreads=t), want), f"host route, nthreads={t}"
    pinned = cuda.from_numpy(vol.view(np.int16)).pin_memory()
    assert np.array_equal(sq.encode(pipeline, pinned.numpy().view(np.uint16), nthreads=4), want)
    assert np.array_equal(sq.encode_device(pipeline, d_vol).cpu().numpy(), want)


def test_pitch_hint_changes_bytes_not_voxels(sq, cuda, port):
    """the row pitch is only a hint to the match finder: any value gives a valid stream"""
    planes = port.bitswap_encode(1, port.rmestbkrd(numpy_volume((16, 512, 512), "scmos", index=8), 2 << 20)[0]).view(np.uint8)
    d = dev(cuda, planes)
    sizes = {}
    for pitch in (0, 32, 64, 96, 4096, 8192, 100, 16384):     # 100 and 16384 do not qualify: same as 0
        p = sq.lz4_encode_device(d, pitch=pitch)
        out = cuda.zeros(planes.size, dtype=cuda.uint8, device="cuda")
        assert sq.lz4_decode_device(p, out) == planes.size
        assert np.array_equal(out.cpu().numpy(), planes), pitch
        sizes[pitch] = int(p.numel())
    assert sizes[100] == sizes[0] and sizes[16384] == sizes[0]
    assert sizes[64] <= sizes[0], sizes     # 512-voxel rows: 64 bytes of a plane


def test_no_noise_hint_changes_no_byte(sq, cuda, port):
    """SQYX_LZ4_HINT_NO_NOISE (include/sqeazy_b200.h) only reorders the work inside a block: with it all sixteen warps
    analyse at once, without it twelve wait for the early-store verdict of the four sampled ones. Noise, sparse planes and
    streams that mix both must come out byte for byte the same either way."""
    NO_NOISE = 0x80000000
    rng = np.random.default_rng(5)
    inputs = dict(_inputs(port))
    inputs["noise"] = (rng.integers(0, 256, 3 << 20, dtype=np.uint8), 0)
    codes = np.clip(rng.normal(40, 3, 4 << 20), 0, 255).astype(np.uint8)      # 8-bit quantiser codes of a noise floor
    inputs["narrow noise"] = (codes, 2048)
    mixed = np.concatenate([inputs["bit planes after rmestbkrd"][0][: 1 << 20], inputs["noise"][0][: (1 << 20) + 777],
                            np.zeros(50_000, np.uint8), codes[: 1 << 19]])
    inputs["mixed"] = (mixed, 64)
    for name, (a, pitch) in inputs.items():
        d = dev(cuda, a)
        plain = sq.lz4_encode_device(d, pitch=pitch).cpu().numpy().copy()
        hinted = sq.lz4_encode_device(d, pitch=pitch | NO_NOISE).cpu().numpy()
        assert hinted.size == plain.size and np.array_equal(hinted, plain), name
