"""HDF5 filter plugin entry points (reference: inc/sqeazy_h5_filter.hpp:28-226, hdf5_utils.hpp:705-740): what
HDF5 asks a plugin for, the filter called with HDF5's conventions (malloc()ed chunk that the filter replaces, header text in
cd_values). No libhdf5 in this image: the tests play HDF5's part through ctypes."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as orc
from sqeazy_b200.synth import numpy_volume


def test_plugin_info_is_what_hdf5_expects(sq):
    L = sq.lib()
    assert L.H5PLget_plugin_type() == 0                              # H5PL_TYPE_FILTER
    info = sq.h5_plugin_info()
    assert info.version == 1 and info.id == 0o1307 == 711            # sqeazy_h5_filter.hpp:211: `01307` is octal
    assert info.encoder_present == 1 and info.decoder_present == 1
    assert info.name == b"HDF5 sqy filter; see https://github.org/sqeazy/sqeazy"
    assert info.can_apply is None and info.set_local is None
    assert info.filter == ctypes.cast(L.H5Z_filter_sqy, ctypes.c_void_p).value


def test_filter_passes_an_encoded_chunk_through(sq):
    """sqeazy_h5_filter.hpp:118-132: a chunk that already starts with a sqeazy header is stored as header + encoded bytes"""
    payload = np.arange(1000, dtype=np.uint8)
    hdr = orc.pack_header((4, 5, 25), "bitswap1(num_bits_per_plane=1)->lz4(accel=1,blocksize_kb=256,framestep_kb=256,n_chunks_of_input=0)",
                          payload.size).encode("latin-1")
    chunk = np.concatenate([np.frombuffer(hdr, dtype=np.uint8), payload, np.zeros(77, dtype=np.uint8)])   # slack behind the blob
    out = sq.h5_filter(chunk, hdr)
    assert out is not None and out.size == len(hdr) + payload.size
    assert np.array_equal(out, chunk[: out.size])
    # blobs of this library: blanks in front of the header (a slot of 256-byte multiples), created from a shorter header
    padded = np.concatenate([np.full(700, 0x20, dtype=np.uint8), chunk])
    cd_short = orc.pack_header((4, 5, 25), "bitswap1->lz4", 2000).encode("latin-1")
    out = sq.h5_filter(padded, cd_short)
    assert out is not None and np.array_equal(out, padded[: 700 + len(hdr) + payload.size])
    short = chunk[: len(hdr) + 10]                                   # header promises more bytes than the chunk has
    assert sq.h5_filter(short, hdr) is None


def test_filter_failures_return_zero_and_leave_the_chunk(sq):
    vol = np.zeros((4, 8, 8), dtype=np.uint16)
    good = orc.pack_header(vol.shape, "bitswap1->lz4", vol.nbytes).encode("latin-1")
    assert sq.h5_filter(vol, b"") is None                            # no header in cd_values
    assert sq.h5_filter(vol, b"not a header at all.") is None
    assert sq.h5_filter(vol, orc.pack_header(vol.shape, "no_such_stage->lz4", vol.nbytes).encode()) is None
    assert sq.h5_filter(vol, orc.pack_header((4, 8, 9), "bitswap1->lz4", vol.nbytes).encode()) is None     # shape != chunk bytes
    assert sq.h5_filter(vol, orc.pack_header(vol.shape, "bitswap1->lz4", vol.nbytes, raw_type="float").encode()) is None
    assert sq.h5_filter(vol, good, reverse=True) is None             # reading a chunk that is no blob
    assert sq.h5_filter(np.zeros(0, dtype=np.uint8), good, reverse=True) is None


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline", ["bitswap1->lz4", "rmestbkrd->bitswap1->lz4", "quantiser->lz4", "lz4", "diff3x3x1->bitswap1->lz4"])
def test_filter_write_then_read_uint16(sq, cuda, port, pipeline):
    vol = numpy_volume((12, 96, 160), "scmos", index=21)
    cd = orc.pack_header(vol.shape, pipeline, vol.nbytes).encode("latin-1")
    blob = sq.h5_filter(vol, cd)
    assert blob is not None and blob.size <= sq.max_compressed_length(pipeline, vol.nbytes)
    hdr = orc.unpack_header(blob.tobytes())
    assert tuple(hdr["shape"]) == vol.shape and hdr["raw_type"] == "uint16" and hdr["size"] + hdr["bytes"] == blob.size
    want = sq.decode(sq.encode(pipeline, vol)).reshape(vol.shape)    # what the C API gives for the same stack (lossy stages included)
    assert np.array_equal(sq.decode(blob).reshape(vol.shape), want)
    # HDF5 hands the stored chunk back in a buffer that may be larger than the blob
    stored = np.concatenate([blob, np.zeros(100, dtype=np.uint8)])
    back = sq.h5_filter(stored, cd, reverse=True)
    assert back is not None and back.size == vol.nbytes
    assert np.array_equal(back.view(np.uint16).reshape(vol.shape), want)
    if pipeline in ("bitswap1->lz4", "lz4", "diff3x3x1->bitswap1->lz4"):
        assert np.array_equal(want, vol)
    # written twice: the second pass sees the header and stores the chunk as it is
    again = sq.h5_filter(blob, cd)
    assert again is not None and np.array_equal(again, blob)


@pytest.mark.gpu
def test_filter_write_then_read_uint8(sq, cuda):
    rng = np.random.default_rng(5)
    vol = np.clip(np.rint(30 + 3 * rng.standard_normal((6, 64, 128))), 0, 255).astype(np.uint8)
    cd = orc.pack_header(vol.shape, "bitswap1->lz4", vol.nbytes, raw_type="uint8", sizeof_raw=1).encode("latin-1")
    blob = sq.h5_filter(vol, cd)
    assert blob is not None and blob.size < vol.nbytes
    assert orc.unpack_header(blob.tobytes())["raw_type"] == "uint8"
    back = sq.h5_filter(blob, cd, reverse=True)
    assert back is not None and np.array_equal(back.reshape(vol.shape), vol)
    # a stage without a uint8 kernel: refused, the chunk stays
    bad = orc.pack_header(vol.shape, "rmestbkrd->lz4", vol.nbytes, raw_type="uint8", sizeof_raw=1).encode("latin-1")
    assert sq.h5_filter(vol, bad) is None
