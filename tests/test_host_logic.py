"""Host side of the drop-in boundary through the C ABI (no GPU work): symbols, pipeline validation, bounds, header
queries, LUT construction. Modelled on the reference's tests/test_pipeline_interface.cpp and
src/java/.../SqeazyLibraryTests.java."""
import ctypes
import re

import numpy as np
import pytest

from oracle import oracle as orc


def test_library_exports_every_declared_symbol(sq):
    L = sq.lib()
    import os

    declared = set()
    for hdr in ("sqeazy.h", "sqeazy_b200.h"):
        text = open(os.path.join(os.path.dirname(sq.LIB_PATH), "..", "include", hdr)).read()
        declared |= set(re.findall(r"\b(SQY_\w+|sqyx_\w+)\s*\(", text))
    assert declared == set(sq.SQY_SYMBOLS) | set(sq.SQYX_SYMBOLS)
    text = open(os.path.join(os.path.dirname(sq.LIB_PATH), "..", "include", "sqeazy_h5_filter.h")).read()
    h5 = set(re.findall(r"\b(H5Z_filter_sqy|H5PLget_\w+)\s*\(", text))
    assert h5 == set(sq.H5_SYMBOLS)
    for name in declared | h5:
        assert getattr(L, name) is not None


def test_version_triple(sq):
    v = sq.version_triple()
    assert v[1] >= 3  # SqeazyLibraryTests.java:24


@pytest.mark.parametrize("p,ok", [
    ("bitswap1->lz4", True), ("", False), ("bswap1_lz4", False), ("bitswap1->lz4!!", False), ("lz4", True),
    ("rmestbkrd->bitswap1->lz4", True), ("quantiser->lz4", True), ("quantiser", True), ("bitswap1", True),
    ("remove_background(threshold=110)->bitswap1->lz4", True), ("pass_through", True),
    ("lz4(accel=1,blocksize_kb=256,framestep_kb=256,n_chunks_of_input=0)", True),
    ("bitswap1(num_bits_per_plane=1)->lz4", True), ("lz4->lz4", False), ("diff->lz4", False), ("bitswap1 ->lz4", False),
    ("bitshuffle->lz4", True), ("rmestbkrd->bitshuffle->lz4", True), ("bitshuffle(block_size=512)->lz4", True), ("bitshuffle", True),
])
def test_pipeline_possible(sq, p, ok):
    """tests/test_pipeline_interface.cpp:28-61 ; names outside the accelerated stages are refused (documented)"""
    assert sq.pipeline_possible(p) is ok
    assert sq.pipeline_possible(p, 4) is False
    if ok or p in ("", "bswap1_lz4", "bitswap1->lz4!!", "bitswap1 ->lz4"):
        assert orc.can_be_built_from(p) is ok  # same verdict as the restated reference rule


def test_aliases_are_extensions(sq):
    for p in ("rmbkrd(threshold=3)->bitswap4->lz4", "bitswap2", "bitswap8->lz4"):
        assert sq.pipeline_possible(p)


def test_max_compressed_length(sq):
    """tests/test_pipeline_interface.cpp:245-272: bound > raw bytes; >= the reference's own formula (lz4.hpp:166-188)"""
    raw = 8 * 8 * 8 * 2
    assert sq.max_compressed_length("bitswap1->lz4", raw) > raw
    assert sq.max_compressed_length_3d("bitswap1->lz4", (8, 8, 8)) == sq.max_compressed_length("bitswap1->lz4", raw)
    n = 1 << 27
    assert sq.max_compressed_length("bitswap1->lz4", n) >= 512 * 262171
    assert sq.max_compressed_length("bitswap1->lz4", n) >= sq.lz4_bound(n)
    with pytest.raises(sq.SqeazyError):
        sq.max_compressed_length("nope", 10)
    # >= 2^31 voxels: 64-bit arithmetic (the reference overflows an int here)
    assert sq.max_compressed_length_3d("rmestbkrd->bitswap1->lz4", (512, 2048, 2048)) > 4 * (1 << 30)


def test_lz4_bound_matches_reference_formula(sq, golden):
    """262171 bytes per 256 KiB chunk (LZ4F_compressBound(262144)=262152 + 19) is inside our bound's stage term"""
    assert int(golden["lz4_max_encoded_size_256KiB_1t"][0]) == 262152 + 19
    assert sq.max_compressed_length("lz4", 1 << 18) >= 262171
    assert int(golden["lz4_max_encoded_size_1MiB_1t"][0]) == 4 * 262171
    assert sq.max_compressed_length("lz4", 1 << 20) >= 4 * 262171


def test_header_queries_on_reference_style_blob(sq):
    """tests/test_pipeline_interface.cpp:274-386 on a header packed by the oracle's restatement of header::pack"""
    name = "bitswap1(num_bits_per_plane=1)->lz4(accel=1,blocksize_kb=256,framestep_kb=256,n_chunks_of_input=0)"
    h = orc.pack_header([128, 1024, 256], name, 777)
    blob = np.frombuffer(h.encode() + b"\x01" * 777, dtype=np.uint8).copy()
    assert sq.header_size(blob) == len(h)
    assert sq.decompressed_shape(blob) == (128, 1024, 256)
    assert sq.decompressed_length(blob) == 128 * 1024 * 256 * 2
    assert sq.decompressed_sizeof(blob) == 2


def test_header_escapes_and_verbatim(sq):
    lut = orc.lut_to_verbatim(np.arange(256, dtype=np.uint16) * 255)
    name = "quantiser(decode_lut_string=%s)->lz4(accel=1,blocksize_kb=256,framestep_kb=256,n_chunks_of_input=0)" % lut
    assert "/" in lut
    h = orc.pack_header([2, 3, 4], name, 5)
    assert "\\/" in h
    blob = np.frombuffer(h.encode() + b"12345", dtype=np.uint8).copy()
    assert sq.header_size(blob) == len(h) and sq.decompressed_shape(blob) == (2, 3, 4)


def test_no_header_is_reported_as_empty(sq):
    blob = np.frombuffer(b"this is not a sqeazy blob at all", dtype=np.uint8).copy()
    assert sq.header_size(blob) == 0
    assert sq.decompressed_shape(blob) == ()


def test_quantiser_luts_host(sq, golden):
    for name in ("q_small", "q_big", "q_ramp"):
        hist = np.zeros(65536, dtype=np.uint32)
        hist[golden[name + "_hist_nonzero_idx"]] = golden[name + "_hist_nonzero_val"]
        enc, dec = sq.quantiser_luts(hist)
        assert np.array_equal(enc, golden[name + "_enc"]) and np.array_equal(dec, golden[name + "_dec"])
    enc, dec = sq.quantiser_luts(np.zeros(65536, dtype=np.uint32))
    assert not enc.any() and not dec.any()


def test_host_l2_probe_matches_compass(sq, ref):
    assert sq.host_l2_bytes() == ref.l2_cache_bytes()


def test_hdf5_entry_points_fail_cleanly(sq):
    L = sq.lib()
    assert L.SQY_h5_query_ndims(b"a.h5", b"d", None) == 1


@pytest.mark.parametrize("p,ok", [
    ("bitswap1->lz4", True), ("lz4", True), ("bitswap1", True), ("bitswap4->lz4", True), ("pass_through", True), ("pass_through->lz4", True),
    ("remove_background(threshold=7)->bitswap1->lz4", True), ("rmbkrd(threshold=7)->lz4", True),
    # stages without uint8 kernels here: refused (documented), like every other unaccelerated stage name
    ("rmestbkrd->lz4", False), ("quantiser->lz4", False), ("bitswap8->lz4", False), ("", False), ("lz4->lz4", False),
    ("bitshuffle->lz4", True), ("bitshuffle(block_size=64)", True),
])
def test_pipeline_possible_uint8(sq, p, ok):
    """dypeline<uint8_t>::can_be_built_from, src/sqeazy.cpp:243-268"""
    assert bool(sq.lib().SQY_Pipeline_Possible_UI8(p.encode())) is ok
    assert sq.pipeline_possible(p, 1) is ok


def test_max_compressed_length_uint8(sq):
    """src/sqeazy.cpp:144-163, 209-231: sizes in bytes of uint8 voxels"""
    raw = 8 * 8 * 8
    assert sq.max_compressed_length_u8("bitswap1->lz4", raw) > raw
    assert sq.max_compressed_length_3d_u8("bitswap1->lz4", (8, 8, 8)) == sq.max_compressed_length_u8("bitswap1->lz4", raw)
    assert sq.max_compressed_length_u8("lz4", 1 << 27) >= sq.lz4_bound(1 << 27)
    with pytest.raises(sq.SqeazyError):
        sq.max_compressed_length_u8("quantiser->lz4", 10)


def test_compute_without_gpu_fails_loudly(sq):
    """no CPU fallback: on a box without a CUDA device the encode entry point returns 1"""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(sq.SqeazyError):
        sq.encode("bitswap1->lz4", np.zeros((8, 8, 8), dtype=np.uint16))
