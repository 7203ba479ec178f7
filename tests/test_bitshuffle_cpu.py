"""bitshuffle head filter without a GPU: the oracle's restatement of the third-party bitshuffle library
(oracle/sqy_oracle.c: orc_bitshuffle — PARITY UNPINNED: the library is downloaded by the reference's cmake and is in neither
tree; the reference's tests hold round trips only) against an independent numpy formulation, the reference's own test
sizes, and the CPU replay of the fast kernels' thread program against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def numpy_bitshuffle(a, bs=0):
    """row r of a block = bit r of every element (least significant first), element e of the block at byte e // 8, bit e % 8;
    whole blocks of bs elements, one block of the rest rounded down to a multiple of 8, the last < 8 elements verbatim"""
    a = np.ascontiguousarray(a).ravel()
    es, n = a.itemsize, a.size
    if bs == 0:
        bs = max(128, (8192 // es) // 8 * 8)
    out = np.empty(n * es, dtype=np.uint8)

    def block(lo, cnt):
        bits = np.unpackbits(a[lo:lo + cnt].view(np.uint8).reshape(cnt, es), axis=1, bitorder="little")
        out[lo * es:(lo + cnt) * es] = np.packbits(bits.T, axis=1, bitorder="little").reshape(-1)

    for b in range(n // bs):
        block(b * bs, bs)
    pos = (n // bs) * bs
    last = (n - pos) // 8 * 8
    if last:
        block(pos, last)
        pos += last
    out[pos * es:] = a.view(np.uint8)[pos * es:]
    return out.view(a.dtype)


@pytest.mark.parametrize("n", [0, 7, 8, 9, 127, 4096, 4097, 4104, 32003, 32 * 1024, 32 * 1024 + 1, 100003])
@pytest.mark.parametrize("bs", [0, 8, 24, 64, 1000, 4096])
def test_oracle_matches_numpy_formulation(port, n, bs):
    rng = np.random.default_rng(n + bs)
    a = rng.integers(0, 65536, size=n, dtype=np.uint16)
    enc = port.bitshuffle(a, bs)
    assert np.array_equal(enc, numpy_bitshuffle(a, bs))
    assert np.array_equal(port.bitshuffle(enc, bs, decode=True), a)


def test_reference_roundtrip_cases(port):
    """tests/test_bitshuffle_scheme_impl.cpp:21-190: iota of 32 Ki, 32 Ki + 1 and 32003 uint16 items; :200-380 the same for uint8"""
    for n in (32 * 1024, 32 * 1024 + 1, 32003):
        for dt in (np.uint16, np.uint8):
            a = np.arange(n).astype(dt)
            enc = port.bitshuffle(a)
            assert enc.size == a.size                      # max_encoded_size == input size (bitshuffle_scheme_impl.hpp:86-89)
            assert np.array_equal(port.bitshuffle(enc, decode=True), a)
            assert np.array_equal(enc, numpy_bitshuffle(a))


def test_block_size_must_be_a_multiple_of_eight(port):
    with pytest.raises(ValueError):
        port.bitshuffle(np.zeros(64, np.uint16), 12)       # the library's error -81


def test_known_small_block(port):
    """16 elements 0x0001, 0x0002, ... one bit each: row r holds exactly element r's bit"""
    a = (1 << np.arange(16)).astype(np.uint16)
    enc = port.bitshuffle(a, 16).view(np.uint8)            # 16 rows of 2 bytes
    rows = enc.reshape(16, 2)
    for r in range(16):
        want = np.zeros(2, np.uint8)
        want[r // 8] = 1 << (r % 8)
        assert np.array_equal(rows[r], want)


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(ROOT, "oracle", "_build", "libbitshuffle_sim.so")
    src = os.path.join(ROOT, "tests", "helpers", "bitshuffle_sim.cpp")
    inc = os.path.join(ROOT, "sqeazy_b200", "csrc", "device")
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", inc, src, "-o", so])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("bs", [32, 64, 1024, 4096, 8192])
def test_fast_kernel_thread_program_matches_oracle(port, sim, bs):
    """tests/helpers/bitshuffle_sim.cpp replays bitshuffle16_{encode,decode}_fast thread by thread with the kernel's own
    register transpose (csrc/device/bit_transpose16.h), PRMT selectors and addressing"""
    n = bs * 3
    a = np.random.default_rng(bs).integers(0, 65536, size=n, dtype=np.uint16)
    want = port.bitshuffle(a, bs)
    got = np.zeros(n, np.uint16)
    sim.sim_bitshuffle16_encode_fast(a.ctypes.data_as(ctypes.c_void_p), got.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n // 32), ctypes.c_uint32(bs))
    assert np.array_equal(got, want)
    back = np.zeros(n, np.uint16)
    sim.sim_bitshuffle16_decode_fast(want.ctypes.data_as(ctypes.c_void_p), back.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n // 32), ctypes.c_uint32(bs))
    assert np.array_equal(back, a)


@pytest.mark.parametrize("bs", [32, 64, 1024, 8192])
def test_fast_uint8_kernel_thread_program_matches_oracle(port, sim, bs):
    n = bs * 3
    a = np.random.default_rng(bs + 1).integers(0, 256, size=n, dtype=np.uint8)
    want = port.bitshuffle(a, bs)
    got = np.zeros(n, np.uint8)
    sim.sim_bitshuffle8_encode_fast(a.ctypes.data_as(ctypes.c_void_p), got.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n // 32), ctypes.c_uint32(bs))
    assert np.array_equal(got, want)
    back = np.zeros(n, np.uint8)
    sim.sim_bitshuffle8_decode_fast(want.ctypes.data_as(ctypes.c_void_p), back.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(n // 32), ctypes.c_uint32(bs))
    assert np.array_equal(back, a)


@pytest.mark.parametrize("n", [0, 7, 8, 9, 8191, 8192, 8200, 32003, 100003])
@pytest.mark.parametrize("bs", [0, 8, 24, 1000, 8192])
def test_oracle_matches_numpy_formulation_uint8(port, n, bs):
    a = np.random.default_rng(n + bs + 5).integers(0, 256, size=n, dtype=np.uint8)
    enc = port.bitshuffle(a, bs)
    assert np.array_equal(enc, numpy_bitshuffle(a, bs))
    assert np.array_equal(port.bitshuffle(enc, bs, decode=True), a)
