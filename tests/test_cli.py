"""sqy command line tool (sqeazy_b200/bin/sqy, a client of the C ABI only): the reference tool's verbs, aliases, flags
and output conventions for the hot path (src/sqy.cpp:182-396, verbs/compress.hpp, decompress.hpp, bench.hpp)."""
import os
import subprocess

import numpy as np
import pytest

from sqeazy_b200.synth import numpy_volume
from tiff_helper import read_tiff, write_tiff

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SQY = os.path.join(ROOT, "sqeazy_b200", "bin", "sqy")


def run(*args, cwd=None):
    return subprocess.run([SQY, *args], capture_output=True, text=True, cwd=cwd, timeout=300)


@pytest.fixture(scope="module", autouse=True)
def built(sq):
    if not os.path.exists(SQY):
        import __graft_entry__

        __graft_entry__.build()
    assert os.path.exists(SQY)


def test_usage_version_and_unknown_verb():
    r = run()
    assert r.returncode == 0 and "usage: sqy" in r.stdout
    for alias in ("compress|enc|encode|comp", "decompress|dec|decode|rec", "ben|bench", "compare|cmp"):
        assert alias in r.stdout
    v = run("--version")
    assert v.returncode == 1 and v.stdout.startswith("sqy 0.7.")      # src/sqy.cpp:318-324 prints and returns 1
    u = run("frobnicate", "x.tif")
    assert u.returncode == 1 and "unable to find matching verb for frobnicate" in u.stderr
    h = run("compress", "-h")
    assert h.returncode == 1 and "--pipeline" in h.stdout and "pipeline builder" in h.stdout
    assert run("compress").returncode == 1                           # no input files


def test_tiff_reader_accepts_what_libtiff_writes(tmp_path):
    """compare needs no GPU: both files go through the TIFF reader — little/big endian, one/many strips, 8/16 bit"""
    vol = numpy_volume((5, 33, 47), "scmos", index=1)
    a, b, c, d = (str(tmp_path / n) for n in ("a.tif", "b.tif", "c.tif", "d.tif"))
    write_tiff(a, vol)
    write_tiff(b, vol, big_endian=True, rows_per_strip=7)
    assert run("compare", a, b).returncode == 0
    other = vol.copy()
    other[3, 20, 11] ^= 1
    write_tiff(c, other, rows_per_strip=4)
    r = run("cmp", a, c)
    assert r.returncode == 1 and "!=" in r.stdout
    write_tiff(d, (vol >> 4).astype(np.uint8), big_endian=True)
    assert run("compare", d, d).returncode == 0
    assert run("compare", a, d).returncode == 1                      # different depth
    bad = tmp_path / "bad.tif"
    bad.write_bytes(b"II*\0garbage")
    assert run("compare", a, str(bad)).returncode == 1


def test_compress_without_gpu_fails_loudly(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = str(tmp_path / "a.tif")
    write_tiff(a, numpy_volume((4, 16, 32), "scmos"))
    r = run("compress", a)
    assert r.returncode == 1 and "errors occurred while processing" in r.stderr
    assert not os.path.exists(str(tmp_path / "a.sqy")) or os.path.getsize(str(tmp_path / "a.sqy")) == 0
    assert run("compress", "-p", "no_such_stage->lz4", a).returncode == 1


@pytest.mark.gpu
@pytest.mark.parametrize("pipeline,dtype", [("bitswap1->lz4", np.uint16), ("rmestbkrd->bitswap1->lz4", np.uint16),
                                            ("quantiser->lz4", np.uint16), ("bitswap1->lz4", np.uint8)])
def test_compress_decompress_roundtrip(sq, cuda, port, tmp_path, pipeline, dtype):
    vol = numpy_volume((12, 64, 96), "scmos", index=2)
    if dtype == np.uint8:
        vol = np.minimum(vol >> 2, 255).astype(np.uint8)
    src = str(tmp_path / "stack.tif")
    write_tiff(src, vol)
    r = run("compress", "-p", pipeline, "-v", src)
    assert r.returncode == 0, r.stderr
    sqy_file = str(tmp_path / "stack.sqy")
    blob = np.fromfile(sqy_file, dtype=np.uint8)
    assert sq.decompressed_shape(blob) == vol.shape and sq.decompressed_sizeof(blob) == vol.dtype.itemsize
    expect = vol
    if pipeline.startswith("rmestbkrd"):
        expect, _ = port.rmestbkrd(vol, sq.host_l2_bytes())
    if pipeline.startswith("quantiser"):
        enc, dec = port.quantiser_luts(port.histogram(vol))
        expect = dec[enc[vol]]
    back = sq.decode(blob) if dtype == np.uint16 else sq.decode_u8(blob)
    assert np.array_equal(back, expect)
    d = run("decompress", sqy_file, "-o", str(tmp_path / "back.tif"))
    assert d.returncode == 0, d.stderr
    assert np.array_equal(read_tiff(str(tmp_path / "back.tif")), expect)
    if expect is vol:
        assert run("compare", src, str(tmp_path / "back.tif")).returncode == 0


@pytest.mark.gpu
def test_output_naming_and_many_files(sq, cuda, tmp_path):
    """verbs/compress.hpp:268-283: a suffix with a period replaces the extension; --output_name is ignored for > 1 file"""
    names = []
    for i in range(3):
        p = str(tmp_path / f"t{i}.tif")
        write_tiff(p, numpy_volume((4, 32, 32), "scmos", index=i))
        names.append(p)
    r = run("enc", "-e", ".sqz", "-o", str(tmp_path / "ignored.sqy"), *names)
    assert "multiple input files detected" in r.stdout
    # .sqz is not a native target: nothing may be written for it
    assert r.returncode == 1 and not os.path.exists(str(tmp_path / "t0.sqz"))
    r = run("comp", *names)
    assert r.returncode == 0 and all(os.path.exists(str(tmp_path / f"t{i}.sqy")) for i in range(3))
    r = run("dec", "-e", "_rec.tif", str(tmp_path / "t1.sqy"), cwd=str(tmp_path))
    assert r.returncode == 0 and os.path.exists(str(tmp_path / "t1_rec.tif"))   # suffix without period: stem + suffix
    assert run("cmp", names[1], str(tmp_path / "t1_rec.tif")).returncode == 0


@pytest.mark.gpu
def test_bench_table(sq, cuda, tmp_path):
    """verbs/bench.hpp:83-131: id,shape,time_mus,final_bytes,ingest_bw_mbps,sizeof_pixel,n_elements,filename,comment"""
    p = str(tmp_path / "b.tif")
    write_tiff(p, numpy_volume((8, 64, 64), "scmos"))
    r = run("bench", "-c", "-r", "3", "--comment", "hello", "-p", "bitswap1->lz4", p)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[0] == "id,shape,time_mus,final_bytes,ingest_bw_mbps,sizeof_pixel,n_elements,filename,comment"
    assert len(lines) == 4
    f = lines[1].split(",")
    assert f[0] == "0" and f[1] == "8x64x64" and f[5] == "2" and f[6] == str(8 * 64 * 64) and f[8] == '"hello"' and float(f[4]) > 0
    r2 = run("ben", "--as-csv", "--noheader", "-r", "1", p)
    assert r2.returncode == 0 and len(r2.stdout.strip().splitlines()) == 1 and "bitswap1->lz4|1threads|" in r2.stdout
