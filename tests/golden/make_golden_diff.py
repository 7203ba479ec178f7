"""Golden vectors of the reference's diff_scheme (encoders/diff_scheme_impl.hpp:78-199, name "diff3x3x1"), made by the
reference's own code through oracle/_ref (only where /root/reference exists):  python tests/golden/make_golden_diff.py
Cubes, stacks with Z < X (the x range follows the Z extent), stacks with Z > X (runs spill into the next row), uint8."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

CASES = [("u16_cube", (8, 8, 8), np.uint16, 65536), ("u16_flat", (5, 7, 9), np.uint16, 400), ("u16_spill", (12, 9, 8), np.uint16, 65536),
         ("u16_vec", (9, 16, 24), np.uint16, 3000), ("u16_tall", (17, 9, 8), np.uint16, 65536), ("u8_cube", (8, 8, 8), np.uint8, 256),
         ("u8_spill", (20, 5, 10), np.uint8, 256), ("u8_vec", (6, 16, 32), np.uint8, 60)]


def main():
    ref = oracle.ref()
    assert ref.available, "oracle/_ref not built"
    out = {"name": np.frombuffer(ref.diff_name().encode(), dtype=np.uint8)}
    for i, (name, shape, dt, hi) in enumerate(CASES):
        a = np.random.default_rng(100 + i).integers(0, hi, size=shape).astype(dt)
        enc = ref.diff(a)
        assert np.array_equal(ref.diff(enc, decode=True), a)
        out[name + "_in"] = a
        out[name + "_enc"] = enc
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_diff_v1.npz"), **out)


if __name__ == "__main__":
    main()
