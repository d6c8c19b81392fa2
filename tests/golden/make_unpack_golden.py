#!/usr/bin/env python
"""Generates tests/golden/unpack_golden.json: CRC-32 of what the REFERENCE'S OWN code (oracle/_ref/libref_unpack.so = RawConverter.cpp,
oracle/_ref/libref_isp.so = CameraIsp.h; built by oracle/ref_build.sh from /root/reference) produces for the cases of
tests/test_footage.py and tests/test_gpu_unpack.py. Run in the container that holds /root/reference:
python tests/golden/make_unpack_golden.py"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def crc(a):
    return format(zlib.crc32(np.ascontiguousarray(a).tobytes()), "08x")


def main():
    import ref_bind as rb
    import test_gpu_unpack as t
    golden = {}
    for bits in (8, 12):
        rng = np.random.default_rng(100 + bits)
        packed = rng.integers(0, 256, 96 * 128 * bits // 8, dtype=np.uint8)
        golden[f"convert{bits}_96x128"] = crc(rb.raw_convert(packed, bits, 128, 96))
    for name, bits, (h, w), cfg in t.FOOTAGE_CASES:
        packed = t.footage_frame(name, bits, h, w)
        raw16, img = t.reference_chain(packed, bits, w, h, cfg)
        golden[name + "_raw"] = crc(raw16)
        golden[name + "_bgr16"] = crc(img)
    json.dump(golden, open(os.path.join(HERE, "unpack_golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(golden, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
