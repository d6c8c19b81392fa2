#!/usr/bin/env python
"""Generates tests/golden/isp_configs.json (ISP configurations in the reference's JSON schema, res/config/isp/README.txt: written
here, not copied) and tests/golden/isp_golden.json (CRC-32 of the output the REFERENCE'S OWN CameraIsp -- oracle/_ref/libref_isp.so,
built by oracle/ref_build.sh from /root/reference -- produces for every case of tests/test_gpu_isp.py). Run in the container that
holds /root/reference:  python tests/golden/make_isp_golden.py"""
import json
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def config(**over):
    c = {
        "serial": 0, "name": "synthetic 16-bit sensor", "bitsPerPixel": 16,
        "compandingLut": [[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [1.0, 1.0, 0.0]],
        "blackLevel": [1310.0, 1290.0, 1275.0],
        "vignetteRollOffH": [[1.12, 1.10, 1.08], [1.02, 1.01, 1.0], [1.0, 1.0, 1.0], [1.01, 1.02, 1.0], [1.09, 1.11, 1.13]],
        "vignetteRollOffV": [[1.15, 1.12, 1.1], [1.0, 1.0, 1.0], [1.0, 1.0, 1.0], [1.07, 1.09, 1.11]],
        "whiteBalanceGain": [1.17, 1.0, 1.58],
        "stuckPixelThreshold": 5, "stuckPixelDarknessThreshold": 0.11, "stuckPixelRadius": 0,
        "ccm": [[1.04, -0.06, 0.02], [0.15, 1.12, -0.27], [-0.14, -0.09, 1.23]],
        "sharpening": [0.5, 0.45, 0.55], "saturation": 1.2, "contrast": 1.0,
        "lowKeyBoost": [-0.2, -0.15, -0.2], "highKeyBoost": [0.2, 0.2, 0.15], "gamma": [0.4545, 0.46, 0.45], "bayerPattern": "GBRG",
    }
    c.update(over)
    return json.dumps({"CameraIsp": c}, indent=2)


def configs():
    return {
        "sensor_a": config(),
        "sensor_b": config(name="second lens", blackLevel=[1270.0, 1270.0, 1270.0], whiteBalanceGain=[1.3, 1.0, 1.4], contrast=1.15,
                                 saturation=1.05, sharpeningSupport=0.01, vignetteRollOffH=[[1.3, 1.3, 1.3], [1.0, 1.0, 1.0], [1.25, 1.2, 1.3]],
                                 stuckPixelRadius=1),   # radius > 0 with threshold 5: the reference's stuck-pixel loop never fires
        "passthrough": json.dumps({"CameraIsp": {"bitsPerPixel": 16}}),
        "variant_rggb": config(bayerPattern="RGGB"),
        "variant_grbg": config(bayerPattern="GRBG", contrast=0.9),
        "variant_bggr_clamp": config(bayerPattern="BGGR", clampMin=[0.02, 0.01, 0.03], clampMax=[0.9, 0.95, 0.85]),
        "variant_nosharpen": config(sharpening=[0.5, 0.0, 0.5]),
        "variant_lowcore": config(noiseCore=0.00002, sharpening=[1.5, 1.2, 0.8]),   # small coring constant: expf arguments of every size
    }


def main():
    import ref_bind as rb
    from surround360_b200 import synth
    import test_gpu_isp as t
    cfgs = configs()
    json.dump(cfgs, open(os.path.join(HERE, "isp_configs.json"), "w"), indent=1)
    t.CONFIGS = cfgs
    golden = {}
    for name, cfg, (h, w), dt, dm, rs, bpp, tone, off, sw in t.CASES:
        raw = synth.bayer_frame(h, w, dt, seed=zlib.crc32(name.encode()) & 0xffff)
        golden[name] = format(zlib.crc32(t.run_ref(cfg, raw, dm, rs, bpp, tone, off, sw).tobytes()), "08x")
        print(name, golden[name])
    raw = synth.bayer_frame(2048, 2048, np.uint16, seed=7)
    golden["a_edge_16_2048"] = format(zlib.crc32(t.run_ref("sensor_a", raw, 2, 1, 16).tobytes()), "08x")
    json.dump(golden, open(os.path.join(HERE, "isp_golden.json"), "w"), indent=1)
    print(golden)


if __name__ == "__main__":
    main()
