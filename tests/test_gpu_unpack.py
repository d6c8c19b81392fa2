"""GPU tier for the Unpacker's front end: raw-sample conversion, packed footage straight into the ISP, and the Unpacker
tool, against the REFERENCE'S OWN RawConverter + CameraIsp compiled unchanged (oracle/_ref/libref_unpack.so, libref_isp.so), the numpy
port (oracle/ounpack.py) and the golden checksums of tests/golden/unpack_golden.json. Bar: bit-exact."""
import json
import os
import subprocess
import sys
import zlib

import numpy as np
import pytest

import ref_bind as rb
from surround360_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ounpack  # noqa: E402

pytestmark = pytest.mark.gpu
CONFIGS = json.load(open(os.path.join(HERE, "golden", "isp_configs.json")))
GOLDEN_PATH = os.path.join(HERE, "golden", "unpack_golden.json")
GOLDEN = json.load(open(GOLDEN_PATH)) if os.path.exists(GOLDEN_PATH) else {}

FOOTAGE_CASES = [  # (name, bits, (h, w), ISP configuration)
    ("footage12_a", 12, (136, 200), "sensor_a"),
    ("footage12_b_odd_rows", 12, (97, 130), "sensor_b"),
    ("footage8_a", 8, (120, 168), "sensor_a"),
    ("footage12_rggb", 12, (64, 96), "variant_rggb"),
]


def footage_frame(name, bits, h, w):
    """The packed bytes of one synthetic camera frame, serial number stamped into bytes 4..7 like the capture software does."""
    raw = synth.bayer_frame(h, w, np.uint16, seed=zlib.crc32(name.encode()) & 0xffff)
    p = synth.pack12(raw >> 4) if bits == 12 else (raw >> 8).astype(np.uint8).reshape(-1).copy()
    p[4:8] = np.frombuffer(np.uint32(16400123).tobytes(), np.uint8)
    return p


def reference_chain(packed, bits, w, h, cfg):
    """Unpacker.cpp:141-181 with the reference's code: RawConverter, then CameraIsp(json, 16 output bits) with the tone map on."""
    raw16 = rb.raw_convert(packed, bits, w, h)
    return raw16, rb.isp_process(CONFIGS[cfg], raw16, demosaic_filter=2, resize=1, output_bpp=16, tone_curve=True)


def crc(a):
    return format(zlib.crc32(np.ascontiguousarray(a).tobytes()), "08x")


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def unpacker_isp(cfg, bits):
    isp = api.CameraIsp(CONFIGS[cfg], 16)   # Unpacker.cpp:174-178
    isp.setBitsPerPixel(bits)
    isp.enableToneMap()
    return isp


@pytest.mark.parametrize("bits", [8, 12])
def test_raw_convert_matches_port_and_reference(ctx, bits):
    rng = np.random.default_rng(bits)
    conv = api.RawConverter.convert12Frame if bits == 12 else api.RawConverter.convert8Frame
    port = ounpack.convert12_frame if bits == 12 else ounpack.convert8_frame
    for (h, w) in [(2, 2), (5, 6), (33, 64), (96, 128), (2048, 2048)]:   # (5, 6): pixel count not a multiple of 8 -> the pair kernel
        packed = rng.integers(0, 256, h * w * bits // 8, dtype=np.uint8)
        got = conv(ctx, packed, w, h)
        assert np.array_equal(got, port(packed, w, h)), (h, w)
        if rb.unpack_available() and h < 2048:
            assert np.array_equal(got, rb.raw_convert(packed, bits, w, h))
    packed = np.random.default_rng(100 + bits).integers(0, 256, 96 * 128 * bits // 8, dtype=np.uint8)
    assert crc(conv(ctx, packed, 128, 96)) == GOLDEN[f"convert{bits}_96x128"]
    if bits == 12:
        v = np.arange(4096, dtype=np.uint16).reshape(64, 64)
        assert np.array_equal(conv(ctx, synth.pack12(v), 64, 64), v << 4 | v >> 8)
        with pytest.raises(api.VrCamException, match="even width"):
            conv(ctx, np.zeros(64, np.uint8), 5, 4)


@pytest.mark.parametrize("case", FOOTAGE_CASES, ids=[c[0] for c in FOOTAGE_CASES])
def test_footage_through_the_isp(ctx, case):
    name, bits, (h, w), cfg = case
    packed = footage_frame(name, bits, h, w)
    isp = unpacker_isp(cfg, bits)
    img, raw16 = isp.footageImage(ctx, packed, bits, w, h, want_raw=True)
    only = isp.footageImage(ctx, packed, bits, w, h)
    assert np.array_equal(only, img)
    # the fused read equals the two-step path of the product: convert, then the plain 16-bit load
    isp2 = unpacker_isp(cfg, bits)
    assert np.array_equal(isp2.getImage(ctx, raw16), img)
    isp.close(); isp2.close()
    assert crc(raw16) == GOLDEN[name + "_raw"] and crc(img) == GOLDEN[name + "_bgr16"], "differs from the reference build's checksums"
    if rb.unpack_available() and rb.available():
        ref_raw, ref_img = reference_chain(packed, bits, w, h, cfg)
        assert np.array_equal(raw16, ref_raw)
        assert np.array_equal(img, ref_img), f"{(img != ref_img).sum()} of {img.size} samples differ"


def test_footage_frames_on_device(ctx):
    """The n cameras of a frame in shared launches, packed bytes resident on the device; 16-bit output and the renderer's BGR8."""
    torch = pytest.importorskip("torch")
    h, w, bits = 136, 200, 12
    packs = [footage_frame(f"dev{k}", bits, h, w) for k in range(5)]
    isp = unpacker_isp("sensor_a", bits)
    want = [isp.footageImage(ctx, p, bits, w, h) for p in packs]
    d_in = [torch.from_numpy(p).cuda() for p in packs]
    d16 = [torch.zeros((h, w, 3), dtype=torch.int16, device="cuda") for _ in packs]
    d8 = [torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda") for _ in packs]
    isp.footageImagesDev(ctx, [t.data_ptr() for t in d_in], bits, w, h, [t.data_ptr() for t in d16])
    isp.footageImagesDev(ctx, [t.data_ptr() for t in d_in], bits, w, h, [t.data_ptr() for t in d8], high_byte=True)
    ctx.synchronize()
    for k in range(5):
        got16 = d16[k].cpu().numpy().view(np.uint16)
        assert np.array_equal(got16, want[k])
        assert np.array_equal(d8[k].cpu().numpy(), (want[k] >> 8).astype(np.uint8))   # imread(CV_LOAD_IMAGE_COLOR) of the 16-bit PNG
    isp.close()


def read_tiff16(path):
    b = open(path, "rb").read()
    assert b[:4] == b"II*\0"
    ifd = int.from_bytes(b[4:8], "little")
    n = int.from_bytes(b[ifd:ifd + 2], "little")
    tags = {}
    for i in range(n):
        e = b[ifd + 2 + 12 * i: ifd + 14 + 12 * i]
        tag, typ = int.from_bytes(e[0:2], "little"), int.from_bytes(e[2:4], "little")
        tags[tag] = int.from_bytes(e[8:10] if typ == 3 else e[8:12], "little")
    assert tags[258] == 16 and tags[259] == 1 and tags[277] == 1
    w, h, off = tags[256], tags[257], tags[273]
    return np.frombuffer(b[off:off + w * h * 2], np.uint16).reshape(h, w)


def test_unpacker_cli(ctx, tmp_path):
    """surround360_b200/bin/Unpacker with the reference's flags (Unpacker.cpp:36-41) on two footage files: PNGs hold what the library (and
    the reference chain) produce, raw TIFFs the converter's output, a camera without <serial>.json is unpacked to raw only, directories
    are renamed cam0.. in ascending serial order, --start_frame / --frame_count select frames."""
    cv2 = pytest.importorskip("cv2")
    exe = os.environ.get("S360_UNPACKER_EXE", os.path.join(os.path.dirname(HERE), "surround360_b200", "bin", "Unpacker"))   # override: tools/emu
    h, w, bits = 72, 104, 12
    serials = [16400900, 16400100, 16400500]   # file order != serial order
    frames = [[synth.bayer_frame(h, w, np.uint16, seed=10 * f + c) >> 4 for c in range(3)] for f in range(3)]
    bin0, bin1 = str(tmp_path / "0.bin"), str(tmp_path / "1.bin")
    packed0 = synth.write_footage_bin(bin0, frames, bits, serials, file_index=0, file_count=2)
    frames8 = [[(synth.bayer_frame(h, w, np.uint16, seed=77) >> 8).astype(np.uint16)]]
    packed1 = synth.write_footage_bin(bin1, frames8, 8, [16400300], file_index=1, file_count=2)
    isp_dir = tmp_path / "isp"; isp_dir.mkdir()
    cfg_of = {16400900: "sensor_a", 16400100: "sensor_b", 16400300: "sensor_a"}   # 16400500 has no configuration
    for s, c in cfg_of.items():
        open(isp_dir / f"{s}.json", "w").write(CONFIGS[c])
    out, raw = tmp_path / "out", tmp_path / "raw"; out.mkdir(); raw.mkdir()
    p = subprocess.run([exe, "--isp_dir", str(isp_dir), "--output_dir", str(out), "--output_raw_dir", str(raw), "--bin_list", f"{bin0},{bin1}",
                        "--start_frame", "0", "--frame_count", "2"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-600:]
    assert "magic = faceb00c" in p.stdout and "numberOfCameras = 3" in p.stdout and "End frame (1) larger than total number of frames (0)" in p.stderr
    order = sorted(serials + [16400300])
    assert sorted(os.listdir(out)) == [f"cam{i}" for i in range(4)]
    for cam, serial in enumerate(order):
        is8 = serial == 16400300
        for f in range(1 if is8 else 2):
            packed = packed1[0][0] if is8 else packed0[f][serials.index(serial)]
            b = 8 if is8 else bits
            tiff = read_tiff16(raw / str(serial) / f"{f:06d}.tiff")   # the raw directories keep their serial names (only output_dir is renamed)
            assert np.array_equal(tiff, (ounpack.convert8_frame if is8 else ounpack.convert12_frame)(packed, w, h))
            png = out / f"cam{cam}" / f"{f:06d}.png"
            if serial not in cfg_of:
                assert not png.exists()
                continue
            got = cv2.imread(str(png), cv2.IMREAD_UNCHANGED)
            isp = unpacker_isp(cfg_of[serial], b)
            want = isp.footageImage(ctx, packed, b, w, h)
            isp.close()
            assert got.dtype == np.uint16 and np.array_equal(got, want)
            if rb.unpack_available() and rb.available():
                assert np.array_equal(got, reference_chain(packed, b, w, h, cfg_of[serial])[1])
        assert not (out / f"cam{cam}" / "000002.png").exists()
    p = subprocess.run([exe, "--isp_dir", str(isp_dir), "--output_dir", str(out), "--bin_list", bin0, "--start_frame", "7"], capture_output=True, text=True)
    assert p.returncode != 0 and "Start frame (7) larger than total number of frames (2)" in p.stderr
    p = subprocess.run([exe, "--output_dir", str(out), "--bin_list", bin0], capture_output=True, text=True)
    assert p.returncode != 0 and "isp_dir" in p.stderr
