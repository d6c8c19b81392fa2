"""CPU tier for the Unpacker's front end: the footage container reader of the product (host code behind s360_footage_*)
and the numpy restatement of RawConverter (oracle/ounpack.py), both against the REFERENCE'S OWN BinaryFootageFile / RawConverter
compiled unchanged (oracle/_ref/libref_unpack.so), and the golden checksums the GPU tier uses on the box without /root/reference."""
import json
import os
import sys
import zlib

import numpy as np
import pytest

import ref_bind as rb
from surround360_b200 import api, synth

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import ounpack  # noqa: E402

GOLDEN = json.load(open(os.path.join(HERE, "golden", "unpack_golden.json")))
needs_ref = pytest.mark.skipif(not rb.unpack_available(), reason="oracle/_ref/libref_unpack.so not built (needs /root/reference)")


def footage_case(bits, n_frames=3, n_cams=4, h=64, w=128, seed=5):   # frame sizes are page multiples: the reference madvise()s whole frames
    rng = np.random.default_rng(seed)
    top = 4096 if bits == 12 else 256
    frames = [[rng.integers(0, top, (h, w)).astype(np.uint16) for _ in range(n_cams)] for _ in range(n_frames)]
    serials = [16400000 + 37 * c for c in range(n_cams)]
    return frames, serials


@needs_ref
@pytest.mark.parametrize("bits", [8, 12])
def test_port_matches_reference_converter(bits):
    rng = np.random.default_rng(bits)
    for (h, w) in [(2, 2), (5, 6), (33, 64), (128, 200)]:
        packed = rng.integers(0, 256, h * w * bits // 8, dtype=np.uint8)
        port = (ounpack.convert12_frame if bits == 12 else ounpack.convert8_frame)(packed, w, h)
        assert np.array_equal(port, rb.raw_convert(packed, bits, w, h))
    # every 12-bit value in both pixel positions survives pack -> convert as its 16-bit replicate
    if bits == 12:
        v = np.arange(4096, dtype=np.uint16).reshape(64, 64)
        got = rb.raw_convert(synth.pack12(v), 12, 64, 64)
        assert np.array_equal(got, (v << 4 | v >> 8))
        assert np.array_equal(rb.raw_convert(synth.pack12(v.T.copy()), 12, 64, 64), (v.T << 4 | v.T >> 8))


def test_port_matches_golden_checksums():
    """The same port against checksums of the reference build's output (tests/golden/make_unpack_golden.py): holds on the GPU box too."""
    for bits in (8, 12):
        rng = np.random.default_rng(100 + bits)
        packed = rng.integers(0, 256, 96 * 128 * bits // 8, dtype=np.uint8)
        port = (ounpack.convert12_frame if bits == 12 else ounpack.convert8_frame)(packed, 128, 96)
        assert format(zlib.crc32(port.tobytes()), "08x") == GOLDEN[f"convert{bits}_96x128"]


@pytest.mark.parametrize("bits", [8, 12])
def test_container_reader_matches_reference(tmp_path, bits):
    frames, serials = footage_case(bits)
    path = tmp_path / f"footage{bits}.bin"
    packed = synth.write_footage_bin(str(path), frames, bits, serials, timestamp=1234, file_index=1, file_count=2)
    f = api.BinaryFootageFile(path)
    hd = f.getMetadata()
    assert (hd.magic, hd.timestamp, hd.fileIndex, hd.fileCount, hd.width, hd.height, hd.bitsPerPixel, hd.numberOfCameras) == (0xfaceb00c, 1234, 1, 2, 128, 64, bits, 4)
    assert f.getNumberOfFrames() == 3 and f.getFrameSize() == 128 * 64 * bits // 8 and f.getBitsPerPixel() == bits and f.getNumberOfCameras() == 4
    lay = ounpack.footage_layout(open(path, "rb").read())
    assert lay["frames"] == 3 and lay["frameSize"] == f.getFrameSize()
    ref = rb.FootageFile(path) if rb.unpack_available() else None
    if ref:
        assert ref.meta == [0xfaceb00c, 1234, 1, 2, 128, 64, bits, 4] and ref.frames == 3 and ref.frame_size == f.getFrameSize()
    blob = open(path, "rb").read()
    for fi in range(3):
        for c in range(4):
            got = f.getFrame(fi, c)
            assert np.array_equal(got, packed[fi][c])
            assert got.tobytes() == blob[lay["offset"](fi, c):lay["offset"](fi, c) + lay["frameSize"]]
            assert int(np.frombuffer(got[4:8].tobytes(), np.uint32)[0]) == serials[c]   # Unpacker.cpp:125
            if ref:
                assert np.array_equal(got, ref.frame(fi, c))
    with pytest.raises(api.VrCamException, match="Camera number out of range"):   # BinaryFootageFile.cpp:183-185
        f.getFrame(0, 4)
    with pytest.raises(api.VrCamException, match="out of range for this file"):
        f.getFrame(3, 0)
    if ref:
        with pytest.raises(RuntimeError, match="Camera number out of range"):
            ref.frame(0, 4)
        ref.close()
    f.close()


def test_container_trailing_bytes_and_empty(tmp_path):
    """getNumberOfFrames floors (BinaryFootageFile.cpp:160-168): a truncated last frame does not count; zero cameras -> zero frames."""
    frames, serials = footage_case(12, n_frames=2, n_cams=2, h=24, w=40)   # 1440-byte frames: not page aligned (the reference's madvise
    path = tmp_path / "t.bin"                                              # rejects those; this reader releases only whole pages)
    synth.write_footage_bin(str(path), frames, 12, serials)
    blob = open(path, "rb").read()
    open(path, "wb").write(blob[:-100])
    f = api.BinaryFootageFile(path)
    assert f.getNumberOfFrames() == 1
    with pytest.raises(api.VrCamException, match="out of range"):
        f.getFrame(1, 1)   # straddles the end of the file
    f.getFrame(1, 0)
    f.close()
    if rb.unpack_available():
        r = rb.FootageFile(path); assert r.frames == 1; r.close()
    hd = np.array([0xfaceb00c, 0, 0, 1, 40, 24, 12, 0], np.uint32)
    open(path, "wb").write(hd.tobytes() + b"\0" * (4096 - 32))
    f = api.BinaryFootageFile(path)
    assert f.getNumberOfFrames() == 0 and f.getNumberOfCameras() == 0
    f.close()
    with pytest.raises(api.VrCamException, match="Error opening file"):
        api.BinaryFootageFile(tmp_path / "missing.bin")
