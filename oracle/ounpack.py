"""ORACLE — TEST INFRASTRUCTURE ONLY. Not part of the product path.

numpy restatement of the Unpacker's front end, pinned bit for bit to the reference's own code (oracle/_ref/libref_unpack.so =
camera_isp/RawConverter.cpp + BinaryFootageFile.cpp compiled unchanged) by tests/test_footage.py."""
import numpy as np


def convert8_frame(raw_frame, width, height):
    """RawConverter::convert8Frame, camera_isp/RawConverter.cpp:14-32: every byte times 0x101."""
    f = np.frombuffer(np.ascontiguousarray(raw_frame, np.uint8), np.uint8)[:width * height].astype(np.uint32)
    return (f * 0x101).astype(np.uint16).reshape(height, width)


def convert12_frame(raw_frame, width, height):
    """RawConverter::convert12Frame, camera_isp/RawConverter.cpp:34-59, for even widths (the running byte index p then advances by
    exactly three bytes per pixel pair and rows do not straddle bytes): even x reads lo = b0, hi = b1 -> lo << 4 | hi & 0xF; odd x
    reads lo = b1, hi = b2 -> hi << 4 | lo >> 4; the 12-bit value u is replicated to 16 bits as uint16(u << 4 | u >> 8)."""
    assert width % 2 == 0
    f = np.frombuffer(np.ascontiguousarray(raw_frame, np.uint8), np.uint8)[:width * height * 3 // 2].astype(np.uint32).reshape(-1, 3)
    even = f[:, 0] << 4 | (f[:, 1] & 0xF)
    odd = f[:, 2] << 4 | f[:, 1] >> 4
    u = np.stack([even, odd], 1).reshape(height, width)
    return ((u << 4 | u >> 8) & 0xFFFF).astype(np.uint16)


def footage_layout(file_bytes):
    """BinaryFootageFile (camera_isp/BinaryFootageFile.cpp:119-121, 160-201): header words, frame size, number of frames and the byte
    offset of frame f of camera c."""
    hd = np.frombuffer(file_bytes[:32], np.uint32)
    magic, timestamp, file_index, file_count, w, h, bpp, cams = [int(v) for v in hd]
    frame_size = w * h * bpp // 8
    frames = 0 if cams == 0 else (len(file_bytes) - 4096) // frame_size // cams
    return dict(magic=magic, timestamp=timestamp, fileIndex=file_index, fileCount=file_count, width=w, height=h, bitsPerPixel=bpp,
                numberOfCameras=cams, frameSize=frame_size, frames=frames,
                offset=lambda f, c: 4096 + (cams * f + c) * frame_size)
