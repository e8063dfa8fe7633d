"""Image output for the progressive render: linear radiance as PFM, the presented frame as PNG.

The reference has no on-disk image format of its own (it only *reads* one JPEG through stb_image, mos9527/Foundation
src/Renderer/Renderer.cpp:198-202), so both writers are dependency-free: PFM is a text header plus raw floats, PNG needs only zlib.
Row 0 of every array here is the TOP row, matching the Vulkan pixel origin the renderer uses (Renderer.cpp:373-380).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


def radiance_from_accum(accum: np.ndarray) -> np.ndarray:
    """float4 accumulation buffer (sum of radiance in rgb, number of samples in a) -> mean linear radiance (H, W, 3) float32.
    Pixels with no samples (tiles owned by another rank) resolve to 0."""
    a = np.asarray(accum, np.float32)
    if a.ndim != 3 or a.shape[2] != 4:
        raise ValueError("accum must be (H, W, 4)")
    n = a[..., 3:4]
    return np.where(n > 0, a[..., :3] / np.where(n > 0, n, 1), 0).astype(np.float32)


def write_pfm(path: str, rgb: np.ndarray) -> None:
    """Colour PFM ('PF'), little-endian (negative scale), rows stored bottom-up as the format requires."""
    img = np.asarray(rgb, np.float32)
    if img.ndim != 3 or img.shape[2] != 3:
        raise ValueError("rgb must be (H, W, 3)")
    h, w, _ = img.shape
    with open(path, "wb") as f:
        f.write(b"PF\n%d %d\n-1.0\n" % (w, h))
        f.write(np.ascontiguousarray(img[::-1]).astype("<f4").tobytes())


def read_pfm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        magic = f.readline().strip()
        if magic not in (b"PF", b"Pf"):
            raise ValueError("not a PFM file")
        w, h = (int(x) for x in f.readline().split())
        scale = float(f.readline())
        ch = 3 if magic == b"PF" else 1
        data = np.frombuffer(f.read(w * h * ch * 4), "<f4" if scale < 0 else ">f4")
    if data.size != w * h * ch:
        raise ValueError("truncated PFM")
    return data.reshape(h, w, ch)[::-1].astype(np.float32)


def _chunk(tag: bytes, payload: bytes) -> bytes:
    return struct.pack(">I", len(payload)) + tag + payload + struct.pack(">I", zlib.crc32(tag + payload) & 0xFFFFFFFF)


def write_png(path: str, rgba8: np.ndarray, level: int = 6) -> None:
    """8-bit RGB or RGBA PNG (filter 0 on every row) of the R8G8B8A8_UNORM frame `resolve_rgba8` returns (Renderer.cpp:40)."""
    img = np.ascontiguousarray(rgba8)
    if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] not in (3, 4):
        raise ValueError("rgba8 must be (H, W, 3|4) uint8")
    h, w, c = img.shape
    rows = np.zeros((h, 1 + w * c), np.uint8)
    rows[:, 1:] = img.reshape(h, w * c)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(_chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6 if c == 4 else 2, 0, 0, 0)))
        f.write(_chunk(b"IDAT", zlib.compress(rows.tobytes(), level)))
        f.write(_chunk(b"IEND", b""))


def read_png(path: str) -> np.ndarray:
    """Reader for the subset `write_png` emits (8-bit RGB/RGBA, non-interlaced, filter 0) — used by the round-trip test."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"\x89PNG\r\n\x1a\n":
        raise ValueError("not a PNG file")
    pos, idat, hdr = 8, b"", None
    while pos < len(raw):
        (n,), tag = struct.unpack(">I", raw[pos:pos + 4]), raw[pos + 4:pos + 8]
        body = raw[pos + 8:pos + 8 + n]
        if struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] != (zlib.crc32(tag + body) & 0xFFFFFFFF):
            raise ValueError("bad PNG chunk CRC")
        if tag == b"IHDR":
            hdr = struct.unpack(">IIBBBBB", body)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    w, h, depth, ctype, _, _, interlace = hdr
    if depth != 8 or ctype not in (2, 6) or interlace:
        raise ValueError("unsupported PNG variant")
    c = 4 if ctype == 6 else 3
    rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * c)
    if rows[:, 0].any():
        raise ValueError("unsupported PNG row filter")
    return rows[:, 1:].reshape(h, w, c).copy()
