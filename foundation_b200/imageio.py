"""Image output for the progressive render: linear radiance as PFM or OpenEXR, the presented frame as PNG.

The reference has no on-disk image format of its own (it only *reads* one JPEG through stb_image, mos9527/Foundation
src/Renderer/Renderer.cpp:198-202), so all writers are dependency-free: PFM is a text header plus raw floats, PNG needs only zlib, the OpenEXR subset (uncompressed float scan lines) only struct.
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


# ---------------------------------------------------------------------------------------------------------------------
# OpenEXR, the subset a linear-radiance frame needs: version 2, single-part scan-line image, no compression, 32-bit float
# channels, increasing-y line order.  File layout per the published OpenEXR file-layout document: magic 20000630, version field,
# attributes (name\0 type\0 size value) closed by a zero byte, one 64-bit offset per scan line, then per line
# [y : int32][bytes : int32][channel 0 row]...[channel n-1 row] with the channels in the (alphabetical) order of the chlist.
# ---------------------------------------------------------------------------------------------------------------------
_EXR_MAGIC = 20000630


def _exr_attr(name: str, typ: str, payload: bytes) -> bytes:
    return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload


def write_exr(path: str, img: np.ndarray) -> None:
    """Linear radiance (H, W, 3) -> channels B, G, R; (H, W, 4) -> A, B, G, R; float32, uncompressed scan lines, row 0 = top."""
    a = np.asarray(img, np.float32)
    if a.ndim != 3 or a.shape[2] not in (3, 4):
        raise ValueError("img must be (H, W, 3|4)")
    h, w, c = a.shape
    names = ["B", "G", "R"] if c == 3 else ["A", "B", "G", "R"]
    src = {"R": 0, "G": 1, "B": 2, "A": 3}
    chlist = b"".join(n.encode() + b"\0" + struct.pack("<iBBBBii", 2, 0, 0, 0, 0, 1, 1) for n in names) + b"\0"   # pixel type 2 = FLOAT, pLinear 0, sampling 1 x 1
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<ii", _EXR_MAGIC, 2)
    header += _exr_attr("channels", "chlist", chlist)
    header += _exr_attr("compression", "compression", b"\0")                    # NO_COMPRESSION
    header += _exr_attr("dataWindow", "box2i", box) + _exr_attr("displayWindow", "box2i", box)
    header += _exr_attr("lineOrder", "lineOrder", b"\0")                        # INCREASING_Y
    header += _exr_attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += _exr_attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    header += _exr_attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    line_bytes = len(names) * w * 4
    first = len(header) + 8 * h
    offsets = (first + np.arange(h, dtype=np.uint64) * np.uint64(8 + line_bytes)).astype("<u8")
    planes = np.ascontiguousarray(np.stack([a[..., src[n]] for n in names], axis=1)).astype("<f4")     # (H, channels, W)
    with open(path, "wb") as f:
        f.write(header)
        f.write(offsets.tobytes())
        for y in range(h):
            f.write(struct.pack("<ii", y, line_bytes))
            f.write(planes[y].tobytes())


def read_exr(path: str) -> np.ndarray:
    """Reader for the subset `write_exr` emits (and any other uncompressed single-part scan-line file with FLOAT R, G, B[, A] channels)."""
    with open(path, "rb") as f:
        raw = f.read()
    magic, version = struct.unpack("<ii", raw[:8])
    if magic != _EXR_MAGIC or (version & 0xFF) != 2 or (version & 0x1E00):
        raise ValueError("not a single-part scan-line OpenEXR 2 file")
    pos, attrs = 8, {}
    while raw[pos] != 0:
        e = raw.index(b"\0", pos); name = raw[pos:e].decode(); pos = e + 1
        e = raw.index(b"\0", pos); typ = raw[pos:e].decode(); pos = e + 1
        (n,) = struct.unpack("<i", raw[pos:pos + 4]); pos += 4
        attrs[name] = (typ, raw[pos:pos + n]); pos += n
    pos += 1
    if attrs["compression"][1] != b"\0" or attrs["lineOrder"][1] != b"\0":
        raise ValueError("unsupported OpenEXR variant")
    x0, y0, x1, y1 = struct.unpack("<iiii", attrs["dataWindow"][1])
    w, h = x1 - x0 + 1, y1 - y0 + 1
    names, p, ch = [], 0, attrs["channels"][1]
    while ch[p] != 0:
        e = ch.index(b"\0", p); names.append(ch[p:e].decode()); p = e + 1
        if struct.unpack("<i", ch[p:p + 4])[0] != 2:
            raise ValueError("only FLOAT channels are supported")
        p += 16
    offsets = np.frombuffer(raw[pos:pos + 8 * h], "<u8")
    out = np.zeros((h, w, len(names)), np.float32)
    order = [n for n in ("R", "G", "B", "A") if n in names]
    for row in range(h):
        o = int(offsets[row])
        y, nbytes = struct.unpack("<ii", raw[o:o + 8])
        if nbytes != len(names) * w * 4:
            raise ValueError("unexpected scan-line size")
        line = np.frombuffer(raw[o + 8:o + 8 + nbytes], "<f4").reshape(len(names), w)
        for k, n in enumerate(order):
            out[y - y0, :, k] = line[names.index(n)]
    return out[..., :len(order)]
