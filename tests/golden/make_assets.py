"""Generates the small texture / mesh fixtures under tests/golden/assets (deterministic; files are committed).

    python tests/golden/make_assets.py
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

OUT = Path(__file__).resolve().parent / "assets"


def write_png(path: Path, a: np.ndarray, palette: np.ndarray | None = None) -> None:
    """a: (H, W) or (H, W, C) uint8 / uint16; C in {1, 2, 3, 4}.  With `palette` (N, 3) uint8, `a` holds indices."""
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    depth = 16 if a.dtype == np.uint16 else 8
    ctype = 3 if palette is not None else {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = a.astype(">u2" if depth == 16 else np.uint8).tobytes()
    stride = w * c * depth // 8
    rows = b"".join(b"\x00" + raw[y * stride:(y + 1) * stride] for y in range(h))  # filter type 0

    def chunk(tag: bytes, body: bytes) -> bytes:
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body) & 0xFFFFFFFF)

    data = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 0))
    if palette is not None:
        data += chunk(b"PLTE", palette.astype(np.uint8).tobytes())
    data += chunk(b"IDAT", zlib.compress(rows, 9)) + chunk(b"IEND", b"")
    path.write_bytes(data)


def checker_rgb8(n=64, cells=8) -> np.ndarray:
    y, x = np.mgrid[0:n, 0:n]
    c = ((x // (n // cells)) + (y // (n // cells))) % 2
    img = np.zeros((n, n, 3), np.uint8)
    img[c == 0] = (200, 60, 40)
    img[c == 1] = (40, 90, 210)
    img[:, :, 1] += (x * 40 // n).astype(np.uint8)  # a gradient so that bilinear weights matter
    return img


def rough_gray8(n=32) -> np.ndarray:
    y, x = np.mgrid[0:n, 0:n]
    return (40 + (x * 5 + y * 2) % 180).astype(np.uint8)


def ramp_rgba16(w=16, h=8) -> np.ndarray:
    y, x = np.mgrid[0:h, 0:w]
    img = np.zeros((h, w, 4), np.uint16)
    img[..., 0] = x * 4096 + 17
    img[..., 1] = y * 8000 + 255
    img[..., 2] = (x + y) * 2500
    img[..., 3] = 65535
    return img


def alpha_gray8(n=16) -> np.ndarray:
    """Cut-out mask: opaque discs on a transparent ground, with a soft rim so that the stochastic test sees fractional alphas."""
    y, x = np.mgrid[0:n, 0:n]
    d = np.hypot((x % 8) - 3.5, (y % 8) - 3.5)
    return np.clip((3.6 - d) * 160, 0, 255).astype(np.uint8)


def normal_rgb8(n=32) -> np.ndarray:
    """Tangent-space normal map of a sine bump field, encoded rgb = (n + 1) / 2."""
    y, x = np.mgrid[0:n, 0:n]
    nx, ny = 0.45 * np.cos(x * 2 * np.pi / 8), 0.45 * np.sin(y * 2 * np.pi / 16)
    nz = np.sqrt(np.maximum(1 - nx * nx - ny * ny, 0))
    return np.round((np.stack([nx, ny, nz], -1) + 1) * 127.5).astype(np.uint8)


def sky_pfm(w=32, h=16) -> np.ndarray:
    """A small HDR sky (top row = zenith): blue-ish gradient, a warm horizon band and one very bright 'sun' texel."""
    y, x = np.mgrid[0:h, 0:w]
    t = (y + 0.5) / h
    img = np.zeros((h, w, 3), np.float32)
    img[..., 0] = 0.15 + 0.5 * np.exp(-((t - 0.5) * 6) ** 2)
    img[..., 1] = 0.25 + 0.4 * np.exp(-((t - 0.5) * 6) ** 2)
    img[..., 2] = 0.6 * (1 - t) + 0.1
    img[t > 0.55] *= 0.15  # dark ground half
    img[4, 22] = (60.0, 52.0, 40.0)
    img[4, 23] = (30.0, 26.0, 20.0)
    return img


CUBE_OBJ = """# unit cube centred at the origin, quads, per-face uvs, no normals (smooth-normal generation is exercised)
v -0.5 -0.5 -0.5
v  0.5 -0.5 -0.5
v  0.5  0.5 -0.5
v -0.5  0.5 -0.5
v -0.5 -0.5  0.5
v  0.5 -0.5  0.5
v  0.5  0.5  0.5
v -0.5  0.5  0.5
vt 0 0
vt 1 0
vt 1 1
vt 0 1
f 5/1 6/2 7/3 8/4
f 2/1 1/2 4/3 3/4
f 6/1 2/2 3/3 7/4
f 1/1 5/2 8/3 4/4
f 8/1 7/2 3/3 4/4
f 1/1 2/2 6/3 5/4
"""


def write_tetra_ply(path: Path, binary: bool) -> None:
    verts = np.array([[0, 0, 0, 0, 0], [1, 0, 0, 1, 0], [0, 1, 0, 0, 1], [0, 0, 1, 1, 1]], np.float32)  # x y z s t
    faces = [(0, 2, 1), (0, 1, 3), (0, 3, 2), (1, 2, 3)]
    head = ["ply", f"format {'binary_little_endian' if binary else 'ascii'} 1.0", "comment tetrahedron fixture", "element vertex 4",
            "property float x", "property float y", "property float z", "property float s", "property float t",
            "element face 4", "property list uchar int vertex_indices", "end_header"]
    data = ("\n".join(head) + "\n").encode()
    if binary:
        data += verts.tobytes()
        for f in faces:
            data += struct.pack("<Biii", 3, *f)
    else:
        data += "".join(" ".join(repr(float(x)) for x in v) + "\n" for v in verts).encode()
        data += "".join("3 " + " ".join(str(i) for i in f) + "\n" for f in faces).encode()
    path.write_bytes(data)


def main():
    OUT.mkdir(exist_ok=True)
    write_png(OUT / "checker_rgb8.png", checker_rgb8())
    write_png(OUT / "rough_gray8.png", rough_gray8())
    write_png(OUT / "ramp_rgba16.png", ramp_rgba16())
    write_png(OUT / "alpha_gray8.png", alpha_gray8())
    write_png(OUT / "normal_rgb8.png", normal_rgb8())
    pal = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [255, 255, 0]], np.uint8)
    idx = (np.add.outer(np.arange(8), np.arange(8)) % 4).astype(np.uint8)
    write_png(OUT / "palette4.png", idx, palette=pal)
    (OUT / "cube.obj").write_text(CUBE_OBJ)
    write_tetra_ply(OUT / "tetra_ascii.ply", False)
    write_tetra_ply(OUT / "tetra_binary.ply", True)
    sky = sky_pfm()
    (OUT / "sky.pfm").write_bytes(b"PF\n%d %d\n-1.0\n" % (sky.shape[1], sky.shape[0]) + sky[::-1].astype("<f4").tobytes())
    # a PFM (bottom-up, little endian) and a P6 PPM of the same 4x2 picture
    pic = np.arange(4 * 2 * 3, dtype=np.float32).reshape(2, 4, 3) / 23.0
    (OUT / "tiny.pfm").write_bytes(b"PF\n4 2\n-1.0\n" + pic[::-1].astype("<f4").tobytes())
    (OUT / "tiny.ppm").write_bytes(b"P6\n4 2\n255\n" + np.round(pic * 255).astype(np.uint8).tobytes())
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
