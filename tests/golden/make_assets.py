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


# ---- BMP / TGA fixtures: every storage variant the readers of luisarender_b200/csrc/host/imageload.cpp accept ----------------------
def bmp_tga_pictures() -> dict[str, np.ndarray]:
    """name -> the (H, W, C) uint8 picture a reader must return (row 0 = top), C as the reference stores it (1, 2 or 4)."""
    r = np.random.default_rng(20240923)
    rgb = checker_rgb8(32, 4)[:17, :30]  # 30 pixels x 3 bytes: rows need 2 padding bytes in a BMP
    opaque = lambda a: np.concatenate([a, np.full(a.shape[:2] + (1,), 255, np.uint8)], axis=2)
    pal16 = r.integers(0, 256, (16, 3), dtype=np.uint8)
    idx8 = r.integers(0, 16, (9, 10), dtype=np.uint8)
    idx4 = r.integers(0, 12, (6, 7), dtype=np.uint8)  # core header: stb_image reads only 12 of the 16 palette entries (see imageload.cpp)
    idx1 = r.integers(0, 2, (5, 13), dtype=np.uint8)
    rgba = r.integers(0, 256, (7, 5, 4), dtype=np.uint8)
    rgba[..., 3] = np.maximum(rgba[..., 3], 1)
    f565 = np.stack([r.integers(0, 32, (6, 9)), r.integers(0, 64, (6, 9)), r.integers(0, 32, (6, 9))], axis=2).astype(np.uint16)
    f555 = r.integers(0, 32, (5, 6, 3)).astype(np.uint16)
    rep = lambda v, bits: ((v << (8 - bits)) | (v >> (2 * bits - 8))).astype(np.uint8)  # bit replication, bits in {5, 6}
    runs = np.repeat(r.integers(0, 256, (12, 4), dtype=np.uint8), r.integers(1, 9, 12), axis=0)[:48].reshape(6, 8, 4)  # runs across rows
    grey = rough_gray8(12)[:7]
    grey_alpha = r.integers(0, 256, (4, 6, 2), dtype=np.uint8)
    return {
        "bmp_rgb24": opaque(rgb), "bmp_pal8": opaque(pal16[idx8]), "bmp_pal4": opaque(pal16[idx4]), "bmp_pal1": opaque(pal16[idx1]),
        "bmp_rgba32_topdown": rgba, "bmp_rgbx32": opaque(rgba[..., :3]),
        "bmp_rgb565": opaque(np.stack([rep(f565[..., 0], 5), rep(f565[..., 1], 6), rep(f565[..., 2], 5)], axis=2)),
        "bmp_rgb555": opaque(rep(f555, 5)),
        "tga_rgb24": opaque(rgb), "tga_rgba32_rle_topdown": runs, "tga_grey8": grey[..., None], "tga_grey_alpha16": grey_alpha,
        "tga_mapped8": opaque(pal16[idx8]), "tga_rgb15": opaque((f555 * 255 // 31).astype(np.uint8)),
        "_pal16": pal16, "_idx8": idx8, "_idx4": idx4, "_idx1": idx1, "_f565": f565, "_f555": f555,
    }


def write_bmp_tga(out: Path) -> None:
    pic = bmp_tga_pictures()
    pal16, idx8, idx4, idx1, f565, f555 = (pic[k] for k in ("_pal16", "_idx8", "_idx4", "_idx1", "_f565", "_f555"))

    def bmp(name, w, h, bpp, rows, *, palette=b"", compression=0, masks=b"", top_down=False, core=False):
        body = b"".join(row + b"\0" * (-len(row) % 4) for row in (rows if top_down else rows[::-1]))
        if core:
            header = struct.pack("<IHHHH", 12, w, h, 1, bpp)
        else:
            header = struct.pack("<IiiHHIIiiII", 40, w, -h if top_down else h, 1, bpp, compression, len(body), 2835, 2835, 0, 0) + masks
        offset = 14 + len(header) + len(palette)
        (out / f"{name}.bmp").write_bytes(b"BM" + struct.pack("<IHHI", offset + len(body), 0, 0, offset) + header + palette + body)

    bgr = lambda a: a[..., ::-1].astype(np.uint8)
    rgb = pic["bmp_rgb24"][..., :3]
    h, w = rgb.shape[:2]
    bmp("bmp_rgb24", w, h, 24, [bgr(rgb)[y].tobytes() for y in range(h)])
    pal4 = b"".join(bytes([c[2], c[1], c[0], 0]) for c in pal16)
    bmp("bmp_pal8", idx8.shape[1], idx8.shape[0], 8, [idx8[y].tobytes() for y in range(idx8.shape[0])], palette=pal4)
    nib = lambda row: bytes((int(row[i]) << 4) | (int(row[i + 1]) if i + 1 < len(row) else 0) for i in range(0, len(row), 2))
    bmp("bmp_pal4", idx4.shape[1], idx4.shape[0], 4, [nib(idx4[y]) for y in range(idx4.shape[0])],
        palette=b"".join(bytes([c[2], c[1], c[0]]) for c in pal16), core=True)  # OS/2 core header: 3-byte palette entries
    bmp("bmp_pal1", idx1.shape[1], idx1.shape[0], 1, [np.packbits(idx1[y]).tobytes() for y in range(idx1.shape[0])], palette=pal4[:8])
    rgba = pic["bmp_rgba32_topdown"]
    bgra = lambda a: a[..., [2, 1, 0, 3]].astype(np.uint8)
    bmp("bmp_rgba32_topdown", rgba.shape[1], rgba.shape[0], 32, [bgra(rgba)[y].tobytes() for y in range(rgba.shape[0])], top_down=True)
    rgbx = rgba.copy()
    rgbx[..., 3] = 0  # alpha zero everywhere: "no alpha channel"
    bmp("bmp_rgbx32", rgbx.shape[1], rgbx.shape[0], 32, [bgra(rgbx)[y].tobytes() for y in range(rgbx.shape[0])])
    v565 = ((f565[..., 0] << 11) | (f565[..., 1] << 5) | f565[..., 2]).astype("<u2")
    bmp("bmp_rgb565", v565.shape[1], v565.shape[0], 16, [v565[y].tobytes() for y in range(v565.shape[0])], compression=3,
        masks=struct.pack("<III", 0xF800, 0x07E0, 0x001F))
    v555 = ((f555[..., 0] << 10) | (f555[..., 1] << 5) | f555[..., 2]).astype("<u2")
    bmp("bmp_rgb555", v555.shape[1], v555.shape[0], 16, [v555[y].tobytes() for y in range(v555.shape[0])])

    def tga(name, image_type, w, h, bits, body, *, descriptor=0, palette=b"", map_len=0, map_bits=0, ident=b""):
        head = struct.pack("<BBBHHBHHHHBB", len(ident), 1 if palette else 0, image_type, 0, map_len, map_bits, 0, 0, w, h, bits, descriptor)
        (out / f"{name}.tga").write_bytes(head + ident + palette + body)

    tga("tga_rgb24", 2, w, h, 24, bgr(rgb)[::-1].tobytes(), ident=b"lrk")  # bottom-up, with an image id to skip
    runs = pic["tga_rgba32_rle_topdown"]
    flat = bgra(runs).reshape(-1, 4)
    body, i = b"", 0
    while i < len(flat):  # run-length packets for repeats (they cross row ends), raw packets otherwise
        j = i
        while j + 1 < len(flat) and j - i < 127 and (flat[j + 1] == flat[i]).all():
            j += 1
        if j > i:
            body += bytes([0x80 | (j - i)]) + flat[i].tobytes()
            i = j + 1
        else:
            body += bytes([0]) + flat[i].tobytes()
            i += 1
    tga("tga_rgba32_rle_topdown", 10, runs.shape[1], runs.shape[0], 32, body, descriptor=0x28)
    grey = pic["tga_grey8"][..., 0]
    tga("tga_grey8", 3, grey.shape[1], grey.shape[0], 8, grey[::-1].tobytes())
    ga = pic["tga_grey_alpha16"]
    tga("tga_grey_alpha16", 3, ga.shape[1], ga.shape[0], 16, ga.tobytes(), descriptor=0x28)
    tga("tga_mapped8", 1, idx8.shape[1], idx8.shape[0], 8, idx8[::-1].tobytes(), palette=bgr(pal16).tobytes(), map_len=16, map_bits=24)
    tga("tga_rgb15", 2, v555.shape[1], v555.shape[0], 16, v555[::-1].tobytes())


# ---- JPEG fixtures: every decoding path of luisarender_b200/csrc/host/jpegload.cpp -------------------------------------------------
# Written with Pillow / OpenCV (libjpeg-turbo) - present in this image; the files are committed, and what a reader must return for
# them is what the reference's stb_image decodes (tests/golden/jpeg_texels.npz, tools/gen_jpeg_pins.py).
def jpeg_picture(w: int, h: int, channels: int = 3, noise: int = 40) -> np.ndarray:
    """smooth colour gradients with edges and some noise: chroma upsampling, the IDCT's rounding and the clamps all matter"""
    r = np.random.default_rng(w * 1000 + h * 10 + channels)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    base = np.stack([127 + 120 * np.sin(x / 3.1 + y / 7.0), 127 + 120 * np.cos(x / 5.3 - y / 2.9), 255.0 * ((x // 4 + y // 3) % 2)], axis=2)
    pic = np.clip(base + r.integers(-noise, noise + 1, (h, w, 3)), 0, 255).astype(np.uint8)
    return pic if channels == 3 else pic[..., 0]


def write_jpegs(out: Path) -> None:
    import io

    import cv2
    from PIL import Image

    def pil(name, a, **kw):
        b = io.BytesIO()
        Image.fromarray(a).save(b, "JPEG", **kw)
        (out / f"{name}.jpg").write_bytes(b.getvalue())
        return b.getvalue()

    def cv(name, a, *flags):
        ok, b = cv2.imencode(".jpg", a[..., ::-1].copy(), [int(f) for f in flags])
        assert ok
        (out / f"{name}.jpg").write_bytes(b.tobytes())

    pil("jpg_grey", jpeg_picture(21, 13, 1), quality=85)
    pil("jpg_444", jpeg_picture(19, 13), quality=90, subsampling=0)
    pil("jpg_422", jpeg_picture(35, 11), quality=88, subsampling=1)
    pil("jpg_420", jpeg_picture(37, 23), quality=90, subsampling=2)
    pil("jpg_420_restart_optimized", jpeg_picture(41, 35), quality=80, subsampling=2, restart_marker_blocks=2, optimize=True)
    pil("jpg_420_low_quality", jpeg_picture(40, 24, noise=90), quality=12, subsampling=2)
    pil("jpg_444_q100_noise", jpeg_picture(16, 16, noise=127), quality=100, subsampling=0)
    pil("jpg_420_one_pixel_wide", jpeg_picture(1, 9), quality=90, subsampling=2)
    pil("jpg_422_one_pixel_wide", jpeg_picture(1, 5), quality=90, subsampling=1)
    pil("jpg_progressive_420", jpeg_picture(45, 27), quality=85, subsampling=2, progressive=True)
    pil("jpg_progressive_444_restart", jpeg_picture(26, 18), quality=92, subsampling=0, progressive=True, restart_marker_blocks=3)
    pil("jpg_progressive_grey", jpeg_picture(30, 17, 1), quality=75, progressive=True)
    cv("jpg_440", jpeg_picture(18, 29), cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440)
    cv("jpg_411", jpeg_picture(43, 10), cv2.IMWRITE_JPEG_QUALITY, 90, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411)
    cv("jpg_progressive_440_restart", jpeg_picture(22, 33), cv2.IMWRITE_JPEG_QUALITY, 80, cv2.IMWRITE_JPEG_SAMPLING_FACTOR,
       cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_PROGRESSIVE, 1, cv2.IMWRITE_JPEG_RST_INTERVAL, 2)

    # four components: Pillow writes CMYK pictures with an Adobe marker (transform 0: CMYK as such, inverted); the same scan data under
    # transform 2 reads as YCCK
    cmyk = np.concatenate([jpeg_picture(23, 15), jpeg_picture(23, 15, noise=10)[..., :1]], axis=2)
    b = io.BytesIO()
    Image.fromarray(cmyk, "CMYK").save(b, "JPEG", quality=90)
    data = b.getvalue()
    (out / "jpg_cmyk.jpg").write_bytes(data)
    at = data.index(b"Adobe") + 11
    assert data[at] in (0, 2)
    (out / "jpg_cmyk_other_transform.jpg").write_bytes(data[:at] + bytes([2 - data[at]]) + data[at + 1:])
    b = io.BytesIO()
    Image.fromarray(cmyk, "CMYK").save(b, "JPEG", quality=85, progressive=True, subsampling=2)
    (out / "jpg_cmyk_progressive_subsampled.jpg").write_bytes(b.getvalue())

    # three components stored as RGB: by their ids 'R','G','B', or by an Adobe marker with transform 0 and no JFIF marker
    def segments(data):
        pos, found = 2, {}
        while data[pos + 1] != 0xDA:
            length = int.from_bytes(data[pos + 2:pos + 4], "big")
            found.setdefault(data[pos + 1], []).append((pos, length + 2))
            pos += length + 2
        found[0xDA] = [(pos, int.from_bytes(data[pos + 2:pos + 4], "big") + 2)]
        return found

    src = bytearray(pil("jpg_rgb_by_ids", jpeg_picture(20, 12), quality=90, subsampling=0))
    seg = segments(src)
    sof, sos = seg[0xC0][0][0], seg[0xDA][0][0]
    for k, letter in enumerate(b"RGB"):
        assert src[sof + 10 + 3 * k] == k + 1 and src[sos + 5 + 2 * k] == k + 1
        src[sof + 10 + 3 * k] = letter
        src[sos + 5 + 2 * k] = letter
    (out / "jpg_rgb_by_ids.jpg").write_bytes(bytes(src))
    src = pil("jpg_rgb_by_adobe_marker", jpeg_picture(17, 14), quality=90, subsampling=0)
    app0, n0 = segments(src)[0xE0][0]
    adobe = b"\xff\xee\x00\x0eAdobe\x00\x64\x00\x00\x00\x00\x00"
    (out / "jpg_rgb_by_adobe_marker.jpg").write_bytes(src[:app0] + adobe + src[app0 + n0:])


def main():
    OUT.mkdir(exist_ok=True)
    write_bmp_tga(OUT)
    write_jpegs(OUT)
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
