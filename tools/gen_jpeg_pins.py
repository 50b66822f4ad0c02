#!/usr/bin/env python
"""tests/golden/jpeg_texels.npz: what the REFERENCE's own stb_image decodes for every JPEG fixture of tests/golden/assets.

stb_image is compiled, unmodified, into oracle/_ref/bin/libluisa-ref.so (oracle/ref/Makefile: src/compute/src/ext/stb); this
script calls its stbi_info / stbi_load exactly as LoadedImage::load does (src/util/imageio.cpp:347-470: the file's component count
decides the storage, 1 -> one channel, 3 -> four) and stores the bytes.  Runs only where /root/reference was built
(python __graft_entry__.py build); the .npz is committed, tests/test_textures_meshes.py compares the host reader with it."""
from __future__ import annotations

import ctypes
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
ASSETS = REPO / "tests" / "golden" / "assets"
OUT = REPO / "tests" / "golden" / "jpeg_texels.npz"
LIB = REPO / "oracle" / "_ref" / "bin" / "libluisa-ref.so"


def reference_decode(lib, path: Path) -> np.ndarray:
    w, h, c = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    if not lib.stbi_info(str(path).encode(), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c)):
        raise RuntimeError(f"stbi_info failed for {path}")
    desired = 4 if c.value >= 3 else c.value
    p = lib.stbi_load(str(path).encode(), ctypes.byref(w), ctypes.byref(h), ctypes.byref(c), desired)
    if not p:
        raise RuntimeError(f"stbi_load failed for {path}")
    a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(h.value, w.value, desired)).copy()
    lib.stbi_image_free(p)
    return a


def load_reference_stb():
    lib = ctypes.CDLL(str(LIB))
    lib.stbi_info.restype = ctypes.c_int
    lib.stbi_load.restype = ctypes.c_void_p
    lib.stbi_image_free.argtypes = [ctypes.c_void_p]
    return lib


def decode_all() -> dict[str, np.ndarray]:
    lib = load_reference_stb()
    return {p.stem: reference_decode(lib, p) for p in sorted(ASSETS.glob("jpg_*.jpg"))}


def main() -> int:
    if not LIB.exists():
        print(f"{LIB} is missing: build the reference first (python __graft_entry__.py build)", file=sys.stderr)
        return 1
    data = decode_all()
    np.savez_compressed(OUT, **data)
    print(f"wrote {OUT}: {len(data)} pictures, {sum(a.size for a in data.values())} bytes")
    return 0


if __name__ == "__main__":
    sys.exit(main())
