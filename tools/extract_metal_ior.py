"""Writes luisarender_b200/data/metal_ior.bin: the complex refractive indices of the Metal surface's eleven named metals.

The reference's Metal surface (src/surfaces/metal.cpp:73-104) accepts `eta { "Au" }` and the like and looks the name up in
measured optical constants - (n, k) at 5 nm steps over 360 - 830 nm, 95 samples per metal - that it ships as a data table
(src/surfaces/metal_ior.inl.h).  They are measurements (published optical constants of the elements and two compounds,
resampled), data and not code, and cannot be regenerated from anything else in the repository.  This script reads the reference's
copy where it lies and stores the numbers as one binary file; no source text is copied.

    python tools/extract_metal_ior.py [/root/reference]

File layout (little endian): magic "LRMI", u32 version = 1, u32 metal count, u32 samples per metal (95: 360 nm .. 830 nm, step 5),
then per metal: 8 bytes name (NUL padded, as the table is called in the reference: Ag Al Au Cu CuZn Fe Ti V VN Li Cr),
f32[samples] n, f32[samples] k.
"""
from __future__ import annotations

import re
import struct
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
OUT = REPO / "luisarender_b200" / "data" / "metal_ior.bin"


def main() -> int:
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    text = (ref / "src" / "surfaces" / "metal_ior.inl.h").read_text()
    text = re.sub(r"//[^\n]*", "", text)
    tables = re.findall(r"std::array\s+(\w+)\s*\{(.*?)\};", text, flags=re.S)
    out = bytearray(b"LRMI" + struct.pack("<III", 1, len(tables), 95))
    for name, body in tables:
        pairs = re.findall(r"make_float2\(\s*([-+0-9.eE]+)f?\s*,\s*([-+0-9.eE]+)f?\s*\)", body)
        assert len(pairs) == 95, (name, len(pairs))
        nk = np.array(pairs, dtype=np.float32)
        out += name.encode().ljust(8, b"\0") + nk[:, 0].tobytes() + nk[:, 1].tobytes()
        print(f"{name:5s} n(550nm) = {nk[38, 0]:.4f}  k(550nm) = {nk[38, 1]:.4f}")
    OUT.write_bytes(bytes(out))
    print(f"wrote {OUT} ({len(out)} bytes, {len(tables)} metals)")
    return 0


if __name__ == "__main__":
    sys.exit(main())
