"""Writes luisarender_b200/data/sampler_tables.bin: the numerical tables of the quasi-Monte-Carlo samplers.

The PMJ02BN, Sobol', PaddedSobol and ZSobol samplers of the reference (src/samplers/{pmj02bn,sobol,padded_sobol,zsobol}.cpp) are
table-driven: pmj02bn sample sets, blue-noise textures, Sobol' generator matrices and the van-der-Corput matrices that map a
pixel to its Sobol' index.  The tables are DATA of the published algorithms - they come from pbrt-v4 (Copyright (c) 1998-2020
Matt Pharr, Wenzel Jakob, Greg Humphreys; Apache-2.0; the blue-noise textures from Christoph Peters, momentsingraphics.de/?p=127;
the Sobol' matrices from Joe & Kuo) and cannot be regenerated: pmj02bn sets are the outcome of a stochastic optimisation.  This
script reads the reference's copies where they lie (src/util/{pmj02tables,bluenoise,sobolmatrices}.cpp) and stores the numbers
as one binary file; no source text is copied.

    python tools/extract_sampler_tables.py [/root/reference]

File layout (little endian): magic "LRST", u32 version = 1, then five sections, each {u32 tag, u32 element size, u64 count, data}:
  1 SobolMatrices32     u32[1024 * 52]          2 VdCSobolMatrices  u64[25 * 52]     3 VdCSobolMatricesInv u64[26 * 52]
  4 PMJ02bnSamples      u32[5 * 65536 * 2]      5 BlueNoiseTextures u16[48 * 128 * 128]
"""
from __future__ import annotations

import re
import struct
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[1]
OUT = REPO / "luisarender_b200" / "data" / "sampler_tables.bin"


def initializer(text: str, marker: str) -> str:
    """The brace-balanced initialiser that follows `marker`, comments removed."""
    start = text.index("{", text.index(marker))
    depth, i = 0, start
    while True:
        c = text[i]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    body = text[start:i + 1]
    return re.sub(r"//[^\n]*", "", body)


def literals(body: str) -> list[int]:
    return [int(t, 16) if t[:2] in ("0x", "0X") else int(t) for t in re.findall(r"0[xX][0-9a-fA-F]+|\d+", body)]


def flat_table(text: str, marker: str, count: int, dtype) -> np.ndarray:
    vals = literals(initializer(text, marker))
    assert len(vals) == count, (marker, len(vals), count)
    return np.array(vals, dtype=np.uint64).astype(dtype)


def ragged_rows(text: str, marker: str, rows: int, cols: int, dtype) -> np.ndarray:
    """A 2D array whose rows list fewer than `cols` values (the rest is zero, as C aggregate initialisation makes it)."""
    body = initializer(text, marker)
    inner = re.findall(r"\{([^{}]*)\}", body[1:-1])
    assert len(inner) == rows, (marker, len(inner), rows)
    out = np.zeros((rows, cols), dtype=np.uint64)
    for r, row in enumerate(inner):
        v = literals(row)
        assert len(v) <= cols
        out[r, :len(v)] = v
    return out.astype(dtype).reshape(-1)


def main() -> int:
    ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    util = ref / "src" / "util"
    sobol = (util / "sobolmatrices.cpp").read_text()
    pmj = (util / "pmj02tables.cpp").read_text()
    bn = (util / "bluenoise.cpp").read_text()
    sections = [
        (1, flat_table(sobol, "SobolMatrices32[NSobolDimensions * SobolMatrixSize] =", 1024 * 52, np.uint32)),
        (2, ragged_rows(sobol, "VdCSobolMatrices[VdCSobolMatrixSize][SobolMatrixSize] =", 25, 52, np.uint64)),
        (3, ragged_rows(sobol, "VdCSobolMatricesInv[VdCSobolMatrixInvSize][SobolMatrixSize] =", 26, 52, np.uint64)),
        (4, flat_table(pmj, "PMJ02bnSamples[nPMJ02bnSets][nPMJ02bnSamples][2] =", 5 * 65536 * 2, np.uint32)),
        (5, flat_table(bn, "BlueNoiseTextures[NumBlueNoiseTextures][BlueNoiseResolution][BlueNoiseResolution] =", 48 * 128 * 128, np.uint16)),
    ]
    OUT.parent.mkdir(parents=True, exist_ok=True)
    with open(OUT, "wb") as f:
        f.write(b"LRST" + struct.pack("<I", 1))
        for tag, arr in sections:
            f.write(struct.pack("<IIQ", tag, arr.dtype.itemsize, arr.size))
            f.write(arr.tobytes())
    print(f"wrote {OUT} ({OUT.stat().st_size} bytes)")
    for tag, arr in sections:
        print(tag, arr.dtype, arr.size, arr[:4], int(arr.astype(np.uint64).sum() % (1 << 32)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
