"""PNG output for the sweep CLIs.  Uses matplotlib when it is installed (as the reference does,
run_fer_sweep.py:175-191); otherwise rasterises the same semilog-y curves with a few lines of NumPy and writes
the PNG with zlib, so the CLIs keep producing their plot file on boxes without matplotlib."""

from __future__ import annotations

import struct
import zlib
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

Series = Tuple[str, Sequence[float], Sequence[float]]
_COLORS = [(31, 119, 180), (255, 127, 14), (44, 160, 44), (214, 39, 40)]


def _png(path: Path, img: np.ndarray) -> None:
    h, w, _ = img.shape
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    path.write_bytes(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                     chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def _raster(series: List[Series], w: int = 900, h: int = 600) -> np.ndarray:
    img = np.full((h, w, 3), 255, np.uint8)
    xs = [x for _, sx, _ in series for x in sx]
    ys = [y for _, _, sy in series for y in sy if y and y > 0 and np.isfinite(y)]
    if not xs or not ys:
        return img
    x0, x1 = min(xs), max(xs)
    ly0, ly1 = np.floor(np.log10(min(ys))), np.ceil(np.log10(max(ys)))
    if x1 == x0:
        x1 = x0 + 1.0
    if ly1 == ly0:
        ly1 = ly0 + 1.0
    m = 60
    px = lambda x: int(m + (x - x0) / (x1 - x0) * (w - 2 * m))
    py = lambda y: int(h - m - (np.log10(y) - ly0) / (ly1 - ly0) * (h - 2 * m))
    img[m:h - m, m] = 0
    img[h - m, m:w - m] = 0
    for d in range(int(ly0), int(ly1) + 1):                # decade grid lines
        img[py(10.0 ** d), m:w - m:4] = 160
    for k, (_, sx, sy) in enumerate(series):
        col = _COLORS[k % len(_COLORS)]
        pts = [(px(x), py(y)) for x, y in zip(sx, sy) if y and y > 0 and np.isfinite(y)]
        for (ax, ay), (bx, by) in zip(pts, pts[1:]):
            n = max(abs(bx - ax), abs(by - ay), 1)
            for t in range(n + 1):
                xx, yy = ax + (bx - ax) * t // n, ay + (by - ay) * t // n
                img[max(yy - 1, 0):yy + 2, max(xx - 1, 0):xx + 2] = col
        for (ax, ay) in pts:
            img[max(ay - 4, 0):ay + 5, max(ax - 4, 0):ax + 5] = col
    return img


def semilogy_plot(path: Path, series: List[Series], xlabel: str, ylabel: str) -> None:
    path.parent.mkdir(parents=True, exist_ok=True)
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except Exception:
        _png(path, _raster(series))
        return
    marks = ["^-", "o-", "s-", "d-"]
    plt.figure(figsize=(6, 4))
    for k, (label, sx, sy) in enumerate(series):
        plt.semilogy(list(sx), list(sy), marks[k % len(marks)], label=label)
    plt.xlabel(xlabel)
    plt.ylabel(ylabel)
    plt.grid(True, which="both", ls="--", alpha=0.4)
    plt.legend()
    plt.tight_layout()
    plt.savefig(path, dpi=200)
    plt.close()
