"""Dependency-free line plots written as PNG (matplotlib is not in the image).

tools/benchmark.py of the reference draws its figures with matplotlib
(``time_loss.png``, ``time_step.png``, ``time_precision.png``, ``step_losses.png``,
``time_cdfs.png``; tools/benchmark.py:165-263).  This module rasterises the same
kind of figure -- multiple labelled series, linear or log y axis, ticks, legend --
with numpy + zlib only.  If matplotlib is importable it is used instead.
"""
from __future__ import annotations

import math
import struct
import zlib
from typing import List, Sequence, Tuple

import numpy as np

# 5x7 bitmap font (digits, lower-case letters, a few symbols); each glyph = 7 rows of 5 bits.
_FONT = {
    "0": "01110 10001 10011 10101 11001 10001 01110", "1": "00100 01100 00100 00100 00100 00100 01110",
    "2": "01110 10001 00001 00010 00100 01000 11111", "3": "11110 00001 00001 01110 00001 00001 11110",
    "4": "00010 00110 01010 10010 11111 00010 00010", "5": "11111 10000 11110 00001 00001 10001 01110",
    "6": "00110 01000 10000 11110 10001 10001 01110", "7": "11111 00001 00010 00100 01000 01000 01000",
    "8": "01110 10001 10001 01110 10001 10001 01110", "9": "01110 10001 10001 01111 00001 00010 01100",
    ".": "00000 00000 00000 00000 00000 01100 01100", "-": "00000 00000 00000 11111 00000 00000 00000",
    "+": "00000 00100 00100 11111 00100 00100 00000", "e": "00000 00000 01110 10001 11111 10000 01110",
    "(": "00010 00100 01000 01000 01000 00100 00010", ")": "01000 00100 00010 00010 00010 00100 01000",
    "%": "11001 11010 00010 00100 01000 01011 10011", "_": "00000 00000 00000 00000 00000 00000 11111",
    "=": "00000 00000 11111 00000 11111 00000 00000", "/": "00001 00010 00010 00100 01000 01000 10000",
    " ": "00000 00000 00000 00000 00000 00000 00000", "<": "00010 00100 01000 10000 01000 00100 00010",
    "a": "00000 00000 01110 00001 01111 10001 01111", "b": "10000 10000 10110 11001 10001 10001 11110",
    "c": "00000 00000 01110 10000 10000 10001 01110", "d": "00001 00001 01101 10011 10001 10001 01111",
    "f": "00110 01001 01000 11100 01000 01000 01000", "g": "00000 01111 10001 10001 01111 00001 01110",
    "h": "10000 10000 10110 11001 10001 10001 10001", "i": "00100 00000 01100 00100 00100 00100 01110",
    "j": "00010 00000 00110 00010 00010 10010 01100", "k": "10000 10000 10010 10100 11000 10100 10010",
    "l": "01100 00100 00100 00100 00100 00100 01110", "m": "00000 00000 11010 10101 10101 10001 10001",
    "n": "00000 00000 10110 11001 10001 10001 10001", "o": "00000 00000 01110 10001 10001 10001 01110",
    "p": "00000 11110 10001 10001 11110 10000 10000", "q": "00000 01101 10011 10001 01111 00001 00001",
    "r": "00000 00000 10110 11001 10000 10000 10000", "s": "00000 00000 01110 10000 01110 00001 11110",
    "t": "01000 01000 11100 01000 01000 01001 00110", "u": "00000 00000 10001 10001 10001 10011 01101",
    "v": "00000 00000 10001 10001 10001 01010 00100", "w": "00000 00000 10001 10001 10101 10101 01010",
    "x": "00000 00000 10001 01010 00100 01010 10001", "y": "00000 10001 10001 01111 00001 10001 01110",
    "z": "00000 00000 11111 00010 00100 01000 11111",
}
_PALETTE = [(31, 119, 180), (255, 127, 14), (44, 160, 44), (214, 39, 40), (148, 103, 189), (140, 86, 75),
            (227, 119, 194), (127, 127, 127), (188, 189, 34), (23, 190, 207), (0, 0, 128), (128, 0, 0)]


class Canvas:
    def __init__(self, w: int, h: int):
        self.w, self.h = w, h
        self.px = np.full((h, w, 3), 255, np.uint8)

    def text(self, x: int, y: int, s: str, color=(0, 0, 0), scale: int = 1) -> None:
        for ch in s.lower():
            g = _FONT.get(ch, _FONT[" "]).split()
            for r, row in enumerate(g):
                for c, bit in enumerate(row):
                    if bit == "1":
                        y0, x0 = y + r * scale, x + c * scale
                        if 0 <= y0 < self.h - scale and 0 <= x0 < self.w - scale:
                            self.px[y0:y0 + scale, x0:x0 + scale] = color
            x += 6 * scale

    def line(self, x0: float, y0: float, x1: float, y1: float, color, width: int = 1) -> None:
        n = int(max(abs(x1 - x0), abs(y1 - y0), 1)) + 1
        xs = np.clip(np.round(np.linspace(x0, x1, n)).astype(int), 0, self.w - 1)
        ys = np.clip(np.round(np.linspace(y0, y1, n)).astype(int), 0, self.h - 1)
        for d in range(width):
            self.px[np.clip(ys + d, 0, self.h - 1), xs] = color
            self.px[ys, np.clip(xs + d, 0, self.w - 1)] = color

    def png(self) -> bytes:
        raw = b"".join(b"\x00" + self.px[r].tobytes() for r in range(self.h))

        def chunk(tag: bytes, data: bytes) -> bytes:
            return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", self.w, self.h, 8, 2, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def _ticks(lo: float, hi: float, n: int = 6) -> List[float]:
    if hi <= lo:
        hi = lo + 1.0
    step = 10 ** math.floor(math.log10((hi - lo) / n))
    for m in (1, 2, 5, 10):
        if (hi - lo) / (step * m) <= n:
            step *= m
            break
    t0 = math.ceil(lo / step) * step
    return [t0 + i * step for i in range(int((hi - t0) / step) + 1)]


def _fmt(v: float) -> str:
    if v == 0:
        return "0"
    if abs(v) >= 1e4 or abs(v) < 1e-2:
        return ("%.1e" % v).replace("e+0", "e").replace("e-0", "e-")
    return ("%.3f" % v).rstrip("0").rstrip(".")


def line_plot(path: str, series: Sequence[Tuple[str, Sequence[float], Sequence[float]]], xlabel: str = "",
              ylabel: str = "", logy: bool = False, title: str = "", size: Tuple[int, int] = (800, 520)) -> str:
    """``series`` = [(label, xs, ys), ...] -> PNG at ``path``."""
    try:  # pragma: no cover - matplotlib is optional
        import matplotlib
        matplotlib.use("Agg")
        from matplotlib import pyplot as plt
        plt.figure(figsize=(size[0] / 100, size[1] / 100))
        for label, xs, ys in series:
            plt.plot(xs, ys, label=label)
        if logy:
            plt.yscale("log")
        plt.xlabel(xlabel), plt.ylabel(ylabel), plt.title(title)
        if series:
            plt.legend(fontsize=8)
        plt.savefig(path)
        plt.close()
        return path
    except ImportError:
        pass
    W, H = size
    c = Canvas(W, H)
    L, R, T, Bm = 70, 20, 30, 50
    pts = [(np.asarray(xs, float), np.asarray(ys, float)) for _, xs, ys in series]
    pts = [(x, np.log10(np.maximum(y, 1e-12)) if logy else y) for x, y in pts]
    allx = np.concatenate([x for x, _ in pts]) if pts and any(len(x) for x, _ in pts) else np.array([0.0, 1.0])
    ally = np.concatenate([y for _, y in pts]) if pts and any(len(y) for _, y in pts) else np.array([0.0, 1.0])
    x0, x1, y0, y1 = float(allx.min()), float(allx.max()), float(ally.min()), float(ally.max())
    if x1 <= x0:
        x1 = x0 + 1
    if y1 <= y0:
        y1 = y0 + 1
    pad = 0.05 * (y1 - y0)
    y0, y1 = y0 - pad, y1 + pad

    def X(v):
        return L + (v - x0) / (x1 - x0) * (W - L - R)

    def Y(v):
        return H - Bm - (v - y0) / (y1 - y0) * (H - T - Bm)
    c.line(L, H - Bm, W - R, H - Bm, (0, 0, 0))
    c.line(L, T, L, H - Bm, (0, 0, 0))
    for t in _ticks(x0, x1):
        c.line(X(t), H - Bm, X(t), H - Bm + 4, (0, 0, 0))
        c.text(int(X(t)) - 3 * len(_fmt(t)), H - Bm + 8, _fmt(t))
    for t in _ticks(y0, y1):
        c.line(L - 4, Y(t), L, Y(t), (0, 0, 0))
        c.line(L, Y(t), W - R, Y(t), (225, 225, 225))
        lab = _fmt(10 ** t) if logy else _fmt(t)
        c.text(max(L - 8 - 6 * len(lab), 0), int(Y(t)) - 3, lab)
    for i, ((label, _, _), (x, y)) in enumerate(zip(series, pts)):
        col = _PALETTE[i % len(_PALETTE)]
        for j in range(len(x) - 1):
            c.line(X(x[j]), Y(y[j]), X(x[j + 1]), Y(y[j + 1]), col, width=2)
        if len(x) == 1:
            c.line(X(x[0]) - 2, Y(y[0]), X(x[0]) + 2, Y(y[0]), col, width=3)
        c.line(W - R - 190, T + 6 + 12 * i, W - R - 175, T + 6 + 12 * i, col, width=2)
        c.text(W - R - 170, T + 3 + 12 * i, label[:27])
    c.text(L + (W - L - R) // 2 - 3 * len(xlabel), H - 22, xlabel)
    c.text(4, 8, ylabel)
    c.text(L + (W - L - R) // 2 - 3 * len(title), 8, title)
    with open(path, "wb") as f:
        f.write(c.png())
    return path
