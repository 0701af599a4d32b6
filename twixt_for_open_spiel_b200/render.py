"""Host-side board renderer: the ASCII/ANSI picture the reference returns from
`ToString()` / `ObservationString()` / `InformationStateString()`
(twixtboard.cc:278-448, twixt.h:65-75), produced from an exported state record
(include/twixt_b200.h, "State record").  String formatting is not data-parallel
work, so it stays on the host; it is SURVEY section 8(f) row 1 and makes the
playthrough strings reproducible byte for byte.

Layout of the picture: a header line of column letters, then three text lines
per board row (top row first): the line above the pegs, the peg line, the line
below.  Every cell contributes three character SLOTS per line; a slot shows the
link glyphs of the links passing through it, or one space when there are none.
The tables below list, per slot, which (cell offset, direction) links draw which
glyph.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

ANSI_RED = "\x1b[91m"
ANSI_BLUE = "\x1b[94m"
ANSI_DEFAULT = "\x1b[0m"

NNE, ENE, ESE, SSE, SSW, WSW, WNW, NNW = range(8)
_DX = (1, 2, 2, 1, -1, -2, -2, -1)
_DY = (2, 1, -1, -2, -2, -1, 1, 2)
HEADER_WORDS = 4
P_RED, P_BLUE, P_LINK0 = 0, 1, 2

# slot tables: ((dx, dy, direction, glyph), ...) drawn in this order; "fallback"
# entries are tried, in order, only while the slot is still empty
_BEFORE = (
    (((-1, 0, ENE, "/"), (-1, -1, NNE, "/"), (0, 0, WNW, "_")), ()),
    (((0, 0, NNE, "|"),), ((0, 0, NNW, "|"),)),
    (((1, 0, WNW, "\\"), (1, -1, NNW, "\\"), (0, 0, ENE, "_")), ()),
)
_PEG_LEFT = (((-1, -1, NNE, "|"), (0, 0, WSW, "_")), ())
_PEG_RIGHT = (((1, -1, NNW, "|"), (0, 0, ESE, "_")), ())
_AFTER = (
    (((1, -1, WNW, "\\"), (0, -1, NNW, "\\")), ()),
    (((-1, -1, ENE, "_"), (1, -1, WNW, "_"), (0, 0, SSW, "|")), ((0, 0, SSE, "|"),)),
    (((-1, -1, ENE, "/"), (0, -1, NNE, "/")), ()),
)


class _Board:
    """Cell colours and 8-direction link masks decoded from a record."""

    def __init__(self, record: Sequence[int], n: int):
        rec = np.asarray(record, dtype=np.uint32)
        self.n = n
        pl = rec[HEADER_WORDS:HEADER_WORDS + 9 * n].reshape(9, n)
        self.ply = int(rec[0])
        self.result = int(rec[1]) & 3
        self.swapped = bool((int(rec[1]) >> 2) & 1)
        self.color = [[2] * n for _ in range(n)]  # 0 red, 1 blue, 2 empty (twixtboard.h:50)
        self.links = [[0] * n for _ in range(n)]
        for x in range(n):
            for y in range(n):
                if (int(pl[P_RED, x]) >> y) & 1:
                    self.color[x][y] = 0
                elif (int(pl[P_BLUE, x]) >> y) & 1:
                    self.color[x][y] = 1
                for d in range(4):  # stored once at the west endpoint; mirror to the east endpoint
                    if (int(pl[P_LINK0 + d, x]) >> y) & 1:
                        self.links[x][y] |= 1 << d
                        self.links[x + _DX[d]][y + _DY[d]] |= 1 << (d + 4)

    def off_board(self, x: int, y: int) -> bool:
        n = self.n
        return x < 0 or y < 0 or x >= n or y >= n or ((x in (0, n - 1)) and (y in (0, n - 1)))


def _paint(ansi: bool, color: str, text: str) -> str:
    return (color + text + ANSI_DEFAULT) if ansi else text


def _link_glyph(b: _Board, ansi: bool, x: int, y: int, d: int, glyph: str) -> str:
    if b.off_board(x, y) or not (b.links[x][y] >> d) & 1:
        return ""
    c = b.color[x][y]
    if c == 0:
        return _paint(ansi, ANSI_RED, glyph)
    if c == 1:
        return _paint(ansi, ANSI_BLUE, glyph)
    return glyph


def _slot(b: _Board, ansi: bool, x: int, y: int, spec) -> str:
    always, fallback = spec
    s = "".join(_link_glyph(b, ansi, x + dx, y + dy, d, g) for dx, dy, d, g in always)
    for dx, dy, d, g in fallback:
        if s:
            break
        s = _link_glyph(b, ansi, x + dx, y + dy, d, g)
    return s or " "


def _peg(b: _Board, ansi: bool, x: int, y: int) -> str:
    n = b.n
    c = b.color[x][y]
    if c == 0:
        return _paint(ansi, ANSI_RED, "x")
    if c == 1:
        return _paint(ansi, ANSI_BLUE, "o")
    if b.off_board(x, y):
        return " "
    if x in (0, n - 1):
        return _paint(ansi, ANSI_BLUE, ".")
    if y in (0, n - 1):
        return _paint(ansi, ANSI_RED, ".")
    return "."


def board_to_string(record: Sequence[int], board_size: int, ansi_color_output: bool = True) -> str:
    b = _Board(record, board_size)
    n, ansi = board_size, ansi_color_output
    out: List[str] = ["     "]
    for x in range(n):
        out.append(_paint(ansi, ANSI_RED, chr(ord("a") + x) + "  "))
    out.append("\n")
    for y in range(n - 1, -1, -1):
        out.append("    ")
        for x in range(n):
            out.extend(_slot(b, ansi, x, y, spec) for spec in _BEFORE)
        out.append("\n")
        out.append("  " if n - y < 10 else " ")
        out.append(_paint(ansi, ANSI_BLUE, str(n - y) + " "))
        for x in range(n):
            out.append(_slot(b, ansi, x, y, _PEG_LEFT))
            out.append(_peg(b, ansi, x, y))
            out.append(_slot(b, ansi, x, y, _PEG_RIGHT))
        out.append("\n")
        out.append("    ")
        for x in range(n):
            out.extend(_slot(b, ansi, x, y, spec) for spec in _AFTER)
        out.append("\n")
    out.append("\n")
    if b.swapped:
        out.append("[swapped]")
    out.append({0: "", 1: "[x has won]", 2: "[o has won]", 3: "[draw]"}[b.result])
    return "".join(out)
