"""Minimal reader for R ``.rda`` fixtures (RDX2 / XDR serialisation, version 2).

Only what the hot path's fixtures need: data frames of
INTSXP / REALSXP columns with ``names`` / ``row.names`` / ``class`` / ``levels``
attributes, as written by ``save()`` for ``data/es.mef.small.rda`` (xz) and
``data/o.ifm.rda`` / ``data/knn.rda`` (bz2).  Standard library + numpy only.

Format reference: R Internals, "Serialization Formats" (public documentation);
nothing here is derived from the reference package's sources, which only ship
the binary files.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct
from typing import Any

import numpy as np

_NILVALUE, _GLOBALENV, _EMPTYENV, _BASEENV = 254, 253, 242, 241
_REFSXP, _NAMESPACESXP, _PACKAGESXP, _PERSISTSXP = 255, 249, 250, 247
_MISSINGARG, _UNBOUND, _BASENAMESPACE = 251, 252, 247
_ALTREP = 238
NA_INTEGER = -2147483648


class RObject:
    """A decoded R vector plus its attribute dict."""

    __slots__ = ("value", "attributes")

    def __init__(self, value: Any, attributes: dict | None = None):
        self.value = value
        self.attributes = attributes or {}

    def __repr__(self) -> str:  # pragma: no cover - debugging aid
        return f"RObject({type(self.value).__name__}, attrs={list(self.attributes)})"


class _Reader:
    def __init__(self, buf: bytes):
        self.buf = buf
        self.pos = 0
        self.refs: list[Any] = []

    def _int(self) -> int:
        v = struct.unpack_from(">i", self.buf, self.pos)[0]
        self.pos += 4
        return v

    def _length(self) -> int:
        n = self._int()
        if n == -1:  # long vector
            hi, lo = self._int(), self._int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def _bytes(self, n: int) -> bytes:
        b = self.buf[self.pos:self.pos + n]
        self.pos += n
        return b

    def item(self) -> Any:
        flags = self._int()
        sxp = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if sxp == _NILVALUE:
            return None
        if sxp in (_GLOBALENV, _EMPTYENV, _BASEENV, _MISSINGARG, _UNBOUND):
            return None
        if sxp == _REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self._int()
            return self.refs[idx - 1]
        if sxp == 1:  # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if sxp in (2, 6):  # LISTSXP / LANGSXP -> python list of (tag, value)
            out = []
            while True:
                attrs = self.item() if has_attr else None  # noqa: F841
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self._int()
                nxt = flags & 0xFF
                if nxt == _NILVALUE:
                    break
                if nxt not in (2, 6):
                    raise ValueError(f"unexpected pairlist tail type {nxt}")
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if sxp == 9:  # CHARSXP
            n = self._int()
            return None if n == -1 else self._bytes(n).decode("utf-8", "replace")
        if sxp == 10:  # LGLSXP
            n = self._length()
            val = np.frombuffer(self._bytes(4 * n), dtype=">i4").astype(np.int32)
        elif sxp == 13:  # INTSXP
            n = self._length()
            val = np.frombuffer(self._bytes(4 * n), dtype=">i4").astype(np.int32)
        elif sxp == 14:  # REALSXP
            n = self._length()
            val = np.frombuffer(self._bytes(8 * n), dtype=">f8").astype(np.float64)
        elif sxp == 16:  # STRSXP
            n = self._length()
            val = [self.item() for _ in range(n)]
        elif sxp == 19:  # VECSXP
            n = self._length()
            val = [self.item() for _ in range(n)]
        else:
            raise ValueError(f"unsupported SEXP type {sxp} at byte {self.pos}")
        attrs = {}
        if has_attr:
            for tag, car in self.item() or []:
                attrs[tag] = car
        return RObject(val, attrs)


def _decompress(raw: bytes) -> bytes:
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    return raw


def read_rda(path: str) -> dict[str, Any]:
    """Return ``{object name: decoded RObject}`` for an ``.rda`` written by ``save()``."""
    with open(path, "rb") as fh:
        buf = _decompress(fh.read())
    if buf[:5] != b"RDX2\n" or buf[5:7] != b"X\n":
        raise ValueError(f"{path}: not an RDX2/XDR file")
    rd = _Reader(buf)
    rd.pos = 7
    version, _writer, _minreader = rd._int(), rd._int(), rd._int()
    if version != 2:
        raise ValueError(f"{path}: serialisation version {version} unsupported")
    top = rd.item()
    return {tag: car for tag, car in top}


def _row_names(obj: RObject, nrow: int) -> list[str]:
    rn = obj.attributes.get("row.names")
    if rn is None:
        return [str(i + 1) for i in range(nrow)]
    v = rn.value
    if isinstance(v, list):
        return list(v)
    if len(v) == 2 and v[0] == NA_INTEGER:  # compact c(NA, -n)
        return [str(i + 1) for i in range(abs(int(v[1])))]
    return [str(int(x)) for x in v]


def as_data_frame(obj: RObject):
    """Convert a decoded R data.frame to ``pandas.DataFrame`` (attributes kept in ``.attrs``)."""
    import pandas as pd

    names = obj.attributes["names"].value
    cols = {}
    for name, col in zip(names, obj.value):
        v = col.value
        if "levels" in col.attributes:  # factor
            lev = col.attributes["levels"].value
            v = pd.Categorical.from_codes(np.asarray(v) - 1, categories=lev)
        cols[name] = v
    nrow = len(next(iter(cols.values()))) if cols else 0
    df = pd.DataFrame(cols, index=_row_names(obj, nrow))
    for k, a in obj.attributes.items():
        if k in ("names", "row.names", "class"):
            continue
        if isinstance(a, RObject) and "levels" in a.attributes:
            lev = a.attributes["levels"].value
            df.attrs[k] = pd.Categorical.from_codes(np.asarray(a.value) - 1, categories=lev)
        elif isinstance(a, RObject):
            df.attrs[k] = a.value
    return df
