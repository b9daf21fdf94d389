"""Minimal reader for R's `save()` files (RDX3 header + XDR serialisation, version 2/3) — just enough to pull
numeric matrices and Matrix::dgCMatrix objects out of the reference's `data/*.rda` without R.

TEST INFRASTRUCTURE ONLY (fixture generation, tests/golden/make_golden.py). Written from the published format
description in "R Internals", section 1.8 "Serialization Formats"; no R code involved.
"""
from __future__ import annotations

import bz2
import gzip
import lzma
import struct

import numpy as np

NILVALUE, REFSXP, PERSISTSXP, PACKAGESXP, NAMESPACESXP = 254, 255, 247, 248, 249
GLOBALENV, UNBOUND, MISSINGARG, BASENAMESPACE, EMPTYENV, BASEENV = 253, 252, 251, 250, 242, 241
ATTRLANGSXP, ATTRLISTSXP, ALTREP = 240, 239, 238
SYMSXP, LISTSXP, CLOSXP, ENVSXP, PROMSXP, LANGSXP, CHARSXP, LGLSXP, INTSXP, REALSXP, CPLXSXP, STRSXP = \
    1, 2, 3, 4, 5, 6, 9, 10, 13, 14, 15, 16
DOTSXP, VECSXP, EXPRSXP, BCODESXP, EXTPTRSXP, WEAKREFSXP, RAWSXP, S4SXP = 17, 19, 20, 21, 22, 23, 24, 25


class RObject:
    """A value with attributes (S4 objects: slots are attributes)."""

    def __init__(self, value, attrs=None, kind=""):
        self.value, self.attrs, self.kind = value, attrs or {}, kind

    def __repr__(self):
        return f"RObject({self.kind}, attrs={list(self.attrs)})"


class Symbol(str):
    pass


class _Reader:
    def __init__(self, buf: bytes):
        self.b, self.o, self.refs = buf, 0, []

    def int(self) -> int:
        v = struct.unpack_from(">i", self.b, self.o)[0]
        self.o += 4
        return v

    def length(self) -> int:
        n = self.int()
        if n == -1:                                   # long vector: two 32-bit halves
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        return n

    def bytes(self, n: int) -> bytes:
        v = self.b[self.o:self.o + n]
        self.o += n
        return v

    def array(self, dtype: str, n: int) -> np.ndarray:
        a = np.frombuffer(self.b, dtype=dtype, count=n, offset=self.o)
        self.o += a.nbytes
        return a.astype(a.dtype.newbyteorder("="))

    def pairlist(self, flags: int):
        """LISTSXP chain -> list of (tag, value); iterative over the cdr."""
        out = []
        while True:
            t = flags & 0xFF
            if t == NILVALUE:
                return out
            if t not in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
                out.append((None, self.item(flags)))          # dotted tail
                return out
            has_attr, has_tag = bool(flags & (1 << 9)), bool(flags & (1 << 10))
            if has_attr or t in (ATTRLISTSXP, ATTRLANGSXP):
                self.item()
            tag = self.item() if has_tag else None
            out.append((tag, self.item()))
            flags = self.int()

    def attributes(self) -> dict:
        return {str(k): v for k, v in self.pairlist(self.int())}

    def item(self, flags: int | None = None):
        if flags is None:
            flags = self.int()
        t = flags & 0xFF
        has_attr = bool(flags & (1 << 9))
        if t in (NILVALUE, 0):
            return None
        if t in (GLOBALENV, UNBOUND, MISSINGARG, BASENAMESPACE, EMPTYENV, BASEENV):
            return Symbol(f"<special {t}>")
        if t == REFSXP:
            idx = flags >> 8
            if idx == 0:
                idx = self.int()
            return self.refs[idx - 1]
        if t == SYMSXP:
            s = Symbol(self.item())
            self.refs.append(s)
            return s
        if t in (PERSISTSXP, PACKAGESXP, NAMESPACESXP):
            v = self.item() if t == PERSISTSXP else self._strvec_body()
            s = Symbol(f"<env {v}>")
            self.refs.append(s)
            return s
        if t == ENVSXP:
            env = RObject({}, kind="env")
            self.refs.append(env)
            self.int()                                         # locked
            for _ in range(4):                                 # enclos, frame, hashtab, attrib
                self.item()
            return env
        if t in (LISTSXP, LANGSXP, ATTRLISTSXP, ATTRLANGSXP):
            return self.pairlist(flags)
        if t in (CLOSXP, PROMSXP, DOTSXP):
            if has_attr:
                self.item()
            tag = self.item() if flags & (1 << 10) else None
            return RObject((tag, self.item(), self.item()), kind="closure")
        if t == CHARSXP:
            n = self.int()
            return None if n == -1 else self.bytes(n).decode("utf-8", "replace")
        if t == ALTREP:
            info, state, attr = self.item(), self.item(), self.item()
            cls = str(info[0][1]) if info else ""
            v = self._altrep(cls, state)
            if isinstance(attr, list) and attr:
                v = RObject(v, {str(k): a for k, a in attr}, kind="altrep")
            return v
        if t in (LGLSXP, INTSXP):
            v = self.array(">i4", self.length())
        elif t == REALSXP:
            v = self.array(">f8", self.length())
        elif t == CPLXSXP:
            v = self.array(">c16", self.length())
        elif t == RAWSXP:
            v = self.bytes(self.length())
        elif t == STRSXP:
            v = [self.item() for _ in range(self.length())]
        elif t in (VECSXP, EXPRSXP):
            v = [self.item() for _ in range(self.length())]
        elif t == S4SXP:
            v = None
        else:
            raise NotImplementedError(f"SEXPTYPE {t} at offset {self.o}")
        if has_attr:
            return RObject(v, self.attributes(), kind={S4SXP: "S4"}.get(t, "vector"))
        return v

    def _strvec_body(self):
        self.int()
        return [self.item() for _ in range(self.int())]

    @staticmethod
    def _altrep(cls: str, state):
        if cls == "compact_intseq":
            n, start, step = (float(x) for x in state)
            return (start + step * np.arange(int(n))).astype(np.int32)
        if cls == "compact_realseq":
            n, start, step = (float(x) for x in state)
            return start + step * np.arange(int(n))
        if cls.startswith("wrap_"):
            return state[0]
        if cls == "deferred_string":
            arg = state[0][1] if isinstance(state, list) and isinstance(state[0], tuple) else state
            vals = arg.value if isinstance(arg, RObject) else arg
            return [str(int(v)) if float(v).is_integer() else repr(float(v)) for v in np.asarray(vals)]
        raise NotImplementedError(f"ALTREP class {cls}")


def read_rda(path: str) -> dict:
    """Returns {name: object} for every object saved in an .rda / .RData file."""
    raw = open(path, "rb").read()
    for opener in (lzma.decompress, gzip.decompress, bz2.decompress):
        try:
            raw = opener(raw)
            break
        except Exception:
            continue
    assert raw[:5] in (b"RDX3\n", b"RDX2\n"), raw[:8]
    r = _Reader(raw[5:])
    assert r.bytes(2) == b"X\n", "only XDR serialisation is supported"
    version = r.int()
    r.int(); r.int()                                           # writer / minimal reader versions
    if version >= 3:
        r.bytes(r.int())                                       # native encoding
    return {str(k): v for k, v in r.pairlist(r.int())}


def as_csc(obj: RObject):
    """Matrix::dgCMatrix (slots i, p, Dim, x) -> (indptr, indices, data, (m, n))."""
    a = obj.attrs
    unwrap = lambda v: v.value if isinstance(v, RObject) else v
    dim = np.asarray(unwrap(a["Dim"]))
    return (np.asarray(unwrap(a["p"]), np.int32), np.asarray(unwrap(a["i"]), np.int32),
            np.asarray(unwrap(a["x"]), np.float64), (int(dim[0]), int(dim[1])))
