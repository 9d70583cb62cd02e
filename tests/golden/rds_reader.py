"""Minimal reader for R's XDR serialisation format (RDS version 2/3).

Test infrastructure only.  It decodes just the SEXP types that occur in the
golden tables shipped with the reference data package
(ChicdiffData/inst/extdata/CD4_Mono_results/test_results.Rds, test_settings.Rds):
lists, atomic vectors, strings, pairlist attributes, symbols and the external
pointer that data.table stores as ``.internal.selfref``.
"""
import gzip
import struct
import numpy as np

NA_INT = -2147483648


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    def i32(self):
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n):
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self):
        n = self.i32()
        if n == -1:
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + lo
        return n

    def item(self):
        flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:            # NILVALUE_SXP
            return None
        if t == 253:            # GLOBALENV
            return "<globalenv>"
        if t == 242:            # EMPTYENV
            return "<emptyenv>"
        if t == 255:            # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t == 1:              # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t == 2 or t == 6:    # LISTSXP / LANGSXP -> python list of (tag, value)
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t2 = flags & 0xFF
                if t2 == 254:
                    break
                if t2 not in (2, 6):
                    raise ValueError("unexpected pairlist tail type %d" % t2)
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if t == 9:              # CHARSXP
            n = self.i32()
            if n == -1:
                return None
            return self.take(n).decode("utf-8", "replace")
        if t == 22:             # EXTPTRSXP
            obj = {"extptr": True}
            self.refs.append(obj)
            self.item()
            self.item()
            if has_attr:
                self.item()
            return obj
        if t == 10 or t == 13:  # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:           # REALSXP
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == 16:           # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t == 19 or t == 20:  # VECSXP / EXPRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        else:
            raise ValueError("unsupported SEXP type %d at offset %d" % (t, self.p))
        attrs = None
        if has_attr:
            attrs = {k: val for k, val in self.item()}
        return RObj(v, attrs) if attrs else v


class RObj:
    """A value with R attributes."""

    def __init__(self, value, attrs):
        self.value = value
        self.attrs = attrs

    def __repr__(self):
        return "RObj(%r, attrs=%r)" % (type(self.value), list(self.attrs))


def read_rds(path):
    raw = open(path, "rb").read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:2] != b"X\n":
        raise ValueError("only XDR RDS supported")
    r = _Reader(raw)
    r.p = 2
    version = r.i32()
    r.i32()
    r.i32()
    if version == 3:
        n = r.i32()
        r.take(n)
    return r.item()


def unwrap(x):
    return x.value if isinstance(x, RObj) else x


def data_frame(obj):
    """RObj(list of columns) -> dict name -> numpy array / list."""
    names = unwrap(obj.attrs["names"])
    return {nm: unwrap(col) for nm, col in zip(names, obj.value)}
