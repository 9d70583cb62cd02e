"""Test tooling: a minimal writer of R's XDR serialisation (version 3), enough to build the objects chicdiff_b200.rds is
meant to read -- a chicagoData-like S4 object with a data.table in its x slot -- so that the reader can be exercised
beyond the two .Rds files the reference ships (tests/test_golden.py reads those).  Follows src/main/serialize.c."""
import gzip
import struct

import numpy as np

NA_INT = -2147483648


class W:
    def __init__(self):
        self.out = [b"X\n", struct.pack(">iii", 3, 0x00040300, 0x00030500), struct.pack(">i", 5), b"UTF-8"]
        self.syms = {}

    def i32(self, v):
        self.out.append(struct.pack(">i", v))

    def flags(self, t, obj=False, attr=False, tag=False, levels=0):
        self.i32(t | (levels << 12) | (0x100 if obj else 0) | (0x200 if attr else 0) | (0x400 if tag else 0))

    def null(self):
        self.i32(254)

    def charsxp(self, s):
        if s is None:
            self.flags(9); self.i32(-1)
            return
        b = s.encode("utf-8")
        self.flags(9, levels=(1 << 6) if all(c < 128 for c in b) else (1 << 3))
        self.i32(len(b)); self.out.append(b)

    def sym(self, name):
        if name in self.syms:
            self.i32(255 | (self.syms[name] << 8))
            return
        self.flags(1)
        self.charsxp(name)
        self.syms[name] = len(self.syms) + 1          # reference indices are shared with environments; none are written here

    def attrs(self, pairs):
        for k, emit in pairs:
            self.flags(2, tag=True)
            self.sym(k)
            emit()
        self.null()

    def ints(self, v, attrs=None, obj=False, logical=False):
        v = np.asarray(v, dtype=np.int32)
        self.flags(10 if logical else 13, obj=obj, attr=bool(attrs))
        self.i32(len(v)); self.out.append(v.astype(">i4").tobytes())
        if attrs:
            self.attrs(attrs)

    def reals(self, v, attrs=None):
        v = np.asarray(v, dtype=np.float64)
        self.flags(14, attr=bool(attrs))
        self.i32(len(v)); self.out.append(v.astype(">f8").tobytes())
        if attrs:
            self.attrs(attrs)

    def strs(self, v, attrs=None):
        self.flags(16, attr=bool(attrs))
        self.i32(len(v))
        for s in v:
            self.charsxp(s)
        if attrs:
            self.attrs(attrs)

    def veclist(self, emitters, attrs=None, obj=False):
        self.flags(19, obj=obj, attr=bool(attrs))
        self.i32(len(emitters))
        for e in emitters:
            e()
        if attrs:
            self.attrs(attrs)

    def compact_intseq(self, n, first=1, step=1):
        """ALTREP compact integer sequence (how R >= 3.5 writes 1:n, e.g. row names): info, state, attributes"""
        self.flags(238)
        self.flags(2); self.sym("compact_intseq")
        self.flags(2); self.sym("base")
        self.flags(2); self.ints([13]); self.null()
        self.reals([float(n), float(first), float(step)])
        self.null()

    def deferred_string(self, ints):
        """ALTREP deferred string: as.character(<integer vector>) not yet materialised; state = CONS(arg, scipen)"""
        self.flags(238)
        self.flags(2); self.sym("deferred_string")
        self.flags(2); self.sym("base")
        self.flags(2); self.ints([16]); self.null()
        self.flags(2); self.ints(ints); self.ints([0])            # dotted pair: the cdr is the scipen scalar
        self.null()

    def bytes(self):
        return b"".join(self.out)


def data_table(w, columns, altrep_rownames=True):
    """columns: list of (name, kind, values) with kind in int / real / lgl / str / factor (values, levels)"""
    names = [c[0] for c in columns]
    n = len(columns[0][2][0]) if columns[0][1] == "factor" else len(columns[0][2])

    def col(kind, vals):
        if kind == "int":
            return lambda: w.ints(vals)
        if kind == "lgl":
            return lambda: w.ints(vals, logical=True)
        if kind == "real":
            return lambda: w.reals(vals)
        if kind == "str":
            return lambda: w.strs(vals)
        if kind == "factor":
            codes, levels = vals
            return lambda: w.ints(codes, obj=True, attrs=[("levels", lambda: w.strs(levels)), ("class", lambda: w.strs(["factor"]))])
        if kind == "deferred":
            return lambda: w.deferred_string(vals)
        raise ValueError(kind)
    rn = (lambda: w.compact_intseq(n)) if altrep_rownames else (lambda: w.ints([NA_INT, -n]))
    w.veclist([col(k, v) for _, k, v in columns], obj=True,
              attrs=[("names", lambda: w.strs(names)), ("row.names", rn),
                     ("class", lambda: w.strs(["data.table", "data.frame"]))])


def chicago_data(path, columns, params, settings, compress=True):
    """an S4 object of class chicagoData with slots x (data.table), params, settings (named lists of scalars)"""
    w = W()

    def named(d):
        def emit():
            ems = []
            for v in d.values():
                if isinstance(v, str):
                    ems.append(lambda v=v: w.strs([v]))
                elif isinstance(v, bool):
                    ems.append(lambda v=v: w.ints([int(v)], logical=True))
                elif isinstance(v, int):
                    ems.append(lambda v=v: w.ints([v]))
                else:
                    ems.append(lambda v=v: w.reals([v]))
            w.veclist(ems, attrs=[("names", lambda: w.strs(list(d.keys())))])
        return emit
    w.flags(25, obj=True, attr=True, levels=1 << 4)             # S4SXP, S4 bit set
    w.attrs([("x", lambda: data_table(w, columns)), ("params", named(params)), ("settings", named(settings)),
             ("class", lambda: w.strs(["chicagoData"], attrs=[("package", lambda: w.strs(["Chicago"]))]))])
    raw = w.bytes()
    with open(path, "wb") as fh:
        fh.write(gzip.compress(raw) if compress else raw)
