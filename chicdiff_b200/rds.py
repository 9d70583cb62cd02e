"""Reader for R's serialisation format (.Rds, XDR, versions 2 and 3; gzip / bzip2 / xz or plain) -- the input codec of
setChicdiffExperiment()'s CHiCAGO objects and peak matrices (chicdiff.R:517-534, 614-623: `readRDS(file)`, then `@x` if the
object is a chicagoData).  Host code: decompression and XDR decoding are byte work on the CPU; the atomic vectors, which
are all but a few hundred bytes of such a file, are decoded in bulk with NumPy (one byte swap per column), so reading
costs about what gunzip costs.  Together with chicdiff_b200.cache (Arrow files that are memory-mapped on the next run)
this is the "input codecs" row of the survey's next steps.

What is decoded: NULL, symbols, pairlists (with attributes and dotted tails), language objects, environments, atomic
vectors (logical, integer, double, complex, character, raw), lists, expression vectors, S4 objects, external pointers,
references, the special environments, byte code is refused, and the ALTREP classes base R writes (compact integer / real
sequences, deferred strings, the wrap_* classes).  Values with attributes come back as RObj(value, attrs); data_frame()
and chicago_table() turn the usual containers into dicts of NumPy columns.
"""
import bz2
import gzip
import lzma
import struct

import numpy as np

NA_INTEGER = -2147483648
NA_LOGICAL = NA_INTEGER


class RObj:
    """A value with R attributes (names, class, levels, dim, slots of an S4 object ...)."""

    def __init__(self, value, attrs, kind=None):
        self.value = value
        self.attrs = attrs
        self.kind = kind            # "S4" for S4 objects, "env" for environments, None otherwise

    def __repr__(self):
        return "RObj(%s, attrs=%r)" % (self.kind or type(self.value).__name__, list(self.attrs))

    def klass(self):
        c = self.attrs.get("class")
        return list(unwrap(c)) if c is not None else []


class RdsError(ValueError):
    pass


def unwrap(x):
    return x.value if isinstance(x, RObj) else x


class _Reader:
    def __init__(self, buf):
        self.b = buf
        self.p = 0
        self.refs = []

    # -- primitives ---------------------------------------------------------------------------
    def i32(self):
        if self.p + 4 > len(self.b):
            raise RdsError("truncated file")
        v = struct.unpack_from(">i", self.b, self.p)[0]
        self.p += 4
        return v

    def take(self, n):
        if n < 0 or self.p + n > len(self.b):
            raise RdsError("truncated file")
        v = self.b[self.p:self.p + n]
        self.p += n
        return v

    def length(self):
        n = self.i32()
        if n == -1:                                  # long vector: two more words
            hi, lo = self.i32(), self.i32()
            n = (hi << 32) + (lo & 0xFFFFFFFF)
        if n < 0:
            raise RdsError("negative vector length")
        return n

    def attrs_of(self, item):
        """pairlist as read by item() -> dict"""
        return {} if item is None else {k: v for k, v in item}

    # -- one serialised item ------------------------------------------------------------------
    def item(self, flags=None):
        if flags is None:
            flags = self.i32()
        t = flags & 0xFF
        has_attr = bool(flags & 0x200)
        has_tag = bool(flags & 0x400)
        if t == 254:                                 # NILVALUE_SXP
            return None
        if t in (253, 242, 241, 247):                # global / empty / base environments, base namespace
            return {253: "<globalenv>", 242: "<emptyenv>", 241: "<baseenv>", 247: "<basenamespace>"}[t]
        if t in (252, 251):                          # unbound value, missing argument
            return None
        if t == 255:                                 # REFSXP
            idx = flags >> 8
            if idx == 0:
                idx = self.i32()
            return self.refs[idx - 1]
        if t in (249, 250, 248):                     # namespace / package / persistent reference: InStringVec (0, n, n CHARSXPs)
            if self.i32() != 0:
                raise RdsError("names in persistent strings are not supported")
            info = [self.item() for _ in range(self.i32())]
            obj = RObj(info, {}, kind={249: "namespace", 250: "package", 248: "persistent"}[t])
            self.refs.append(obj)
            return obj
        if t == 1:                                   # SYMSXP
            name = self.item()
            self.refs.append(name)
            return name
        if t in (2, 6, 239, 240):                    # LISTSXP / LANGSXP (239 / 240: with attributes, version 2 writers)
            out = []
            while True:
                if t in (239, 240) or has_attr:
                    self.item()                      # attributes of the cons cell itself: not used by any consumer here
                tag = self.item() if has_tag else None
                car = self.item()
                out.append((tag, car))
                flags = self.i32()
                t = flags & 0xFF
                if t == 254:
                    break
                if t not in (2, 6, 239, 240):        # dotted pair: the tail is an ordinary value (ALTREP deferred strings)
                    out.append((".tail", self.item(flags)))
                    break
                has_attr = bool(flags & 0x200)
                has_tag = bool(flags & 0x400)
            return out
        if t in (3, 5, 17):                          # CLOSXP / PROMSXP / DOTSXP: attributes?, tag (environment)?, car, cdr
            if has_attr:
                self.item()
            if has_tag:
                self.item()
            self.item(); self.item()
            return {3: "<closure>", 5: "<promise>", 17: "<dots>"}[t]
        if t == 4:                                   # ENVSXP: locked, enclosure, frame, hash table, attributes
            obj = RObj({}, {}, kind="env")
            self.refs.append(obj)
            self.i32()
            self.item()
            frame = self.item()
            hashtab = self.item()
            attr = self.item()
            if frame:
                obj.value.update({k: v for k, v in frame if k is not None})
            for bucket in (unwrap(hashtab) or []):
                if bucket:
                    obj.value.update({k: v for k, v in bucket if k is not None})
            if attr:
                obj.attrs.update(self.attrs_of(attr))
            return obj
        if t in (7, 8):                              # SPECIALSXP / BUILTINSXP: the name
            n = self.i32()
            return "<builtin %s>" % self.take(n).decode("latin-1")
        if t == 9:                                   # CHARSXP
            n = self.i32()
            if n == -1:
                return None                          # NA_character_
            raw = self.take(n)
            enc = "latin-1" if (flags >> 12) & 4 else "utf-8"      # gp bit 2 (LATIN1_MASK) of the levels field
            return raw.decode(enc, "replace")
        if t == 22:                                  # EXTPTRSXP (data.table's .internal.selfref)
            obj = RObj(None, {}, kind="extptr")
            self.refs.append(obj)
            self.item()
            self.item()
            if has_attr:
                obj.attrs.update(self.attrs_of(self.item()))
            return obj
        if t == 23:                                  # WEAKREFSXP
            obj = RObj(None, {}, kind="weakref")
            self.refs.append(obj)
            if has_attr:
                self.item()
            return obj
        if t == 21:
            raise RdsError("byte code is not supported")
        if t == 25:                                  # S4SXP: nothing but attributes (the slots and the class)
            attrs = self.attrs_of(self.item()) if has_attr else {}
            return RObj(None, attrs, kind="S4")
        if t == 238:                                 # ALTREP_SXP: info, state, attributes
            info = self.item()
            state = self.item()
            attr = self.item()
            v = _altrep(info, state)
            a = self.attrs_of(attr)
            return RObj(v, a) if a else v
        if t in (10, 13):                            # LGLSXP / INTSXP
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == 14:                                # REALSXP
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == 15:                                # CPLXSXP
            n = self.length()
            v = np.frombuffer(self.take(16 * n), dtype=">c16").astype(np.complex128)
        elif t == 24:                                # RAWSXP
            n = self.length()
            v = np.frombuffer(self.take(n), dtype=np.uint8).copy()
        elif t == 16:                                # STRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        elif t in (19, 20):                          # VECSXP / EXPRSXP
            n = self.length()
            v = [self.item() for _ in range(n)]
        else:
            raise RdsError("unsupported SEXP type %d at offset %d" % (t, self.p))
        if has_attr:
            attrs = self.attrs_of(self.item())
            if attrs:
                return RObj(v, attrs)
        return v


def _altrep(info, state):
    """the ALTREP classes base R serialises (src/main/altclasses.c)"""
    cls = info[0][1] if info else None
    if cls in ("compact_intseq", "compact_realseq"):
        n, first, incr = (float(x) for x in unwrap(state)[:3])
        seq = first + incr * np.arange(int(n), dtype=np.float64)
        return seq.astype(np.int32) if cls == "compact_intseq" else seq
    if cls == "deferred_string":                     # state = CONS(the numeric vector, scipen)
        arg = unwrap(state[0][1])
        if np.issubdtype(np.asarray(arg).dtype, np.integer):
            return [None if int(x) == NA_INTEGER else str(int(x)) for x in arg]
        return [None if np.isnan(x) else ("%.15g" % x) for x in arg]
    if cls is not None and cls.startswith("wrap_"):  # state = list(x, metadata)
        return unwrap(state)[0]
    raise RdsError("unsupported ALTREP class %r" % (cls,))


def _decompress(raw):
    if raw[:2] == b"\x1f\x8b":
        return gzip.decompress(raw)
    if raw[:3] == b"BZh":
        return bz2.decompress(raw)
    if raw[:6] == b"\xfd7zXZ\x00":
        return lzma.decompress(raw)
    return raw


def read_rds(path):
    """readRDS(path) -> Python objects: None, str, NumPy arrays, lists, lists of (tag, value) for pairlists, RObj."""
    with open(path, "rb") as fh:
        raw = _decompress(fh.read())
    if raw[:2] != b"X\n":
        raise RdsError("%s: not an XDR serialisation (ascii and native-binary saves are not supported)" % path)
    r = _Reader(raw)
    r.p = 2
    version = r.i32()
    r.i32()                                          # R version that wrote the file
    r.i32()                                          # minimal R version to read it
    if version == 3:
        r.take(r.i32())                              # native encoding
    elif version != 2:
        raise RdsError("%s: serialisation version %d is not supported" % (path, version))
    return r.item()


def as_column(col):
    """one column of a data.frame -> NumPy array (factors and character vectors: object arrays of str, None = NA)"""
    if isinstance(col, RObj):
        if "factor" in col.klass():
            lev = unwrap(col.attrs["levels"])
            codes = np.asarray(col.value)
            out = np.empty(len(codes), dtype=object)
            ok = codes != NA_INTEGER
            out[ok] = np.asarray(lev, dtype=object)[codes[ok] - 1]
            out[~ok] = None
            return out
        col = col.value
    if isinstance(col, list):
        return np.asarray(col, dtype=object)
    return np.asarray(col)


def data_frame(obj):
    """data.frame / data.table (a list with a names attribute) -> dict name -> NumPy column"""
    if not isinstance(obj, RObj) or "names" not in obj.attrs or not isinstance(obj.value, list):
        raise RdsError("not a data.frame-like list")
    names = unwrap(obj.attrs["names"])
    return {nm: as_column(col) for nm, col in zip(names, obj.value)}


def named_list(obj):
    """a named list -> dict (values left as they are)"""
    if not isinstance(obj, RObj) or "names" not in obj.attrs:
        return {}
    return dict(zip(unwrap(obj.attrs["names"]), obj.value))


def chicago_table(path):
    """What setChicdiffExperiment() takes from a CHiCAGO .Rds (chicdiff.R:517-534, 614-623): the interaction table -- the
    `x` slot when the file holds a chicagoData object, the object itself when it holds the data.table -- as a dict of
    columns, plus the object's `params` / `settings` lists when present."""
    obj = read_rds(path)
    params = settings = None
    if isinstance(obj, RObj) and obj.kind == "S4":
        if "x" not in obj.attrs:
            raise RdsError("%s: an S4 object without an x slot (class %s)" % (path, obj.klass()))
        params, settings = named_list(obj.attrs.get("params")), named_list(obj.attrs.get("settings"))
        obj = obj.attrs["x"]
    return dict(columns=data_frame(obj), params=params, settings=settings)
