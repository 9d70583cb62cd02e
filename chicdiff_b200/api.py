"""Host-side mirror of the reference's R functions on the hot path, over the C ABI.

The reference is R (Chicdiff/R/chicdiff.R); no R interpreter exists in this image, so the host side above
the C ABI is written in Python with the reference's names, argument meaning and error behaviour:

    defaultChicdiffSettings()                          chicdiff.R:3-24
    DESeq2Wrap(chicdiff_settings, RU, FullRegionData, suffix="", theta=None)      chicdiff.R:1494-1777

Tables are dicts of equally long NumPy columns (the stand-in for data.table).  `R/chicdiff_b200.R` is the
same adapter written in R for a real drop-in (it can not be executed here); both only marshal columns and
call the library -- no numerics live on this side.
"""
import sys

import numpy as np

from . import engine

_ENGINE = None


def _get_engine(device=0):
    global _ENGINE
    if _ENGINE is None:
        _ENGINE = engine.Engine(device)
    return _ENGINE


def defaultChicdiffSettings():
    """chicdiff.R:3-24 plus the keys that select this backend."""
    return {
        "inputfiles": None, "peakfiles": None, "chicagoData": None, "countData": None, "rmapfile": None,
        "targetColumns": None, "baitmapfile": None, "RUexpand": 5, "score": 5, "norm": "combined", "theta": None,
        "theta_grid": np.arange(0, 1.0001, 0.25), "saveAuxData": False, "parallel": False, "device": "png",
        "printMemory": False, "outprefix": "",
        # additions of this backend
        "backend": "cuda", "gpu": 0, "dispPriorVar": None, "dispPriorVarGrid": None, "batch": None,
    }


def message(*a):
    print(*a, file=sys.stderr)


def region_rows(RU, FullRegionData, sample_order=None):
    """Long table (chicdiff.R:912-925) -> region-contiguous per-sample row columns for the C ABI.

    Rows of each sample are ordered by (regionID, otherEndID): the order in which data.table sums them
    after `setkey(fragData, otherEndID)` and `by = (baitID, regionID, sample)` (chicdiff.R:1526, 1540-1547).
    Returns (samples, conditions, region_ids, region_bait, row_off, N[S, R], FullMean[S, R])."""
    frd = FullRegionData
    sample = np.asarray(frd["sample"])
    if sample_order is None:
        _, first = np.unique(sample, return_index=True)
        sample_order = [sample[i] for i in sorted(first)]           # first-appearance order, like unique()
    S = len(sample_order)
    region = np.asarray(frd["regionID"], dtype=np.int64)
    oe = np.asarray(frd["otherEndID"], dtype=np.int64)
    bait = np.asarray(frd["baitID"], dtype=np.int64)
    N = np.asarray(frd["N"])
    FMc = np.asarray(frd["FullMean"], dtype=np.float64)
    cond = np.asarray(frd["condition"])
    cols_N, cols_F, conditions = [], [], []
    base = None
    for s in sample_order:
        sel = np.flatnonzero(sample == s)
        o = sel[np.lexsort((oe[sel], region[sel]))]
        key = (region[o], oe[o])
        if base is None:
            base = key
            row_region, row_bait = region[o], bait[o]
        elif not (np.array_equal(key[0], base[0]) and np.array_equal(key[1], base[1])):
            raise ValueError("FullRegionData: samples do not cover the same (regionID, otherEndID) rows")
        cols_N.append(N[o].astype(np.int32))
        cols_F.append(FMc[o])
        conditions.append(str(cond[o[0]]))
    region_ids, start = np.unique(row_region, return_index=True)
    row_off = np.concatenate([start, [len(row_region)]]).astype(np.int64)
    return (list(sample_order), conditions, region_ids, row_bait[start].astype(np.int32), row_off,
            np.stack(cols_N), np.stack(cols_F))


def model_matrix(conditions, batch=None):
    """model.matrix(~ condition) (chicdiff.R:1559) or (~ batch + condition); levels alphabetical."""
    lv = sorted(set(conditions))
    if len(lv) != 2:
        raise ValueError("exactly two conditions are required")
    cols = [np.ones(len(conditions))]
    if batch is not None and len(set(batch)) > 1:
        bl = sorted(set(batch))
        for b in bl[1:]:
            cols.append(np.array([1.0 if x == b else 0.0 for x in batch]))
    cols.append(np.array([1.0 if c == lv[1] else 0.0 for c in conditions]))
    return np.stack(cols, axis=1), lv


def DESeq2Wrap(chicdiff_settings, RU, FullRegionData, suffix="", theta=None, rmap=None):
    """chicdiff.R:1494-1777 on the CUDA backend.

    RU: dict(baitID, regionID, otherEndID); FullRegionData: the long table of getFullRegionData() with at
    least baitID, otherEndID, regionID, sample, N, FullMean, condition.  rmap: dict(chr, start, end, ID)
    (read from settings["rmapfile"] when omitted).  Returns the output table (dict of columns in the
    reference's column order, rows ordered by regionID) with key "attr_theta" standing in for
    attributes(out)$theta (present only for norm == "combined", chicdiff.R:1759)."""
    st = chicdiff_settings
    Grid = st["theta_grid"]
    if theta is None and st.get("theta") is not None:
        theta = st["theta"]
    norm = st["norm"]
    if norm not in ("standard", "fullmean", "combined"):
        raise ValueError("DESeq2Wrap error: Unknown normalisation method.")
    if theta is not None:
        if theta == 1 and norm != "standard":
            message('Warning: Mixing parameter theta set to 1, equivalent to norm = "standard". The norm method has been reset accordingly.')
            norm = "standard"
        if not theta and norm != "fullmean":
            message('Warning: Mixing parameter theta set to 0, equivalent to norm = "fullmean". The norm method has been reset accordingly.')
            norm = "fullmean"
    samples, conditions, region_ids, region_bait, row_off, N, FMr = region_rows(RU, FullRegionData)
    X, levels = model_matrix(conditions, st.get("batch"))
    S, p = X.shape
    eng = _get_engine(st.get("gpu", 0) or 0)
    eng.set_design(X)
    eng.set_regions(row_off)
    for s in range(S):
        eng.set_sample_rows(s, N[s], FMr[s])
    eng.aggregate(fetch=False)
    if norm == "combined" and theta is None:
        message("Optimising scaling factors...")
    res = eng.region_test(norm=norm, theta=None if norm != "combined" else theta, theta_grid=Grid,
                          disp_prior_var=st.get("dispPriorVar"), disp_prior_var_grid=st.get("dispPriorVarGrid"), fetch="table")
    if res["deviances"] is not None:
        message("Total deviances by theta (Fullmean --> Standard):")
        message(" ".join("%f" % d for d in res["deviances"]))
    if norm == "combined":
        message("Theta=%s" % res["theta"])
    message("Processing model output")
    adj = engine.results_adjust(res["baseMean"], res["maxCooks"], res["flags"], res["pvalue"], S, p)
    label = {"standard": "Standard DESeq2 normalisation", "fullmean": "Chicago full mean-based normalisation",
             "combined": "combined normalisation"}[norm]
    with np.errstate(invalid="ignore"):
        message("%s: # unweighted interactions with padj<0.05: %d" % (label, int(np.sum(adj["padj"] < 0.05))))
    # annotation (chicdiff.R:1700-1717)
    ru_region = np.asarray(RU["regionID"], dtype=np.int64)
    ru_oe = np.asarray(RU["otherEndID"], dtype=np.int64)
    ru_bait = np.asarray(RU["baitID"], dtype=np.int64)
    order = np.argsort(ru_region, kind="stable")
    rr, ro, rb = ru_region[order], ru_oe[order], ru_bait[order]
    ids, start = np.unique(rr, return_index=True)
    minOE = np.minimum.reduceat(ro, start)
    maxOE = np.maximum.reduceat(ro, start)
    baitID = rb[start]
    if not np.array_equal(ids, np.arange(1, len(ids) + 1)):
        raise ValueError("stopifnot(identical(1:nrow(annoData), annoData$regionID)) failed")       # chicdiff.R:1717
    if not np.array_equal(ids, region_ids):
        raise ValueError("RU and FullRegionData do not describe the same regions")
    if rmap is None:
        rmap = read_rmap(st["rmapfile"])
    rid = np.asarray(rmap["ID"], dtype=np.int64)
    lut = {k: np.asarray(rmap[k]) for k in ("chr", "start", "end")}
    pos = np.searchsorted(rid, np.concatenate([minOE, maxOE, baitID]))
    if np.any(pos >= len(rid)) or np.any(rid[np.minimum(pos, len(rid) - 1)] != np.concatenate([minOE, maxOE, baitID])):
        raise ValueError("fragment ID missing from the rmap")
    a, b, c = np.split(pos, 3)
    out = {
        "baseMean": res["baseMean"], "log2FoldChange": res["log2FoldChange"], "lfcSE": res["lfcSE"], "stat": res["stat"],
        "pvalue": adj["pvalue"], "padj": adj["padj"], "baitID": baitID, "maxOE": maxOE, "minOE": minOE, "regionID": ids,
        "OEchr": lut["chr"][a], "OEstart": lut["start"][a], "OEend": lut["end"][b],
        "baitchr": lut["chr"][c], "baitstart": lut["start"][c], "baitend": lut["end"][c],
    }
    if norm == "combined":
        out["attr_theta"] = res["theta"]
    return out


def read_rmap(path):
    """Chicago:::.readRmap: whitespace separated chr start end fragID (chr possibly quoted), sorted by ID."""
    chr_, start, end, fid = [], [], [], []
    with open(path) as fh:
        for line in fh:
            f = line.split()
            if len(f) < 4:
                continue
            chr_.append(f[0].strip('"'))
            start.append(int(f[1])); end.append(int(f[2])); fid.append(int(f[3]))
    o = np.argsort(np.asarray(fid), kind="stable")
    return {"chr": np.asarray(chr_)[o], "start": np.asarray(start)[o], "end": np.asarray(end)[o], "ID": np.asarray(fid)[o]}
