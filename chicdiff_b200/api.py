"""Host-side mirror of the reference's R functions on the hot path, over the C ABI.

The reference is R (Chicdiff/R/chicdiff.R); no R interpreter exists in this image, so the host side above
the C ABI is written in Python with the reference's names, argument meaning and error behaviour:

    defaultChicdiffSettings()                                                      chicdiff.R:3-24
    getFullRegionData(chicdiff_settings, RU, RUcontrol, ...) / getFullRegionData1   chicdiff.R:1460-1478, 577-948
    .chicEstimateDistFun -> chicEstimateDistFun(distbin, refBinMean)               chicdiff.R:538-573
    DESeq2Wrap(chicdiff_settings, RU, FullRegionData, suffix="", theta=None)       chicdiff.R:1494-1777

Tables are dicts of equally long NumPy columns (the stand-in for data.table).  `R/chicdiff_b200.R` is the
same adapter written in R for a real drop-in (it can not be executed here); both only marshal columns and
call the library -- no numerics live on this side.
"""
import os
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
                          disp_prior_var=st.get("dispPriorVar"), disp_prior_var_grid=st.get("dispPriorVarGrid"), fetch="table",
                          prior_var_fn=st.get("dispPriorVarFn"))
    if res["deviances"] is not None:
        message("Total deviances by theta (Fullmean --> Standard):")
        message(" ".join("%f" % d for d in res["deviances"]))
    if norm == "combined":
        message("Theta=%s" % res["theta"])
    message("Processing model output")
    adj = eng.results_resident()            # results() on the arrays still in device memory (chicdiff.R:1721,1730,1739)
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


# ---------------------------------------------------------------------------------------------------------
# getFullRegionData (chicdiff.R:1460-1478 -> getFullRegionData1, :577-948)
# ---------------------------------------------------------------------------------------------------------

def chicEstimateDistFun(distbin, refBinMean, binsize=20000):
    """.chicEstimateDistFun (chicdiff.R:538-573): unique non-NA (distbin, refBinMean) pairs ordered by decreasing
    refBinMean get the midpoints round(binsize/2) + k*binsize; cubic least squares of log(refBinMean) on
    log(midpoint); C1 linear head and tail.  Returns the 10 numbers cubicFit[4], obs.min, obs.max, head.coef[2],
    tail.coef[2].  (Host side: ~75 points per replicate; binsize = Chicago::defaultSettings()$binsize.)"""
    db = np.asarray(distbin)
    rb = np.asarray(refBinMean, dtype=np.float64)
    ok = ~np.isnan(rb)
    pairs = sorted(set(zip(db[ok].tolist(), rb[ok].tolist())))
    ref = np.array(sorted((v for _, v in pairs), reverse=True))
    mid = np.rint(binsize / 2.0) + binsize * np.arange(len(ref))
    l = np.log(mid)
    A = np.stack([np.ones_like(l), l, l * l, l ** 3], axis=1)
    fit, *_ = np.linalg.lstsq(A, np.log(ref), rcond=None)
    obs = (np.log(mid.min()), np.log(mid.max()))
    out = list(fit) + [obs[0], obs[1]]
    for x in obs:
        beta = fit[1] + 2 * fit[2] * x + 3 * fit[3] * x * x
        alpha = fit[0] + (fit[1] - beta) * x + fit[2] * x * x + fit[3] * x ** 3
        out += [alpha, beta]
    return np.asarray(out, dtype=np.float64)


def read_chicago_tables(files, use_cache=True, columns=None):
    """The readRDS loop of setChicdiffExperiment() / getFullRegionData1() (chicdiff.R:517-534, 614-623):
    {condition: [path of a CHiCAGO .Rds (a chicagoData object or its data.table), ...]} -> {condition: [dict of columns,
    ...]}.  With use_cache the decoded columns are kept as an Arrow file beside each .Rds and memory-mapped on the next
    run (chicdiff_b200.cache); columns=None keeps every column of the table."""
    from . import cache, rds
    out = {}
    for cond, paths in files.items():
        out[cond] = []
        for p in paths:
            if use_cache:
                t = dict(cache.load_or_build(p, columns=columns))
            else:
                t = rds.chicago_table(p)["columns"]
                if columns is not None:
                    t = {k: v for k, v in t.items() if k in columns}
            t["name"] = os.path.splitext(os.path.basename(p))[0]
            out[cond].append(t)
    return out


def label_codes(lab):
    """bin labels (R factor / character; None or NaN = NA) -> (int32 codes with -1 = NA, sorted levels)"""
    lab = np.asarray(lab, dtype=object)
    na = np.array([v is None or (isinstance(v, float) and np.isnan(v)) for v in lab])
    levels = sorted(set(lab[~na].tolist()))
    lut = {v: k for k, v in enumerate(levels)}
    return np.array([-1 if m else lut[v] for v, m in zip(lab, na)], dtype=np.int32), levels


def chicago_columns(x, counts=None, binsize=20000):
    """One replicate's CHiCAGO table (dict of columns: baitID, otherEndID, N, s_j, s_i, tblb, tlb, Tmean, distbin,
    refBinMean) -> the raw columns cd_build_sample_tables takes.  No sort, no join, no first-per-group pass happens
    here: those are the device's (csrc/tables.cu); the host only turns the bin labels into codes and fits the ~75-point
    distance function (.chicEstimateDistFun, chicdiff.R:538-573)."""
    tb, tb_levels = label_codes(x["tblb"])
    tl, tl_levels = label_codes(x["tlb"])
    t = dict(baitID=np.asarray(x["baitID"], dtype=np.int32), otherEndID=np.asarray(x["otherEndID"], dtype=np.int32),
             s_j=np.asarray(x["s_j"], dtype=np.float64), s_i=np.asarray(x["s_i"], dtype=np.float64), tblb=tb, tlb=tl,
             Tmean=np.asarray(x["Tmean"], dtype=np.float64), N=np.asarray(x["N"], dtype=np.int32),
             n_tblb=max(1, len(tb_levels)), n_tlb=max(1, len(tl_levels)),
             distfun=chicEstimateDistFun(x["distbin"], x["refBinMean"], binsize), tblb_levels=tb_levels, tlb_levels=tl_levels)
    if counts is not None:
        t["cnt_baitID"] = np.asarray(counts["baitID"], dtype=np.int32)
        t["cnt_otherEndID"] = np.asarray(counts["otherEndID"], dtype=np.int32)
        t["cnt_N"] = np.asarray(counts["N"], dtype=np.int32)
    return t


def reconstruct_count_tables(reps):
    """The reference's branch for countData = NULL (chicdiff.R:738-746, 774-786): the (baitID, otherEndID, N) columns of
    the replicates' CHiCAGO tables are merged with Reduce(merge, ...), an INNER join, before the counts are read back, so
    a pair keeps its counts only if every replicate has a row for it; everywhere else the later left join fills 0 (:802).
    Returns one count table per replicate, restricted to the pairs common to all of them."""
    keys = []
    for t in reps:
        b = np.asarray(t["baitID"], dtype=np.int64)
        o = np.asarray(t["otherEndID"], dtype=np.int64)
        keys.append((b << 32) | o)
    common = keys[0]
    for k in keys[1:]:
        common = np.intersect1d(common, k)
    out = []
    for t, k in zip(reps, keys):
        keep = np.isin(k, common)
        out.append({"baitID": np.asarray(t["baitID"])[keep], "otherEndID": np.asarray(t["otherEndID"])[keep],
                    "N": np.asarray(t["N"])[keep]})
    return out


def getFullRegionData1(chicdiff_settings, RU, rmap, chicago_tables, count_tables=None, is_control=False, engine_obj=None):
    """getFullRegionData1 (chicdiff.R:577-948) on the CUDA backend for ONE region universe.

    chicago_tables: {condition: [replicate table, ...]} in the order of settings$chicagoData; count_tables: same
    shape with (baitID, otherEndID, N) chinput tables, or None to take N from the CHiCAGO tables the way the reference
    does for countData = NULL (pairs present in every replicate only, see reconstruct_count_tables).
    Returns the long table (dict of columns baitID, otherEndID, regionID, distSign, sample, N, s_j, Bmean, Tmean,
    score, FullMean, condition; sample-major blocks, then stable-sorted by regionID like setkey(recast, regionID))
    and, for the test set, countput."""
    ids = np.asarray(rmap["ID"], dtype=np.int64)
    id0, F = int(ids[0]), len(ids)
    chr_labels = np.asarray(rmap["chr"])
    _, chr_codes = np.unique(chr_labels, return_inverse=True)
    start = np.asarray(rmap["start"], dtype=np.int64); end = np.asarray(rmap["end"], dtype=np.int64)
    ru_region = np.asarray(RU["regionID"], dtype=np.int64)
    ru_bait = np.asarray(RU["baitID"], dtype=np.int64)
    ru_oe = np.asarray(RU["otherEndID"], dtype=np.int64)
    o = np.lexsort((ru_oe, ru_region))
    ru_region, ru_bait, ru_oe = ru_region[o], ru_bait[o], ru_oe[o]
    region_ids, first = np.unique(ru_region, return_index=True)
    row_off = np.concatenate([first, [len(ru_region)]]).astype(np.int64)
    names, conditions, reps, cnts = [], [], [], []
    for cond, tabs in chicago_tables.items():
        for k, t in enumerate(tabs):
            names.append("%s.%s" % (cond, t.get("name", "rep%d" % (k + 1))))
            conditions.append(cond)
            reps.append(t)
            cnts.append(None if count_tables is None else count_tables[cond][k])
    S = len(reps)
    if count_tables is None:
        message("Reconstructing countData")
        cnts = reconstruct_count_tables(reps)
    eng = engine_obj or _get_engine(chicdiff_settings.get("gpu", 0) or 0)
    X, _ = model_matrix(conditions, chicdiff_settings.get("batch"))
    eng.set_design(X)
    eng.set_rmap(chr_codes, start, end, id0)
    eng.set_regions(row_off)
    eng.set_region_rows(ru_bait, ru_oe)
    tabs = []
    for s in range(S):
        message("\nReading Chicago dataset %d of %d : %s" % (s + 1, S, names[s]))
        eng.build_sample_tables(s, chicago_columns(reps[s], cnts[s]))          # joins / first-per-key tables on the device
        tabs.append(eng.get_sample_tables(s))
    message("Processing count data")
    eng.assemble(keep_rows=True, fetch=False)
    R = len(ru_oe)
    mid = np.rint(0.5 * (start + end))                                     # chicdiff.R:871 (R rounds half to even)
    same = chr_codes[ru_oe - id0] == chr_codes[ru_bait - id0]
    dist = np.where(same, mid[ru_oe - id0] - mid[ru_bait - id0], np.nan)  # :878-881
    out = {k: [] for k in ("baitID", "otherEndID", "regionID", "distSign", "sample", "N", "s_j", "Bmean", "Tmean", "score",
                           "FullMean", "condition")}
    for s in range(S):
        N, FM = eng.get_sample_rows(s, R)
        Bm = eng.get_sample_bmean(s, R)
        t = tabs[s]
        x = reps[s]
        # score of the pair in this replicate's CHiCAGO table, NA where the pair is absent (:634)
        xb = np.asarray(x["baitID"], dtype=np.int64); xo = np.asarray(x["otherEndID"], dtype=np.int64)
        xkey = xb * (1 << 32) + xo
        xo_ = np.argsort(xkey, kind="stable")
        pos = np.searchsorted(xkey[xo_], ru_bait * (1 << 32) + ru_oe)
        pos = np.minimum(pos, len(xkey) - 1)
        hit = xkey[xo_][pos] == ru_bait * (1 << 32) + ru_oe
        score = np.where(hit, np.asarray(x["score"], dtype=np.float64)[xo_][pos], np.nan)
        out["baitID"].append(ru_bait); out["otherEndID"].append(ru_oe); out["regionID"].append(ru_region)
        out["distSign"].append(dist); out["sample"].append(np.repeat(names[s], R)); out["N"].append(N)
        tb, tl = t["tblb"][ru_bait - id0], t["tlb"][ru_oe - id0]
        with np.errstate(invalid="ignore"):
            tmin = np.array([np.nan if np.isnan(r).all() else np.nanmin(r) for r in t["tmean"]])
        Tm = np.where(tb >= 0, np.where(tl >= 0, t["tmean"][np.maximum(tb, 0), np.maximum(tl, 0)], tmin[np.maximum(tb, 0)]), np.nan)
        out["s_j"].append(t["s_j"][ru_bait - id0]); out["Bmean"].append(Bm); out["Tmean"].append(Tm)
        out["score"].append(score); out["FullMean"].append(FM); out["condition"].append(np.repeat(conditions[s], R))
    long = {k: np.concatenate(v) for k, v in out.items()}
    key = np.argsort(long["regionID"], kind="stable")                     # setkey(recast, regionID) (:925)
    long = {k: v[key] for k, v in long.items()}
    if is_control:
        return long
    message("Saving counts\n")
    countput = {k: [] for k in ("baitID", "otherEndID", "Nav", "Bav", "score", "oeID_mid", "condition")}
    for cond in dict.fromkeys(conditions):
        rows = []
        for s in range(S):
            if conditions[s] != cond:
                continue
            x = reps[s]
            keep = ~np.isnan(np.asarray(x["distSign"], dtype=np.float64))                  # :715
            inr = (np.asarray(x["otherEndID"], dtype=np.int64) >= id0) & (np.asarray(x["otherEndID"], dtype=np.int64) < id0 + F)
            keep &= inr                                                                    # inner merge with the rmap (:724)
            xb = np.asarray(x["baitID"], dtype=np.int64)[keep]; xo = np.asarray(x["otherEndID"], dtype=np.int64)[keep]
            oo = np.lexsort((xo, xb))
            rows.append(dict(baitID=xb[oo], otherEndID=xo[oo], N=np.asarray(x["N"])[keep][oo],
                             Bmean=np.asarray(x["Bmean"], dtype=np.float64)[keep][oo],
                             score=np.asarray(x["score"], dtype=np.float64)[keep][oo]))
        cp = eng.countput(rows)
        for k in ("baitID", "otherEndID", "Nav", "Bav", "score", "oeID_mid"):
            countput[k].append(cp[k])
        countput["condition"].append(np.repeat(cond, len(cp["baitID"])))
    countput = {k: np.concatenate(v) for k, v in countput.items()}
    return [long, None, countput]


def getFullRegionData(chicdiff_settings, RU, RUcontrol, rmap, chicago_tables, count_tables=None):
    """chicdiff.R:1460-1478: list(FullRegionData, FullControlRegionData, countput)."""
    message("Reading data for significant interactions")
    res = getFullRegionData1(chicdiff_settings, RU, rmap, chicago_tables, count_tables, is_control=False)
    message("\nReading data for control interactions")
    res[1] = getFullRegionData1(chicdiff_settings, RUcontrol, rmap, chicago_tables, count_tables, is_control=True)
    return res


def IHWapply(out, distLookup, engine_obj=None):
    """The "apply to test data" block of IHWcorrection() (chicdiff.R:2038-2049).

    out: dict of columns of the DESeq2Wrap table plus "avDist" (:1965-1967); distLookup: dict(minLogDist, maxLogDist,
    avWeights) as learned on the control set (:2013-2033).  Returns a copy of `out` with avgLogDist, group, avWeights,
    weight, weighted_pvalue, weighted_padj, rows ordered by group (NA first) like the reference's merge() leaves them."""
    if engine_obj is not None:            # the n-sized work on that context's GPU (cd_ihw_apply_device)
        r = engine_obj.ihw_apply_device(out["pvalue"], distLookup["minLogDist"], distLookup["maxLogDist"], distLookup["avWeights"],
                                        avDist=out["avDist"])
    else:
        r = engine.ihw_apply(out["avDist"], out["pvalue"], distLookup["minLogDist"], distLookup["maxLogDist"], distLookup["avWeights"])
    res = {k: np.asarray(v) for k, v in out.items() if not str(k).startswith("attr_")}
    with np.errstate(divide="ignore", invalid="ignore"):
        res["avgLogDist"] = np.log(np.abs(np.asarray(out["avDist"], dtype=np.float64)))
    na = r["group"] < 0
    w = np.asarray(distLookup["avWeights"], dtype=np.float64)
    res["group"] = r["group"]
    res["avWeights"] = np.where(na, np.nan, w[np.clip(r["group"], 1, len(w)) - 1])
    res["weight"], res["weighted_pvalue"], res["weighted_padj"] = r["weight"], r["weighted_pvalue"], r["weighted_padj"]
    order = np.argsort(np.where(na, 0, r["group"]), kind="stable")
    return {k: v[order] for k, v in res.items()}
