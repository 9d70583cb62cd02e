"""Pins the oracle (and the product's host-side results() step) against the golden table shipped with the
reference: ChicdiffData/inst/extdata/CD4_Mono_results/test_results.Rds (fixture tests/golden/chr19_golden.npz,
made by tests/golden/make_golden.py).  The inputs of that run are not in the mount, so what the table pins is
the output identities of SURVEY.md Appendix D."""
import numpy as np
import pytest

from oracle import oracle as O


def test_stat_is_lfc_over_se(golden):
    assert np.array_equal(golden["stat"], golden["log2FoldChange"] / golden["lfcSE"])


def test_wald_pvalue_formula(golden):
    L = O.lib()
    pv = np.array([L.orc_wald_pvalue(z) for z in golden["stat"]])
    rel = np.abs(pv - golden["pvalue"]) / golden["pvalue"]
    assert golden["pvalue"].min() < 1e-50          # deep tail is exercised
    assert rel.max() < 1e-12


def test_independent_filtering_reproduces_golden_padj(golden):
    f = O.independent_filtering(golden["baseMean"], golden["pvalue"])
    assert f["j"] == 5                               # k = 6 in R's 1-based indexing
    assert abs(f["theta"] - 0.09693877551020408) < 1e-15
    assert abs(f["cutoff"] - 4.796776) < 1e-6
    gp = golden["padj"]
    assert np.array_equal(np.isnan(gp), np.isnan(f["padj"]))
    assert np.isnan(gp).sum() == 2411
    ok = ~np.isnan(gp)
    assert np.max(np.abs(f["padj"][ok] - gp[ok]) / gp[ok]) < 1e-14
    assert (gp[ok] < 0.05).sum() == 2792


def test_bh_on_weighted_pvalues(golden):
    wp = O.p_adjust_bh(golden["weighted_pvalue"])
    assert np.nanmax(np.abs(wp - golden["weighted_padj"])) < 1e-15
    assert (golden["weighted_padj"] < 0.05).sum() == 2759


def test_ihw_weight_identities(golden):
    w = golden["avWeights"] / golden["avWeights"].mean()
    assert np.max(np.abs(w - golden["weight"])) < 1e-12
    assert np.max(np.abs(golden["pvalue"] / golden["weight"] - golden["weighted_pvalue"])) < 1e-15


def _lookup_from_golden(golden):
    """A distLookup (chicdiff.R:2013-2033) consistent with the golden table: per stratum the range of log|avDist|
    and the weight; first minimum 0, last maximum Inf as at :2030-2031."""
    grp, x = golden["group"], golden["avgLogDist"]
    G = int(grp.max())
    lo = np.array([x[grp == g].min() for g in range(1, G + 1)])
    hi = np.array([x[grp == g].max() for g in range(1, G + 1)])
    w = np.array([golden["avWeights"][grp == g][0] for g in range(1, G + 1)])
    for g in range(1, G + 1):
        assert np.all(golden["avWeights"][grp == g] == w[g - 1])          # one weight per stratum
    assert np.all(lo[1:] > hi[:-1])                                       # strata are distance bands
    lo[0], hi[-1] = 0.0, np.inf
    return lo, hi, w


@pytest.mark.parametrize("impl", ["oracle", "product"])
def test_ihw_weight_application_reproduces_golden_columns(golden, impl):
    """chicdiff.R:2038-2049 from avDist + pvalue + the stratum lookup: group, weight, weighted p, weighted padj."""
    lo, hi, w = _lookup_from_golden(golden)
    if impl == "oracle":
        r = O.ihw_apply(golden["avDist"], golden["pvalue"], lo, hi, w)
    else:
        from chicdiff_b200 import engine
        r = engine.ihw_apply(golden["avDist"], golden["pvalue"], lo, hi, w)
    assert np.array_equal(r["group"], golden["group"])
    assert np.max(np.abs(np.log(np.abs(golden["avDist"])) - golden["avgLogDist"])) < 1e-12
    assert np.max(np.abs(r["weight"] - golden["weight"]) / golden["weight"]) < 1e-15
    assert np.max(np.abs(r["weighted_pvalue"] - golden["weighted_pvalue"]) / golden["weighted_pvalue"]) < 1e-15
    assert np.max(np.abs(r["weighted_padj"] - golden["weighted_padj"]) / golden["weighted_padj"]) < 1e-14
    assert (r["weighted_padj"] < 0.05).sum() == 2759


def test_IHWapply_mirror_restores_golden_row_order(golden):
    """The reference's merge(by = "group") leaves the output sorted by stratum, regionID order inside (:2043-2046)."""
    from chicdiff_b200 import api
    lo, hi, w = _lookup_from_golden(golden)
    by_region = np.argsort(golden["regionID"], kind="stable")             # the order DESeq2Wrap returns (:1752-1754)
    table = {k: golden[k][by_region] for k in ("regionID", "pvalue", "avDist", "baitID")}
    out = api.IHWapply(table, dict(minLogDist=lo, maxLogDist=hi, avWeights=w))
    for k in ("regionID", "baitID", "group", "avWeights", "weight", "weighted_pvalue"):
        assert np.allclose(out[k], golden[k], rtol=1e-15, atol=0), k
    assert np.max(np.abs(out["weighted_padj"] - golden["weighted_padj"]) / golden["weighted_padj"]) < 1e-14


def test_ihw_weight_application_na_rules():
    """avDist == 0 or NA falls outside every stratum; mean(avWeights) without na.rm then poisons every weight."""
    from chicdiff_b200 import engine
    lo, hi, w = np.array([0.0, 10.0]), np.array([9.0, np.inf]), np.array([2.0, 0.5])
    av = np.array([np.exp(5.0), -np.exp(12.0), np.exp(9.5), np.exp(9.6)])
    p = np.array([0.01, 0.02, 0.5, np.nan])
    for f in (O.ihw_apply, engine.ihw_apply):
        r = f(av, p, lo, hi, w)
        assert list(r["group"]) == [1, 2, 1, 2]                           # break at (10 + 9) / 2, right-closed
        mean = (2.0 + 0.5 + 2.0 + 0.5) / 4
        assert np.allclose(r["weight"], np.array([2.0, 0.5, 2.0, 0.5]) / mean, rtol=0, atol=1e-16)
        assert np.isnan(r["weighted_padj"][3]) and not np.isnan(r["weighted_padj"][:3]).any()
        r = f(np.array([0.0, np.exp(5.0)]), np.array([0.1, 0.2]), lo, hi, w)
        assert r["group"][0] < 0 and np.isnan(r["weight"]).all() and np.isnan(r["weighted_padj"]).all()
    for f in (O.ihw_apply, engine.ihw_apply):
        with pytest.raises(Exception):                                    # cut(): 'breaks' are not unique
            f(av, p, np.array([0.0, 0.0]), np.array([0.0, np.inf]), w)


def test_annotation_lookups(golden):
    rid, rs, re_ = golden["rmap_id"], golden["rmap_start"], golden["rmap_end"]
    assert np.array_equal(np.diff(rid), np.ones(len(rid) - 1, dtype=rid.dtype))     # contiguous IDs
    base = rid[0]
    assert np.array_equal(golden["OEstart"], rs[golden["minOE"] - base])
    assert np.array_equal(golden["OEend"], re_[golden["maxOE"] - base])
    assert np.array_equal(golden["baitstart"], rs[golden["baitID"] - base])
    assert np.array_equal(golden["baitend"], re_[golden["baitID"] - base])


def test_region_width_law(golden):
    """.expandAvoidBait (chicdiff.R:353-367) with RUexpand = 5: widths 6..11, never touching bait +- 1."""
    from chicdiff_b200 import synth
    s = int(golden["settings_RUexpand"][0])
    width = golden["maxOE"] - golden["minOE"] + 1
    assert width.min() >= s + 1 and width.max() == 2 * s + 1
    assert (width == 2 * s + 1).sum() == 24449
    bait = golden["baitID"].astype(np.int64)
    assert not np.any((golden["minOE"] <= bait + 1) & (golden["maxOE"] >= bait - 1))
    # every golden window is what expand_avoid_bait yields for some seed, up to chromosome-end trimming
    lo, hi = golden["minOE"].astype(np.int64), golden["maxOE"].astype(np.int64)
    right = lo > bait
    seed = np.where(right, hi - s, lo + s)
    elo, ehi = synth.expand_avoid_bait(bait, seed, s)
    last = golden["rmap_id"].max()
    ehi = np.minimum(ehi, last)
    assert np.array_equal(elo[right], lo[right]) or np.all((elo[right] == lo[right]) | (lo[right] == bait[right] + 2))
    assert np.array_equal(ehi[~right], hi[~right]) or np.all((ehi[~right] == hi[~right]) | (hi[~right] == bait[~right] - 2))


def test_product_results_adjust_matches_golden(golden, built):
    """cd_results_adjust is host code (no GPU needed): independent filtering + BH of the product."""
    from chicdiff_b200 import engine
    adj = engine.results_adjust(golden["baseMean"], None, None, golden["pvalue"], 4, 2)
    gp = golden["padj"]
    assert adj["filterIndex"] == 6
    assert abs(adj["filterThreshold"] - 4.796776) < 1e-6
    assert np.array_equal(np.isnan(gp), np.isnan(adj["padj"]))
    ok = ~np.isnan(gp)
    assert np.max(np.abs(adj["padj"][ok] - gp[ok]) / gp[ok]) < 1e-14


def test_region_universe_restatement_reproduces_golden_windows(golden):
    """getRegionUniverse (chicdiff.R:369-426): the oracle restatement must give back every (minOE, maxOE) of the
    golden table from the seeds it implies, including the windows cut short next to the bait and at the
    chromosome end."""
    s = int(golden["settings_RUexpand"][0])
    bait = golden["baitID"].astype(np.int64)
    lo, hi = golden["minOE"].astype(np.int64), golden["maxOE"].astype(np.int64)
    first, last = int(golden["rmap_id"].min()), int(golden["rmap_id"].max())
    right = lo > bait
    # seed implied by the window: s fragments inside the far edge, unless that edge was cut at the chromosome end
    seed = np.where(right, np.where(hi == last, np.maximum(lo + s, bait + 2), hi - s),
                    np.where(lo == first, np.minimum(hi - s, bait - 2), lo + s))
    seed = np.where(right & (lo == bait + 2), hi - s, seed)
    seed = np.where(~right & (hi == bait - 2), lo + s, seed)
    chr_ = np.ones(len(golden["rmap_id"]), np.int32)
    row_off, row_bait, row_oe = O.region_universe(bait, seed, s, chr_, frag_id0=first)
    got_lo = np.minimum.reduceat(row_oe, row_off[:-1])
    got_hi = np.maximum.reduceat(row_oe, row_off[:-1])
    ok = (got_lo == lo) & (got_hi == hi)
    assert ok.mean() > 0.999, ok.mean()          # a window cut on both sides leaves its seed ambiguous
    assert np.array_equal(np.diff(row_off)[ok], (hi - lo + 1)[ok])
    assert np.array_equal(row_bait, np.repeat(bait, np.diff(row_off)))
