"""Parity of the CUDA path (through the C ABI) with the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): aggregated counts bit-exact; dispersions, log2FC and p-values within 1e-6
relative, checked per region.  How that is checked at scale -- shared global scalars to take the coupling through
the trend fit out of the per-region comparison, and the oracle's own record of which line searches are decided by
rounding -- is described in tests/parity.py; the gate is parity.assert_parity:
  * every region whose search decisions are not within rounding noise: 1e-6 in EVERY column, no exception;
  * the others: 1e-6 as well, except a bounded number of branch flips (<= 1e-4 n + 2), none on the small inputs;
  * free run: same theta, same filter, identical significance calls, trend coupling < 1e-4.
"""
import numpy as np
import pytest

import parity
from chicdiff_b200 import engine, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-6


def frac_ok(a, b, scale=None, tol=TOL):
    a, b = np.asarray(a, float), np.asarray(b, float)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NA pattern differs"
    ref = np.abs(b) if scale is None else np.maximum(np.abs(b), scale)
    with np.errstate(invalid="ignore"):
        ok = (np.abs(a - b) <= tol * ref) | np.isnan(b)
    return ok.mean(), ok


_C3_FULL = []


def c3_full():
    """the full-size bench workload, generated once per test session (a minute of NumPy)"""
    if not _C3_FULL:
        _C3_FULL.append(synth.generate("c3"))
    return _C3_FULL[0]


def run_and_check(d, all_rows=False, **kw):
    t = parity.run_three(d, **kw)
    report = []
    S = parity.compare(d, t, report)
    try:
        parity.assert_parity(S, all_rows=all_rows)
    except AssertionError:
        print("\n".join(report))
        raise
    return t, S


def test_tiny_3v3_all_regions_within_tolerance():
    d = synth.generate("tiny")
    t, S = run_and_check(d, all_rows=True)
    r, ro = t["shared"], t["oracle"]
    assert S["rows_bad"] == 0
    # one launch per stage of a batch, not per fit: the five theta-grid fits run as one problem (round 1: 330 launches)
    assert 30 < t["launches"] / 2 < 120
    # same search paths: identical IRLS iteration counts; the MAP line search may stop one trip apart on a
    # region whose last gain sits at the 1e-6 stopping threshold (values then agree to ~1e-7)
    assert np.array_equal(r["betaIter"], ro["betaIter"])
    assert (r["dispIter"] != ro["dispIter"]).mean() <= 0.002
    assert np.array_equal(r["flags"] & 63, ro["flags"])
    # the free run as well (no branch flip on this input: the coupling is zero)
    f = t["free"]
    p = d.X.shape[1]
    for k, b, sc in (("dispersion", ro["dispersion"], None), ("lfcSE", ro["betaSE"][p - 1], None), ("stat", ro["stat"], 1.0),
                     ("pvalue", ro["pvalue"], None), ("log2FoldChange", ro["beta"][p - 1], ro["betaSE"][p - 1])):
        assert frac_ok(f[k], b, sc)[0] == 1.0, k


def test_c1_shape_2v2_with_given_prior_variance():
    """chr19-shaped 2-vs-2 (BASELINE configs[0] shape).  S - p = 2: DESeq2's seeded Monte-Carlo prior-variance
    estimator is not restated, so both sides receive the same dispPriorVar."""
    run_and_check(synth.generate("c1"), prior=0.5, prior_grid=0.5)


def test_c1_2v2_stand_alone_with_the_small_df_rule():
    """S - p <= 3 without a given prior variance: the library's restatement of DESeq2's Monte-Carlo rule (csrc/priorvar.cpp)
    supplies it.  The value must be what the independent restatement (oracle/priorvar.py) makes of the same residuals, the
    run must equal a run that is handed that value, and that run is held to the oracle as usual.  With the theta grid the
    five intercept-only fits (S - p = 3) use the rule as well."""
    from oracle import priorvar as PV
    d = synth.generate("c1")
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    e.aggregate(fetch=False)
    r = e.region_test(theta=0.5)                                         # no grid: one fit, S - p = 2
    pv = r["dispPriorVar"]
    ok = np.isfinite(r["dispGeneEst"]) & (r["dispGeneEst"] >= 1e-6)
    resid = np.log(r["dispGeneEst"][ok]) - np.log(r["dispFit"][ok])
    assert pv == PV.prior_var_small_df(2, resid) and pv >= 0.25
    r2 = e.region_test(theta=0.5, disp_prior_var=pv)
    for k in ("dispersion", "pvalue", "lfcSE"):
        assert np.array_equal(r[k], r2[k], equal_nan=True), k
    rg = e.region_test()                                                 # theta grid: six fits, all through the rule
    assert rg["theta"] in (0.0, 0.25, 0.5, 0.75, 1.0) and np.isfinite(rg["dispPriorVar"]) and np.isfinite(rg["pvalue"]).any()
    e.close()
    run_and_check(d, prior=pv, theta=0.5)


def test_c2_one_chromosome_2v2():
    run_and_check(synth.generate("c2"), prior=0.6, prior_grid=0.6)


def test_c3_subset_3v3_100k():
    run_and_check(synth.generate("c3", n_regions=100000))


def test_c3_full_size_against_the_oracle():
    """BASELINE configs[2] at full size (2.1 M regions, 23 M rows, 3-vs-3), the configuration the bench line is quoted on,
    against the oracle (about a minute on the box's cores); the per-column table goes to gpurun_out/ when it exists
    (committed copy: profiles/r02_parity_c3_full.txt)."""
    import os
    d = c3_full()
    assert d.n > 2_000_000
    t = parity.run_three(d)
    report = ["== c3 full size: n = %d regions, R = %d rows, S = %d, p = %d" % (d.n, d.R, d.S, d.X.shape[1])]
    S = parity.compare(d, t, report)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.isdir(os.path.join(root, "gpurun_out")):
        with open(os.path.join(root, "gpurun_out", "parity_c3_full_from_pytest.txt"), "w") as fh:
            fh.write("\n".join(report) + "\n")
    print("\n".join(report[:24]))
    parity.assert_parity(S)


def test_c4_batch_covariate_8v8_three_column_glm():
    """~ batch + condition, 8-vs-8: 4 design cells != 3 columns, so mu of the gene-wise step comes from the
    NB GLM (IRLS) and the Wald fit is the 3-column IRLS."""
    d = synth.generate("c4", n_regions=20000)
    assert d.X.shape == (16, 3)
    t, S = run_and_check(d)
    assert frac_ok(t["shared"]["mu"], t["oracle"]["mu"])[0] >= 0.999


@pytest.mark.parametrize("norm,theta", [("standard", None), ("fullmean", None), ("combined", 0.5), ("combined", 1.0), ("combined", 0.0)])
def test_norm_modes_and_fixed_theta(norm, theta):
    d = synth.generate("tiny", seed_offset=3)
    run_and_check(d, norm=norm, theta=theta)


def test_all_zero_region_gives_na_and_poisons_theta_grid():
    d = synth.generate("tiny", seed_offset=5)
    lo, hi = d.row_off[11], d.row_off[12]
    d.N_rows[:, lo:hi] = 0
    # rows shared with neighbouring regions keep their counts there; region 11 itself must be all zero
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K, FM = e.aggregate()
    assert K[:, 11].sum() == 0
    with pytest.raises(engine.ChicdiffError) as ei:
        e.region_test()
    assert "NA" in str(ei.value)
    r = e.region_test(theta=0.25)
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    ro = O.region_test(Ko, FMo, d.X, theta=0.25)
    for k in ("dispGeneEst", "dispersion", "pvalue", "log2FoldChange", "stat"):
        assert np.isnan(r[k][11])
    assert r["flags"][11] & engine.FLAG_ALLZERO
    assert frac_ok(r["pvalue"], ro["pvalue"])[0] >= 0.999
    e.close()


def test_aggregate_edge_cases_on_device():
    e = engine.Engine(0)
    e.set_design(np.array([[1, 0], [1, 0], [1, 1], [1, 1]], float))
    # empty regions, single-row regions, a very wide region (several shared-memory chunks), NA and overflow
    widths = np.array([0, 1, 3, 9000, 0, 11, 2, 7000, 1], np.int64)
    row_off = np.concatenate([[0], np.cumsum(widths)])
    R = int(row_off[-1])
    rng = np.random.default_rng(0)
    N = rng.integers(0, 50, (4, R)).astype(np.int32)
    FMr = rng.random((4, R)) + 0.01
    FMr[1, 5] = np.nan
    N[2, row_off[6]] = 2147483647
    N[2, row_off[6] + 1] = 5
    e.set_regions(row_off)
    for s in range(4):
        e.set_sample_rows(s, N[s], FMr[s])
    K, FM = e.aggregate()
    Ko, FMo = O.aggregate(row_off, N, FMr)
    assert np.array_equal(K, Ko)
    assert K[2, 6] == -2147483648 and K[0, 0] == 0 and K[0, 4] == 0
    assert np.array_equal(np.isnan(FM), np.isnan(FMo)) and np.isnan(FM[1, 3])
    okm = ~np.isnan(FMo)
    assert np.max(np.abs(FM[okm] - FMo[okm]) / np.maximum(np.abs(FMo[okm]), 1e-300)) < 1e-13
    e.close()


def test_aggregate_unaligned_and_ragged_sizes():
    """row counts that are not multiples of the 16-byte copy granule, n not a multiple of the CTA tile"""
    rng = np.random.default_rng(1)
    for n in (1, 7, 255, 256, 257, 1000):
        widths = rng.integers(1, 12, n)
        row_off = np.concatenate([[0], np.cumsum(widths)]).astype(np.int64)
        R = int(row_off[-1])
        N = rng.integers(0, 1000, (3, R)).astype(np.int32)
        FMr = rng.random((3, R))
        e = engine.Engine(0)
        e.set_design(np.array([[1, 0], [1, 1], [1, 1]], float))
        e.set_regions(row_off)
        for s in range(3):
            e.set_sample_rows(s, N[s], FMr[s])
        K, FM = e.aggregate()
        assert np.array_equal(K, np.add.reduceat(N.astype(np.int64), row_off[:-1], axis=1))
        Ko, FMo = O.aggregate(row_off, N, FMr)
        assert np.max(np.abs(FM - FMo) / FMo) < 1e-14
        e.close()


def test_error_paths():
    e = engine.Engine(0)
    with pytest.raises(engine.ChicdiffError):
        e.set_regions(np.array([0, 3], np.int64))                       # design first
    with pytest.raises(engine.ChicdiffError):
        e.set_design(np.ones((2, 2)))                                    # S <= p
    with pytest.raises(engine.ChicdiffError):
        e.set_design(np.array([[1, 1], [1, 1], [1, 1]], float))          # not full rank
    e.set_design(np.array([[1, 0], [1, 0], [1, 1], [1, 1]], float))
    with pytest.raises(engine.ChicdiffError):
        e.set_regions(np.array([0, 3, 2], np.int64))                     # decreasing offsets
    e.set_regions(np.array([0, 2, 4], np.int64))
    with pytest.raises(engine.ChicdiffError):
        e.aggregate()                                                    # rows never set
    with pytest.raises(engine.ChicdiffError):
        e.set_sample_rows(0, np.zeros(3, np.int32), np.zeros(3))         # wrong row count
    for s in range(4):
        e.set_sample_rows(s, np.array([3, 4, 5, 6], np.int32), np.ones(4))
    e.aggregate()
    with pytest.raises(engine.ChicdiffError):
        e.region_test()            # two regions: whatever fails first (trend fit, S - p = 2 rule on an empty histogram), it must say so
    e.close()


def test_two_gpu_sharded_run_matches_single_gpu():
    """Regions sharded by bait over 2 ranks (NCCL all-gathers / all-reduces inside cd_region_test) must give
    the single-GPU answer.  Skipped on a 1-GPU box."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "scripts", "multi_gpu_check.py"), "c3", "60000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_single_process_multi_gpu_matches_single_gpu():
    """cd_multi_*: one process, one host thread per GPU inside the library, peer mailboxes mapped with
    cudaDeviceEnablePeerAccess -- what a single R session uses to reach several GPUs (chicdiff.R:301-347).  Must reproduce
    the single-GPU run of the same set.  Skipped on a 1-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    d = synth.generate("c3", n_regions=60000)
    e1 = engine.Engine(0)
    e1.set_design(d.X); e1.set_regions(d.row_off)
    for s in range(d.S):
        e1.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K1, FM1 = e1.aggregate()
    r1 = e1.region_test()
    m = engine.MultiEngine(2)
    m.set_design(d.X)
    m.set_regions(d.row_off, d.region_bait)
    b = m.shards()
    assert b[0] == 0 and b[-1] == d.n and 0 < b[1] < d.n and d.region_bait[b[1]] != d.region_bait[b[1] - 1]
    for s in range(d.S):
        m.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K2, FM2 = m.aggregate()
    assert np.array_equal(K1, K2) and np.array_equal(FM1, FM2, equal_nan=True)
    r2 = m.region_test()
    assert r2["theta"] == r1["theta"] and r2["n_nonzero"] == r1["n_nonzero"]
    assert np.max(np.abs(r2["sizeFactors"] - r1["sizeFactors"]) / r1["sizeFactors"]) < 1e-14
    assert abs(r2["trend_a0"] - r1["trend_a0"]) < 1e-4 * r1["trend_a0"]
    # per region with the single-GPU run's global scalars (the sharded sums add up in another order: 1e-16 in the trend)
    rs = m.region_test(theta_grid=[r1["theta"]], trend=(r1["trend_a0"], r1["trend_a1"]), var_log_disp=r1["varLogDispEsts"],
                       disp_prior_var=r1["dispPriorVar"])
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    ro = O.region_test(Ko, FMo, d.X, margins=True)
    with np.errstate(invalid="ignore"):
        noisy = (ro["geneMargin"] < parity.MARGIN_NOISE) | (ro["mapMargin"] < parity.MARGIN_NOISE)
    bad = np.zeros(d.n, bool)
    for k in ("baseMean", "dispGeneEst", "dispFit", "dispMAP", "dispersion", "log2FoldChange", "lfcSE", "stat", "pvalue", "deviance", "maxCooks"):
        bad |= parity.rel(rs[k], r1[k]) > 1e-6
    for k in ("normFactors", "mu", "beta"):
        bad |= (parity.rel(rs[k], r1[k]) > 1e-6).any(axis=0)
    assert np.all(noisy[bad]) and bad.sum() <= 2 + 1e-4 * d.n, (int(bad.sum()), np.flatnonzero(bad)[:10])
    m.close(); e1.close()


def test_DESeq2Wrap_mirror_table_matches_oracle():
    """The host mirror of DESeq2Wrap (chicdiff.R:1494-1777): column set and order, regionID ordering, annotation
    lookups, theta attribute, padj -- against the oracle on the reference-shaped tables."""
    from chicdiff_b200 import api
    d = synth.generate("tiny", seed_offset=11)
    RU, frd, rmap = synth.to_reference_tables(d)
    # shuffle the rows inside every sample block: the adapter must restore (regionID, otherEndID) order itself
    rng = np.random.default_rng(0)
    perm = np.concatenate([s * d.R + rng.permutation(d.R) for s in range(d.S)])
    frd = {k: v[perm] for k, v in frd.items()}
    st = api.defaultChicdiffSettings()
    out = api.DESeq2Wrap(st, RU, frd, rmap=rmap)
    assert list(out)[:16] == ["baseMean", "log2FoldChange", "lfcSE", "stat", "pvalue", "padj", "baitID", "maxOE", "minOE",
                              "regionID", "OEchr", "OEstart", "OEend", "baitchr", "baitstart", "baitend"]
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    ro = O.region_test(Ko, FMo, d.X)
    res_o = O.results(ro, Ko, d.X)
    assert out["attr_theta"] == ro["theta"]
    assert np.array_equal(out["regionID"], np.arange(1, d.n + 1))
    assert np.array_equal(out["baitID"], d.region_bait)
    lo = np.minimum.reduceat(d.row_oe, d.row_off[:-1]); hi = np.maximum.reduceat(d.row_oe, d.row_off[:-1])
    assert np.array_equal(out["minOE"], lo) and np.array_equal(out["maxOE"], hi)
    assert np.array_equal(out["OEstart"], d.frag_start[lo - 1]) and np.array_equal(out["OEend"], d.frag_end[hi - 1])
    assert np.array_equal(out["baitstart"], d.frag_start[d.region_bait - 1])
    for k in ("baseMean", "log2FoldChange", "lfcSE", "stat", "pvalue", "padj"):
        f, _ = frac_ok(out[k], res_o[k], ro["betaSE"][1] if k == "log2FoldChange" else None)
        assert f == 1.0, k
    # theta = 1 silently becomes "standard" and then carries no theta attribute (chicdiff.R:1511-1515, 1759)
    out1 = api.DESeq2Wrap(st, RU, frd, theta=1, rmap=rmap)
    assert "attr_theta" not in out1
    st_bad = dict(st, norm="nonsense")
    with pytest.raises(ValueError):
        api.DESeq2Wrap(st_bad, RU, frd, rmap=rmap)


def _assemble_on_device(d, keep_rows=False):
    e = engine.Engine(0)
    e.set_design(d.X)
    e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    e.set_regions(d.row_off)
    e.set_region_rows(d.row_bait, d.row_oe)
    for s in range(d.S):
        e.set_sample_tables(s, d.extra["tables"][s])
    return e, e.assemble(keep_rows=keep_rows)


def test_fused_assembly_matches_oracle_and_long_table_path():
    """getFullRegionData1's per-replicate joins + Bmean/Tmean reconstruction + count merge (chicdiff.R:609-702,
    820-910) fused with the region sums, against the oracle's row-by-row restatement and against the long-table
    route through cd_set_sample_rows / cd_aggregate."""
    d = synth.generate("c1")
    e, (K, FM, av) = _assemble_on_device(d, keep_rows=True)
    Nr = np.empty_like(d.N_rows); Fr = np.empty_like(d.FM_rows)
    dist2 = None
    for s in range(d.S):
        N_o, FM_o, dist_o, bm_o, tm_o = O.assemble_sample(d.row_bait, d.row_oe, d.frag_chr, d.frag_start, d.frag_end,
                                                           d.extra["tables"][s], want_all=True)
        Nr[s], Fr[s] = N_o, FM_o
        Ng, Fg = e.get_sample_rows(s, d.R)
        assert np.array_equal(Ng, N_o)                                         # counts bit-exact
        assert np.array_equal(np.isnan(Fg), np.isnan(FM_o))
        okm = ~np.isnan(FM_o)
        assert np.max(np.abs(Fg[okm] - FM_o[okm]) / FM_o[okm]) < 1e-12
        assert np.isnan(FM_o).any() or s > 0
    assert np.array_equal(Nr, d.N_rows)                                        # oracle == generator's NumPy restatement
    Ko, FMo = O.aggregate(d.row_off, Nr, Fr)
    assert np.array_equal(K, Ko)
    assert np.array_equal(np.isnan(FM), np.isnan(FMo))
    okm = ~np.isnan(FMo)
    assert np.max(np.abs(FM[okm] - FMo[okm]) / FMo[okm]) < 1e-12
    # avDist: mean over the region's rows of round(.5(s+e))_oe - round(.5(s+e))_bait  (chicdiff.R:871-881, 1965)
    mid = np.rint(0.5 * (d.frag_start + d.frag_end))
    dist_rows = mid[d.row_oe - 1] - mid[d.row_bait - 1]
    av_ref = np.add.reduceat(dist_rows, d.row_off[:-1]) / np.diff(d.row_off)
    assert np.max(np.abs(av - av_ref)) < 1e-6
    # the region test downstream of the fused path equals the one downstream of the long-table path
    r1 = e.region_test(disp_prior_var=0.5, disp_prior_var_grid=0.5)
    e2 = engine.Engine(0)
    e2.set_design(d.X); e2.set_regions(d.row_off)
    for s in range(d.S):
        e2.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    e2.aggregate(fetch=False)
    r2 = e2.region_test(disp_prior_var=0.5, disp_prior_var_grid=0.5)
    assert r1["theta"] == r2["theta"]
    assert frac_ok(r1["pvalue"], r2["pvalue"], np.abs(r2["pvalue"]) * np.maximum(1, r2["stat"] ** 2), 1e-6 + 1e-4)[0] >= 0.999
    e.close(); e2.close()


def test_replicate_tables_built_on_device():
    """cd_build_sample_tables (chicdiff.R:632-634, 659-692, 828-853): the per-bait / per-other-end / per-bin-pair "first in
    key order" tables and the sparse count rows, built by atomicMin + one radix sort from the raw CHiCAGO columns in ANY
    row order, against the sort-then-first restatement in the oracle -- bit for bit (the tables are gathers)."""
    from chicdiff_b200 import api
    d = synth.generate("c1")
    ids = np.arange(1, len(d.frag_chr) + 1)
    e = engine.Engine(0)
    e.set_design(d.X)
    e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    rng = np.random.default_rng(3)
    for s in range(d.S):
        x = synth.chicago_table(d, s)
        cnt = synth.chinput_table(d, s) if s % 2 == 0 else None               # .chinput rows / the table's own N column
        perm = rng.permutation(len(x["baitID"]))
        xs = {k: (np.asarray(v)[perm] if hasattr(v, "__len__") and len(v) == len(perm) else v) for k, v in x.items()}
        if cnt is not None:
            pc = rng.permutation(len(cnt["baitID"]))
            cnt = {k: np.asarray(v)[pc] for k, v in cnt.items()}
        e.build_sample_tables(s, api.chicago_columns(xs, cnt))
        got = e.get_sample_tables(s)
        ref = O.replicate_tables(x, ids, None if cnt is None else synth.chinput_table(d, s))
        for k in ("s_j", "s_i", "tmean"):
            assert np.array_equal(got[k], ref[k], equal_nan=True), (s, k)
        for k in ("tblb", "tlb", "cnt_off", "cnt_oe", "cnt_N"):
            assert np.array_equal(got[k], ref[k]), (s, k)
        assert np.isnan(ref["s_j"]).any() and (ref["tlb"] >= 0).any() and ref["cnt_off"][-1] > 1000
    # the tables feed the fused assembly exactly like the host-prepared ones (counts from the .chinput for every replicate)
    for s in range(d.S):
        e.build_sample_tables(s, api.chicago_columns(synth.chicago_table(d, s), synth.chinput_table(d, s)))
    e.set_regions(d.row_off)
    e.set_region_rows(d.row_bait, d.row_oe)
    K, FM, av = e.assemble()
    e2 = engine.Engine(0)
    e2.set_design(d.X)
    e2.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    e2.set_regions(d.row_off)
    e2.set_region_rows(d.row_bait, d.row_oe)
    for s in range(d.S):
        e2.set_sample_tables(s, O.replicate_tables(synth.chicago_table(d, s), ids, synth.chinput_table(d, s)))
    K2, FM2, av2 = e2.assemble()
    assert np.array_equal(K, K2) and np.array_equal(FM, FM2, equal_nan=True) and np.array_equal(av, av2, equal_nan=True)
    e2.close()
    # ties: the same pair twice -> the earlier input row wins; a fragment outside the rmap is refused; an empty table works
    t = dict(baitID=[5, 5, 5, 9], otherEndID=[7, 7, 6, 7], s_j=[1.5, 2.5, 3.5, 4.5], s_i=[0.1, 0.2, 0.3, 0.4], tblb=[0, 1, 1, 0],
             tlb=[1, 1, 0, 1], Tmean=[0.01, 0.02, 0.03, 0.04], N=[3, 4, 5, 6], n_tblb=2, n_tlb=2, distfun=np.zeros(10))
    e.build_sample_tables(0, t)
    g = e.get_sample_tables(0)
    assert g["s_j"][4] == 3.5 and g["tblb"][4] == 1            # bait 5: first in (otherEndID, row) order is (5, 6)
    assert g["s_i"][6] == 0.1 and g["tlb"][6] == 1             # other end 7: first in (baitID, row) order is row 0
    assert g["tmean"].tolist() == [[np.nan, 0.01], [0.03, 0.02]] or (np.isnan(g["tmean"][0, 0]) and g["tmean"][0, 1] == 0.01
                                                                       and g["tmean"][1, 0] == 0.03 and g["tmean"][1, 1] == 0.02)
    assert g["cnt_off"][4] == 0 and g["cnt_off"][5] == 3 and g["cnt_oe"].tolist() == [6, 7, 7, 7] and g["cnt_N"].tolist() == [5, 3, 4, 6]
    with pytest.raises(engine.ChicdiffError):
        e.build_sample_tables(0, dict(t, otherEndID=[7, 7, 6, 10 ** 6]))
    e.build_sample_tables(1, dict(baitID=[], otherEndID=[], s_j=[], s_i=[], tblb=[], tlb=[], Tmean=[], N=[], n_tblb=1, n_tlb=1, distfun=np.zeros(10)))
    g = e.get_sample_tables(1)
    assert np.isnan(g["s_j"]).all() and g["cnt_off"][-1] == 0
    e.close()


def test_assembly_edge_cases():
    """trans rows (Bmean = 0), bait unknown to a replicate (s_j NA -> FullMean NA), unseen other end
    (s_i -> 1, Tmean -> lowest of the tblb), (tblb, tlb) combination never observed (Tmean NA), empty count table."""
    F = 40
    chr_ = np.array([1] * 20 + [2] * 20, np.int32)
    start = (np.arange(F) * 1000 + 1).astype(np.int32); end = (start + 998).astype(np.int32)
    row_off = np.array([0, 3, 5, 8], np.int64)
    row_bait = np.array([5, 5, 5, 12, 12, 30, 30, 30], np.int32)
    row_oe = np.array([8, 9, 10, 25, 26, 33, 34, 35], np.int32)            # region 2 is trans (chr1 bait, chr2 other ends)
    tab = dict(s_j=np.full(F, np.nan), tblb=np.full(F, -1, np.int32), s_i=np.full(F, 1.3), tlb=np.zeros(F, np.int32),
               tmean=np.array([[0.01, 0.02], [0.03, np.nan]]), distfun=synth.dist_fun_params((-3.0, 2.6, -0.33, 0.009), (np.log(1e4), np.log(1.5e6))),
               cnt_off=np.zeros(F + 1, np.int64), cnt_oe=np.zeros(0, np.int32), cnt_N=np.zeros(0, np.int32))
    tab["s_j"][4] = 2.0; tab["tblb"][4] = 0            # bait 5 known
    tab["s_j"][11] = 1.5; tab["tblb"][11] = 1          # bait 12 known, tblb 1
    # bait 30 unknown to this replicate: s_j NA
    tab["s_i"][8] = np.nan; tab["tlb"][8] = -1         # other end 9 never seen
    tab["tlb"][24] = 1                                 # (tblb 1, tlb 1) never observed -> Tmean NA
    tab2 = dict(tab)
    cnt = {(5, 8): 4, (5, 10): 1, (5, 11): 9, (12, 26): 2, (30, 33): 7}
    keys = sorted(cnt)
    off = np.zeros(F + 1, np.int64)
    for (b, o) in keys:
        off[b:] += 1
    tab2["cnt_off"] = off; tab2["cnt_oe"] = np.array([o for _, o in keys], np.int32); tab2["cnt_N"] = np.array([cnt[k] for k in keys], np.int32)
    e = engine.Engine(0)
    e.set_design(np.array([[1, 0], [1, 0], [1, 1]], float))
    e.set_rmap(chr_, start, end, 1)
    e.set_regions(row_off)
    e.set_region_rows(row_bait, row_oe)
    for s, t in enumerate((tab, tab2, tab2)):
        e.set_sample_tables(s, t)
    K, FM, av = e.assemble(keep_rows=True)
    for s, t in enumerate((tab, tab2, tab2)):
        N_o, FM_o = O.assemble_sample(row_bait, row_oe, chr_, start, end, t)
        Ng, Fg = e.get_sample_rows(s, len(row_oe))
        assert np.array_equal(Ng, N_o)
        assert np.array_equal(np.isnan(Fg), np.isnan(FM_o))
        assert np.allclose(Fg[~np.isnan(FM_o)], FM_o[~np.isnan(FM_o)], rtol=1e-13, atol=0)
    assert K[0].tolist() == [0, 0, 0] and K[1].tolist() == [5, 2, 7]
    N1, F1 = e.get_sample_rows(1, len(row_oe))
    assert F1[4] == 0.0 + 0.03 and np.isnan(F1[3])           # trans: Bmean 0 + Tmean(1,0); Tmean(1,1) absent -> NA
    assert np.isnan(F1[5:]).all()                            # bait 30: s_j NA -> NA
    d9 = np.rint(((start[8] + end[8]) - (start[4] + end[4])) / 2.0)
    assert abs(F1[1] - (2.0 * 1.0 * synth.dist_fun(np.array([abs(d9)]))[0] + 0.01)) < 1e-13     # s_i -> 1, Tmean -> min of row 0
    assert np.isnan(av[1])                                   # trans region has no distance
    with pytest.raises(engine.ChicdiffError):
        e.set_region_rows(np.array([5, 5, 5, 12, 12, 30, 30, 99], np.int32), row_oe)   # accepted here ...
        e.assemble()                                                                     # ... rejected by the kernel
    e.close()


def test_region_universe_on_device():
    """getRegionUniverse (chicdiff.R:369-426) on the device against the oracle restatement, then straight into
    the fused assembly without a host round trip of the rows."""
    d = synth.generate("c3", n_regions=50000)
    e = engine.Engine(0)
    e.set_design(d.X)
    e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    off, rb, ro = e.region_universe(d.region_bait, d.region_seed, 5)
    off_o, rb_o, ro_o = O.region_universe(d.region_bait, d.region_seed, 5, d.frag_chr)
    assert np.array_equal(off, off_o) and np.array_equal(rb, rb_o) and np.array_equal(ro, ro_o)
    assert np.array_equal(off, d.row_off) and np.array_equal(ro, d.row_oe)
    for s in range(d.S):
        e.set_sample_tables(s, d.extra["tables"][s])
    K, FM, av = e.assemble()
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    assert np.array_equal(K, Ko)
    # peaks at the genome edge and next to the bait, other RUexpand values
    F = len(d.frag_chr)
    pb = np.array([5, 5, 5, F - 1, 2, 100, 100], np.int32)
    po = np.array([7, 3, 11, F, 1, 98, 102], np.int32)
    for s_ in (1, 3, 5):
        got = e.region_universe(pb, po, s_)
        ref = O.region_universe(pb, po, s_, d.frag_chr)
        for a, b in zip(got, ref):
            assert np.array_equal(a, b)
    with pytest.raises(engine.ChicdiffError):
        e.region_universe(np.array([9], np.int32), np.array([9], np.int32), 5)
    e.close()


def test_countput_on_device():
    """countput (chicdiff.R:755-770): per-condition means / max over the replicates in which a pair occurs,
    first-appearance row order."""
    d = synth.generate("c1")
    e = engine.Engine(0)
    e.set_design(d.X)
    e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    for cond in sorted(set(d.conditions)):
        reps = [synth.chicago_rows(d, s) for s in range(d.S) if d.conditions[s] == cond]
        got = e.countput(reps)
        ref = O.countput(reps, d.frag_start, d.frag_end)
        assert len(got["baitID"]) == len(ref["baitID"]) > 0
        assert np.array_equal(got["baitID"], ref["baitID"]) and np.array_equal(got["otherEndID"], ref["otherEndID"])
        assert np.array_equal(got["oeID_mid"], ref["oeID_mid"])
        assert np.max(np.abs(got["Nav"] - ref["Nav"])) < 1e-12
        assert np.max(np.abs(got["Bav"] - ref["Bav"]) / ref["Bav"]) < 1e-14
        assert np.array_equal(np.isnan(got["score"]), np.isnan(ref["score"])) and np.isnan(ref["score"]).any()
        okm = ~np.isnan(ref["score"])
        assert np.array_equal(got["score"][okm], ref["score"][okm])
    # one replicate, and an empty replicate in the middle
    r0 = synth.chicago_rows(d, 0)
    empty = {k: v[:0] for k, v in r0.items()}
    got = e.countput([r0, empty])
    assert np.array_equal(got["baitID"], r0["baitID"]) and np.array_equal(got["Nav"], r0["N"].astype(float))
    e.close()


@pytest.mark.parametrize("reps,extra_cols", [((6, 6), 0), ((5, 7), 0), ((16, 16), 0), ((8, 8), 2)])
def test_other_designs(reps, extra_cols):
    """more replicates (register/shared-memory paths with S = 12, 32), unbalanced groups, and a 4-column design
    (intercept + two nuisance covariates + condition: the IRLS-mu variant of the gene-wise step with p = 4)."""
    d = synth.generate("tiny", reps=reps, seed_offset=17 + reps[0])
    X = d.X
    if extra_cols:
        S = X.shape[0]
        c1 = (np.arange(S) % 2).astype(float)
        c2 = ((np.arange(S) // 2) % 2).astype(float)
        X = np.column_stack([X[:, 0], c1, c2, X[:, 1]])
        d.X = X
    run_and_check(d)


def test_full_size_c3_properties():
    """BASELINE configs[2] at full size (2.1 M regions, 23 M rows, 3-vs-3), where the oracle would take minutes:
    size-independent properties instead.  Checksum of checksums for the aggregation (exact), idempotence of the
    whole region test (bitwise), output identities (stat = LFC / SE, p = 2 Phi(-|stat|)), value ranges, and the
    fused-assembly route giving the same counts."""
    from scipy import special
    d = c3_full()
    assert d.n > 2_000_000
    e = engine.Engine(0)
    e.set_design(d.X)
    e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K, FM = e.aggregate()
    assert np.array_equal(K.astype(np.int64).sum(axis=1), d.N_rows.astype(np.int64).sum(axis=1))      # nothing lost, nothing double counted
    # spot-check 1000 random regions against a direct sum
    rng = np.random.default_rng(5)
    for i in rng.integers(0, d.n, 1000):
        lo, hi = d.row_off[i], d.row_off[i + 1]
        assert np.array_equal(K[:, i], d.N_rows[:, lo:hi].sum(axis=1))
    r1 = e.region_test()
    r2 = e.region_test()
    for k in ("dispGeneEst", "dispersion", "log2FoldChange", "lfcSE", "stat", "pvalue", "deviance"):
        assert np.array_equal(r1[k], r2[k], equal_nan=True), k                                         # deterministic reductions, no fp atomics
    assert r1["theta"] == r2["theta"] and np.array_equal(r1["deviances"], r2["deviances"])
    ok = ~np.isnan(r1["pvalue"])
    assert ok.all()                                                                                    # generator leaves no all-zero region
    assert np.max(np.abs(r1["stat"] - r1["log2FoldChange"] / r1["lfcSE"])) == 0.0
    p_ref = special.erfc(np.abs(r1["stat"]) / np.sqrt(2.0))
    assert np.max(np.abs(r1["pvalue"] - p_ref) / np.maximum(p_ref, 1e-300)) < 1e-12
    assert np.all((r1["dispersion"] >= 1e-8) & (r1["dispersion"] <= 10.0))
    assert np.all((r1["dispGeneIter"] >= 1) & (r1["dispGeneIter"] <= 100)) and np.all(r1["betaIter"] < 100)
    assert r1["n_nonzero"] == d.n
    # true effects are found: planted regions dominate the top of the ranking
    top = np.argsort(r1["pvalue"])[:2000]
    assert (d.true_lfc[top] != 0).mean() > 0.5          # base rate is 0.1; the 3-vs-3 Wald tail is anti-conservative
    # the fused assembly route reproduces the same aggregated counts from the per-replicate tables
    e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    e.set_regions(d.row_off)
    e.set_region_rows(d.row_bait, d.row_oe)
    for s in range(d.S):
        e.set_sample_tables(s, d.extra["tables"][s])
    K2, FM2, av = e.assemble()
    assert np.array_equal(K2, K)
    assert np.array_equal(np.isnan(FM2), np.isnan(FM))
    okm = ~np.isnan(FM)
    assert np.max(np.abs(FM2[okm] - FM[okm]) / FM[okm]) < 1e-12
    e.close()


def test_getFullRegionData_mirror_and_pipeline():
    """The second boundary function (chicdiff.R:1460-1478): getFullRegionData on reference-shaped inputs (CHiCAGO
    tables + .chinput tables per condition, RU, RUcontrol, rmap) -> long tables + countput, then DESeq2Wrap on the
    result, i.e. the body of chicdiffPipeline between the region universes and IHW (chicdiff.R:315-332)."""
    from chicdiff_b200 import api
    d = synth.generate("tiny", seed_offset=23)
    RU, _, rmap = synth.to_reference_tables(d)
    conds = list(dict.fromkeys(d.conditions))
    chicago = {c: [synth.chicago_table(d, s) for s in range(d.S) if d.conditions[s] == c] for c in conds}
    chinput = {c: [synth.chinput_table(d, s) for s in range(d.S) if d.conditions[s] == c] for c in conds}
    # a control universe: same baits, windows shifted by 40 fragments (stays on the single chromosome)
    F = len(d.frag_chr)
    seeds = np.clip(d.region_seed.astype(np.int64) + 40, 1, F)
    okc = np.abs(seeds - d.region_bait) > 1
    off_c, rb_c, ro_c = O.region_universe(d.region_bait[okc], seeds[okc], 5, d.frag_chr)
    RUc = {"baitID": rb_c.astype(np.int64), "otherEndID": ro_c.astype(np.int64),
           "regionID": np.repeat(np.arange(1, len(off_c)), np.diff(off_c)).astype(np.int64)}
    st = api.defaultChicdiffSettings()
    frd, frd_control, countput = api.getFullRegionData(st, RU, RUc, rmap, chicago, chinput)
    assert list(frd) == ["baitID", "otherEndID", "regionID", "distSign", "sample", "N", "s_j", "Bmean", "Tmean", "score",
                         "FullMean", "condition"]
    assert len(frd["N"]) == d.R * d.S and np.all(np.diff(frd["regionID"]) >= 0)
    ids = np.arange(1, F + 1)
    for uni, table in ((RU, frd), (RUc, frd_control)):
        o = np.lexsort((uni["otherEndID"], uni["regionID"]))
        rb, ro = uni["baitID"][o], uni["otherEndID"][o]
        for s in range(d.S):
            name = "%s.rep%d" % (d.conditions[s], s + 1)
            sel = np.flatnonzero(table["sample"] == name)
            oo = sel[np.lexsort((table["otherEndID"][sel], table["regionID"][sel]))]
            tabs = O.replicate_tables(synth.chicago_table(d, s), ids, synth.chinput_table(d, s))
            N_o, FM_o, dist_o, bm_o, tm_o = O.assemble_sample(rb, ro, d.frag_chr, d.frag_start, d.frag_end, tabs, want_all=True)
            assert np.array_equal(table["N"][oo], N_o)
            for col, ref in (("FullMean", FM_o), ("Bmean", bm_o), ("Tmean", tm_o)):
                assert np.array_equal(np.isnan(table[col][oo]), np.isnan(ref)), col
                okm = ~np.isnan(ref)
                assert np.allclose(table[col][oo][okm], ref[okm], rtol=1e-12, atol=0), col
    assert set(countput) == {"baitID", "otherEndID", "Nav", "Bav", "score", "oeID_mid", "condition"}
    assert set(countput["condition"]) == set(conds)
    # downstream: the region test on the assembled long table (the test set has no all-zero region here? use theta)
    out = api.DESeq2Wrap(st, RU, frd, rmap=rmap, theta=0.5)
    assert len(out["pvalue"]) == d.n and out["attr_theta"] == 0.5
    out_c = api.DESeq2Wrap(st, RUc, frd_control, rmap=rmap, theta=out["attr_theta"])       # chicdiff.R:331
    assert len(out_c["pvalue"]) == len(off_c) - 1


def test_chinput_codec_on_device():
    """the .chinput text parser (reference: fread, chicdiff.R:828) against a plain Python parse, including the comment
    and header lines, NA distances, CRLF line ends, a missing final newline and blank lines."""
    d = synth.generate("c3", n_regions=20000)
    e = engine.Engine(0)
    txt = synth.chinput_text(d, 0, max_rows=200000)
    got = e.parse_chinput(txt)
    ref = O.parse_chinput(txt)
    assert len(ref["N"]) > 50000
    for k in ("baitID", "otherEndID", "N", "otherEndLen"):
        assert np.array_equal(got[k], ref[k]), k
    assert np.array_equal(np.isnan(got["distSign"]), np.isnan(ref["distSign"]))
    assert np.array_equal(np.nan_to_num(got["distSign"]), np.nan_to_num(ref["distSign"]))
    t = synth.chinput_table(d, 0)
    assert np.array_equal(got["baitID"], t["baitID"][:200000]) and np.array_equal(got["N"], t["N"][:200000])
    weird = b"# comment\r\nbaitID\totherEndID\tN\totherEndLen\tdistSign\r\n5\t9\t3\t1200\tNA\r\n\r\n7 12 1 800 -4500\n8\t13\t2\t77\t15000"
    got = e.parse_chinput(weird)
    assert got["baitID"].tolist() == [5, 7, 8] and got["otherEndID"].tolist() == [9, 12, 13] and got["N"].tolist() == [3, 1, 2]
    assert np.isnan(got["distSign"][0]) and got["distSign"][1:].tolist() == [-4500.0, 15000.0]
    assert e.parse_chinput(b"")["N"].size == 0 and e.parse_chinput(b"# only a comment\n")["N"].size == 0
    # numbers written with a fraction or an exponent are read as numbers; a row whose ID or count is not a whole number
    # is dropped (never truncated to 123 or 1)
    odd = b"baitID\totherEndID\tN\totherEndLen\tdistSign\n1e3\t12.0\t3\t1.2e3\t-4.5e3\n123.5\t7\t1\t10\t5\n9\t1e5x\t1\t10\t5\n8\t2\t4\t7.5\t1250.5\n"
    got, ref = e.parse_chinput(odd), O.parse_chinput(odd)
    assert got["baitID"].tolist() == [1000, 8] and got["otherEndID"].tolist() == [12, 2] and got["N"].tolist() == [3, 4]
    assert got["otherEndLen"].tolist() == [1200, -2147483648] and got["distSign"].tolist() == [-4500.0, 1250.5]
    for k in ("baitID", "otherEndID", "N", "otherEndLen", "distSign"):
        assert np.array_equal(got[k], ref[k]), k
    e.close()


def _resident_vs_host(e, r, S, p):
    dev = e.results_resident()
    host = engine.results_adjust(r["baseMean"], r["maxCooks"], r["flags"], r["pvalue"], S, p)
    assert np.array_equal(dev["pvalue"], host["pvalue"], equal_nan=True), "Cook's filter differs"
    assert np.array_equal(dev["padj"], host["padj"], equal_nan=True), "adjusted p-values differ"
    for k in ("cooksCutoff", "filterThreshold", "filterTheta", "filterIndex"):
        assert dev[k] == host[k] or (np.isnan(dev[k]) and np.isnan(host[k])), k
    return dev, host


def test_results_on_resident_arrays_match_host_routine():
    """cd_results_resident (radix sorts + prefix counts on the device) against cd_results_adjust (host) on the same
    columns: same Cook's filter, same filtering cut-off, bit-identical BH values (chicdiff.R:1721,1730,1739)."""
    # an all-zero region (baseMean 0, NA p-value) and a fixed theta
    d = synth.generate("tiny", seed_offset=5)
    lo, hi = d.row_off[11], d.row_off[12]
    d.N_rows[:, lo:hi] = 0
    e = engine.Engine(0)
    with pytest.raises(engine.ChicdiffError):
        e.results_resident()                                             # nothing tested yet
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    e.aggregate()
    r = e.region_test(theta=0.25)
    dev, _ = _resident_vs_host(e, r, d.S, d.X.shape[1])
    assert np.isnan(dev["padj"][11])
    e.close()
    # 100k regions (49 chunks of sorted rows), Cook's outliers present, and the oracle's results() as third opinion
    d = synth.generate("c3", n_regions=100000)
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K, _ = e.aggregate()
    r = e.region_test()
    dev, host = _resident_vs_host(e, r, d.S, d.X.shape[1])
    assert np.isnan(dev["pvalue"]).sum() >= np.isnan(r["pvalue"]).sum()
    with np.errstate(invalid="ignore"):
        assert (dev["padj"] < 0.05).sum() > 100
    # an 8-vs-8 design with a covariate (p = 3: no two-level heuristic)
    d = synth.generate("c4", n_regions=20000)
    e2 = engine.Engine(0)
    e2.set_design(d.X); e2.set_regions(d.row_off)
    for s in range(d.S):
        e2.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    e2.aggregate()
    r = e2.region_test()
    _resident_vs_host(e2, r, d.S, d.X.shape[1])
    e2.close()
    e.close()


def _rule(df, resid):
    """a deterministic stand-in for DESeq2's Monte-Carlo rule: any function of (df, residuals) will do here"""
    mad = 1.4826 * np.median(np.abs(resid - np.median(resid)))
    return max(mad ** 2 - 0.05 * df, 0.25)


def _oracle_residuals(ro):
    ge, fit = ro["dispGeneEst"], ro["dispFit"]
    with np.errstate(invalid="ignore"):
        keep = ge >= 1e-6
    return np.log(ge[keep]) - np.log(fit[keep])


def test_prior_variance_callback_for_small_df():
    """2-vs-2 (S - p = 2; theta-grid fits S - 1 = 3): DESeq2 estimates dispPriorVar with R's RNG.  The library asks the
    caller instead (cd_options.prior_var_fn) and must hand over exactly the residuals DESeq2's rule sees."""
    d = synth.generate("c1")
    p = d.X.shape[1]
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    K, FM = e.aggregate()
    # one fit at a fixed theta: the oracle runs twice (any prior first: gene-wise estimates and trend do not depend on it)
    seen = []
    def rule(df, resid):
        seen.append((df, resid))
        return _rule(df, resid)
    r = e.region_test(theta=0.5, prior_var_fn=rule)
    assert [c[0] for c in seen] == [d.S - p]
    Ko, FMo = O.aggregate(d.row_off, d.N_rows, d.FM_rows)
    ro = O.region_test(Ko, FMo, d.X, theta=0.5, prior_var=1.0)
    res_o = _oracle_residuals(ro)
    assert abs(len(seen[0][1]) - len(res_o)) <= 2                     # rows at the 1e-6 threshold may fall either side
    v = _rule(d.S - p, res_o)
    assert abs(r["dispPriorVar"] - v) <= 1e-4 * v
    assert abs(r["dispPriorVar"] - _rule(d.S - p, seen[0][1])) == 0.0
    # the fit that used the rule's value equals the oracle's fit with the same value (per region, with the oracle's trend
    # handed over so that the comparison is not blurred by the coupling through the global fit; see tests/parity.py)
    ro = O.region_test(Ko, FMo, d.X, theta=0.5, prior_var=r["dispPriorVar"], margins=True)
    assert abs(r["trend_a0"] - ro["trend_a0"]) < 1e-4 * ro["trend_a0"] and abs(r["trend_a1"] - ro["trend_a1"]) < 1e-4 * ro["trend_a1"]
    rs = e.region_test(theta=0.5, disp_prior_var=r["dispPriorVar"], trend=(ro["trend_a0"], ro["trend_a1"]), var_log_disp=ro["varLogDispEsts"])
    with np.errstate(invalid="ignore"):
        clean = ~((ro["geneMargin"] < parity.MARGIN_NOISE) | (ro["mapMargin"] < parity.MARGIN_NOISE)) & ~np.isnan(ro["geneMargin"])
    err = parity.rel(rs["dispersion"], ro["dispersion"])
    assert err[clean].max() <= 1e-6 and (err > 1e-6).sum() <= 2 + 1e-4 * d.n
    # theta grid: five intercept-only fits (df = S - 1) and the final fit (df = S - p), in that order
    seen.clear()
    r = e.region_test(prior_var_fn=rule)
    assert [c[0] for c in seen] == [d.S - 1] * 5 + [d.S - p]
    assert r["dispPriorVar"] == _rule(d.S - p, seen[-1][1])
    # a rule that fails: NaN is refused, an exception travels to the caller
    with pytest.raises(engine.ChicdiffError):
        e.region_test(theta=0.5, prior_var_fn=lambda df, resid: float("nan"))
    def boom(df, resid):
        raise KeyError("rule failed")
    with pytest.raises(KeyError):
        e.region_test(theta=0.5, prior_var_fn=boom)
    # without a caller's rule the library's own restatement of DESeq2's rule answers (csrc/priorvar.cpp)
    r = e.region_test(theta=0.5)
    assert r["dispPriorVar"] >= 0.25 and np.isfinite(r["pvalue"]).any()
    e.close()


def test_results_into_caller_buffers():
    """region_test(out=...) fills the caller's (here page-locked) arrays with exactly what it returns otherwise"""
    import torch
    d = synth.generate("tiny")
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    for s in range(d.S):
        e.set_sample_rows(s, d.N_rows[s], d.FM_rows[s])
    e.aggregate(fetch=False)
    ref = e.region_test(fetch="table")
    out = {k: torch.empty(d.n, dtype=torch.float64).pin_memory().numpy() for k in ("pvalue", "lfcSE")}
    got = e.region_test(fetch="table", out=out)
    assert got["pvalue"] is out["pvalue"]
    for k in ("baseMean", "log2FoldChange", "lfcSE", "stat", "pvalue", "maxCooks", "flags"):
        assert np.array_equal(got[k], ref[k], equal_nan=True), k
    with pytest.raises(ValueError):
        e.region_test(fetch="table", out={"pvalue": np.empty(d.n - 1)})
    e.close()


def test_next_batch_uploads_overlap_the_region_test():
    """The pipelined call order of bench.py's e2e steps: the rows (or the replicate tables) of batch k+1 are handed over
    between the aggregation (assembly) and the region test of batch k and travel on the context's copy stream.  Two
    different batches, alternated: every region test must return exactly what the same batch returns when it is run
    alone (the upload in flight belongs to the NEXT aggregation and must not touch the matrices of this one)."""
    import torch
    d = synth.generate("c1")
    batches = []
    perm = [d.S - 1] + list(range(1, d.S - 1)) + [0]                      # second batch: first and last replicate swapped
    for b in range(2):
        order = perm if b == 1 else list(range(d.S))
        batches.append((torch.from_numpy(np.ascontiguousarray(d.N_rows[order])).pin_memory().numpy(),
                        torch.from_numpy(np.ascontiguousarray(d.FM_rows[order])).pin_memory().numpy()))
    kw = dict(disp_prior_var=0.5, disp_prior_var_grid=0.5, fetch="table")

    def alone(N, FM):
        e = engine.Engine(0)
        e.set_design(d.X); e.set_regions(d.row_off)
        for s in range(d.S):
            e.set_sample_rows(s, N[s], FM[s])
        e.aggregate(fetch=False)
        r = e.region_test(**kw)
        e.close()
        return r
    ref = [alone(*b) for b in batches]
    e = engine.Engine(0)
    e.set_design(d.X); e.set_regions(d.row_off)
    upload = lambda b: [e.set_sample_rows(s, batches[b][0][s], batches[b][1][s]) for s in range(d.S)]
    upload(0)
    for k in range(4):
        e.aggregate(fetch=False)
        upload((k + 1) % 2)                                               # next batch on its way ...
        r = e.region_test(**kw)                                           # ... under this batch's region test
        for col in ("baseMean", "pvalue", "lfcSE", "flags"):
            assert np.array_equal(r[col], ref[k % 2][col], equal_nan=True), (k, col)
    e.close()
    # the same with the replicate tables of the assembly path
    tabs = [d.extra["tables"], [d.extra["tables"][s] for s in perm]]

    def alone_asm(tables):
        e = engine.Engine(0)
        e.set_design(d.X); e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
        e.set_regions(d.row_off); e.set_region_rows(d.row_bait, d.row_oe)
        for s in range(d.S):
            e.set_sample_tables(s, tables[s])
        e.assemble(fetch=False)
        r = e.region_test(**kw)
        e.close()
        return r
    ref = [alone_asm(t) for t in tabs]
    assert not np.array_equal(ref[0]["baseMean"], ref[1]["baseMean"])
    e = engine.Engine(0)
    e.set_design(d.X); e.set_rmap(d.frag_chr, d.frag_start, d.frag_end, 1)
    e.set_regions(d.row_off); e.set_region_rows(d.row_bait, d.row_oe)
    upload = lambda b: [e.set_sample_tables(s, tabs[b][s]) for s in range(d.S)]
    upload(0)
    for k in range(4):
        e.assemble(fetch=False)
        upload((k + 1) % 2)
        r = e.region_test(**kw)
        for col in ("baseMean", "pvalue", "lfcSE", "flags"):
            assert np.array_equal(r[col], ref[k % 2][col], equal_nan=True), (k, col)
    e.close()


def test_ihw_weight_application_on_the_device():
    """cd_ihw_apply_device (chicdiff.R:2038-2049 on the GPU: group look-up, weights, weighted p-values, BH by one radix
    sort and a minimum scan) against the host routine cd_ihw_apply, which the golden table pins: bit for bit, NA rules
    included, on the golden rows and on a 1.2 M-row synthetic table with ties and NAs."""
    import os
    e = engine.Engine(0)
    cases = []
    # the golden table of the reference's data package, with the stratum lookup test_golden.py derives from it
    from test_golden import _lookup_from_golden
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chr19_golden.npz"))
    glo, ghi, gw = _lookup_from_golden(g)
    cases.append((g["avDist"], g["pvalue"], glo, ghi, gw))
    rng = np.random.default_rng(12)
    n = 1200000
    av = np.rint(np.exp(rng.uniform(np.log(5e3), np.log(5e7), n))) * rng.choice([-1.0, 1.0], n)
    pv = rng.uniform(0, 1, n) ** 3
    pv[rng.integers(0, n, 5000)] = np.nan                               # filtered rows
    pv[rng.integers(0, n, 2000)] = pv[0]                                # ties
    pv[rng.integers(0, n, 50)] = 0.0
    lo = np.array([0.0, 9.5, 11.0, 12.5, 14.0, 15.5])
    hi = np.array([9.4, 10.9, 12.4, 13.9, 15.4, np.inf])
    w = np.array([2.1, 1.7, 1.2, 0.8, 0.4, 0.1])
    cases.append((av, pv, lo, hi, w))
    av2 = av.copy(); av2[7] = 0.0; av2[11] = np.nan                     # rows without a group poison the mean weight
    cases.append((av2[:50000], pv[:50000], lo, hi, w))
    for avd, p, lo_, hi_, w_ in cases:
        ref = engine.ihw_apply(avd, p, lo_, hi_, w_)
        got = e.ihw_apply_device(p, lo_, hi_, w_, avDist=avd)
        for k in ("group", "weight", "weighted_pvalue", "weighted_padj"):
            assert np.array_equal(got[k], ref[k], equal_nan=(k != "group")), (k, len(p))
    got = e.ihw_apply_device(g["pvalue"], glo, ghi, gw, avDist=g["avDist"])
    assert np.array_equal(got["group"], g["group"]) and int((got["weighted_padj"] < 0.05).sum()) == 2759
    with pytest.raises(engine.ChicdiffError):
        e.ihw_apply_device(pv[:10], np.array([0.0, 2.0, 2.0]), np.array([2.0, 2.0, np.inf]), w[:3], avDist=av[:10])   # 'breaks' are not unique
    with pytest.raises(engine.ChicdiffError):
        e.ihw_apply_device(pv[:10], lo, hi, w)                            # no avDist on the device
    e.close()

