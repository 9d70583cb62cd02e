"""Deterministic synthetic PCHi-C generator.

Produces the flat arrays the C ABI consumes -- the region universe in the layout of
getRegionUniverse() (chicdiff.R:369-426: every peak (bait, oe) becomes the fragment window
``.expandAvoidBait(bait, oe, RUexpand)``, chicdiff.R:353-367) and, per replicate, the per-row
count ``N`` and expected background ``FullMean = Bmean + Tmean`` columns of the long table
built by getFullRegionData1() (chicdiff.R:577-948).  Rows are region-contiguous: sorted by
(regionID, otherEndID).

Counts for a (bait, other end) pair are drawn once and shared by every region that contains
the pair, as in real data.  Seeds: ``numpy.random.Generator(PCG64(20261018 + config_index))``.
"""
from dataclasses import dataclass, field
import numpy as np

CONFIGS = {
    # name: (config_index, chromosomes, fragments, baits, regions, reps per condition, batches)
    "tiny": (9, 1, 3000, 120, 2000, (3, 3), 1),
    "c1": (0, 1, 11542, 1052, 24863, (2, 2), 1),
    "c2": (1, 1, 40000, 1000, 100000, (2, 2), 1),
    "c3": (2, 23, 840000, 22000, 2000000, (3, 3), 1),
    "c4": (3, 23, 840000, 22000, 2000000, (8, 8), 2),
}


@dataclass
class SynthData:
    name: str
    S: int
    conditions: list                 # condition label per sample (reference level = alphabetically first)
    batch: np.ndarray                # batch index per sample
    X: np.ndarray                    # S x p design (Intercept[, batch], condition)
    frag_chr: np.ndarray             # per fragment (ID = index + 1)
    frag_start: np.ndarray
    frag_end: np.ndarray
    bait_ids: np.ndarray
    region_bait: np.ndarray          # n
    region_seed: np.ndarray          # n
    row_off: np.ndarray              # n + 1 (int64), CSR into rows
    row_oe: np.ndarray               # R otherEndID per row
    row_bait: np.ndarray             # R baitID per row
    N_rows: np.ndarray               # S x R int32
    FM_rows: np.ndarray              # S x R float64, NaN = NA
    true_lfc: np.ndarray             # n
    extra: dict = field(default_factory=dict)

    @property
    def n(self):
        return len(self.region_bait)

    @property
    def R(self):
        return len(self.row_oe)


def design_matrix(conditions, batch=None):
    """model.matrix(~ condition) or (~ batch + condition); reference level alphabetical."""
    lv = sorted(set(conditions))
    assert len(lv) == 2
    cond = np.array([1.0 if c == lv[1] else 0.0 for c in conditions])
    cols = [np.ones(len(conditions))]
    if batch is not None and len(set(batch.tolist())) > 1:
        assert len(set(batch.tolist())) == 2
        cols.append((np.asarray(batch) == max(batch)).astype(np.float64))
    cols.append(cond)
    return np.stack(cols, axis=1)


def expand_avoid_bait(bait, oe, s):
    """Vectorised .expandAvoidBait (chicdiff.R:353-367): returns inclusive [lo, hi] fragment IDs."""
    far = np.abs(bait - oe) > s + 1
    lo = np.where(far | (oe < bait), oe - s, bait + 2)
    hi = np.where(far | (oe > bait), oe + s, bait - 2)
    return lo, hi


def dist_fun(d, cubic=(-3.0, 2.6, -0.33, 0.009), obs=(np.log(1e4), np.log(1.5e6))):
    """Chicago-style distance function: cubic in log d on [obs.min, obs.max], C1 linear tails."""
    l = np.log(np.maximum(d, 1.0))
    c0, c1, c2, c3 = cubic

    def cub(x):
        return c0 + c1 * x + c2 * x * x + c3 * x ** 3

    def dcub(x):
        return c1 + 2 * c2 * x + 3 * c3 * x * x

    lo, hi = obs
    out = cub(l)
    out = np.where(l < lo, cub(lo) + dcub(lo) * (l - lo), out)
    out = np.where(l > hi, cub(hi) + dcub(hi) * (l - hi), out)
    return np.exp(out)


def dist_fun_params(cubic, obs):
    """The 10 numbers .chicEstimateDistFun returns (chicdiff.R:559-569): cubicFit[4], obs.min, obs.max,
    head.coef[2], tail.coef[2] with the C1 continuation beta = f'(l), alpha = f(l) - beta l."""
    c0, c1, c2, c3 = cubic
    out = list(cubic) + [obs[0], obs[1]]
    for l in obs:
        beta = c1 + 2 * c2 * l + 3 * c3 * l * l
        alpha = c0 + (c1 - beta) * l + c2 * l * l + c3 * l ** 3
        out += [alpha, beta]
    return np.asarray(out, dtype=np.float64)


def r_round_half_even(x):
    return np.rint(x)


def generate(name="tiny", n_regions=None, reps=None, ru_expand=5, ensure_nonzero=True, seed_offset=0):
    idx, n_chr, F, B, n_target, reps_default, n_batch = CONFIGS[name]
    if n_regions is not None:
        scale = n_regions / n_target
        n_target = int(n_regions)
        # keep regions-per-bait roughly constant when resizing
        B = max(20, int(round(B * scale)))
        F = max(40 * 5, int(round(F * scale))) if scale < 1 else int(F * max(1.0, scale))
    if reps is None:
        reps = reps_default
    rng = np.random.Generator(np.random.PCG64(20261018 + idx + seed_offset))

    # 1. genome
    n_chr = max(1, min(n_chr, F // 2000)) if F < 2000 * n_chr else n_chr
    chr_sizes = np.full(n_chr, F // n_chr)
    chr_sizes[: F - chr_sizes.sum()] += 1
    frag_chr = np.repeat(np.arange(1, n_chr + 1), chr_sizes).astype(np.int32)
    length = np.clip(np.exp(rng.normal(7.9, 0.9, F)), 100, 60000).astype(np.int64)
    frag_end = np.empty(F, np.int64)
    frag_start = np.empty(F, np.int64)
    pos = 0
    for c in range(n_chr):
        sl = slice(pos, pos + chr_sizes[c])
        e = np.cumsum(length[sl])
        frag_end[sl] = e
        frag_start[sl] = e - length[sl] + 1
        pos += chr_sizes[c]
    chr_first = np.concatenate([[0], np.cumsum(chr_sizes)[:-1]]) + 1      # first fragment ID per chr
    chr_last = np.cumsum(chr_sizes)
    mid_even = frag_start + frag_end                                         # 2 * midpoint

    # 2. baits: Bernoulli, never adjacent
    is_bait = rng.random(F) < (B / F)
    is_bait[1:] &= ~is_bait[:-1]
    bait_ids = (np.flatnonzero(is_bait) + 1).astype(np.int64)
    B = len(bait_ids)

    # 3. seeds
    per_bait = rng.geometric(min(1.0, B / (1.33 * n_target)), B)      # ~25 % of draws are rejected below
    bait_rep = np.repeat(bait_ids, per_bait)
    c_of = frag_chr[bait_rep - 1] - 1
    span = np.maximum(chr_sizes[c_of] / 2.0, 4.0)
    dist = np.rint(np.exp(rng.uniform(np.log(3.0), np.log(span)))).astype(np.int64)
    dist *= np.where(rng.random(len(dist)) < 0.5, -1, 1)
    oe = bait_rep + dist
    ok = (np.abs(dist) > 1) & (oe >= chr_first[c_of]) & (oe <= chr_last[c_of])
    bait_rep, oe, c_of = bait_rep[ok], oe[ok], c_of[ok]
    key = bait_rep * (F + 2) + oe
    _, first = np.unique(key, return_index=True)
    bait_rep, oe, c_of = bait_rep[first], oe[first], c_of[first]            # sorted by (bait, seed)
    n = len(bait_rep)

    # 4. expand to rows
    lo, hi = expand_avoid_bait(bait_rep, oe, ru_expand)
    lo = np.maximum(lo, chr_first[c_of])
    hi = np.minimum(hi, chr_last[c_of])
    width = (hi - lo + 1).astype(np.int64)
    assert (width > 0).all()
    row_off = np.concatenate([[0], np.cumsum(width)]).astype(np.int64)
    R = int(row_off[-1])
    row_region = np.repeat(np.arange(n), width)
    row_oe = (np.arange(R) - row_off[row_region] + lo[row_region]).astype(np.int64)
    row_bait = bait_rep[row_region]

    # 5. unique (bait, oe) pairs; owner region = first region that contains the pair
    pkey = row_bait * (F + 2) + row_oe
    upair, pfirst, pinv = np.unique(pkey, return_index=True, return_inverse=True)
    U = len(upair)
    owner = row_region[pfirst]
    p_bait = row_bait[pfirst]
    p_oe = row_oe[pfirst]
    # distance (chicdiff.R:648): round(((oe.start+oe.end) - (bait.start+bait.end))/2), half to even
    p_dist = r_round_half_even((mid_even[p_oe - 1] - mid_even[p_bait - 1]) / 2.0)

    # region-level truth
    true_lfc = np.where(rng.random(n) < 0.10, rng.normal(0, 1.5, n), 0.0)
    enrich = np.exp(rng.normal(np.log(40.0), 1.2, n))

    S = sum(reps)
    conditions = ["A_ctrl"] * reps[0] + ["B_test"] * reps[1]
    cond01 = np.array([0] * reps[0] + [1] * reps[1])
    batch = np.zeros(S, np.int64)
    if n_batch > 1:
        for g in (0, 1):
            k = np.flatnonzero(cond01 == g)
            batch[k[len(k) // 2:]] = 1
    X = design_matrix(conditions, batch if n_batch > 1 else None)

    bait_index = np.searchsorted(bait_ids, p_bait)
    N_rows = np.empty((S, R), np.int32)
    FM_rows = np.empty((S, R), np.float64)
    batch_eff = np.exp2(rng.normal(0, 0.3, (n, 2))) if n_batch > 1 else None
    fm_u_mean = np.zeros(U)
    fm_u_all = []
    cubic = (-3.0, 2.6, -0.33, 0.009)
    obs = (np.log(1e4), np.log(1.5e6))
    distfun = dist_fun_params(cubic, obs)
    tables = []
    for s in range(S):
        libsize = np.exp(rng.normal(0, 0.25))
        s_j = np.exp(rng.normal(0, 0.5, B))
        s_j_na = rng.random(B) < 0.02
        s_i = np.exp(rng.normal(0, 0.3, F))
        tblb = np.minimum((np.argsort(np.argsort(s_j)) * 5) // B, 4)
        tlb = np.minimum((np.argsort(np.argsort(s_i)) * 5) // F, 4)
        tmean = np.exp(rng.uniform(np.log(1e-3), np.log(1e-1), (5, 5)))
        # other ends CHiCAGO never saw in this replicate: s_i = NA (-> 1, chicdiff.R:672) and tlb = NA
        # (-> lowest Tmean of the bait's tblb, chicdiff.R:689-692)
        oe_unseen = rng.random(F) < 0.01
        # per-fragment lookup tables in the form the C ABI takes (cd_sample_tables)
        sj_frag = np.full(F, np.nan)
        sj_frag[bait_ids - 1] = np.where(s_j_na, np.nan, s_j)
        tblb_frag = np.full(F, -1, np.int32)
        tblb_frag[bait_ids - 1] = tblb
        si_frag = np.where(oe_unseen, np.nan, s_i)
        tlb_frag = np.where(oe_unseen, -1, tlb).astype(np.int32)
        tables.append(dict(s_j=sj_frag, tblb=tblb_frag, s_i=si_frag, tlb=tlb_frag, tmean=tmean, distfun=distfun, libsize=libsize))
        # FullMean by the reference's rules (NumPy restatement, independent of the C oracle and the kernel)
        si_eff = np.where(oe_unseen, 1.0, s_i)[p_oe - 1]
        bmean = s_j[bait_index] * si_eff * dist_fun(np.abs(p_dist), cubic, obs)
        tb = tblb[bait_index]
        tm = np.where(oe_unseen[p_oe - 1], tmean.min(axis=1)[tb], tmean[tb, tlb[p_oe - 1]])
        fm = bmean + tm
        fm_u_all.append((fm, s_j_na[bait_index], libsize))
        fm_u_mean += np.log(fm)
    fm_u_mean = np.exp(fm_u_mean / S)
    # NB truth: alpha_i = 0.05 + 2 / mubar_i with mubar_i the region's mean expected count
    mu_pair_base = enrich[owner] * fm_u_mean
    reg_mu = np.add.reduceat(mu_pair_base[pinv], row_off[:-1])
    alpha = 0.05 + 2.0 / np.maximum(reg_mu, 1e-3)
    cnt_u = np.empty((S, U), np.int32)
    for s in range(S):
        fm, na_mask, libsize = fm_u_all[s]
        sign = 0.5 if cond01[s] == 1 else -0.5
        mu = libsize * enrich[owner] * fm * np.exp2(sign * true_lfc[owner])
        if batch_eff is not None:
            mu = mu * batch_eff[owner, batch[s]]
        g_region = rng.gamma(1.0 / alpha, alpha)                    # one multiplier per (region, sample)
        cnt_u[s] = rng.poisson(g_region[owner] * mu)
        FM_rows[s] = np.where(na_mask, np.nan, fm)[pinv]
    if ensure_nonzero:
        # the theta grid sums deviances without na.rm (chicdiff.R:1647): test sets have no all-zero region
        tot_u = cnt_u.astype(np.int64).sum(axis=0)
        tot = np.add.reduceat(tot_u[pinv], row_off[:-1])
        zero = np.flatnonzero(tot == 0)
        cnt_u[0, pinv[row_off[zero]]] += 1
    for s in range(S):
        N_rows[s] = cnt_u[s][pinv]
    # sparse per-replicate count tables (the .chinput / CHiCAGO rows): N > 0 pairs of the region universe
    # plus a halo of pairs outside it, sorted by (baitID, otherEndID), CSR by bait fragment ID
    n_halo = int(0.3 * U)
    hb = bait_ids[rng.integers(0, B, n_halo)]
    ho = hb + np.rint(rng.normal(0, 40, n_halo)).astype(np.int64)
    okh = (ho >= 1) & (ho <= F) & (np.abs(ho - hb) > 1)
    hkey = np.unique(hb[okh] * (F + 2) + ho[okh])
    hkey = hkey[~np.isin(hkey, upair)]
    for s in range(S):
        nz = cnt_u[s] > 0
        keep_h = rng.random(len(hkey)) < 0.5
        keys = np.concatenate([upair[nz], hkey[keep_h]])
        vals = np.concatenate([cnt_u[s][nz], 1 + rng.poisson(1.0, int(keep_h.sum()))]).astype(np.int32)
        o = np.argsort(keys, kind="stable")
        keys, vals = keys[o], vals[o]
        tables[s]["cnt_off"] = np.searchsorted(keys, np.arange(1, F + 2) * (F + 2)).astype(np.int64)
        tables[s]["cnt_oe"] = (keys % (F + 2)).astype(np.int32)
        tables[s]["cnt_N"] = vals
    return SynthData(name=name, S=S, conditions=conditions, batch=batch, X=X,
                     frag_chr=frag_chr, frag_start=frag_start, frag_end=frag_end,
                     bait_ids=bait_ids, region_bait=bait_rep.astype(np.int32), region_seed=oe.astype(np.int32),
                     row_off=row_off, row_oe=row_oe.astype(np.int32), row_bait=row_bait.astype(np.int32),
                     N_rows=N_rows, FM_rows=FM_rows, true_lfc=true_lfc,
                     extra=dict(alpha=alpha, pair_index=pinv, n_pairs=U, tables=tables))


def to_reference_tables(d):
    """The same data in the shape of the reference's R objects: RU (chicdiff.R:392-397), the long
    FullRegionData table after melt (chicdiff.R:912-925, sample-major blocks) and the rmap columns."""
    n, R, S = d.n, d.R, d.S
    region_of_row = np.repeat(np.arange(1, n + 1), np.diff(d.row_off))
    RU = {"baitID": d.row_bait.astype(np.int64), "regionID": region_of_row.astype(np.int64), "otherEndID": d.row_oe.astype(np.int64)}
    names = []
    counts = {}
    for c in d.conditions:
        counts[c] = counts.get(c, 0) + 1
        names.append("%s.rep%d" % (c, counts[c]))
    frd = {
        "baitID": np.tile(RU["baitID"], S), "otherEndID": np.tile(RU["otherEndID"], S), "regionID": np.tile(RU["regionID"], S),
        "sample": np.repeat(np.asarray(names), R), "N": d.N_rows.reshape(-1), "FullMean": d.FM_rows.reshape(-1),
        "condition": np.repeat(np.asarray(d.conditions), R),
    }
    F = len(d.frag_chr)
    rmap = {"chr": d.frag_chr.astype(str), "start": d.frag_start, "end": d.frag_end, "ID": np.arange(1, F + 1)}
    return RU, frd, rmap


def chicago_rows(d, s, seed=0):
    """The rows of replicate s's CHiCAGO table that countput uses (chicdiff.R:715-721): (baitID, otherEndID, N,
    Bmean, score) for every observed pair, sorted by (baitID, otherEndID) like the keyed table."""
    t = d.extra["tables"][s]
    F = len(d.frag_chr)
    per_bait = np.diff(t["cnt_off"])
    bait = np.repeat(np.arange(1, F + 1), per_bait).astype(np.int32)
    rng = np.random.Generator(np.random.PCG64(777 + 31 * s + seed))
    m = len(bait)
    bmean = np.exp(rng.normal(-2.0, 1.0, m))
    score = np.maximum(0.0, rng.normal(2.0, 3.0, m))
    score[rng.random(m) < 0.001] = np.nan
    return dict(baitID=bait, otherEndID=t["cnt_oe"].copy(), N=t["cnt_N"].copy(), Bmean=bmean, score=score)


def chicago_table(d, s, seed=0, binsize=20000, nbins=75):
    """Replicate s as a reference-shaped CHiCAGO table (the columns getFullRegionData1 reads, chicdiff.R:614-696):
    one row per observed pair whose other end CHiCAGO has seen, with s_j / tblb per bait, s_i / tlb per other end,
    Tmean per (tblb, tlb), distSign, distbin / refBinMean (the distance function sampled at the bin midpoints),
    plus the replicate's own Bmean and score.  The matching .chinput is `chinput_table`."""
    t = d.extra["tables"][s]
    F = len(d.frag_chr)
    per_bait = np.diff(t["cnt_off"])
    bait = np.repeat(np.arange(1, F + 1), per_bait).astype(np.int64)
    oe = t["cnt_oe"].astype(np.int64)
    seen = ~np.isnan(t["s_i"][oe - 1])                       # other ends without s_i were never seen by CHiCAGO
    bait, oe, N = bait[seen], oe[seen], t["cnt_N"][seen]
    mid2 = d.frag_start + d.frag_end
    same = d.frag_chr[oe - 1] == d.frag_chr[bait - 1]
    dist = np.where(same, np.rint((mid2[oe - 1] - mid2[bait - 1]) / 2.0), np.nan)
    tb = t["tblb"][bait - 1]; tl = t["tlb"][oe - 1]
    mids = np.rint(binsize / 2.0) + binsize * np.arange(nbins)
    k = np.minimum(np.floor(np.abs(np.nan_to_num(dist)) / binsize).astype(np.int64), nbins)
    cubic = t["distfun"][:4]
    ref_at = np.exp(cubic[0] + cubic[1] * np.log(mids) + cubic[2] * np.log(mids) ** 2 + cubic[3] * np.log(mids) ** 3)
    inb = (k < nbins) & same
    refbin = np.where(inb, ref_at[np.minimum(k, nbins - 1)], np.nan)
    distbin = np.array(["(%d,%d]" % (binsize * kk, binsize * (kk + 1)) if ok else None for kk, ok in zip(k, inb)], dtype=object)
    rng = np.random.Generator(np.random.PCG64(4242 + 17 * s + seed))
    m = len(bait)
    return dict(baitID=bait, otherEndID=oe, N=N, distSign=dist, s_j=t["s_j"][bait - 1], s_i=t["s_i"][oe - 1],
                tblb=np.array(["tb%d" % v for v in tb], dtype=object), tlb=np.array(["tl%d" % v for v in tl], dtype=object),
                Tmean=t["tmean"][tb, tl], Bmean=np.exp(rng.normal(-2, 1, m)), score=np.maximum(0.0, rng.normal(2, 3, m)),
                distbin=distbin, refBinMean=refbin, name="rep%d" % (s + 1))


def chinput_table(d, s):
    """The replicate's .chinput (chicdiff.R:828): every pair with N >= 1, seen by CHiCAGO or not."""
    t = d.extra["tables"][s]
    F = len(d.frag_chr)
    bait = np.repeat(np.arange(1, F + 1), np.diff(t["cnt_off"])).astype(np.int64)
    return dict(baitID=bait, otherEndID=t["cnt_oe"].astype(np.int64), N=t["cnt_N"])


def chinput_text(d, s, max_rows=None):
    """The replicate's .chinput as text: '#' comment line, header, then baitID otherEndID N otherEndLen distSign
    (tab separated; distSign NA across chromosomes), as written by CHiCAGO's bam2chicago."""
    t = chinput_table(d, s)
    bait, oe, N = t["baitID"], t["otherEndID"], t["N"]
    if max_rows is not None:
        bait, oe, N = bait[:max_rows], oe[:max_rows], N[:max_rows]
    ln = (d.frag_end - d.frag_start + 1)[oe - 1]
    mid2 = d.frag_start + d.frag_end
    same = d.frag_chr[oe - 1] == d.frag_chr[bait - 1]
    dist = np.rint((mid2[oe - 1] - mid2[bait - 1]) / 2.0).astype(np.int64)
    lines = ["##\tsamplename=rep%d\tbamname=synthetic.bam\tbaitmapfile=x.baitmap\tdigestfile=x.rmap" % (s + 1),
             "baitID\totherEndID\tN\totherEndLen\tdistSign"]
    for b, o, n, l, sm, dd in zip(bait.tolist(), oe.tolist(), N.tolist(), ln.tolist(), same.tolist(), dist.tolist()):
        lines.append("%d\t%d\t%d\t%d\t%s" % (b, o, n, l, dd if sm else "NA"))
    return ("\n".join(lines) + "\n").encode()
