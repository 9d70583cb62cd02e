"""Host-side logic of the Python mirror of the reference API (chicdiff_b200/api.py), no GPU needed: marshalling of the
long table into region-contiguous rows, the model matrix, settings handling, the restriction-map reader."""
import numpy as np
import pytest

from chicdiff_b200 import api, synth


def test_region_rows_orders_by_region_then_other_end_whatever_the_input_order():
    d = synth.generate("tiny")
    RU, frd, _ = synth.to_reference_tables(d)
    rng = np.random.default_rng(0)
    perm = rng.permutation(len(frd["sample"]))                          # data.table row order is not guaranteed
    shuffled = {k: np.asarray(v)[perm] for k, v in frd.items()}
    order = list(dict.fromkeys(frd["sample"]))                          # the caller knows the replicate order
    samples, conditions, region_ids, region_bait, row_off, N, FM = api.region_rows(RU, shuffled, sample_order=order)
    assert samples == order and conditions == list(d.conditions)
    assert np.array_equal(region_ids, np.arange(1, d.n + 1))
    assert np.array_equal(row_off, d.row_off)
    assert np.array_equal(region_bait, d.region_bait)
    assert np.array_equal(N, d.N_rows)
    assert np.array_equal(FM, d.FM_rows, equal_nan=True)
    # first-appearance order of the samples when none is given (R's unique())
    s2 = api.region_rows(RU, frd)[0]
    assert s2 == order


def test_region_rows_rejects_samples_with_different_rows():
    d = synth.generate("tiny")
    RU, frd, _ = synth.to_reference_tables(d)
    keep = np.ones(len(frd["sample"]), bool)
    keep[3] = False                                                     # one replicate misses a fragment
    broken = {k: np.asarray(v)[keep] for k, v in frd.items()}
    with pytest.raises(ValueError):
        api.region_rows(RU, broken)


def test_model_matrix_levels_are_alphabetical_and_batch_columns_come_first():
    X, lv = api.model_matrix(["Mono", "Mono", "CD4", "CD4"])
    assert lv == ["CD4", "Mono"]                                        # reference level = alphabetically first (chicdiff.R:1559)
    assert np.array_equal(X, np.array([[1, 1], [1, 1], [1, 0], [1, 0]], float))
    X, _ = api.model_matrix(["a", "a", "b", "b"], batch=["x", "y", "x", "y"])
    assert np.array_equal(X, np.array([[1, 0, 0], [1, 1, 0], [1, 0, 1], [1, 1, 1]], float))
    X, _ = api.model_matrix(["a", "a", "b", "b"], batch=["x", "x", "x", "x"])
    assert X.shape == (4, 2)                                            # a single batch adds no column
    with pytest.raises(ValueError):
        api.model_matrix(["a", "b", "c"])


def test_settings_defaults_follow_the_reference():
    st = api.defaultChicdiffSettings()
    assert st["RUexpand"] == 5 and st["score"] == 5 and st["norm"] == "combined" and st["theta"] is None   # chicdiff.R:3-24
    assert np.allclose(st["theta_grid"], [0, 0.25, 0.5, 0.75, 1.0])
    assert st["saveAuxData"] is False and st["backend"] == "cuda"


def test_unknown_normalisation_is_refused_before_any_device_work():
    st = api.defaultChicdiffSettings()
    st["norm"] = "median"
    with pytest.raises(ValueError) as ei:
        api.DESeq2Wrap(st, {}, {})
    assert "Unknown normalisation method" in str(ei.value)              # the reference's message (chicdiff.R:1508)


def test_read_rmap_sorts_by_fragment_id_and_strips_quotes(tmp_path):
    f = tmp_path / "x.rmap"
    f.write_text('"chr1"\t100\t200\t3\nchr1\t1\t99\t2\n\n"chrX" 201 300 4\n')
    r = api.read_rmap(str(f))
    assert list(r["ID"]) == [2, 3, 4] and list(r["chr"]) == ["chr1", "chr1", "chrX"]
    assert list(r["start"]) == [1, 100, 201] and list(r["end"]) == [99, 200, 300]


def test_counts_reconstructed_from_chicago_tables_keep_only_pairs_seen_in_every_replicate():
    """countData = NULL in the reference: Reduce(merge, ...) over the replicates' (baitID, otherEndID, N) columns is an
    inner join (chicdiff.R:778), so a pair missing from one replicate loses its counts in all of them."""
    reps = [
        {"baitID": np.array([1, 1, 2, 3]), "otherEndID": np.array([10, 11, 20, 30]), "N": np.array([5, 6, 7, 8])},
        {"baitID": np.array([1, 2, 2, 3]), "otherEndID": np.array([10, 20, 21, 30]), "N": np.array([1, 2, 3, 4])},
        {"baitID": np.array([3, 1, 2]), "otherEndID": np.array([30, 10, 20]), "N": np.array([9, 9, 9])},
    ]
    out = api.reconstruct_count_tables(reps)
    for t in out:
        pairs = sorted(zip(t["baitID"].tolist(), t["otherEndID"].tolist()))
        assert pairs == [(1, 10), (2, 20), (3, 30)]
    assert out[0]["N"].tolist() == [5, 7, 8] and out[1]["N"].tolist() == [1, 2, 4] and sorted(out[2]["N"].tolist()) == [9, 9, 9]
    # a random check against a dictionary-based inner join
    rng = np.random.default_rng(1)
    reps = []
    for _ in range(4):
        m = 500
        b = rng.integers(1, 20, m); o = rng.integers(1, 60, m)
        key = np.unique(b * 1000 + o)
        reps.append({"baitID": key // 1000, "otherEndID": key % 1000, "N": rng.integers(1, 50, len(key))})
    out = api.reconstruct_count_tables(reps)
    common = set.intersection(*[set(zip(t["baitID"].tolist(), t["otherEndID"].tolist())) for t in reps])
    for t_in, t_out in zip(reps, out):
        want = {(b, o): n for b, o, n in zip(t_in["baitID"].tolist(), t_in["otherEndID"].tolist(), t_in["N"].tolist()) if (b, o) in common}
        got = {(b, o): n for b, o, n in zip(t_out["baitID"].tolist(), t_out["otherEndID"].tolist(), t_out["N"].tolist())}
        assert got == want
