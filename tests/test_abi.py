"""CPU-side checks of the boundary: the shared library loads without a GPU, exports every symbol that
include/chicdiff_b200.h declares, fails loudly (no CPU fallback) when there is no device, and its host-only
entry points (shard planner, results adjustment) behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "chicdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cd_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from chicdiff_b200 import engine
    L = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 18
    for nm in names:
        assert hasattr(L, nm), nm
    assert sorted(engine.EXPORTED) == names
    assert b"sm_100a" in L.cd_version()


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from chicdiff_b200 import engine
    with pytest.raises(engine.ChicdiffError) as ei:
        engine.Engine(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_the_oracle():
    """Nothing under chicdiff_b200/ may import, load, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "chicdiff_b200")
    pat = re.compile(r"(^|\s)(import\s+oracle|from\s+oracle|from\s+\.\.?oracle)|liboracle|oracle[/\\]|orc_[a-z_]+\s*\(")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not pat.search(txt), (dp, f, pat.search(txt).group(0))


def test_plan_shards_is_bait_aligned_and_balanced(built):
    from chicdiff_b200 import engine, synth
    d = synth.generate("c1")
    for k in (1, 2, 3, 8):
        b = engine.plan_shards(d.region_bait, d.row_off, k)
        assert b[0] == 0 and b[-1] == d.n and np.all(np.diff(b) >= 0)
        for c in b[1:-1]:
            assert 0 < c < d.n and d.region_bait[c] != d.region_bait[c - 1]      # cuts only between baits
        rows = np.diff(d.row_off[b])
        assert rows.max() <= 1.15 * rows.mean() + 11 * 600                       # balanced up to one bait
    # degenerate inputs
    assert engine.plan_shards(np.zeros(0, np.int32), np.zeros(1, np.int64), 4).tolist() == [0, 0, 0, 0, 0]
    one_bait = engine.plan_shards(np.full(10, 7, np.int32), np.arange(11, dtype=np.int64) * 3, 2)
    assert one_bait.tolist() in ([0, 10, 10], [0, 0, 10])


def test_results_adjust_matches_oracle_restatement(built):
    from chicdiff_b200 import engine
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    n = 5000
    base = np.exp(rng.normal(3, 1.5, n))
    base[:40] = 0
    z = rng.normal(0, 1, n) + (rng.random(n) < 0.1) * rng.normal(0, 4, n) * np.sqrt(np.minimum(base, 50) / 10)
    p = np.array([O.lib().orc_wald_pvalue(v) for v in z])
    p[:40] = np.nan
    mc = rng.random(n) * 10
    mc[100:110] = 50.0
    flags = np.zeros(n, np.uint8)
    flags[105:110] = engine.FLAG_COOKS_KEEP
    adj = engine.results_adjust(base, mc, flags, p, 6, 2)
    assert abs(adj["cooksCutoff"] - 18.0) < 1e-9
    assert np.isnan(adj["pvalue"][100:105]).all() and not np.isnan(adj["pvalue"][105:110]).any()
    f = O.independent_filtering(base, adj["pvalue"])
    assert adj["filterIndex"] == f["j"] + 1
    assert np.array_equal(np.isnan(adj["padj"]), np.isnan(f["padj"]))
    ok = ~np.isnan(f["padj"])
    assert np.max(np.abs(adj["padj"][ok] - f["padj"][ok])) < 1e-15


def _split_top_level(args):
    out, depth, cur, quote = [], 0, "", None
    for ch in args:
        if quote:
            cur += ch
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
        if ch in "([{":
            depth += 1
        if ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _calls(text, opener):
    """argument strings of every `opener(...)` occurrence, parentheses balanced"""
    res, i = [], 0
    while True:
        i = text.find(opener + "(", i)
        if i < 0:
            return res
        j, depth = i + len(opener), 0
        while True:
            depth += text[j] == "("
            depth -= text[j] == ")"
            j += 1
            if depth == 0:
                break
        res.append(text[i + len(opener) + 1:j - 1])
        i = j


def test_r_glue_compiles_against_the_abi_header():
    """R/r_glue.c cannot be built without R; against declarations-only stand-ins for R's headers (tests/r_mock) the
    compiler still checks every cd_* call in it against include/chicdiff_b200.h."""
    import subprocess
    cmd = ["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-std=c11", "-I" + os.path.join(ROOT, "tests", "r_mock"),
           os.path.join(ROOT, "R", "r_glue.c")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_r_adapter_calls_match_the_glue():
    """every .Call("cdR_x", ...) of R/chicdiff_b200.R has a stub in R/r_glue.c with the same number of arguments, and
    every stub only reaches entry points the header declares"""
    glue = open(os.path.join(ROOT, "R", "r_glue.c")).read()
    glue_nc = re.sub(r"/\*.*?\*/", "", glue, flags=re.S)
    stubs = {m.group(1): len(_split_top_level(m.group(2))) for m in re.finditer(r"^SEXP (cdR_\w+)\(([^)]*)\)", glue_nc, flags=re.M)}
    assert len(stubs) >= 16
    rsrc = open(os.path.join(ROOT, "R", "chicdiff_b200.R")).read()
    rsrc = "\n".join(l.split("##")[0] for l in rsrc.splitlines())
    seen = set()
    for args in _calls(rsrc, ".Call"):
        parts = _split_top_level(args)
        name = parts[0].strip("\"'")
        assert name in stubs, name
        assert len(parts) - 1 == stubs[name], (name, len(parts) - 1, stubs[name])
        seen.add(name)
    assert {"cdR_create", "cdR_region_test", "cdR_results_resident", "cdR_ihw_apply", "cdR_assemble"} <= seen
    used = set(re.findall(r"\b(cd_[a-z_0-9]+)\s*\(", glue_nc))
    assert used <= set(declared_symbols()), used - set(declared_symbols())


def test_c_example_builds_against_the_abi_and_fails_loudly_without_a_device(built, tmp_path):
    """examples/c_abi_example.c drives the boundary from plain C.  It must compile and link against the header and the
    library; on a box without a GPU it has to stop with the library's message (status 3), never compute on the CPU."""
    import subprocess
    exe = str(tmp_path / "c_abi_example")
    libdir = os.path.join(ROOT, "chicdiff_b200")
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L" + libdir, "-lchicdiff_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    from chicdiff_b200 import engine
    try:
        engine.Engine(0).close()
        have_gpu = True
    except engine.ChicdiffError:
        have_gpu = False
    if have_gpu:
        assert r.returncode == 0, (r.stdout, r.stderr)            # at least 25 of the 40 planted regions are called
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


def test_every_context_entry_point_refuses_a_null_context(built):
    """Nothing may crash across the ABI: a NULL context gives CD_EINVAL (or 0 / NULL for the two getters)."""
    from chicdiff_b200 import engine
    engine.load_library()
    L = C.CDLL(engine._LIB_PATH)                                 # a private handle: the prototypes set below stay local
    src = open(os.path.join(ROOT, "include", "chicdiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(int|int64_t|void|const char\*)\s+(cd_[a-z_0-9]+)\s*\(([^;]*?)\)\s*;", src)
    checked = 0
    for ret, name, args in protos:
        params = _split_top_level(args)
        if not params or "cd_ctx*" not in params[0].replace(" *", "*") or "**" in params[0]:
            continue                                             # cd_create, cd_version and the context-free host routines
        f = getattr(L, name)
        f.argtypes = None
        if name == "cd_destroy":
            f.restype = None
            f(C.c_void_p(None))
            checked += 1
            continue
        f.restype = C.c_void_p if ret == "const char*" else (C.c_int64 if ret == "int64_t" else C.c_int)
        # every remaining argument as a zero of pointer width: never dereferenced because the context check comes first
        rc = f(C.c_void_p(None), *[C.c_void_p(None) if "*" in a or "[" in a else
                                   (C.c_double(0.0) if a.strip().startswith("double") else C.c_int64(0)) for a in params[1:]])
        if name == "cd_launch_count":
            assert rc == 0
        elif name == "cd_last_error":
            pass                                                 # NULL context = message of the last cd_create
        else:
            assert rc == -1, (name, rc)
        checked += 1
    assert checked >= 30
