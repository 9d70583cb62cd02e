"""Strict per-region parity report of the CUDA path against the CPU oracle (needs a GPU); the comparison itself is
tests/parity.py (what the GPU tests assert on).  Written to stdout; committed copies live under profiles/.

    python scripts/parity_strict.py [workload] [n_regions|full] [dispPriorVar for 2-vs-2 designs]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from chicdiff_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "c3"
    nreg = None if len(sys.argv) < 3 or sys.argv[2] == "full" else int(sys.argv[2])
    prior = float(sys.argv[3]) if len(sys.argv) > 3 else None
    d = synth.generate(workload, n_regions=nreg)
    t0 = time.time()
    t = parity.run_three(d, prior=prior, prior_grid=prior)
    report = ["== %s: n = %d regions, R = %d rows, S = %d, p = %d ; three runs (CUDA free, oracle on %d threads, CUDA with shared "
              "scalars) %.1f s" % (workload, d.n, d.R, d.S, d.X.shape[1], O.lib().orc_num_threads(), time.time() - t0)]
    S = parity.compare(d, t, report)
    print("\n".join(report))
    try:
        parity.assert_parity(S)
        print("   GATE: pass")
    except AssertionError as ex:
        print("   GATE: FAIL %s" % (ex,))


if __name__ == "__main__":
    main()
