"""Development aid: synth.generate with a pickle cache under /tmp, so that several profiling runs of one GPU session do not
each spend a minute regenerating the same 2.1 M-region set."""
import os
import pickle
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chicdiff_b200 import synth  # noqa: E402


def cached_generate(workload, n_regions=None, seed_offset=0):
    path = "/tmp/chicdiff_synth_%s_%s_%d.pkl" % (workload, n_regions, seed_offset)
    if os.path.exists(path):
        with open(path, "rb") as fh:
            return pickle.load(fh)
    d = synth.generate(workload, n_regions=n_regions, seed_offset=seed_offset)
    try:
        with open(path + ".tmp", "wb") as fh:
            pickle.dump(d, fh, protocol=4)
        os.replace(path + ".tmp", path)
    except Exception:
        pass
    return d
