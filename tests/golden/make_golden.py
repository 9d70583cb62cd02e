"""Regenerates tests/golden/chr19_golden.npz from the reference data package.

Run in the build container only (reads /root/reference, which does not exist on
the GPU box).  Source files:
  ChicdiffData/inst/extdata/CD4_Mono_results/test_results.Rds   (golden output table)
  ChicdiffData/inst/extdata/CD4_Mono_results/test_settings.Rds  (settings of that run)
  ChicdiffData/inst/extdata/designDir/chr19_GRCh37_HindIII.rmap / .baitmap

The golden table is what chicdiffPipeline() returned for the bundled chr19
CD4-vs-monocyte example (R 3.5.1); its inputs are not in the mount, so it pins
output identities (Wald p-value, independent filtering, BH, annotation,
region-width law), not the GLM numbers themselves.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from chicdiff_b200.rds import read_rds, data_frame, unwrap  # noqa: E402

REF = "/root/reference/ChicdiffData/inst/extdata"


def main():
    tab = data_frame(read_rds(os.path.join(REF, "CD4_Mono_results/test_results.Rds")))
    settings = read_rds(os.path.join(REF, "CD4_Mono_results/test_settings.Rds"))
    sv = dict(zip(unwrap(settings.attrs["names"]), settings.value))
    out = {}
    for k in ["baseMean", "log2FoldChange", "lfcSE", "stat", "pvalue", "padj", "avDist",
              "avgLogDist", "avWeights", "weight", "weighted_pvalue", "weighted_padj"]:
        out[k] = np.asarray(tab[k], dtype=np.float64)
    for k in ["group", "baitID", "maxOE", "minOE", "regionID", "OEstart", "OEend",
              "baitstart", "baitend"]:
        out[k] = np.asarray(tab[k], dtype=np.int32)
    rmap = np.loadtxt(os.path.join(REF, "designDir/chr19_GRCh37_HindIII.rmap"),
                      usecols=(1, 2, 3), dtype=np.int64)
    out["rmap_start"] = rmap[:, 0].astype(np.int32)
    out["rmap_end"] = rmap[:, 1].astype(np.int32)
    out["rmap_id"] = rmap[:, 2].astype(np.int32)
    bait_ids = []
    with open(os.path.join(REF, "designDir/chr19_GRCh37_HindIII.baitmap")) as fh:
        for line in fh:
            bait_ids.append(int(line.split()[3]))
    out["baitmap_id"] = np.asarray(bait_ids, dtype=np.int32)
    out["settings_RUexpand"] = np.asarray(unwrap(sv["RUexpand"]), dtype=np.int32)
    out["settings_score"] = np.asarray(unwrap(sv["score"]), dtype=np.float64)
    out["settings_theta_grid"] = np.asarray(unwrap(sv["theta_grid"]), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "chr19_golden.npz"), **out)
    print("wrote chr19_golden.npz with", len(out), "arrays;", len(out["pvalue"]), "rows")


if __name__ == "__main__":
    main()
