## chicdiff_b200.R -- R adapter: DESeq2Wrap() on the CUDA backend.
##
## Drop-in for Chicdiff::DESeq2Wrap (Chicdiff/R/chicdiff.R:1494-1777) when chicdiff.settings$backend == "cuda".
## Same arguments, same messages, same output table (column order of chicdiff.R:1752-1762, rows ordered by
## regionID, attr(out, "theta") only for norm == "combined").  It only marshals columns; every number comes
## from libchicdiff_b200.so through R/r_glue.c.  Could not be executed in the build image (no R there);
## chicdiff_b200/api.py is the same adapter in Python and is what the tests run.
##
## Not produced under this backend: the `_DESeqObj<suffix>.Rds` DESeqDataSet of saveAuxData = TRUE.

## Dispersion prior variance for designs with 1 <= S - p <= 3 (2-vs-2; the intercept-only theta-grid fits of a
## 2-vs-2 run).  The library calls this once per dispersion fit with the residuals log(dispGeneEst) - log(dispFit) of
## the regions whose gene-wise estimate is >= 1e-6.  It restates that branch of DESeq2's estimateDispersionsPriorVar
## (SURVEY.md Appendix A.7) and runs on R's own RNG, hist() and loess(), which is what makes the value identical to
## a DESeq2 run; the RNG state of the session is saved and restored as DESeq2 does.
dispPriorVarSmallDf <- function(df, resid) {
  saved <- if (exists(".Random.seed", envir = .GlobalEnv)) get(".Random.seed", envir = .GlobalEnv) else NULL
  on.exit(if (is.null(saved)) suppressWarnings(rm(".Random.seed", envir = .GlobalEnv))
          else assign(".Random.seed", saved, envir = .GlobalEnv))
  set.seed(2)
  brks <- -20:20 / 2
  inside <- function(v) v[v > min(brks) & v < max(brks)]
  obs <- hist(inside(resid), breaks = brks, plot = FALSE)$density
  varGrid <- seq(from = 0, to = 8, length = 200)
  kl <- sapply(varGrid, function(v) {
    sim <- log(rchisq(1e4, df = df)) + rnorm(1e4, 0, sqrt(v)) - log(df)
    dens <- hist(inside(sim), breaks = brks, plot = FALSE)$density
    both <- c(obs, dens)
    small <- min(both[both > 0])
    sum(obs * (log(obs + small) - log(dens + small)))
  })
  fit <- loess(kl ~ varGrid, span = .2)
  fine <- seq(from = 0, to = 8, length = 1000)
  max(fine[which.min(predict(fit, fine))], 0.25)
}

DESeq2Wrap.cuda <- function(chicdiff.settings, RU, FullRegionData, suffix = "", theta = NULL) {
  Grid <- chicdiff.settings[["theta_grid"]]
  rmapfile <- chicdiff.settings[["rmapfile"]]
  if (is.null(theta) & !is.null(chicdiff.settings[["theta"]])) theta <- chicdiff.settings[["theta"]]
  norm <- chicdiff.settings[["norm"]]
  if (!norm %in% c("standard", "fullmean", "combined")) stop("DESeq2Wrap error: Unknown normalisation method.")
  if (!is.null(theta)) {
    if (theta == 1 & norm != "standard") {
      warning("Mixing parameter theta set to 1, equivalent to norm = \"standard\". The norm method has been reset accordingly.")
      norm <- "standard"
    }
    if (!theta & norm != "fullmean") {
      warning("Mixing parameter theta set to 0, equivalent to norm = \"fullmean\". The norm method has been reset accordingly.")
      norm <- "fullmean"
    }
  }
  ## region-contiguous rows per sample: (regionID, otherEndID) order = the order data.table sums them in
  fragData <- copy(FullRegionData)
  samples <- unique(fragData$sample)
  setkey(fragData, regionID, otherEndID)
  conditions <- sapply(samples, function(s) fragData[sample == s, condition[1]])
  X <- model.matrix(~ condition, data.frame(condition = factor(conditions)))
  one <- fragData[sample == samples[1]]
  row_off <- c(0, cumsum(as.numeric(one[, .N, by = regionID]$N)))
  n <- length(row_off) - 1L; S <- length(samples); p <- ncol(X)

  ctx <- .Call("cdR_create", as.integer(if (is.null(chicdiff.settings[["gpu"]])) 0L else chicdiff.settings[["gpu"]]))
  .Call("cdR_set_design", ctx, X)
  .Call("cdR_set_regions", ctx, row_off)
  for (i in seq_along(samples)) {
    x <- fragData[sample == samples[i]]
    .Call("cdR_set_sample_rows", ctx, i, as.integer(x$N), as.numeric(x$FullMean))
  }
  .Call("cdR_aggregate", ctx, n, S)
  if (norm == "combined" && is.null(theta)) message("Optimising scaling factors...")
  na <- NA_real_
  pv <- chicdiff.settings[["dispPriorVar"]]; pvg <- chicdiff.settings[["dispPriorVarGrid"]]
  fit <- .Call("cdR_region_test", ctx, n, S, p, match(norm, c("standard", "fullmean", "combined")) - 1L,
               if (is.null(theta) || norm != "combined") na else as.numeric(theta), as.numeric(Grid),
               if (is.null(pv)) na else pv, if (is.null(pvg)) na else pvg, dispPriorVarSmallDf)
  if (length(fit$deviances)) {
    message("Total deviances by theta (Fullmean --> Standard):")
    cat(sprintf("%f", fit$deviances), "\n", file = stderr())
  }
  if (norm == "combined") message("Theta=", fit$theta)
  message("Processing model output")
  ## results() on the arrays still in device memory (cdR_results_adjust is the host routine for gathered columns)
  adj <- .Call("cdR_results_resident", ctx, n)
  adj$pvalue[is.nan(adj$pvalue)] <- NA_real_; adj$padj[is.nan(adj$padj)] <- NA_real_

  ## annotation columns of the output table (what chicdiff.R:1700-1717 derives by three merges): one row per region,
  ## coordinates looked up in the restriction map by fragment ID.  Column order as in the reference's cbind (:1752).
  frag <- fread(rmapfile, col.names = c("chr", "start", "end", "ID"))
  ends <- RU[order(regionID), .(bait = baitID[1], lo = min(otherEndID), hi = max(otherEndID)), by = regionID]
  stopifnot(identical(ends$regionID, seq_len(nrow(ends))))
  at <- function(id) match(id, frag$ID)
  annoData <- data.table(baitID = ends$bait, maxOE = ends$hi, minOE = ends$lo, regionID = ends$regionID,
                         OEchr = frag$chr[at(ends$lo)], OEstart = frag$start[at(ends$lo)], OEend = frag$end[at(ends$hi)],
                         baitchr = frag$chr[at(ends$bait)], baitstart = frag$start[at(ends$bait)],
                         baitend = frag$end[at(ends$bait)], key = "regionID")

  label <- c(standard = "Standard DESeq2 normalisation", fullmean = "Chicago full mean-based normalisation",
             combined = "combined normalisation")[[norm]]
  message(label, ": # unweighted interactions with padj<0.05: ", sum(adj$padj < 0.05, na.rm = TRUE))
  results <- data.table(baseMean = fit$baseMean, log2FoldChange = fit$log2FoldChange, lfcSE = fit$lfcSE,
                        stat = fit$stat, pvalue = adj$pvalue, padj = adj$padj)
  out <- cbind(results, annoData)
  if (norm == "combined") attributes(out)$theta <- fit$theta
  out
}


## getFullRegionData1() on the CUDA backend (chicdiff.R:577-948) -- marshalling only.
## Per replicate it builds the same small intermediate tables the reference builds (first s_j/tblb per bait :659,
## first s_i/tlb per other end :668, first Tmean per (tblb, tlb) :680, .chicEstimateDistFun :696) as dense
## per-fragment vectors, hands them and the .chinput counts to cd_set_sample_tables, and lets cd_assemble do the
## joins, Bmean/Tmean reconstruction, count merge and region sums (stubs in r_glue.c).
getFullRegionData1.cuda <- function(chicdiff.settings, RU, is_control = FALSE, ctx) {
  rmap <- Chicago:::.readRmap(list(rmapfile = chicdiff.settings[["rmapfile"]]))
  colnames(rmap) <- c("chr", "start", "end", "ID"); setkey(rmap, ID)
  id0 <- rmap$ID[1]; nF <- nrow(rmap)
  stopifnot(identical(rmap$ID, seq.int(id0, length.out = nF)))
  .Call("cdR_set_rmap", ctx, as.integer(factor(rmap$chr)), as.integer(rmap$start), as.integer(rmap$end), as.integer(id0))
  setkey(RU, regionID, otherEndID)
  .Call("cdR_set_regions", ctx, c(0, cumsum(as.numeric(RU[, .N, by = regionID]$N))))
  .Call("cdR_set_region_rows", ctx, as.integer(RU$baitID), as.integer(RU$otherEndID))
  files <- unlist(chicdiff.settings[["chicagoData"]]); counts <- unlist(chicdiff.settings[["countData"]])
  ## countData = NULL: the reference reads the counts back from Reduce(merge, ...) over the replicates' CHiCAGO tables,
  ## an inner join (chicdiff.R:778): only pairs with a row in every replicate keep their counts
  common <- NULL
  if (is.null(counts)) {
    pairs <- lapply(files, function(f) { x <- readRDSorRDA(f); x <- if ("chicagoData" %in% class(x)) as.data.table(x@x) else setDT(x)
                                         unique(x[, .(baitID, otherEndID)]) })
    common <- Reduce(function(a, b) merge(a, b, by = c("baitID", "otherEndID")), pairs)
    setkey(common, baitID, otherEndID)
  }
  for (i in seq_along(files)) {
    x <- readRDSorRDA(files[i]); x <- if ("chicagoData" %in% class(x)) as.data.table(x@x) else setDT(x)
    setkey(x, baitID, otherEndID)
    bait <- x[, list(s_j = s_j[1], tblb = tblb[1]), by = "baitID"]
    oe <- x[, list(s_i = s_i[1], tlb = tlb[1]), by = "otherEndID"]
    tb.lv <- sort(unique(na.omit(x$tblb))); tl.lv <- sort(unique(na.omit(x$tlb)))
    tm <- x[!is.na(tblb) & !is.na(tlb), list(Tmean = Tmean[1]), by = c("tblb", "tlb")]
    tmean <- matrix(NA_real_, length(tb.lv), length(tl.lv)); tmean[cbind(match(tm$tblb, tb.lv), match(tm$tlb, tl.lv))] <- tm$Tmean
    s_j <- rep(NA_real_, nF); s_j[bait$baitID - id0 + 1L] <- bait$s_j
    tblb <- rep(-1L, nF); tblb[bait$baitID - id0 + 1L] <- ifelse(is.na(bait$tblb), -1L, match(bait$tblb, tb.lv) - 1L)
    s_i <- rep(NA_real_, nF); s_i[oe$otherEndID - id0 + 1L] <- oe$s_i
    tlb <- rep(-1L, nF); tlb[oe$otherEndID - id0 + 1L] <- ifelse(is.na(oe$tlb), -1L, match(oe$tlb, tl.lv) - 1L)
    dfp <- .chicEstimateDistFun(x)
    cnt <- if (is.null(counts)) x[common, .(baitID, otherEndID, N), nomatch = 0L] else fread(counts[i])[, .(baitID, otherEndID, N)]
    # rows whose bait is not a fragment of the rmap would be dropped by tabulate() but stay in cnt$otherEndID / cnt$N,
    # and every offset after them would point at the wrong rows: drop them first
    cnt <- cnt[baitID >= id0 & baitID < id0 + nF]
    setkey(cnt, baitID, otherEndID)
    cnt_off <- c(0, cumsum(tabulate(cnt$baitID - id0 + 1L, nbins = nF)))
    .Call("cdR_set_sample_tables", ctx, i, s_j, tblb, s_i, tlb, t(tmean),
          c(dfp$cubicFit, dfp$obs.min, dfp$obs.max, dfp$head.coef, dfp$tail.coef),
          as.numeric(cnt_off), as.integer(cnt$otherEndID), as.integer(cnt$N))
  }
  n <- length(unique(RU$regionID))
  .Call("cdR_assemble", ctx, n, length(files), TRUE)      # -> list(K, FullMean, avDist); per-row columns via cdR_get_sample_rows
}


## IHWcorrection(), "apply to test data" block (chicdiff.R:2038-2049) -- marshalling only.  `out` is the DESeq2Wrap
## table with avDist attached (:1965-1967), `distLookup` the table learned from the control set (:2013-2033).
## Returns `out` with group, avWeights, weight, weighted_pvalue, weighted_padj, ordered by group like the
## reference's merge() leaves it.
ihwApply.cuda <- function(out, distLookup) {
  r <- .Call("cdR_ihw_apply", as.numeric(out$avDist), as.numeric(out$pvalue), as.numeric(distLookup$minLogDist),
             as.numeric(distLookup$maxLogDist), as.numeric(distLookup$avWeights))
  out[, avgLogDist := log(abs(avDist))]
  out$group <- r$group
  out$avWeights <- distLookup$avWeights[r$group]
  out$weight <- ifelse(is.nan(r$weight), NA_real_, r$weight)
  out$weighted_pvalue <- ifelse(is.nan(r$weighted_pvalue) & !is.nan(out$pvalue), NA_real_, r$weighted_pvalue)
  out$weighted_padj <- ifelse(is.nan(r$weighted_padj), NA_real_, r$weighted_padj)
  setkey(out, group)
  out
}
