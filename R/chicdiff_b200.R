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
  ## chicdiff.R:1559 hard-codes `~ condition`; chicdiff.settings$batch (one label per sample, in sample order) adds the
  ## covariate of BASELINE.json configs[3]: `~ batch + condition`, the tested coefficient stays the last column
  batch <- chicdiff.settings[["batch"]]
  X <- if (is.null(batch)) model.matrix(~ condition, data.frame(condition = factor(conditions)))
       else model.matrix(~ batch + condition, data.frame(batch = factor(batch), condition = factor(conditions)))
  one <- fragData[sample == samples[1]]
  row_off <- c(0, cumsum(as.numeric(one[, .N, by = regionID]$N)))
  n <- length(row_off) - 1L; S <- length(samples); p <- ncol(X)
  norm_code <- match(norm, c("standard", "fullmean", "combined")) - 1L
  na <- NA_real_
  theta_arg <- if (is.null(theta) || norm != "combined") na else as.numeric(theta)
  pv <- chicdiff.settings[["dispPriorVar"]]; pvg <- chicdiff.settings[["dispPriorVarGrid"]]
  gpus <- if (is.null(chicdiff.settings[["gpus"]])) 1L else as.integer(chicdiff.settings[["gpus"]])

  if (gpus > 1L) {
    ## several GPUs from this one R session: the library shards the regions by bait and runs one host thread per GPU
    ## (cd_multi_*).  The Monte-Carlo prior-variance rule needs all residuals in R: pass dispPriorVar for 2-vs-2 designs.
    ctx <- .Call("cdR_multi_create", gpus)
    .Call("cdR_multi_set_design", ctx, X)
    .Call("cdR_multi_set_regions", ctx, row_off, as.integer(one[, baitID[1], by = regionID]$V1))
    for (i in seq_along(samples)) {
      x <- fragData[sample == samples[i]]
      .Call("cdR_multi_set_sample_rows", ctx, i, as.integer(x$N), as.numeric(x$FullMean))
    }
    .Call("cdR_multi_aggregate", ctx, n, S)
    if (norm == "combined" && is.null(theta)) message("Optimising scaling factors...")
    fit <- .Call("cdR_multi_region_test", ctx, n, S, norm_code, theta_arg, as.numeric(Grid),
                 if (is.null(pv)) na else pv, if (is.null(pvg)) na else pvg)
  } else {
    ctx <- .Call("cdR_create", as.integer(if (is.null(chicdiff.settings[["gpu"]])) 0L else chicdiff.settings[["gpu"]]))
    .Call("cdR_set_design", ctx, X)
    .Call("cdR_set_regions", ctx, row_off)
    for (i in seq_along(samples)) {
      x <- fragData[sample == samples[i]]
      .Call("cdR_set_sample_rows", ctx, i, as.integer(x$N), as.numeric(x$FullMean))
    }
    .Call("cdR_aggregate", ctx, n, S)
    if (norm == "combined" && is.null(theta)) message("Optimising scaling factors...")
    fit <- .Call("cdR_region_test", ctx, n, S, p, norm_code, theta_arg, as.numeric(Grid),
                 if (is.null(pv)) na else pv, if (is.null(pvg)) na else pvg, dispPriorVarSmallDf)
  }
  if (length(fit$deviances)) {
    message("Total deviances by theta (Fullmean --> Standard):")
    cat(sprintf("%f", fit$deviances), "\n", file = stderr())
  }
  if (norm == "combined") message("Theta=", fit$theta)
  message("Processing model output")
  ## results() on the arrays still in device memory; with several GPUs the columns were gathered: host routine
  adj <- if (gpus > 1L) .Call("cdR_results_adjust", S, p, fit$baseMean, fit$maxCooks, fit$flags, fit$pvalue)
         else .Call("cdR_results_resident", ctx, n)
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
## Per replicate the raw CHiCAGO columns go to the device as they are (cdR_build_sample_tables): the keyed joins, the
## setkey sorts and the first-per-bait / per-other-end / per-(tblb, tlb) passes of chicdiff.R:632-634, 659-692 happen
## there; R only turns the bin labels into integer codes and fits the ~75-point distance function.  cd_assemble then
## does Bmean/Tmean reconstruction, count merge and region sums.  Returns the reference's long table: a data.table keyed
## by regionID with columns baitID, otherEndID, regionID, distSign, sample, N, s_j, Bmean, Tmean, score, FullMean,
## condition (chicdiff.R:912-925), plus attr "aggregated" = list(K, FullMean, avDist) so that DESeq2Wrap.cuda need
## not re-aggregate.
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
  conds <- rep(names(chicdiff.settings[["chicagoData"]]), lengths(chicdiff.settings[["chicagoData"]]))
  ## countData = NULL: the reference reads the counts back from Reduce(merge, ...) over the replicates' CHiCAGO tables,
  ## an inner join (chicdiff.R:778): only pairs with a row in every replicate keep their counts
  common <- NULL
  if (is.null(counts)) {
    pairs <- lapply(files, function(f) { x <- readRDSorRDA(f); x <- if ("chicagoData" %in% class(x)) as.data.table(x@x) else setDT(x)
                                         unique(x[, .(baitID, otherEndID)]) })
    common <- Reduce(function(a, b) merge(a, b, by = c("baitID", "otherEndID")), pairs)
    setkey(common, baitID, otherEndID)
  }
  scores <- vector("list", length(files)); sj <- vector("list", length(files)); tm <- vector("list", length(files))
  mid <- round(0.5 * (rmap$start + rmap$end))                               # chicdiff.R:871 (R rounds half to even)
  for (i in seq_along(files)) {
    x <- readRDSorRDA(files[i]); x <- if ("chicagoData" %in% class(x)) as.data.table(x@x) else setDT(x)
    tb <- factor(x$tblb); tl <- factor(x$tlb)
    code <- function(f) { v <- as.integer(f) - 1L; v[is.na(v)] <- -1L; v }
    dfp <- .chicEstimateDistFun(x)
    cnt <- if (is.null(counts)) x[common, .(baitID, otherEndID, N), on = c("baitID", "otherEndID"), nomatch = 0L]
           else fread(counts[i])[, .(baitID, otherEndID, N)]
    .Call("cdR_build_sample_tables", ctx, i, as.integer(x$baitID), as.integer(x$otherEndID), as.numeric(x$s_j), as.numeric(x$s_i),
          code(tb), code(tl), as.numeric(x$Tmean), NULL, max(1L, nlevels(tb)), max(1L, nlevels(tl)),
          c(dfp$cubicFit, dfp$obs.min, dfp$obs.max, dfp$head.coef, dfp$tail.coef),
          as.integer(cnt$baitID), as.integer(cnt$otherEndID), as.integer(cnt$N))
    ## columns of the long table that are plain look-ups in x (chicdiff.R:634, 659): score of the pair, s_j of the bait
    scores[[i]] <- x[RU, score, on = c("baitID", "otherEndID")]
    sj[[i]] <- x[, .(s_j = s_j[1]), keyby = baitID][RU, s_j, on = "baitID"]
  }
  n <- length(unique(RU$regionID)); S <- length(files); R <- nrow(RU)
  agg <- .Call("cdR_assemble", ctx, n, S, TRUE)        # list(K, FullMean, avDist); per-row columns stay on the device
  same <- rmap$chr[RU$otherEndID - id0 + 1L] == rmap$chr[RU$baitID - id0 + 1L]
  distSign <- ifelse(same, mid[RU$otherEndID - id0 + 1L] - mid[RU$baitID - id0 + 1L], NA_real_)   # :878-881
  long <- rbindlist(lapply(seq_along(files), function(i) {
    rows <- .Call("cdR_get_sample_rows", ctx, i, R)    # list(N, FullMean, Bmean)
    fm <- rows$FullMean; fm[is.nan(fm)] <- NA_real_; bm <- rows$Bmean; bm[is.nan(bm)] <- NA_real_
    data.table(baitID = RU$baitID, otherEndID = RU$otherEndID, regionID = RU$regionID, distSign = distSign,
               sample = paste0(conds[i], ".", basename(files[i])), N = rows$N, s_j = sj[[i]], Bmean = bm, Tmean = fm - bm,
               score = scores[[i]], FullMean = fm, condition = conds[i])
  }))
  setkey(long, regionID)                                                    # chicdiff.R:925
  attr(long, "aggregated") <- agg
  long
}

## getFullRegionData() (chicdiff.R:1460-1478): list(FullRegionData, FullControlRegionData, countput) and the same files:
## <outprefix>_countput<suffix>.Rds always, the two long tables when saveAuxData.
getFullRegionData.cuda <- function(chicdiff.settings, RU, RUcontrol, suffix = "") {
  ctx <- .Call("cdR_create", as.integer(if (is.null(chicdiff.settings[["gpu"]])) 0L else chicdiff.settings[["gpu"]]))
  files <- chicdiff.settings[["chicagoData"]]
  .Call("cdR_set_design", ctx, model.matrix(~ condition, data.frame(condition = factor(rep(names(files), lengths(files))))))
  FullRegionData <- getFullRegionData1.cuda(chicdiff.settings, RU, FALSE, ctx)
  ## countput (chicdiff.R:708-735, 755-770): per condition, over the replicates' CHiCAGO rows that have a distance
  countput <- rbindlist(lapply(names(files), function(cond) {
    reps <- lapply(files[[cond]], function(f) { x <- readRDSorRDA(f); x <- if ("chicagoData" %in% class(x)) as.data.table(x@x) else setDT(x)
                                                x <- x[!is.na(distSign)]
                                                list(as.integer(x$baitID), as.integer(x$otherEndID), as.integer(x$N), as.numeric(x$Bmean), as.numeric(x$score)) })
    cp <- as.data.table(.Call("cdR_countput", ctx, reps)); cp[, condition := cond]; cp
  }))
  FullControlRegionData <- getFullRegionData1.cuda(chicdiff.settings, RUcontrol, TRUE, ctx)
  outprefix <- chicdiff.settings[["outprefix"]]
  saveRDS(countput, paste0(outprefix, "_countput", suffix, ".Rds"))
  if (isTRUE(chicdiff.settings[["saveAuxData"]])) {
    saveRDS(FullRegionData, paste0(outprefix, "_FullRegionData", suffix, ".Rds"))
    saveRDS(FullControlRegionData, paste0(outprefix, "_FullControlRegionData", suffix, ".Rds"))
  }
  list(FullRegionData, FullControlRegionData, countput)
}


## IHWcorrection(), "apply to test data" block (chicdiff.R:2038-2049) -- marshalling only.  `out` is the DESeq2Wrap
## table with avDist attached (:1965-1967), `distLookup` the table learned from the control set (:2013-2033).
## Returns `out` with group, avWeights, weight, weighted_pvalue, weighted_padj, ordered by group like the
## reference's merge() leaves it.
## ctx: a context (cdR_create) to run the n-sized work on its GPU (cdR_ihw_apply_device); NULL = the host routine.
ihwApply.cuda <- function(out, distLookup, ctx = NULL) {
  r <- if (is.null(ctx)) {
    .Call("cdR_ihw_apply", as.numeric(out$avDist), as.numeric(out$pvalue), as.numeric(distLookup$minLogDist),
          as.numeric(distLookup$maxLogDist), as.numeric(distLookup$avWeights))
  } else {
    .Call("cdR_ihw_apply_device", ctx, as.numeric(out$avDist), as.numeric(out$pvalue), as.numeric(distLookup$minLogDist),
          as.numeric(distLookup$maxLogDist), as.numeric(distLookup$avWeights))
  }
  out[, avgLogDist := log(abs(avDist))]
  out$group <- r$group
  out$avWeights <- distLookup$avWeights[r$group]
  out$weight <- ifelse(is.nan(r$weight), NA_real_, r$weight)
  out$weighted_pvalue <- ifelse(is.nan(r$weighted_pvalue) & !is.nan(out$pvalue), NA_real_, r$weighted_pvalue)
  out$weighted_padj <- ifelse(is.nan(r$weighted_padj), NA_real_, r$weighted_padj)
  setkey(out, group)
  out
}
