## kmer_spans_b200.R -- loads the B200 drop-in for the hot path of lmjakt/kmer_spans into R.
##
## The shared object built from r/src/kmer_spans_glue.c registers the same six .Call names and arities as the
## reference's src/kmer_spans.c (:795-808), so the reference's own wrapper file works on top of it unchanged.
## Two ways to use it:
##   1. with the reference's wrapper:  Sys.setenv(KMER_SPANS_R = "/path/to/kmer_spans/kmer_spans.R") before
##      source("kmer_spans_b200.R").  Its first statement (the dyn.load of the CPU library) is skipped, everything
##      else is evaluated as it is -- kmer.counts(), kmer.regions(), kmer.low.comp.regions(), kmer.seq(),
##      lr.regions(), window.kmer.dist(), kmers.to.file(), read.kmers() then run on the GPU library.
##   2. stand-alone: the four hot-path functions below, same names, arguments and result fields as the reference's
##      (kmer_spans.R:18-27, :41-52, :72-79, :84-86), plus kmer.mode.regions() for the README's other score modes.
## Several GPUs: Sys.setenv(KSPANS_DEVICES = "0,1,2,3") before the first call (see INTEGRATION.md).
## R is not part of the build image: this file has not been run there (r/smoke.R is the first thing to run).

.ksb200.dir <- if (!is.null(sys.frame(1)$ofile)) dirname(sys.frame(1)$ofile) else getwd()
.ksb200.so <- Sys.getenv("KMER_SPANS_SO", file.path(.ksb200.dir, "src", "kmer_spans.so"))
dyn.load(.ksb200.so)

.ksb200.ref <- Sys.getenv("KMER_SPANS_R", "")
if (nzchar(.ksb200.ref)) {
    ## evaluate the reference's wrapper without its dyn.load() call
    .exprs <- parse(.ksb200.ref)
    for (.e in .exprs) {
        if (is.call(.e) && identical(.e[[1]], as.name("dyn.load"))) next
        eval(.e, envir = globalenv())
    }
} else {
    kmer.seq <- function(k) .Call("kmer_seq_r", as.integer(k))

    kmer.counts <- function(seq, k, with.f = TRUE) {
        k <- as.integer(k)
        res <- .Call("kmer_counts", seq, k)
        out <- list(n = c(k = k, n = res[[1]]), counts = res[[2]])
        if (with.f) out$f <- out$counts / sum(out$counts)
        out
    }

    kmer.regions <- function(seq, k, kmer.scores, min.width, min.score) {
        if (length(kmer.scores) != 4^k) stop("There should be a total of 4^k scores")
        ord <- kmer.seq(k)
        if (!all(ord %in% names(kmer.scores))) stop("all kmers not defined")
        res <- .Call("kmer_regions_r", seq, as.integer(k), as.double(kmer.scores[ord]),
                     as.integer(min.width), as.double(min.score))
        setNames(res, c("n", "counts", "pos", "score"))
    }

    kmer.low.comp.regions <- function(seq, k, min.w, min.score, thr = 0.75) {
        res <- .Call("kmer_low_comp_regions", seq, as.integer(k), as.integer(min.w), as.double(min.score), thr)
        res <- setNames(res, c("n", "counts", "w.rank", "pos", "score"))
        res$pos <- t(res$pos)
        res$score <- t(res$score)
        res
    }
}

## extension: counts -> score table (mode) -> scan, resident on the GPU.
##   mode 0 weighted rank - thr, 1 log2(f / f_med), 2 +-1 around f_t (param; NA = median), 3 (r - r_t) / r_t
kmer.mode.regions <- function(seq, k, mode, min.w, min.score, thr = 0, param = NA_real_, want.scores = TRUE) {
    res <- .Call("kmer_mode_regions", seq, as.integer(k), as.integer(mode), as.double(param), as.double(thr),
                 as.integer(min.w), as.double(min.score), as.integer(want.scores))
    res <- setNames(res, c("n", "counts", "scores", "pos", "score"))
    res$pos <- t(res$pos)
    res$score <- t(res$score)
    res
}
