## smoke.R -- first thing to run under a real R: the reference's own known answers (KA1-KA3 of SURVEY.md section 4)
## and one kmer.low.comp.regions() call through the GPU library.  Rscript r/smoke.R
source(file.path(dirname(sub("--file=", "", grep("--file=", commandArgs(FALSE), value = TRUE)[1])), "kmer_spans_b200.R"))

## KA3: k-mers are ordered A, C, T, G (kmer_spans.R:81-83)
stopifnot(identical(kmer.seq(2), c("AA", "AC", "AT", "AG", "CA", "CC", "CT", "CG", "TA", "TC", "TT", "TG",
                                   "GA", "GC", "GT", "GG")))
## KA1: dinucleotide counts of CGCCAATGCG (test.R:365-375)
cn <- kmer.counts("CGCCAATGCG", 2)
names(cn$counts) <- kmer.seq(2)
stopifnot(cn$counts[["CG"]] == 2, cn$counts[["GC"]] == 2, cn$counts[["CC"]] == 1, cn$counts[["CA"]] == 1,
          cn$counts[["AA"]] == 1, cn$counts[["AT"]] == 1, cn$counts[["TG"]] == 1, sum(cn$counts) == 9)
## KA2: N runs split the sequence (test.R:66-77)
set.seed(1)
s <- paste(sample(c("A", "C", "G", "T"), 50000, replace = TRUE), collapse = "")
c1 <- kmer.counts(s, 2)$counts
c2 <- kmer.counts(paste0(s, strrep("N", 36), s), 2)$counts
stopifnot(all(c2 == 2 * c1))
## one full call: a planted (AG)x500 array must come out as a span
s2 <- paste0(substr(s, 1, 20000), strrep("AG", 500), substr(s, 20001, 50000))
r <- kmer.low.comp.regions(s2, 4, 100, 20, 0.75)
stopifnot(nrow(r$pos) >= 1, any(r$pos[, 2] <= 20100 & r$pos[, 3] >= 20900), sum(r$counts) == r$n[1])
cat(sprintf("smoke ok: %d spans, checksum %.6f\n", nrow(r$pos), sum(r$score[, 1])))
