# Dispatch added to flexgsea() (R/flexgsea.R:300-330): when the plugins are the
# package's own built-ins the null loop and the significance run on the GPU;
# anything user-defined keeps the reference's R closures.  Never call this from
# inside a %dopar% worker (forked children must not inherit a CUDA context).
.fgs_kind <- function(gene.score.fn, es.fn) {
    score <- if (identical(gene.score.fn, flexgsea_s2n)) 0L else if (identical(gene.score.fn, flexgsea_lm)) 1L else NA
    es <- if (identical(es.fn, flexgsea_weighted_ks)) 0L else if (identical(es.fn, flexgsea_maxmean)) 1L else
          if (identical(es.fn, flexgsea_mean)) 2L else NA
    list(score=score, es=es)
}

flexgsea_perm_native <- function(x, y, gene.sets, gene.names, nperm, gene.score.fn, es.fn, responses, abs=F,
                                 device=0L) {
    k <- .fgs_kind(gene.score.fn, es.fn)
    stopifnot(!is.na(k$score), !is.na(k$es))
    ctx <- fgs_create_(device)
    # CSR once, instead of match() per set per block (R/flexgsea.R:412)
    idx <- lapply(gene.sets, match, table=gene.names)
    offsets <- c(0L, cumsum(vapply(idx, length, 1L)))
    fgs_upload_(ctx, x, as.integer(offsets), as.integer(unlist(idx)))
    if (k$score == 0L) {
        y_bin <- apply(y, 2, function(ph) ph == sort(unique(ph))[1])     # R/flexgsea.R:704-710
        fgs_response_s2n_(ctx, y_bin)
    } else {
        mm <- if (is.model.matrix(y)) y else model.matrix(~ ., as.data.frame(y, stringsAsFactors=T))
        fgs_response_lm_(ctx, mm, colnames(mm)[[1]] == "(Intercept)")
    }
    # the same RNG stream as flexgsea_perm_sequential: one sample.int(n) per permutation, in order (:384)
    perm <- vapply(seq_len(nperm), function(i) sample.int(nrow(y)), integer(nrow(y)))
    es.null <- fgs_perm_null_(ctx, k$score, k$es, abs, perm, length(gene.sets), length(responses), TRUE)
    dimnames(es.null) <- list(GeneSet=names(gene.sets), Response=responses, perm=NULL)
    attr(es.null, "fgs_ctx") <- ctx      # flexgsea_calc_sig's native branch reuses the resident null
    es.null
}
