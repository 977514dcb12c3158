# GPU path of flexgsea(): what a maintainer adds as R/flexgsea_native.R, together with the four
# small hunks of r-package/flexgsea.R.patch.  When the plugins are the package's own built-ins
# the observed ES, the null loop and the significance run on the GPU(s) through the .Call
# entries of src/flexgsea_native.cpp; anything user-defined keeps the reference's R closures
# (a user gene.score.fn still runs in R, its scores go to the device once).  Never call this
# from inside a %dopar% worker: forked children must not inherit a CUDA context.

.fgs <- new.env(parent=emptyenv())   # the session of the flexgsea() call in progress

.fgs_call <- function(name, ...) .Call(name, PACKAGE='flexgsea', ...)

.fgs_kind <- function(gene.score.fn, es.fn, sig.fun) {
    score <- if (identical(gene.score.fn, flexgsea_s2n)) 0L else
             if (identical(gene.score.fn, flexgsea_lm)) 1L else NA_integer_
    es <- if (identical(es.fn, flexgsea_weighted_ks)) 0L else
          if (identical(es.fn, flexgsea_maxmean)) 1L else
          if (identical(es.fn, flexgsea_mean)) 2L else NA_integer_
    sig <- if (identical(sig.fun, flexgsea_calc_sig)) 1L else
           if (identical(sig.fun, flexgsea_calc_sig_simple)) 0L else NA_integer_
    list(score=score, es=es, sig=sig)
}

# y_bin exactly as flexgsea_s2n makes it (R/flexgsea.R:690-710): for a logical y the classes are
# c(TRUE, FALSE), for a factor its levels, otherwise sort(unique()) per response; class 1 -> TRUE
flexgsea_s2n_ybin <- function(y) {
    if (!is.matrix(y)) {
        y <- matrix(y, dimnames=list(names(y), 'Response'))
    }
    fixed <- if (is.logical(y)) c(TRUE, FALSE) else if (is.factor(y)) levels(y) else NULL
    y_bin <- apply(y, 2, function(phenotype) {
        classes <- if (is.null(fixed)) sort(unique(phenotype)) else fixed
        stopifnot(length(classes) == 2)
        phenotype == classes[1]
    })
    matrix(y_bin, nrow=nrow(y), dimnames=dimnames(y))
}

# flexgsea_lm's design (R/flexgsea.R:648-663)
.fgs_design <- function(y) {
    if (is.model.matrix(y)) y else stats::model.matrix(~ ., as.data.frame(y, stringsAsFactors=T))
}

# Called by flexgsea() before the observed pass (patch hunk 1).  Returns NULL when the GPU path
# does not apply -- es.fn is not a built-in, or no GPU is visible -- and flexgsea() goes on
# exactly as before.  `parallel`: TRUE = every visible GPU, a number = that many GPUs,
# FALSE / NULL = one GPU; never forked workers.
flexgsea_native_begin <- function(x, y, t.x, gene.sets, gene.names, gene.score.fn, es.fn, sig.fun,
                                  parallel, nperm, return_values) {
    k <- .fgs_kind(gene.score.fn, es.fn, sig.fun)
    if (is.na(k$es)) return(NULL)
    n.avail <- .fgs_call('_flexgsea_n_gpu')
    if (n.avail < 1L) return(NULL)
    n.gpu <- if (is.numeric(parallel)) as.integer(parallel) else if (isTRUE(parallel)) n.avail else 1L
    n.gpu <- max(1L, min(n.gpu, n.avail))
    sess <- .fgs_call('_flexgsea_session_create', n.gpu)
    # CSR once, instead of match() per set per block (R/flexgsea.R:279, :412)
    idx <- lapply(gene.sets, match, table=gene.names)
    offsets <- c(0L, cumsum(vapply(idx, length, 1L)))
    E <- if (t.x) x$E else x
    .fgs_call('_flexgsea_session_upload', sess, E, t.x, as.integer(offsets), as.integer(unlist(idx)))
    .fgs_call('_flexgsea_session_option', sess, 'nperm_total', as.numeric(nperm))
    if (identical(k$score, 0L)) {
        .fgs_call('_flexgsea_session_response_s2n', sess, flexgsea_s2n_ybin(y))
    } else if (identical(k$score, 1L)) {
        mm <- .fgs_design(y)
        .fgs_call('_flexgsea_session_response_lm', sess, mm, colnames(mm)[[1]] == "(Intercept)")
    }
    .fgs_call('_flexgsea_session_interruptible', sess, TRUE)
    st <- list(sess=sess, kind=k, n.sets=length(gene.sets), nnz=length(unlist(idx)), offsets=offsets,
               set.names=names(gene.sets),
               # es.null goes back to R only when somebody reads it there
               want.null=('es_null' %in% return_values) || is.na(k$sig))
    .fgs$session <- st
    st
}

flexgsea_native_end <- function() {
    if (!is.null(.fgs$session)) .fgs_call('_flexgsea_session_close', .fgs$session$sess)
    .fgs$session <- NULL
}

# The observed ES with every extra of the es.fn (patch hunk 2, replaces the loop over gene
# sets at R/flexgsea.R:278-292).  gene.scores is the [genes x responses x 1] array flexgsea()
# already made with gene.score.fn.  The observed ES stays on the GPUs: the null loop recomputes,
# in the reference's own operation order, every null value that lands within rounding of it.
flexgsea_native_observed <- function(st, gene.scores, responses, return_stats, return_values, extra.names) {
    sc <- matrix(gene.scores[, , 1], nrow=dim(gene.scores)[1])
    ragged <- length(intersect(extra.names, return_values)) > 0
    o <- .fgs_call('_flexgsea_session_observed_es', st$sess, st$kind$es, 1.0, sc, st$n.sets, st$nnz, ragged)
    extra_stats <- structure(lapply(return_stats, function(n) o[[n]]), names=return_stats)
    extra <- named_full_list(
        named_full_list(named_empty_list(names=st$set.names), names=responses),
        names=extra.names)
    if (ragged) {
        off <- st$offsets
        for (response.i in seq_along(responses)) {
            for (gs.i in seq_len(st$n.sets)) {
                span <- seq.int(off[gs.i] + 1L, length.out=off[gs.i + 1L] - off[gs.i])
                at <- o$running_es_at[span, response.i]
                extra$running_es_at[[response.i]][[gs.i]] <- at
                extra$running_es_pos[[response.i]][[gs.i]] <- o$running_es_pos[span, response.i]
                extra$running_es_neg[[response.i]][[gs.i]] <- o$running_es_neg[span, response.i]
                le <- seq.int(o$le_first[gs.i, response.i] + 1L, length.out=o$le_count[gs.i, response.i])
                extra$leading_edge[[response.i]][[gs.i]] <- at[le]
            }
        }
    }
    list(es=o$es, extra_stats=extra_stats, extra=extra)
}

# perm.fun of the GPU path: the signature flexgsea() calls (R/flexgsea.R:305-307),
#   perm.fun(x, y, gene.sets, gene.names, nperm, block.size, gene.score.fn, es.fn, responses,
#            abs=abs, verbose=verbose)
# block.size: the built-in score functions size their own blocks on the device; with a user-defined
# gene.score.fn it is the number of permutations scored in R per tile handed to the GPUs.
flexgsea_perm_native <- function(x, y, gene.sets, gene.names, nperm, block.size, gene.score.fn, es.fn,
                                 responses, abs=F, verbose=T) {
    st <- .fgs$session
    own <- is.null(st)
    if (own) {   # called on its own, outside flexgsea()
        t.x <- methods::is(x, 'EList')
        st <- flexgsea_native_begin(x, y, t.x, gene.sets, gene.names, gene.score.fn, es.fn,
                                    flexgsea_calc_sig, FALSE, nperm, 'es_null')
        stopifnot(!is.null(st))
        on.exit(flexgsea_native_end())
    }
    n.samples <- nrow(y)
    # the RNG stream of flexgsea_perm_sequential: one sample.int(n) per permutation, in order (:384)
    perm <- vapply(seq_len(nperm), function(i) sample.int(n.samples), integer(n.samples))
    if (!is.na(st$kind$score)) {
        es.null <- .fgs_call('_flexgsea_session_perm_null', st$sess, st$kind$score, st$kind$es, 1.0, abs,
                             perm, 0, as.integer(nperm), st$n.sets, st$want.null)
    } else {
        # user-defined gene.score.fn (flexgsea_limma, say): the R closure per permutation
        # (:383-394), in tiles of block.size permutations.  Handing a tile over only queues its
        # upload, ranking and ES, so the GPUs work on tile k while R scores tile k + 1.
        t.x <- methods::is(x, 'EList')
        pass.t.x <- 't.x' %in% methods::formalArgs(gene.score.fn)
        for (tile.start in seq(1, nperm, by=block.size)) {
            tile <- seq(tile.start, min(nperm, tile.start + block.size - 1))
            gene.scores <- array(NA_real_, c(length(gene.names), length(responses), length(tile)))
            for (k in seq_along(tile)) {
                y.perm <- y[perm[, tile[k]], , drop=F]
                for (a in setdiff(names(attributes(y)), c('dim', 'dimnames'))) {   # :385-392
                    attr(y.perm, a) <- attr(y, a)
                }
                gene.scores[, , k] <- if (pass.t.x) gene.score.fn(x, y.perm, t.x=t.x, abs=abs)
                                      else gene.score.fn(x, y.perm, abs=abs)
            }
            if (any(!is.finite(gene.scores))) {
                stop("Gene scoring function returned NA or Inf.")   # :400-402
            }
            .fgs_call('_flexgsea_session_es_from_scores_tile', st$sess, st$kind$es, 1.0, gene.scores,
                      tile.start - 1, as.numeric(nperm))
        }
        es.null <- .fgs_call('_flexgsea_session_null_finish', st$sess, st$n.sets, length(responses),
                             as.integer(nperm), st$want.null)
    }
    if (length(es.null)) {
        dimnames(es.null) <- list(GeneSet=st$set.names, Response=responses, perm=NULL)   # :363-368
    }
    es.null
}

# sig.fun on the null the GPUs hold (patch hunk 4; R/flexgsea.R:314-330).  One table per response,
# columns in flexgsea_calc_sig's / flexgsea_calc_sig_simple's order.
flexgsea_native_sig <- function(st, es, responses, abs) {
    lapply(seq_along(responses), function(response.i) {
        cols <- .fgs_call('_flexgsea_session_calc_sig', st$sess, es[, response.i], response.i - 1L,
                          st$kind$sig == 1L, st$kind$sig == 1L, abs)
        tibble::as_tibble(cols)
    })
}

# gene.sets given as a GMT file: read_gmt + filter_gene_sets + match(), natively and once
flexgsea_gmt_csr <- function(path, gene.names, gs.size.min=10, gs.size.max=300) {
    .fgs_call('_flexgsea_gmt_read', path, gene.names, as.integer(gs.size.min), as.integer(gs.size.max))
}
