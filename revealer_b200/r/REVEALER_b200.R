# REVEALER_b200.R -- R host side of the B200-native CMI search.
#
# Defines REVEALER.v1 with the reference's twelve formals (REVEALER_library.v5C.R:1504-1516, same names and defaults)
# plus the optional trailing n.grid, bandwidth.mult, bandwidth.shrink, n.gpus.  All scoring, ranking and seed merging is one
# .Call into librevealer_b200.so through r_shim.c; nothing is computed in R and there is no CPU
# fallback (a missing GPU or a non-binary feature matrix is an R error).
#
# Unlike the reference (which returns dev.off()'s value, LIB:2488) this returns the numbers: the
# objects REVEALER.v1 builds at LIB:1837-1842, so the reference's plotting code can consume them.
#
# Not runnable in the build container (no R there); exercised through the same C ABI by the Python
# ctypes host in revealer_b200/api.py, which mirrors this file function for function.

revealer.b200.load <- function(lib.dir = ".") {
   dyn.load(file.path(lib.dir, "librevealer_b200.so"), local = FALSE)
   dyn.load(file.path(lib.dir, "revealer_b200_r.so"))
   invisible(TRUE)
}

# The iteration section of REVEALER.v1 (LIB:1837-1864, 2167-2171) as one call.
#   target: numeric[n] in REVEALER.v1's sample order; m.2: F x n 0/1 matrix; seed: numeric[n] 0/1
#   n.gpus > 1: the rows of m.2 are sharded over devices 0 .. n.gpus-1 of this R process (rvlR_search_multi)
revealer.b200.search <- function(target, m.2, seed, max.n.iter = 5, target.match = "positive",
                                 target.rand = NULL, n.grid = 25, bandwidth.mult = 0.25,
                                 bandwidth.shrink = -0.75, device = 0, ctx = NULL, n.gpus = 1) {
   if (n.gpus > 1) {
      if (is.null(m.2)) stop("n.gpus > 1 shards the R matrix m.2; it cannot be combined with a matrix already resident on one device")
      storage.mode(m.2) <- "double"
      res <- .Call("rvlR_search_multi", as.integer(device + 0:(n.gpus - 1)), as.double(target), m.2, as.double(seed),
                   as.integer(max.n.iter), identical(target.match, "negative"), as.integer(n.grid),
                   as.double(bandwidth.mult), as.double(bandwidth.shrink), target.rand)
      res$ctx <- NULL
      return(res)
   }
   if (is.null(ctx)) ctx <- .Call("rvlR_create", as.integer(device))
   if (!is.null(m.2)) storage.mode(m.2) <- "double"        # NULL: the matrix is already resident in ctx
   res <- .Call("rvlR_search", ctx, as.double(target), m.2, as.double(seed), as.integer(max.n.iter),
                identical(target.match, "negative"), as.integer(n.grid), as.double(bandwidth.mult),
                as.double(bandwidth.shrink), target.rand)
   res$ctx <- ctx
   res
}

# p.val[i] <- sum(cmi.rand > cmi[i]) / length(cmi.rand), LIB:1868-1869: on the scores still resident on the device
# after a single-GPU search; from the gathered R arrays (one upload) after a sharded one.
revealer.b200.pvalues <- function(res, iter) {
   if (!is.null(res$ctx)) return(.Call("rvlR_pvalues_iter", res$ctx, as.integer(iter)))
   tmp <- .Call("rvlR_create", 0L)
   .Call("rvlR_pvalues", tmp, res$cmi[, iter], as.double(res$cmi.rand[, , iter]))
}

# The numerical core of REVEALER_assess_features.v1 (LIB:2870-2890) for a 0/1 feature matrix: IC of every row against
# the target and its permutation p-value, n.perm shuffled targets scored in one device pass per 60000 permutations.
# Same p-value rule as LIB:2886-2890 (">=" for IC >= 0, "<=" otherwise).  set.seed(5209761) is the reference's own
# (LIB:2689), so sample() yields the stream the reference would draw for the FIRST feature; the reference redraws
# for every feature, this draws once and shares the permutations across features (documented in DESIGN.md 8).
# Plots, ROC and the squared error stay in the reference.
REVEALER_assess_features.b200 <- function(target, feature.mat, n.perm = 10000, n.grid = 25, r.seed = 5209761,
                                          device = 0, target.rand = NULL) {
   ctx <- .Call("rvlR_create", as.integer(device))
   if (is.null(target.rand) && n.perm > 0) {
      set.seed(r.seed)
      target.rand <- matrix(0, nrow = n.perm, ncol = length(target))
      for (h in 1:n.perm) target.rand[h, ] <- sample(target)
   }
   storage.mode(feature.mat) <- "double"
   res <- .Call("rvlR_assess", ctx, as.double(target), feature.mat, target.rand, as.integer(n.grid))
   IC <- res$IC
   if (is.null(res$null.IC)) return(list(IC = IC, p.val = rep(NA_real_, length(IC)), null.IC = NULL))
   n.perm <- ncol(res$null.IC)
   p.val <- ifelse(IC >= 0, rowSums(res$null.IC >= IC), rowSums(res$null.IC <= IC)) / n.perm
   names(IC) <- names(p.val) <- row.names(feature.mat)
   list(IC = IC, p.val = p.val, null.IC = res$null.IC)
}

# REVEALER.v1 with the reference's twelve formals (LIB:1504-1516: same names, order and defaults), followed by the
# optional grid / bandwidth arguments and n.gpus.  Sourcing this file after REVEALER_library.v5C.R replaces the
# reference's REVEALER.v1 by this one (R binds the last definition), exactly as the library's own second copy
# replaces its first.  pdf.output.file / locs.table.file are accepted for signature compatibility; plotting stays
# in the reference (INTEGRATION.md section 3 shows the in-place patch that keeps the PDF).
REVEALER.v1 <- function(
   ds1, target.name, target.match = "positive", ds2, seed.names = NULL, exclude.features = NULL,
   max.n.iter = 5, pdf.output.file, count.thres.low = NULL, count.thres.high = NULL,
   n.markers = 30, locs.table.file = NULL,
   n.grid = 25, bandwidth.mult = 0.25, bandwidth.shrink = -0.75, n.gpus = 1,
   n.perm = 10, r.seed = 34578, device = 0,                 # the reference's internal constants LIB:1519, 1531
   gct.ingest = FALSE,                                      # TRUE: ds2 is read by the device GCT tokeniser (opt-in)
   consolidate.identical.features = FALSE)                  # the reference's internal switch, LIB:1532
{
   set.seed(r.seed)                                        # LIB:1557
   if (is.null(seed.names)) seed.names <- "NULLSEED"       # LIB:1554
   null.seed <- identical(seed.names, "NULLSEED")          # (a vector condition in the reference, LIB:1658/1690)
   if (gct.ingest) {
      if (n.gpus > 1) stop("gct.ingest = TRUE loads ds2 onto one device; use gct.ingest = FALSE with n.gpus > 1")
      return(REVEALER.v1.b200.gct(ds1, target.name, target.match, ds2, seed.names, exclude.features,
                                  max.n.iter, count.thres.low, count.thres.high, n.markers, n.grid,
                                  bandwidth.mult, bandwidth.shrink, n.perm, device,
                                  consolidate.identical.features))
   }
   read.gct <- function(f) {                               # same on-disk format as MSIG.Gct2Frame
      d <- read.delim(f, header = TRUE, sep = "\t", skip = 2, row.names = 1, blank.lines.skip = TRUE,
                      comment.char = "", as.is = TRUE, na.strings = "")
      data.matrix(d[-1])
   }
   m.1 <- read.gct(ds1)
   m.2 <- read.gct(ds2)
   m.1 <- m.1[, !is.na(m.1[target.name, ]), drop = FALSE]  # samples with a target value
   common <- intersect(colnames(m.1), colnames(m.2))
   m.1 <- m.1[, common, drop = FALSE]
   m.2 <- m.2[, common, drop = FALSE]
   target <- m.1[target.name, ]
   ord <- order(target, decreasing = !identical(target.match, "negative"))
   target <- target[ord]
   m.2 <- m.2[, ord, drop = FALSE]
   if (target.name %in% rownames(m.2)) m.2 <- m.2[rownames(m.2) != target.name, , drop = FALSE]
   if (!is.null(count.thres.low) && !is.null(count.thres.high)) {
      s <- rowSums(m.2)
      keep <- (s >= count.thres.low & s <= count.thres.high)
      if (!null.seed) keep <- keep | (rownames(m.2) %in% seed.names)
      m.2 <- m.2[keep, , drop = FALSE]
   }
   if (null.seed) {
      seed <- rep(0, ncol(m.2))
   } else {
      seed <- apply(m.2[seed.names, , drop = FALSE], 2, max)
      m.2 <- m.2[!(rownames(m.2) %in% seed.names), , drop = FALSE]
   }
   if (!is.null(exclude.features)) m.2 <- m.2[!(rownames(m.2) %in% exclude.features), , drop = FALSE]

   target.rand <- NULL
   if (n.perm > 0) {                                       # permuted targets, R's own sample() stream
      target.rand <- matrix(target, nrow = n.perm, ncol = length(target), byrow = TRUE)
      for (i in 1:n.perm) target.rand[i, ] <- sample(target.rand[i, ])
   }
   res <- revealer.b200.search(target, m.2, seed, max.n.iter, target.match, target.rand, n.grid,
                               bandwidth.mult, bandwidth.shrink, device, n.gpus = n.gpus)
   res$feature.names <- rownames(m.2)
   res$seed.names.iter <- rownames(m.2)[res$top]           # seed.names.iter
   res$target <- target
   res$cmi.names <- sapply(1:max.n.iter, function(it) {    # display ranking, as order() at LIB:1858-1864
      if (!is.null(res$ctx)) return(rownames(m.2)[.Call("rvlR_topk", res$ctx, as.integer(it), as.integer(n.markers))])
      rownames(m.2)[order(res$cmi[, it], decreasing = !identical(target.match, "negative"))][1:min(n.markers, nrow(m.2))]
   })
   if (n.perm > 0) res$p.val <- sapply(1:max.n.iter, function(it) revealer.b200.pvalues(res, it))
   res
}
REVEALER.v1.b200 <- REVEALER.v1                            # the name this file used in round 1

# The same function with ds2 read by the device GCT ingest: the file is memory-mapped, its text goes over
# PCIe and is tokenised on the GPU straight into the bit matrix -- no read.delim, no F x n matrix of doubles
# (MSIG.Gct2Frame + data.matrix, LIB:2616-2626 / 1582).  R keeps only names, row counts and the seed rows.
REVEALER.v1.b200.gct <- function(ds1, target.name, target.match, ds2, seed.names, exclude.features, max.n.iter,
                                 count.thres.low, count.thres.high, n.markers, n.grid, bandwidth.mult,
                                 bandwidth.shrink, n.perm, device, consolidate.identical.features = FALSE)
{
   ctx <- .Call("rvlR_create", as.integer(device))
   g1 <- .Call("rvlR_gct_open", ds1); i1 <- .Call("rvlR_gct_info", g1)
   g2 <- .Call("rvlR_gct_open", ds2); i2 <- .Call("rvlR_gct_info", g2)
   t.row <- match(target.name, i1$row.names)
   if (is.na(t.row)) stop("target not found in ds1: ", target.name)
   target <- .Call("rvlR_gct_row", g1, as.integer(t.row))
   names(target) <- i1$names
   target <- target[!is.na(target)]                        # LIB:1589-1594
   common <- intersect(names(target), i2$names)            # LIB:1596-1601
   target <- target[common]
   ord <- order(target, decreasing = !identical(target.match, "negative"))   # LIB:1626-1633
   target <- target[ord]
   cols <- match(common, i2$names)[ord]
   counts <- .Call("rvlR_load_gct", ctx, g2, as.integer(cols))   # rowSums(m.2) of the resident matrix
   rows <- i2$row.names
   alive <- rows != target.name | duplicated(rows)          # LIB:1635-1638 (first match only)
   null.seed <- is.null(seed.names) || identical(seed.names, "NULLSEED")
   if (!is.null(count.thres.low) && !is.null(count.thres.high)) {          # LIB:1655-1671
      keep <- (counts >= count.thres.low & counts <= count.thres.high)
      if (!null.seed) keep <- keep | (rows %in% seed.names)
      alive <- alive & keep
   }
   if (null.seed) {
      seed <- rep(0, length(target))
   } else {                                                 # LIB:1690-1707
      locs <- match(seed.names, ifelse(alive, rows, NA))
      if (any(is.na(locs))) stop("seed(s) not found in ds2: ", paste(seed.names[is.na(locs)], collapse = ", "))
      seed <- apply(.Call("rvlR_get_rows", ctx, as.integer(locs)), 2, max)
      alive[locs] <- FALSE
   }
   if (!is.null(exclude.features)) alive <- alive & !(rows %in% exclude.features)   # LIB:1716-1720
   idx <- which(alive)
   .Call("rvlR_select_rows", ctx, as.integer(idx))
   feature.names <- rows[idx]
   if (identical(consolidate.identical.features, "identical")) {             # LIB:1727-1748, on the device
      g <- .Call("rvlR_consolidate_identical", ctx)
      feature.names <- paste(feature.names[g$rows], " (", g$counts, ")", sep = "")
   }
   target.rand <- NULL
   if (n.perm > 0) {
      target.rand <- matrix(target, nrow = n.perm, ncol = length(target), byrow = TRUE)
      for (i in 1:n.perm) target.rand[i, ] <- sample(target.rand[i, ])
   }
   res <- revealer.b200.search(as.double(target), NULL, seed, max.n.iter, target.match, target.rand, n.grid,
                               bandwidth.mult, bandwidth.shrink, device, ctx = ctx)
   res$feature.names <- feature.names
   res$seed.names.iter <- feature.names[res$top]
   res$target <- target
   res$cmi.names <- apply(res$cmi, 2, function(v)
      feature.names[order(v, decreasing = !identical(target.match, "negative"))][1:min(n.markers, length(feature.names))])
   if (n.perm > 0) res$p.val <- sapply(1:max.n.iter, function(it) revealer.b200.pvalues(res, it))
   res
}
