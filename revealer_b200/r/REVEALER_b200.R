# REVEALER_b200.R -- R host side of the B200-native CMI search.
#
# Keeps REVEALER.v1's twelve formals (REVEALER_library.v5C.R:1504-1516, same names and defaults) and
# adds the optional trailing grid/bandwidth arguments.  All scoring, ranking and seed merging is one
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
revealer.b200.search <- function(target, m.2, seed, max.n.iter = 5, target.match = "positive",
                                 target.rand = NULL, n.grid = 25, bandwidth.mult = 0.25,
                                 bandwidth.shrink = -0.75, device = 0, ctx = NULL) {
   if (is.null(ctx)) ctx <- .Call("rvlR_create", as.integer(device))
   storage.mode(m.2) <- "double"
   res <- .Call("rvlR_search", ctx, as.double(target), m.2, as.double(seed), as.integer(max.n.iter),
                identical(target.match, "negative"), as.integer(n.grid), as.double(bandwidth.mult),
                as.double(bandwidth.shrink), target.rand)
   res$ctx <- ctx
   res
}

REVEALER.v1.b200 <- function(
   ds1, target.name, target.match = "positive", ds2, seed.names = NULL, exclude.features = NULL,
   max.n.iter = 5, pdf.output.file = NULL, count.thres.low = NULL, count.thres.high = NULL,
   n.markers = 30, locs.table.file = NULL,
   n.grid = 25, bandwidth.mult = 0.25, bandwidth.shrink = -0.75, n.perm = 10, r.seed = 34578, device = 0)
{
   set.seed(r.seed)                                        # LIB:1557
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
   null.seed <- is.null(seed.names) || identical(seed.names, "NULLSEED")
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
                               bandwidth.mult, bandwidth.shrink, device)
   res$feature.names <- rownames(m.2)
   res$seed.names.iter <- rownames(m.2)[res$top]           # seed.names.iter
   res$target <- target
   res$cmi.names <- apply(res$cmi, 2, function(v)          # display ranking, as order() at LIB:1858-1864
      rownames(m.2)[order(v, decreasing = !identical(target.match, "negative"))][1:min(n.markers, nrow(m.2))])
   if (n.perm > 0) {
      res$p.val <- sapply(1:max.n.iter, function(it)
         .Call("rvlR_pvalues", res$ctx, res$cmi[, it], as.double(res$cmi.rand[, , it])))
   }
   res
}
