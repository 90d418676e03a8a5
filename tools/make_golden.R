# make_golden.R -- run on ANY machine that has R with MASS and misc3d (the build container has none):
#     Rscript tools/make_golden.R /path/to/REVEALER_library.v5C.R tests/golden/r_golden.json
# Dumps full-precision bcv / IC / CIC values and a 3-iteration search (winners, every score, seeds, seed ICs) from the
# real reference functions for the fixed inputs of tests/golden/oracle_golden.json, so that
# tests/test_oracle_cpu.py::test_r_golden_vectors_if_present pins the oracle and
# tests/test_gpu_parity.py::test_r_golden_vectors_through_cuda_if_present pins the CUDA path to real R.  Until that file exists every parity
# statement in this repository is "vs restated oracle", not "vs R".
args <- commandArgs(trailingOnly = TRUE)
lib <- args[1]; out <- args[2]
suppressMessages({ library(MASS); library(misc3d); library(jsonlite) })
# source only the function definitions (the library file ends with a CLI driver that would run)
src <- readLines(lib)
cli <- grep("option_list|parse_args|OptionParser", src)
if (length(cli)) src <- src[1:(min(cli) - 1)]
src <- src[!grepl("^\\s*(library|suppressPackageStartupMessages|install.packages)", src)]
eval(parse(text = src))
g <- fromJSON("tests/golden/oracle_golden.json", simplifyVector = FALSE)
cases <- lapply(g$cases, function(c) {
   x <- unlist(c$x); y <- unlist(c$y); z <- unlist(c$z)
   list(x = x, y = y, z = z, bcv_x = bcv(x), bcv_y = suppressWarnings(bcv(y)), bcv_z = suppressWarnings(bcv(z)),
        IC = suppressWarnings(assoc(x, y, "IC")), CIC = suppressWarnings(cond.assoc(x, y, z, "IC")))
})
# the iteration section of REVEALER.v1 (scoring loop LIB:1846-1856, rank LIB:1858-1864, seed update LIB:2167-2171,
# cmi.seed.iter0 LIB:1833) run with the reference's own functions on the fixture's in-memory inputs
s <- g$search
target <- unlist(s$target)
m.2 <- matrix(unlist(s$m2), nrow = length(s$m2), byrow = TRUE)
seed <- unlist(s$seed)
cmi <- matrix(0, nrow(m.2), s$iters); top <- c(); cmi.seed <- c(); seed.iter <- matrix(0, s$iters, length(seed))
ic0 <- suppressWarnings(assoc(target, seed, "IC"))
for (iter in 1:s$iters) {
   for (k in 1:nrow(m.2)) cmi[k, iter] <- suppressWarnings(cond.assoc(target, m.2[k, ], seed, "IC"))
   ind <- order(cmi[, iter], decreasing = TRUE)
   top <- c(top, ind[1] - 1)                                   # 0-based, as the C ABI reports rows
   seed <- apply(rbind(seed, m.2[ind[1], ]), 2, "max")
   seed.iter[iter, ] <- seed
   cmi.seed <- c(cmi.seed, suppressWarnings(assoc(target, seed, "IC")))
}
search <- list(target = target, m2 = s$m2, seed = unlist(s$seed), iters = s$iters, top = top,
               cmi = lapply(1:s$iters, function(i) cmi[, i]), seed_iter = lapply(1:s$iters, function(i) seed.iter[i, ]),
               cmi_seed = cmi.seed, ic_seed0 = ic0)
writeLines(toJSON(list(provenance = paste("R", R.version.string, "MASS", packageVersion("MASS"), "misc3d",
                                          packageVersion("misc3d")), cases = cases, search = search),
                  digits = NA, auto_unbox = TRUE), out)
