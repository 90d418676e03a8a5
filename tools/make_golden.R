# make_golden.R -- run on ANY machine that has R with MASS and misc3d (the build container has none):
#     Rscript tools/make_golden.R /path/to/REVEALER_library.v5C.R tests/golden/r_golden.json
# Dumps full-precision bcv / IC / CIC values from the real reference functions for the fixed inputs of
# tests/golden/oracle_golden.json, so that tests/test_oracle_cpu.py::test_r_golden_vectors_if_present
# pins the oracle (and through it the CUDA path) to real R.  Until that file exists every parity
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
writeLines(toJSON(list(provenance = paste("R", R.version.string, "MASS", packageVersion("MASS"), "misc3d",
                                          packageVersion("misc3d")), cases = cases),
                  digits = NA, auto_unbox = TRUE), out)
