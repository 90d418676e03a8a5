/* Minimal stand-in for R.h: ONLY for the syntax check of revealer_b200/r/r_shim.c in an image without R. */
#include <stddef.h>
char *R_alloc(size_t n, int size);
void Rf_error(const char *fmt, ...);
