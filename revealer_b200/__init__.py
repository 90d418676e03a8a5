"""revealer_b200 -- B200-native implementation of REVEALER's iterative conditional-mutual-information
search (genepattern/REVEALER, REVEALER_library.v5C.R).  The product is librevealer_b200.so
(include/revealer_b200.h); this package is the Python host mirror of the reference's R interface
for that path.  See DESIGN.md and INTEGRATION.md."""
from .api import (RevealerContext, RevealerError, assoc, cond_assoc, revealer_search, REVEALER_v1,  # noqa: F401
                  pack_bits, shard_rows, assess_features, revealer_search_multi_gpu,
                  revealer_search_target_batch_multi_gpu, GctFile,
                  prepare_inputs, prepare_inputs_device, read_gct)

__all__ = ["RevealerContext", "RevealerError", "assoc", "cond_assoc", "revealer_search", "REVEALER_v1",
           "pack_bits", "shard_rows", "assess_features",
           "revealer_search_multi_gpu", "revealer_search_target_batch_multi_gpu", "GctFile", "prepare_inputs", "prepare_inputs_device", "read_gct"]
