// agg_nchw_tma_kernel<K,PPT,kVarWarpOnly> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarWarpOnly
#include "tma_variant_impl.inc"
