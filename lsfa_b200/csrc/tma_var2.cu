// agg_nchw_tma_kernel<K,PPT,kVarScale> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarScale
#include "tma_variant_impl.inc"
