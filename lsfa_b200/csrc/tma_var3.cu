// agg_nchw_tma_kernel<K,PPT,kVarScaleCur> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarScaleCur
#include "tma_variant_impl.inc"
