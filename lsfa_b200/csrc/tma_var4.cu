// agg_nchw_tma_kernel<K,PPT,kVarResCur> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarResCur
#include "tma_variant_impl.inc"
