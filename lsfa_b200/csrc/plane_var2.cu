// agg_nchw_plane_kernel<K,PPT,kVarScale> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarScale
#include "plane_variant_impl.inc"
