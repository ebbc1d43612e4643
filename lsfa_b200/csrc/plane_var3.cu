// agg_nchw_plane_kernel<K,PPT,kVarScaleCur> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarScaleCur
#include "plane_variant_impl.inc"
