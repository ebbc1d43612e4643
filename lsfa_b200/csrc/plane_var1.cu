// agg_nchw_plane_kernel<K,PPT,kVarWarpOnly> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarWarpOnly
#include "plane_variant_impl.inc"
