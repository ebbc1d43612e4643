// agg_nchw_plane_kernel<K,PPT,kVarResCur> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarResCur
#include "plane_variant_impl.inc"
