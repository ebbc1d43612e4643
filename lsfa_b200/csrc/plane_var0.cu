// agg_nchw_plane_kernel<K,PPT,kVarRuntime> instantiations (one TU per variant: parallel nvcc)
#define LSFA_VAR kVarRuntime
#include "plane_variant_impl.inc"
