// agg_nchw_tma2_kernel<1,PPT,kVarWarpOnly> instantiations (2-CTA cluster, multicast key load)
#define LSFA_VAR kVarWarpOnly
#include "tma2_variant_impl.inc"
