// agg_nchw_tma2_kernel<1,PPT,kVarScale> instantiations (2-CTA cluster, multicast key load)
#define LSFA_VAR kVarScale
#include "tma2_variant_impl.inc"
