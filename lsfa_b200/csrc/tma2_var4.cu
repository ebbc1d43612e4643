// agg_nchw_tma2_kernel<1,PPT,kVarResCur> instantiations (2-CTA cluster, multicast key load)
#define LSFA_VAR kVarResCur
#include "tma2_variant_impl.inc"
