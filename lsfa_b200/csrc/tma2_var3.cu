// agg_nchw_tma2_kernel<1,PPT,kVarScaleCur> instantiations (2-CTA cluster, multicast key load)
#define LSFA_VAR kVarScaleCur
#include "tma2_variant_impl.inc"
