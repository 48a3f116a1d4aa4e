#define DFP_BWD_MODE 2
#include "dfp_dec_bwd.inl"
