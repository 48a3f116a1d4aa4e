#define DFP_BWD_MODE 1
#include "dfp_dec_bwd.inl"
