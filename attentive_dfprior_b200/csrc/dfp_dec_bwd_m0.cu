#define DFP_BWD_MODE 0
#include "dfp_dec_bwd.inl"
