// placeholder until the tcgen05 kernel lands
#include "common.cuh"
extern "C" int lcgan_tapconv_tc_eligible(const lcgan_tapconv*) { return 0; }
extern "C" int lcgan_tapconv_tc(const lcgan_tapconv*, const void*, const void*, void*, const float*, const float*, const void*, void*) { lcgan_set_error("tc path not built"); return 9; }
extern "C" int lcgan_tapconv_wgrad_tc(const lcgan_tapconv*, const void*, const void*, float*, float, void*) { lcgan_set_error("tc path not built"); return 9; }
