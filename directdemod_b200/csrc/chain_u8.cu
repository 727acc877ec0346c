// The unsigned 8-bit input half of the fused chain kernel instantiations (see chain.cu).
#define DDM_CHAIN_PART 1
#include "chain.cu"
