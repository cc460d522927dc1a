// BN254 G2 (coordinates in Fq2): scalar side, accumulate and fixup kernels + launchers.
#define KGR_PART 1
#define KGR_FP2_CALLS 1
#include "launch_impl.cuh"
template struct kgr::Launch<kgr::Bn254G2>;
