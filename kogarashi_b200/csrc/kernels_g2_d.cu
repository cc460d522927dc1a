// BN254 G2 (coordinates in Fq2): precompute / fixed-base / test kernels + launchers.
#define KGR_PART 16
#define KGR_FP2_CALLS 1
#include "launch_impl.cuh"
template struct kgr::Launch<kgr::Bn254G2>;
