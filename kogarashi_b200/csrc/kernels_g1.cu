// BN254 G1 instantiation of every curve kernel + launcher.
#include "launch_impl.cuh"
template struct kgr::Launch<kgr::Bn254G1>;
