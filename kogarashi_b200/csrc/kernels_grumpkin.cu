// Grumpkin instantiation of every curve kernel + launcher.
#include "launch_impl.cuh"
template struct kgr::Launch<kgr::GrumpkinC>;
