// kernels_r1cs.cu — sparse R1CS matrix-vector products, the Nova cross term and the witness fold (row N4), Fq and Fr.
// One thread per row; a row's coefficients stream (32 B each, read once for both z vectors), the z entries are 32-byte gathers.
#include <cuda_runtime.h>

#include "launch.cuh"
#include "r1cs_kernels.cuh"

namespace kgr {

template <class P> __global__ void __launch_bounds__(128) k_spmv(uint32_t m, Csr mat, const uint32_t *z, uint32_t *out) {
    body_spmv<P>(blockIdx.x * blockDim.x + threadIdx.x, m, mat, z, out);
}
template <class P> __global__ void __launch_bounds__(128) k_cross_term(uint32_t m, Csr a, Csr b, Csr c, const uint32_t *z1, const uint32_t *z2, uint32_t *t) {
    body_cross_term<P>(blockIdx.x * blockDim.x + threadIdx.x, m, a, b, c, z1, z2, t);
}
template <class P> __global__ void __launch_bounds__(256) k_vec_fold(uint32_t n, const uint32_t *a, const uint32_t *b, Fp<P> r, uint32_t *out) {
    body_vec_fold<P>(blockIdx.x * blockDim.x + threadIdx.x, n, a, b, r, out);
}

static inline unsigned blocks_for(size_t n, unsigned tpb) { return (unsigned)((n + tpb - 1) / tpb); }

void LaunchR1cs::spmv(cudaStream_t st, int field, uint32_t m, const Csr &mat, const uint32_t *z, uint32_t *out) {
    if (!m) return;
    if (field == 0) k_spmv<FqP><<<blocks_for(m, 128), 128, 0, st>>>(m, mat, z, out);
    else k_spmv<FrP><<<blocks_for(m, 128), 128, 0, st>>>(m, mat, z, out);
}
void LaunchR1cs::cross_term(cudaStream_t st, int field, uint32_t m, const Csr &a, const Csr &b, const Csr &c, const uint32_t *z1, const uint32_t *z2, uint32_t *t) {
    if (!m) return;
    if (field == 0) k_cross_term<FqP><<<blocks_for(m, 128), 128, 0, st>>>(m, a, b, c, z1, z2, t);
    else k_cross_term<FrP><<<blocks_for(m, 128), 128, 0, st>>>(m, a, b, c, z1, z2, t);
}
void LaunchR1cs::vec_fold(cudaStream_t st, int field, uint32_t n, const uint32_t *a, const uint32_t *b, const uint32_t r8[8], uint32_t *out) {
    if (!n) return;
    if (field == 0) {
        Fp<FqP> r;
        for (int i = 0; i < 8; i++) r.v[i] = r8[i];
        k_vec_fold<FqP><<<blocks_for(n, 256), 256, 0, st>>>(n, a, b, r, out);
    } else {
        Fp<FrP> r;
        for (int i = 0; i < 8; i++) r.v[i] = r8[i];
        k_vec_fold<FrP><<<blocks_for(n, 256), 256, 0, st>>>(n, a, b, r, out);
    }
}

}  // namespace kgr
