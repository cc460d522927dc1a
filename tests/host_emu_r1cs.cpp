// tests/host_emu_r1cs.cpp — runs the per-row bodies of the Nova folding kernels (kogarashi_b200/csrc/r1cs_kernels.cuh, compiled for the
// host with the emulated carry chains) row by row on the CPU and compares every output limb with the oracle's restatement of
// SparseMatrix::prod / compute_cross_term / fold (oracle/zkstd_oracle.hpp NovaVec).  usage: host_emu_r1cs <field 0|1> <m> <n_z> <seed>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#define KGR_HOST_EMULATE_CHAINS 1
#include "../kogarashi_b200/csrc/r1cs_kernels.cuh"
#include "../oracle/zkstd_oracle.hpp"

using namespace kgr;

template <class P, class OF> static int run(uint32_t m, uint32_t n_z, uint64_t seed) {
    typedef zko::Field<OF> Fd;
    std::mt19937_64 rng(seed);
    auto rnd = [&]() {
        uint64_t w[8];
        for (auto &x : w) x = rng();
        return Fd::from_u512(w);
    };
    struct Mat {
        std::vector<uint32_t> row_ptr, cols;
        std::vector<uint64_t> coeffs;
    } mats[3];
    for (auto &mt : mats) {
        mt.row_ptr.push_back(0);
        for (uint32_t i = 0; i < m; i++) {
            uint32_t nnz = (rng() % 7 == 0) ? 0 : (uint32_t)(1 + rng() % (i % 31 == 0 ? 40 : 4));
            for (uint32_t j = 0; j < nnz; j++) {
                mt.cols.push_back((uint32_t)(rng() % n_z));
                zko::Limbs c = (rng() & 1) ? Fd::one() : rnd();   // R1CS coefficients are mostly one: exercises the skipped product
                for (int k = 0; k < 4; k++) mt.coeffs.push_back(c[k]);
            }
            mt.row_ptr.push_back((uint32_t)mt.cols.size());
        }
        if (mt.coeffs.empty()) mt.coeffs.resize(4);
        if (mt.cols.empty()) mt.cols.resize(1);
    }
    std::vector<zko::Limbs> z1(n_z), z2(n_z);
    for (auto &v : z1) v = rnd();
    for (auto &v : z2) v = rnd();
    z2[0] = Fd::one();  // the reference's z2 = (1, x2, w2)
    zko::CsrRef ref[3];
    Csr dev[3];
    for (int k = 0; k < 3; k++) {
        ref[k] = zko::CsrRef{mats[k].row_ptr.data(), mats[k].cols.data(), mats[k].coeffs.data()};
        dev[k] = Csr{mats[k].row_ptr.data(), mats[k].cols.data(), reinterpret_cast<const uint32_t *>(mats[k].coeffs.data())};
    }
    const uint32_t *z1w = reinterpret_cast<const uint32_t *>(z1.data()), *z2w = reinterpret_cast<const uint32_t *>(z2.data());
    std::vector<zko::Limbs> out(m + 1);
    uint32_t *outw = reinterpret_cast<uint32_t *>(out.data());
    int bad = 0;
    for (int k = 0; k < 3; k++) {
        std::vector<zko::Limbs> exp = zko::NovaVec<OF>::prod(m, ref[k], z1);
        for (uint32_t i = 0; i < m + 3; i++) body_spmv<P>(i, m, dev[k], z1w, outw);   // indices >= m must be ignored
        for (uint32_t i = 0; i < m; i++) bad += !(out[i] == exp[i]);
    }
    std::vector<zko::Limbs> t_exp = zko::NovaVec<OF>::cross_term(m, ref[0], ref[1], ref[2], z1, z2);
    for (uint32_t i = 0; i < m + 3; i++) body_cross_term<P>(i, m, dev[0], dev[1], dev[2], z1w, z2w, outw);
    for (uint32_t i = 0; i < m; i++) bad += !(out[i] == t_exp[i]);
    zko::Limbs r = rnd();
    Fp<P> rr;
    memcpy(rr.v, r.data(), 32);
    std::vector<zko::Limbs> f_exp = zko::NovaVec<OF>::fold(z1, z2, r), fo(n_z + 1);
    for (uint32_t i = 0; i < n_z + 2; i++) body_vec_fold<P>(i, n_z, z1w, z2w, rr, reinterpret_cast<uint32_t *>(fo.data()));
    for (uint32_t i = 0; i < n_z; i++) bad += !(fo[i] == f_exp[i]);
    printf("%s field=%d m=%u n_z=%u mismatches=%d\n", bad ? "FAIL" : "OK", (int)std::is_same<P, FrP>::value, m, n_z, bad);
    return bad ? 1 : 0;
}

int main(int argc, char **argv) {
    if (argc < 5) { fprintf(stderr, "usage: host_emu_r1cs field m n_z seed\n"); return 2; }
    int field = atoi(argv[1]);
    uint32_t m = (uint32_t)atol(argv[2]), n_z = (uint32_t)atol(argv[3]);
    uint64_t seed = strtoull(argv[4], nullptr, 10);
    return field == 0 ? run<FqP, zko::FqParams>(m, n_z, seed) : run<FrP, zko::FrParams>(m, n_z, seed);
}
