// Float32 instantiations of the resident KL engine (see kl_resident.cuh).
#include "kl_resident.cuh"

namespace nmfk {

cudaError_t launch_kl_resident_f32(const SolveArgs& a, cudaStream_t s) {
    const int kt = resident_template_k(a.k);
    NMFK_DISPATCH_K(launch_resident_k, float, float, kt, a, s)
}

size_t resident_smem_bytes(int n, int m, int Kt, size_t szTC) {
    const int vec = (int)(16 / szTC);
    const int KP = (Kt + vec - 1) / vec * vec;
    return ResidentSmem::make(n, m, KP, szTC, szTC, resident_threads(Kt)).total;
}

bool resident_fits(int n, int m, int k, size_t szTC) {
    const int kt = resident_template_k(k);
    if (kt < 0) return false;
    return resident_smem_bytes(n, m, kt, szTC) <= 227u * 1024u;
}

}  // namespace nmfk
