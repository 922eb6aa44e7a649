// Float64 instantiations of the resident KL engine (see kl_resident.cuh).
#include "kl_resident.cuh"

namespace nmfk {

cudaError_t launch_kl_resident_f64(const SolveArgs& a, cudaStream_t s) {
    const int kt = resident_template_k(a.k);
    NMFK_DISPATCH_K(launch_resident_k, double, double, kt, a, s)
}

}  // namespace nmfk
