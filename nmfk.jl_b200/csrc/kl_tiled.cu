// Tiled KL engine (host-driven, factors streamed): placeholder until the kernels land.
#include "nmfk_internal.h"

namespace nmfk {

bool tiled_supported(int k) { return k >= 1 && k <= kMaxK; }

cudaError_t solve_tiled(const SolveArgs&, int, cudaStream_t, int64_t*) { return cudaErrorNotSupported; }

}  // namespace nmfk
