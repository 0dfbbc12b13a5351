// NCCL plumbing for the one collective of the path: the corpus-level CMVN statistics all-reduce (Normalizer subsystem).
// libnccl is dlopen'ed so that libafe_cuda.so has no link-time dependency on it.
#pragma once
#include <cuda_runtime.h>

namespace afe {
// In place: sums[n_sum] (ncclSum), mins[n_min] (ncclMin), maxs[n_max] (ncclMax), all double, one NCCL group.
void nccl_allreduce_stats(void *comm, double *d_sums, int n_sum, double *d_mins, int n_min, double *d_maxs, int n_max,
                          cudaStream_t st);
}
