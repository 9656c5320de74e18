// TEST INFRASTRUCTURE (CPU tier): meshopticalflow_b200/csrc/dist.cu (NCCL: one mesh over several GPUs) is the one source
// file of the library that is not compiled into the emulated build; these are its entry points for a single "GPU".
#include "emul_cuda_runtime.h"

#include "../../meshopticalflow_b200/csrc/mof_internal.cuh"

namespace mof {
int dist_unique_id(unsigned char*) { return MOF_E_UNSUPPORTED; }
int dist_init(mof_ctx* ctx, int, int, const unsigned char*) { return fail(ctx, MOF_E_UNSUPPORTED, "the emulated build has no communicator"); }
void dist_destroy(mof_ctx*) {}
int dist_setup_mesh(mof_ctx*) { return MOF_OK; }
int dist_p2p_setup(mof_ctx*) { return MOF_OK; }
int dist_p2p_check(mof_ctx*) { return MOF_OK; }
bool dist_active(const mof_ctx*) { return false; }
int dist_world(const mof_ctx*) { return 1; }
void dist_range(const mof_ctx*, int, int* s0, int* s1, int* r0, int* r1) { *s0 = *s1 = *r0 = *r1 = 0; }
int dist_halo_f64(mof_ctx*, int, double*) { return MOF_OK; }
int dist_halo_f32(mof_ctx*, int, float*) { return MOF_OK; }
int dist_allreduce_f64(mof_ctx*, double*, int) { return MOF_OK; }
int dist_allreduce_f32(mof_ctx*, float*, int) { return MOF_OK; }
int dist_allgather_rows(mof_ctx*, int, double*) { return MOF_OK; }
int dist_add_partition(mof_ctx*, int, const int*, int, int* id) { *id = 0; return MOF_OK; }
void dist_clear_partitions(mof_ctx*) {}
long long dist_partition_halo(const mof_ctx*, int) { return 0; }
int dist_halo_part_f32(mof_ctx*, int, float*) { return MOF_OK; }
int dist_allgather_part_f32(mof_ctx*, int, float* const*, int) { return MOF_OK; }
int dist_rank(const mof_ctx*) { return 0; }
void dist_row_starts(const mof_ctx*, int, int* out) { out[0] = out[1] = 0; }
}  // namespace mof
