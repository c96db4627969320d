// mcx_comm.h — slab decomposition across GPUs (one process per GPU, NCCL over NVLink).  DESIGN.md §5.
#pragma once
#include <string>
#include "mcx_internal.h"

struct McxComm;
// nccl_unique_id: the ncclUniqueId bytes created by rank 0 (mcx_comm_unique_id) and distributed by the host
McxComm* mcx_comm_create(const void* nccl_unique_id, uint32_t id_bytes, int rank, int world_size, unsigned int halo_capacity,
                         std::string& err);
void mcx_comm_destroy(McxComm* c);
const char* mcx_comm_error(McxComm* c);
// one iteration: evaluate (owned + halo molecules) -> halo refresh with both neighbours -> sort
int mcx_comm_iteration(McxComm* c, DevParams& p, const StepPlan& plan, cudaStream_t s);
// after an upload: align fresh-id ranges across ranks and fetch the neighbours' halo molecules (no evaluation)
int mcx_comm_refresh(McxComm* c, DevParams& p, const StepPlan& plan, cudaStream_t s);
int mcx_comm_allreduce_u64(McxComm* c, unsigned long long* host_buf, int n, cudaStream_t s);
// true when the halo refresh goes through the neighbours' peer memory (NVLink stores), false on the NCCL path
bool mcx_comm_is_p2p(const McxComm* c);
// device array [world]: fresh molecule ids every rank hands out in the current iteration (filled by mcx_comm_iteration)
const uint32_t* mcx_comm_rank_fresh(const McxComm* c);
// global maximum of next_id at the last mcx_comm_refresh, before the ranks were moved to their congruence classes
unsigned long long mcx_comm_id_floor(const McxComm* c);
