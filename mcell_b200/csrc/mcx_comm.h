// mcx_comm.h — slab decomposition across GPUs (one process per GPU, NCCL over NVLink).
#pragma once
#include <string>
#include "mcx_internal.h"

struct McxComm;
McxComm* mcx_comm_create(const void* nccl_unique_id, uint32_t id_bytes, int rank, int world_size, DevParams& p,
                         std::string& err);
void mcx_comm_destroy(McxComm* c);
const char* mcx_comm_error(McxComm* c);
// one iteration with halo exchange + migration; returns MCX_OK or MCX_ERR_*
int mcx_comm_iteration(McxComm* c, DevParams& p, const StepPlan& plan, cudaStream_t s);
int mcx_comm_allreduce_u64(McxComm* c, unsigned long long* host_buf, int n, cudaStream_t s);
