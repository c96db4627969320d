// mcx_comm.cu — placeholder until the slab exchange lands (single-GPU build path).
#include "mcx_comm.h"
struct McxComm { std::string err; };
McxComm* mcx_comm_create(const void*, uint32_t, int, int, DevParams&, std::string& err) {
  err = "multi-GPU slab exchange not available in this build";
  return nullptr;
}
void mcx_comm_destroy(McxComm* c) { delete c; }
const char* mcx_comm_error(McxComm* c) { return c ? c->err.c_str() : ""; }
int mcx_comm_iteration(McxComm*, DevParams&, const StepPlan&, cudaStream_t) { return MCX_ERR_COMM; }
int mcx_comm_allreduce_u64(McxComm*, unsigned long long*, int, cudaStream_t) { return MCX_ERR_COMM; }
