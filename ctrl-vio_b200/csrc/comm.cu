// NCCL communicator wrappers (multi-GPU landmark sharding: one sum-allreduce of the reduced camera
// system per LM step over NVLink 5 / NVSwitch).
#include <nccl.h>

#include "marginalize.h"

namespace ctvio {

static_assert(sizeof(ncclUniqueId) == 128, "ctvio_nccl_unique_id hands out 128 bytes");

bool comm_unique_id(uint8_t* id128, std::string* err) {
  ncclUniqueId id;
  ncclResult_t r = ncclGetUniqueId(&id);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclGetUniqueId: ") + ncclGetErrorString(r);
    return false;
  }
  memcpy(id128, &id, 128);
  return true;
}

void* comm_create(int rank, int world, const uint8_t* id128, std::string* err) {
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm = nullptr;
  ncclResult_t r = ncclCommInitRank(&comm, world, id, rank);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclCommInitRank: ") + ncclGetErrorString(r);
    return nullptr;
  }
  return comm;
}

void comm_destroy(void* comm) {
  if (comm) ncclCommDestroy(static_cast<ncclComm_t>(comm));
}

bool comm_allreduce_sum(void* comm, double* buf, size_t n, cudaStream_t s, std::string* err) {
  ncclResult_t r = ncclAllReduce(buf, buf, n, ncclDouble, ncclSum, static_cast<ncclComm_t>(comm), s);
  if (r != ncclSuccess) {
    if (err) *err = std::string("ncclAllReduce: ") + ncclGetErrorString(r);
    return false;
  }
  return true;
}

}  // namespace ctvio
