// The slice of the NCCL API the library uses, resolved at run time (nccl_dl.cu).  Types mirror nccl.h (2.x ABI).
#pragma once
#include <cstddef>
#include <string>
#include <cuda_runtime.h>

namespace omr {
struct NcclUniqueId { char internal[128]; };               // ncclUniqueId, NCCL_UNIQUE_ID_BYTES = 128
struct NcclApi {
    typedef int (*get_unique_id_t)(NcclUniqueId*);
    typedef int (*comm_init_rank_t)(void** comm, int nranks, NcclUniqueId id, int rank);
    typedef int (*comm_destroy_t)(void* comm);
    typedef int (*all_reduce_t)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t s);
    typedef const char* (*error_string_t)(int);
    get_unique_id_t get_unique_id = nullptr; comm_init_rank_t comm_init_rank = nullptr; comm_destroy_t comm_destroy = nullptr;
    all_reduce_t all_reduce = nullptr; error_string_t error_string = nullptr;
    static constexpr int UINT64 = 5, SUM = 0;              // ncclUint64, ncclSum
};
const NcclApi* nccl_api(std::string* err);
}  // namespace omr
