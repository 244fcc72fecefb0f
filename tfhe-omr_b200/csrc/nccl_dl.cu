// nccl_dl.cu — NCCL reached through dlopen, so that the single-GPU library has no hard dependency on libnccl.  Only the five
// entry points K7 needs (omr_digest_allreduce).  A communicator must be used with the library instance
// that created it: an already-loaded libnccl (e.g. the one torch brought in) is preferred over loading a second copy.
#include <dlfcn.h>
#include <cstdlib>
#include <mutex>
#include <string>
#include "nccl_dl.hpp"

namespace omr {

const NcclApi* nccl_api(std::string* err) {
    static NcclApi api; static bool tried = false, ok = false; static std::string why; static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        void* h = nullptr;
        const char* names[] = {getenv("OMR_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;       // already in the process
        if (!h) for (const char* n : names) if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) why = std::string("libnccl not found (set OMR_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        else {
            api.get_unique_id = (NcclApi::get_unique_id_t)dlsym(h, "ncclGetUniqueId");
            api.comm_init_rank = (NcclApi::comm_init_rank_t)dlsym(h, "ncclCommInitRank");
            api.comm_destroy = (NcclApi::comm_destroy_t)dlsym(h, "ncclCommDestroy");
            api.all_reduce = (NcclApi::all_reduce_t)dlsym(h, "ncclAllReduce");
            api.error_string = (NcclApi::error_string_t)dlsym(h, "ncclGetErrorString");
            ok = api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce && api.error_string;
            if (!ok) why = "libnccl lacks a required symbol";
        }
    }
    if (!ok && err) *err = why;
    return ok ? &api : nullptr;
}

}  // namespace omr
