// The one collective of the multi-GPU path: a sum-reduce of the raw (sum rgb, sum weight) film of every rank
// to one root, over NCCL / NVLink.  The reference is single-device; its film (src/films/color.cpp:107-130) is the buffer that is
// reduced.  Every pixel is owned by exactly one rank (lrk_tile_owner), all others contribute +0: the sum is exact whatever order
// NCCL adds in, and the reduced film is bit-identical to a single-GPU render.
//
// NCCL is opened at run time, by soname: inside a torch process that is the copy torch already loaded (so that the library's
// communicator and torch's share one NCCL), in the standalone CLI the system's libnccl.so.2.  Nothing here is linked, the
// header only supplies the types.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <string>

namespace lrk {

struct NcclApi {
    void *lib{nullptr};
    ncclResult_t (*get_unique_id)(ncclUniqueId *){nullptr};
    ncclResult_t (*comm_init_rank)(ncclComm_t *, int, ncclUniqueId, int){nullptr};
    ncclResult_t (*comm_destroy)(ncclComm_t){nullptr};
    ncclResult_t (*reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t){nullptr};
    const char *(*error_string)(ncclResult_t){nullptr};
    std::string error;
};

inline NcclApi &nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);// the copy this process already has (torch's), if any
        if (a.lib == nullptr) a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (a.lib == nullptr) {
            a.error = std::string("libnccl.so.2 could not be opened: ") + dlerror();
            return a;
        }
        auto sym = [&](const char *name) {
            void *p = dlsym(a.lib, name);
            if (p == nullptr && a.error.empty()) a.error = std::string("libnccl.so.2 lacks ") + name;
            return p;
        };
        a.get_unique_id = reinterpret_cast<decltype(a.get_unique_id)>(sym("ncclGetUniqueId"));
        a.comm_init_rank = reinterpret_cast<decltype(a.comm_init_rank)>(sym("ncclCommInitRank"));
        a.comm_destroy = reinterpret_cast<decltype(a.comm_destroy)>(sym("ncclCommDestroy"));
        a.reduce = reinterpret_cast<decltype(a.reduce)>(sym("ncclReduce"));
        a.error_string = reinterpret_cast<decltype(a.error_string)>(sym("ncclGetErrorString"));
        return a;
    }();
    return api;
}

}// namespace lrk
