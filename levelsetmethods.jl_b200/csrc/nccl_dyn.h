// nccl_dyn.h — NCCL bound at run time with dlopen, so that liblsm_b200.so has no link-time NCCL
// dependency (single-GPU use and CPU-only symbol checks never touch it) and, inside a Python
// process that already imported torch, resolves to the very libnccl.so.2 torch loaded.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and enums only

namespace lsm {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;

    bool load(const char** why) {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { *why = dlerror(); return false; }
#define LSM_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(handle, sym)); \
        if (!field) { *why = "missing NCCL symbol " sym; dlclose(handle); handle = nullptr; return false; }
        LSM_SYM(GetUniqueId, "ncclGetUniqueId")
        LSM_SYM(CommInitRank, "ncclCommInitRank")
        LSM_SYM(CommInitAll, "ncclCommInitAll")
        LSM_SYM(CommDestroy, "ncclCommDestroy")
        LSM_SYM(Send, "ncclSend")
        LSM_SYM(Recv, "ncclRecv")
        LSM_SYM(AllReduce, "ncclAllReduce")
        LSM_SYM(GroupStart, "ncclGroupStart")
        LSM_SYM(GroupEnd, "ncclGroupEnd")
        LSM_SYM(GetErrorString, "ncclGetErrorString")
#undef LSM_SYM
        return true;
    }
};

inline NcclApi& nccl() { static NcclApi api; return api; }

}  // namespace lsm
