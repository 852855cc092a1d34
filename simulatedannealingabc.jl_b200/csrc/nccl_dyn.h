// nccl_dyn.h -- NCCL bound at run time (dlopen), so libsabc_b200.so loads on a single GPU without
// libnccl and shares the NCCL build already mapped into the process (e.g. the one bundled with
// torch when the host side is Python) when world_size > 1.
#pragma once
#include <nccl.h>
#include <dlfcn.h>
#include "common.h"

namespace sabc {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

inline NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (api.handle) {
#define SABC_SYM(field, sym) api.field = (decltype(api.field))dlsym(api.handle, sym)
            SABC_SYM(GetUniqueId, "ncclGetUniqueId"); SABC_SYM(CommInitRank, "ncclCommInitRank");
            SABC_SYM(CommDestroy, "ncclCommDestroy"); SABC_SYM(AllReduce, "ncclAllReduce");
            SABC_SYM(AllGather, "ncclAllGather"); SABC_SYM(Broadcast, "ncclBroadcast"); SABC_SYM(Send, "ncclSend"); SABC_SYM(Recv, "ncclRecv");
            SABC_SYM(GroupStart, "ncclGroupStart"); SABC_SYM(GroupEnd, "ncclGroupEnd");
            SABC_SYM(GetErrorString, "ncclGetErrorString");
#undef SABC_SYM
            if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.AllGather || !api.Broadcast || !api.Send || !api.Recv ||
                !api.GroupStart || !api.GroupEnd) { dlclose(api.handle); api.handle = nullptr; }
        }
    }
    return api.handle ? &api : nullptr;
}

#define SABC_NCCL(call)                                                                              \
    do {                                                                                             \
        ncclResult_t r__ = (call);                                                                   \
        if (r__ != ncclSuccess)                                                                      \
            return ::sabc::set_error(SABC_ERR_NCCL, "NCCL error %d at %s:%d: %s", (int)r__, __FILE__, \
                                     __LINE__, nccl_api()->GetErrorString ? nccl_api()->GetErrorString(r__) : "?"); \
    } while (0)

inline int nccl_get_unique_id(void* out128) {
    NcclApi* a = nccl_api();
    if (!a) return set_error(SABC_ERR_NCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    SABC_NCCL(a->GetUniqueId(&id));
    memcpy(out128, &id, 128);
    return 0;
}

struct NcclComm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int init(const void* id128, int rank_, int world_) {
        NcclApi* a = nccl_api();
        if (!a) return set_error(SABC_ERR_NCCL, "libnccl.so.2 could not be loaded");
        ncclUniqueId id;
        memcpy(&id, id128, 128);
        rank = rank_; world = world_;
        SABC_NCCL(a->CommInitRank(&comm, world, id, rank));
        return 0;
    }
    void destroy() { if (comm && nccl_api()) nccl_api()->CommDestroy(comm); comm = nullptr; }
};

}  // namespace sabc
