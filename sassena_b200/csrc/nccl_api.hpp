// nccl_api.hpp — the handful of NCCL entry points the library uses, resolved at run time (dlopen) so that
// libsassena_b200.so has no link-time dependency on a particular libnccl: inside a Python process it binds to the
// libnccl.so.2 torch has already loaded, a C++ host picks it up from the loader path (or SASSENA_NCCL_LIB).
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <string>

namespace sass {

struct NcclApi {
    ncclResult_t (*GetVersion)(int *);
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, ncclConfig_t *);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char *(*GetErrorString)(ncclResult_t);
};

// nullptr (and *err set) when no usable libnccl is found
const NcclApi *nccl_api(std::string *err);

}  // namespace sass
