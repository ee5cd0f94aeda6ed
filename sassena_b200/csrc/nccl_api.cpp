// nccl_api.cpp — run-time binding of NCCL (see nccl_api.hpp)
#include "nccl_api.hpp"

#include <dlfcn.h>

#include <cstdlib>
#include <mutex>

namespace sass {

namespace {
NcclApi g_api;
bool g_ok = false;
std::string g_err;
std::once_flag g_once;

void load() {
    const char *names[] = {getenv("SASSENA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) {
        g_err = std::string("libnccl not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "");
        return;
    }
#define SYM(field, name)                                                  \
    g_api.field = reinterpret_cast<decltype(g_api.field)>(dlsym(h, name)); \
    if (!g_api.field) {                                                   \
        g_err = std::string("libnccl lacks ") + name;                     \
        return;                                                           \
    }
    SYM(GetVersion, "ncclGetVersion")
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(CommSplit, "ncclCommSplit")
    SYM(AllReduce, "ncclAllReduce")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_ok = true;
}
}  // namespace

const NcclApi *nccl_api(std::string *err) {
    std::call_once(g_once, load);
    if (!g_ok) {
        if (err) *err = g_err;
        return nullptr;
    }
    return &g_api;
}

}  // namespace sass
