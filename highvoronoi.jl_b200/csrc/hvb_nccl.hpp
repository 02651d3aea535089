// hvb_nccl.hpp -- NCCL, bound at run time.  The library is linked without a libnccl dependency: a process that already
// carries an NCCL (PyTorch bundles its own libnccl.so.2) must not get a second one, and a single-GPU caller needs none.
// The first multi-GPU call dlopen()s "libnccl.so.2" (an already loaded copy wins, same SONAME) and resolves the handful of
// entry points below; failure is reported as HVB_ENCCL.  Types and prototypes come from <nccl.h> (build time only).
#pragma once
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>

namespace hvb {

struct Nccl {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    void* handle = nullptr;
    std::string error;

    static Nccl& get() {
        static Nccl inst;
        static std::once_flag once;
        std::call_once(once, [] { inst.load(); });
        return inst;
    }
    bool ok() const { return handle != nullptr && error.empty(); }

private:
    template <class F>
    void sym(F& f, const char* name) {
        f = reinterpret_cast<F>(dlsym(handle, name));
        if (!f && error.empty()) error = std::string("libnccl lacks ") + name;
    }
    void load() {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) {
            handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) { const char* e = dlerror(); error = std::string("cannot load libnccl.so.2: ") + (e ? e : "?"); return; }
        sym(GetUniqueId, "ncclGetUniqueId"); sym(CommInitRank, "ncclCommInitRank"); sym(CommInitAll, "ncclCommInitAll");
        sym(CommDestroy, "ncclCommDestroy"); sym(AllGather, "ncclAllGather"); sym(AllReduce, "ncclAllReduce");
        sym(GroupStart, "ncclGroupStart"); sym(GroupEnd, "ncclGroupEnd"); sym(GetErrorString, "ncclGetErrorString");
        sym(GetVersion, "ncclGetVersion");
    }
};

}  // namespace hvb
