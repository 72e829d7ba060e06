// Cross-GPU exchange of the hot path: ONE all-reduce of the packed float64 buffer [partials | hist] (and a
// broadcast of the shift) over NCCL / NVLink.  Frames are independent in mapping and measurement
// (pycgtool/mapping.py:446-463 and the mdtraj kernels work frame by frame), so this is the only collective; the
// reference itself is single-process and has no counterpart.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2 -- inside a PyTorch process that is the copy torch already
// loaded), so the library has no link-time dependency on it and single-GPU use never touches it.

#include <dlfcn.h>

#include <mutex>

#include "cgb_common.cuh"

namespace cgb {
namespace nccl {

// The handful of NCCL declarations used here (stable across NCCL 2.x; see nccl.h).
typedef struct ncclComm* comm_t;
typedef struct {
    char internal[CGB_COMM_ID_BYTES];
} unique_id;
enum { kSuccess = 0 };
enum { kUint8 = 1, kFloat64 = 8 };
enum { kSum = 0 };

struct Api {
    void* handle = nullptr;
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

// Immutable after the first call (std::call_once): not tuning state, just the resolved entry points.
static const Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) return;
        a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(a.handle, "ncclGetUniqueId"));
        a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(a.handle, "ncclCommInitRank"));
        a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(a.handle, "ncclCommDestroy"));
        a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(a.handle, "ncclAllReduce"));
        a.Broadcast = reinterpret_cast<decltype(a.Broadcast)>(dlsym(a.handle, "ncclBroadcast"));
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.Broadcast;
    });
    return a;
}

}  // namespace nccl
}  // namespace cgb

extern "C" int cgb_comm_unique_id(uint8_t* id) {
    using namespace cgb::nccl;
    if (!id) return CGB_EINVAL;
    const Api& a = api();
    if (!a.ok) return CGB_ENCCL;
    unique_id u;
    if (a.GetUniqueId(&u) != kSuccess) return CGB_ENCCL;
    for (int i = 0; i < CGB_COMM_ID_BYTES; ++i) id[i] = (uint8_t)u.internal[i];
    return CGB_OK;
}

extern "C" int cgb_comm_init(const uint8_t* id, int32_t rank, int32_t world_size, void** comm) {
    using namespace cgb::nccl;
    if (!id || !comm || world_size < 1 || rank < 0 || rank >= world_size) return CGB_EINVAL;
    const Api& a = api();
    if (!a.ok) return CGB_ENCCL;
    unique_id u;
    for (int i = 0; i < CGB_COMM_ID_BYTES; ++i) u.internal[i] = (char)id[i];
    comm_t c = nullptr;
    if (a.CommInitRank(&c, world_size, u, rank) != kSuccess) return CGB_ENCCL;
    *comm = c;
    return CGB_OK;
}

extern "C" int cgb_comm_destroy(void* comm) {
    using namespace cgb::nccl;
    if (!comm) return CGB_OK;
    const Api& a = api();
    if (!a.ok) return CGB_ENCCL;
    return a.CommDestroy(static_cast<comm_t>(comm)) == kSuccess ? CGB_OK : CGB_ENCCL;
}

extern "C" int cgb_allreduce_partials(void* comm, double* packed, int64_t count, void* stream) {
    using namespace cgb::nccl;
    if (count < 0) return CGB_EINVAL;
    if (count == 0) return CGB_OK;
    if (!comm || !packed) return CGB_EINVAL;
    const Api& a = api();
    if (!a.ok) return CGB_ENCCL;
    const int rc = a.AllReduce(packed, packed, (size_t)count, kFloat64, kSum, static_cast<comm_t>(comm),
                               static_cast<cudaStream_t>(stream));
    return rc == kSuccess ? CGB_OK : CGB_ENCCL;
}

extern "C" int cgb_comm_broadcast(void* comm, void* buffer, int64_t n_bytes, int32_t root, void* stream) {
    using namespace cgb::nccl;
    if (n_bytes < 0 || root < 0) return CGB_EINVAL;
    if (n_bytes == 0) return CGB_OK;
    if (!comm || !buffer) return CGB_EINVAL;
    const Api& a = api();
    if (!a.ok) return CGB_ENCCL;
    const int rc = a.Broadcast(buffer, buffer, (size_t)n_bytes, kUint8, root, static_cast<comm_t>(comm),
                               static_cast<cudaStream_t>(stream));
    return rc == kSuccess ? CGB_OK : CGB_ENCCL;
}
