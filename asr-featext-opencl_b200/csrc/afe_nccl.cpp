// dlopen-based NCCL binding. Only the five entry points the Normalizer needs; enum values from nccl.h (2.x ABI):
// ncclSuccess=0, ncclFloat64=8, ncclSum=0, ncclMax=2, ncclMin=3, ncclUniqueId = 128 bytes.
#include <dlfcn.h>
#include <cstring>
#include <mutex>

#include "afe_internal.h"
#include "afe_nccl.h"

namespace afe {
namespace {
struct UniqueId { char internal[128]; };
typedef int (*get_unique_id_t)(UniqueId *);
typedef int (*comm_init_rank_t)(void **, int, UniqueId, int);
typedef int (*comm_destroy_t)(void *);
typedef int (*all_reduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*group_t)(void);
typedef const char *(*err_str_t)(int);

struct Api {
    void *lib = nullptr;
    get_unique_id_t get_unique_id = nullptr;
    comm_init_rank_t comm_init_rank = nullptr;
    comm_destroy_t comm_destroy = nullptr;
    all_reduce_t all_reduce = nullptr;
    group_t group_start = nullptr, group_end = nullptr;
    err_str_t err_str = nullptr;
};

Api &api()
{
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        // if torch already loaded its bundled libnccl.so.2 the soname resolves to that copy
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            a.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) return;
        a.get_unique_id = (get_unique_id_t)dlsym(a.lib, "ncclGetUniqueId");
        a.comm_init_rank = (comm_init_rank_t)dlsym(a.lib, "ncclCommInitRank");
        a.comm_destroy = (comm_destroy_t)dlsym(a.lib, "ncclCommDestroy");
        a.all_reduce = (all_reduce_t)dlsym(a.lib, "ncclAllReduce");
        a.group_start = (group_t)dlsym(a.lib, "ncclGroupStart");
        a.group_end = (group_t)dlsym(a.lib, "ncclGroupEnd");
        a.err_str = (err_str_t)dlsym(a.lib, "ncclGetErrorString");
    });
    if (!a.lib || !a.get_unique_id || !a.comm_init_rank || !a.all_reduce || !a.group_start || !a.group_end)
        throw Error("NCCL is not available (dlopen libnccl.so.2 failed)");
    return a;
}

void check(int rc, const char *what)
{
    if (rc != 0) {
        Api &a = api();
        throw Error(std::string("NCCL error in ") + what + ": " + (a.err_str ? a.err_str(rc) : "?"));
    }
}
} // namespace

void nccl_allreduce_stats(void *comm, double *d_sums, int n_sum, double *d_mins, int n_min, double *d_maxs, int n_max,
                          cudaStream_t st)
{
    Api &a = api();
    if (!comm) throw Error("allreduce: null communicator");
    const int kF64 = 8, kSum = 0, kMax = 2, kMin = 3;
    check(a.group_start(), "ncclGroupStart");
    // Record the first failure but ALWAYS close the group: an exception between GroupStart and GroupEnd would leave it
    // open on this thread and the next collective (ours or torch's, same libnccl) would hang.
    int rc = 0;
    const char *what = nullptr;
    auto step = [&](int r, const char *w) { if (r != 0 && rc == 0) { rc = r; what = w; } };
    step(a.all_reduce(d_sums, d_sums, (size_t)n_sum, kF64, kSum, comm, st), "ncclAllReduce(sum)");
    if (rc == 0) step(a.all_reduce(d_mins, d_mins, (size_t)n_min, kF64, kMin, comm, st), "ncclAllReduce(min)");
    if (rc == 0) step(a.all_reduce(d_maxs, d_maxs, (size_t)n_max, kF64, kMax, comm, st), "ncclAllReduce(max)");
    const int rc_end = a.group_end();
    if (rc != 0) check(rc, what);
    check(rc_end, "ncclGroupEnd");
}
} // namespace afe

using namespace afe;

extern "C" {

int afe_nccl_get_unique_id(void *id128)
{
    return guarded([&] {
        UniqueId id;
        check(api().get_unique_id(&id), "ncclGetUniqueId");
        memcpy(id128, &id, sizeof id);
    });
}

int afe_nccl_comm_init(const void *id128, int n_ranks, int rank, int cuda_device, void **comm)
{
    return guarded([&] {
        DeviceGuard g(cuda_device);
        UniqueId id;
        memcpy(&id, id128, sizeof id);
        check(api().comm_init_rank(comm, n_ranks, id, rank), "ncclCommInitRank");
    });
}

int afe_nccl_comm_destroy(void *comm)
{
    return guarded([&] { if (comm) check(api().comm_destroy(comm), "ncclCommDestroy"); });
}

} // extern "C"
