// comm.cpp -- NCCL binding by dlopen (see comm.h)
#include "comm.h"
#include <dlfcn.h>
#include <string.h>

namespace cd {

static const int kNcclInt8 = 0, kNcclInt64 = 4, kNcclUint64 = 5, kNcclFloat64 = 8, kNcclSum = 0, kNcclMin = 3;

std::string Comm::load()
{
    if (lib_) return "";
    // prefer an NCCL already mapped into the process (the PyTorch host loads its bundled one)
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        lib_ = dlopen(nm, RTLD_NOW | RTLD_NOLOAD);
        if (lib_) break;
    }
    if (!lib_) for (const char* nm : names) {
        lib_ = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (lib_) break;
    }
    if (!lib_) return std::string("cannot load NCCL: ") + dlerror();
#define CD_SYM(field, name)                                                    \
    *(void**)(&field) = dlsym(lib_, name);                                     \
    if (!field) return std::string("NCCL symbol missing: ") + name;
    CD_SYM(GetUniqueId_, "ncclGetUniqueId")
    CD_SYM(CommInitRank_, "ncclCommInitRank")
    CD_SYM(CommDestroy_, "ncclCommDestroy")
    CD_SYM(AllReduce_, "ncclAllReduce")
    CD_SYM(AllGather_, "ncclAllGather")
    CD_SYM(Broadcast_, "ncclBroadcast")
    CD_SYM(GroupStart_, "ncclGroupStart")
    CD_SYM(GroupEnd_, "ncclGroupEnd")
    CD_SYM(GetErrorString_, "ncclGetErrorString")
#undef CD_SYM
    return "";
}

#define CD_NCCL(call)                                                                          \
    do {                                                                                       \
        int rc_ = (call);                                                                      \
        if (rc_ != 0) return std::string(#call) + ": " + GetErrorString_(rc_);                 \
    } while (0)

Comm::~Comm()
{
    if (comm_ && CommDestroy_) CommDestroy_(comm_);
    if (scratch_) cudaFree(scratch_);
}

std::string Comm::unique_id(char id[128])
{
    std::string e = load();
    if (!e.empty()) return e;
    Uid u;
    CD_NCCL(GetUniqueId_(&u));
    memcpy(id, u.internal, 128);
    return "";
}

std::string Comm::init(int nr, int rk, const char id[128])
{
    if (nr < 1 || rk < 0 || rk >= nr) return "bad nranks/rank";
    // a second init replaces the first communicator
    if (comm_ && CommDestroy_) { CommDestroy_(comm_); comm_ = nullptr; }
    if (scratch_) { cudaFree(scratch_); scratch_ = nullptr; }
    nranks = 1; rank = 0;
    if (nr == 1) return "";
    std::string e = load();
    if (!e.empty()) return e;
    Uid u;
    memcpy(u.internal, id, 128);
    CD_NCCL(CommInitRank_(&comm_, nr, u, rk));
    nranks = nr; rank = rk;
    if (cudaMalloc(&scratch_, sizeof(int64_t) * (size_t)nr) != cudaSuccess) return "cudaMalloc(comm scratch) failed";
    return "";
}

std::string Comm::allreduce_sum(double* buf, size_t count, cudaStream_t st)
{
    if (!active()) return "";
    CD_NCCL(AllReduce_(buf, buf, count, kNcclFloat64, kNcclSum, comm_, st));
    return "";
}

std::string Comm::allreduce_u64(unsigned long long* buf, size_t count, bool min_op, cudaStream_t st)
{
    if (!active()) return "";
    CD_NCCL(AllReduce_(buf, buf, count, kNcclUint64, min_op ? kNcclMin : kNcclSum, comm_, st));
    return "";
}

std::string Comm::allgatherv(const void* send, void* recv, const std::vector<int64_t>& counts,
                             const std::vector<int64_t>& displs, size_t elem_size, cudaStream_t st)
{
    if (!active()) return "";
    CD_NCCL(GroupStart_());
    for (int r = 0; r < nranks; r++) {
        char* dst = (char*)recv + (size_t)displs[r] * elem_size;
        const void* src = (r == rank) ? send : dst;
        int rc = Broadcast_(src, dst, (size_t)counts[r] * elem_size, kNcclInt8, r, comm_, st);
        if (rc != 0) { GroupEnd_(); return std::string("ncclBroadcast: ") + GetErrorString_(rc); }
    }
    CD_NCCL(GroupEnd_());
    return "";
}

std::string Comm::allgather_i64(int64_t mine, std::vector<int64_t>& all, cudaStream_t st)
{
    all.assign((size_t)nranks, mine);
    if (!active()) return "";
    int64_t* d = (int64_t*)scratch_;
    if (cudaMemcpyAsync(d + rank, &mine, sizeof(int64_t), cudaMemcpyHostToDevice, st) != cudaSuccess) return "H2D failed";
    CD_NCCL(AllGather_(d + rank, d, 1, kNcclInt64, comm_, st));
    if (cudaMemcpyAsync(all.data(), d, sizeof(int64_t) * (size_t)nranks, cudaMemcpyDeviceToHost, st) != cudaSuccess) return "D2H failed";
    if (cudaStreamSynchronize(st) != cudaSuccess) return "sync failed";
    return "";
}

}  // namespace cd
