// comm.h -- the collectives of the sharded run (one process per GPU): NCCL over NVLink, loaded
// with dlopen so that libchicdiff_b200.so has no link-time dependency on NCCL (a single-GPU R
// session never needs it).  Nothing on the path is gathered; the exchanges that go through NCCL are
//   * all-reduce of S+1 offset sums (moments estimate) and of one total deviance per dispersion fit,
//   * when peer memory is unavailable: all-reduce of the trend fit's 8 sums per pass and of the median kernels'
//     histogram counters (otherwise those happen inside the kernels over NVLink, see offsets.cu / select.cu),
//   * the one-off exchange of shard sizes and cudaIpc mailbox handles at cd_comm_init.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

namespace cd {

struct Uid { char internal[128]; };     // layout of ncclUniqueId

class Comm {
public:
    int nranks = 1, rank = 0;
    ~Comm();
    bool active() const { return nranks > 1; }
    // returns empty string on success, else an error message
    std::string unique_id(char id[128]);
    std::string init(int nranks, int rank, const char id[128]);
    std::string allreduce_sum(double* buf, size_t count, cudaStream_t st);
    std::string allreduce_u64(unsigned long long* buf, size_t count, bool min_op, cudaStream_t st);
    // every rank contributes counts[rank] elements of elem_size bytes; recv is laid out by displs
    std::string allgatherv(const void* send, void* recv, const std::vector<int64_t>& counts,
                           const std::vector<int64_t>& displs, size_t elem_size, cudaStream_t st);
    // host-visible exchange of one int64 per rank (through a device staging buffer)
    std::string allgather_i64(int64_t mine, std::vector<int64_t>& all, cudaStream_t st);

private:
    std::string load();
    void* lib_ = nullptr;
    void* comm_ = nullptr;
    void* scratch_ = nullptr;      // device, nranks * 8 bytes
    // function pointers (signatures from nccl.h; ncclResult_t is an int enum, 0 = success)
    int (*GetUniqueId_)(void*) = nullptr;
    int (*CommInitRank_)(void**, int, Uid /* ncclUniqueId by value */, int) = nullptr;
    int (*CommDestroy_)(void*) = nullptr;
    int (*AllReduce_)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllGather_)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*Broadcast_)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart_)() = nullptr;
    int (*GroupEnd_)() = nullptr;
    const char* (*GetErrorString_)(int) = nullptr;
};

}  // namespace cd
