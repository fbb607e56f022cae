// s3d_comm.h — the communicator the multi-GPU paths run over (z-slab extraction: s3d_slab.cu, database-sharded
// matching: s3d_match.cu).  Two transports behind one interface:
//   * NcclComm   one process per GPU; ncclSend/ncclRecv groups, ncclAllReduce, ncclAllGather on the caller's stream
//                (libnccl.so.2 is loaded at run time: a single-GPU client has no NCCL dependency)
//   * LocalComm  one process, one host thread per shard; the shards read each other's buffers with
//                cudaMemcpyPeerAsync ordered by events (shards may share a device)
// Every operation only ENQUEUES work on the given stream unless it says "host".
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

namespace s3d {

struct Xfer {
    int peer;
    void* ptr;      // on this rank's device
    size_t bytes;
};

struct Comm {
    int world = 1, rank = 0, device = 0;
    unsigned long long sent = 0, recvd = 0;  // bytes through p2p / collectives, this rank
    virtual ~Comm() {}
    // Point-to-point batch.  The k-th send of rank a to rank b pairs with the k-th recv of b from a.
    virtual int p2p(const std::vector<Xfer>& sends, const std::vector<Xfer>& recvs, cudaStream_t st) = 0;
    // In-place max over the ranks of n unsigned words (non-negative floats order like their bit patterns).
    virtual int allreduce_max_u32(unsigned* d, int n, cudaStream_t st) = 0;
    // d_recv[r * bytes_each ...] = rank r's d_send (device buffers; d_send must stay untouched until finish()).
    virtual int allgather(const void* d_send, void* d_recv, size_t bytes_each, cudaStream_t st) = 0;
    // The same for small HOST buffers; blocks until `all` is filled.
    virtual int allgather_host(const void* mine, void* all, size_t bytes_each, cudaStream_t st) = 0;
    // After this, buffers that peers were given to read may be freed / overwritten in stream order.
    virtual int finish(cudaStream_t st) = 0;
};

// In-process transport (s3d_slab.cu): one LocalGroup per collective job, one LocalComm per shard / host thread.
struct LocalGroup;
LocalGroup* local_group_create(int world);
void local_group_destroy(LocalGroup* g);
void local_group_fail(LocalGroup* g);                          // wake every shard blocked in a barrier
Comm* local_comm_create(LocalGroup* g, int rank, int device);  // nullptr on failure (s3d_last_error)
// shard bounds shared by both sharded paths: contiguous, balanced index ranges, rank r owns [b[r], b[r+1])
inline int shard_lo(long long n, int world, int r) { return (int)(n / world * r + (n % world < r ? n % world : r)); }

}  // namespace s3d

struct s3d_comm {
    s3d::Comm* impl = nullptr;
};
