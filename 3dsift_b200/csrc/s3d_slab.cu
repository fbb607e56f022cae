// s3d_slab.cu — one volume over several GPUs in z-slabs (BASELINE.json configs[2]), and the two
// transports of s3d_comm.h.  The reference runs KpSiftAlgorithm in one address space (Src/cSIFT3D.cc:165-235); the axis
// sharded here is the one its z pass runs along (Src/cSIFT3D.cc:615-617).  The stages themselves (and every kernel of the
// extraction path) live in s3d_extract.cu; this unit only decides WHICH planes each shard computes and moves planes and
// scalars between the shards, all of it enqueued on the shards' streams.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "s3d_comm.h"
#include "s3d_ctx.h"
#include "s3d_devcache.h"

namespace s3d {

// =================================================================================================
// NCCL, loaded at run time
// =================================================================================================
struct NcclApi {
    void* h = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string err;
    bool load() {
        if (h) return true;
        // a process that already holds a libnccl.so.2 (torch's bundled copy) gets that one back by soname
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names)
            if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!h) { err = dlerror() ? dlerror() : "dlopen(libnccl.so.2) failed"; return false; }
#define S3D_NCCL_SYM(name)                                                     \
        name = reinterpret_cast<decltype(name)>(dlsym(h, "nccl" #name));       \
        if (!name) { err = "libnccl: missing symbol nccl" #name; h = nullptr; return false; }
        S3D_NCCL_SYM(GetUniqueId) S3D_NCCL_SYM(CommInitRank) S3D_NCCL_SYM(CommDestroy) S3D_NCCL_SYM(Send) S3D_NCCL_SYM(Recv)
        S3D_NCCL_SYM(GroupStart) S3D_NCCL_SYM(GroupEnd) S3D_NCCL_SYM(AllReduce) S3D_NCCL_SYM(AllGather) S3D_NCCL_SYM(GetErrorString)
#undef S3D_NCCL_SYM
        return true;
    }
};
static NcclApi g_nccl;
static std::mutex g_nccl_mu;

#define S3D_NCCL(expr)                                                                                      \
    do {                                                                                                    \
        ncclResult_t r_ = (expr);                                                                           \
        if (r_ != ncclSuccess)                                                                              \
            return fail(S3D_ERR_CUDA, "%s:%d %s -> NCCL: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r_)); \
    } while (0)

struct NcclComm : Comm {
    ncclComm_t comm = nullptr;
    unsigned* d_stage = nullptr;
    ~NcclComm() override {
        if (comm) g_nccl.CommDestroy(comm);
        if (d_stage) cudaFree(d_stage);
    }
    int p2p(const std::vector<Xfer>& sends, const std::vector<Xfer>& recvs, cudaStream_t st) override {
        if (sends.empty() && recvs.empty()) return S3D_OK;
        S3D_NCCL(g_nccl.GroupStart());
        for (const Xfer& x : sends) { S3D_NCCL(g_nccl.Send(x.ptr, x.bytes, ncclInt8, x.peer, comm, st)); sent += x.bytes; }
        for (const Xfer& x : recvs) { S3D_NCCL(g_nccl.Recv(x.ptr, x.bytes, ncclInt8, x.peer, comm, st)); recvd += x.bytes; }
        S3D_NCCL(g_nccl.GroupEnd());
        return S3D_OK;
    }
    int allreduce_max_u32(unsigned* d, int n, cudaStream_t st) override {
        if (world == 1 || n <= 0) return S3D_OK;
        S3D_NCCL(g_nccl.AllReduce(d, d, (size_t)n, ncclUint32, ncclMax, comm, st));
        sent += 4ull * n; recvd += 4ull * n;
        return S3D_OK;
    }
    int allgather(const void* d_send, void* d_recv, size_t bytes_each, cudaStream_t st) override {
        if (bytes_each == 0) return S3D_OK;
        S3D_NCCL(g_nccl.AllGather(d_send, d_recv, bytes_each, ncclInt8, comm, st));
        sent += bytes_each * (world - 1); recvd += bytes_each * (world - 1);
        return S3D_OK;
    }
    int allgather_host(const void* mine, void* all, size_t bytes_each, cudaStream_t st) override {
        if (world == 1) { memcpy(all, mine, bytes_each); return S3D_OK; }
        unsigned char* d = nullptr;
        S3D_CUDA(cudaMallocAsync((void**)&d, bytes_each * (world + 1), st));
        S3D_CUDA(cudaMemcpyAsync(d, mine, bytes_each, cudaMemcpyHostToDevice, st));
        S3D_NCCL(g_nccl.AllGather(d, d + bytes_each, bytes_each, ncclInt8, comm, st));
        S3D_CUDA(cudaMemcpyAsync(all, d + bytes_each, bytes_each * world, cudaMemcpyDeviceToHost, st));
        S3D_CUDA(cudaStreamSynchronize(st));
        S3D_CUDA(cudaFreeAsync(d, st));
        return S3D_OK;
    }
    int finish(cudaStream_t) override { return S3D_OK; }  // NCCL operations are complete in stream order on both sides
};

// =================================================================================================
// LocalComm: shards of one process (one host thread each)
// =================================================================================================
__global__ void max_rows_u32_kernel(const unsigned* __restrict__ rows, int nrows, int n, unsigned* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned m = 0;
    for (int r = 0; r < nrows; ++r) m = max(m, rows[(size_t)r * n + i]);
    out[i] = m;
}

struct LocalGroup {
    int world;
    std::atomic<int> arrived{0};
    std::atomic<int> generation{0};
    std::atomic<bool> failed{false};
    struct Slot {
        int device = 0;
        cudaEvent_t ready = nullptr, done = nullptr;
        std::vector<Xfer> sends;
        const void* ptr = nullptr;          // collectives: the published buffer
        std::vector<unsigned char> host;    // allgather_host
    };
    std::vector<Slot> slot;
    explicit LocalGroup(int w) : world(w), slot(w) {}
    // sense-reversing spin barrier; returns false when a shard has failed (nobody may block for ever)
    bool barrier() {
        const int gen = generation.load(std::memory_order_acquire);
        if (arrived.fetch_add(1, std::memory_order_acq_rel) + 1 == world) {
            arrived.store(0, std::memory_order_relaxed);
            generation.fetch_add(1, std::memory_order_acq_rel);
        } else {
            int spins = 0;
            while (generation.load(std::memory_order_acquire) == gen) {
                if (failed.load(std::memory_order_relaxed)) return false;
                if (++spins > 200) std::this_thread::yield();
            }
        }
        return !failed.load(std::memory_order_relaxed);
    }
};

struct LocalComm : Comm {
    LocalGroup* g = nullptr;
    unsigned* d_stage = nullptr;
    size_t stage_words = 0;
    ~LocalComm() override {
        if (d_stage) cudaFree(d_stage);
        LocalGroup::Slot& me = g->slot[rank];
        if (me.ready) cudaEventDestroy(me.ready);
        if (me.done) cudaEventDestroy(me.done);
    }
    int init() {
        LocalGroup::Slot& me = g->slot[rank];
        me.device = device;
        S3D_CUDA(cudaEventCreateWithFlags(&me.ready, cudaEventDisableTiming));
        S3D_CUDA(cudaEventCreateWithFlags(&me.done, cudaEventDisableTiming));
        return S3D_OK;
    }
    int bar() { return g->barrier() ? S3D_OK : fail(S3D_ERR_STATE, "another shard of the group failed"); }
    int copy_from(int peer, void* dst, const void* src, size_t bytes, cudaStream_t st) {
        if (bytes == 0) return S3D_OK;
        const int pd = g->slot[peer].device;
        if (pd == device) S3D_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
        else S3D_CUDA(cudaMemcpyPeerAsync(dst, device, src, pd, bytes, st));
        if (peer != rank) recvd += bytes;
        return S3D_OK;
    }
    int p2p(const std::vector<Xfer>& sends, const std::vector<Xfer>& recvs, cudaStream_t st) override {
        LocalGroup::Slot& me = g->slot[rank];
        me.sends = sends;
        for (const Xfer& x : sends) sent += x.bytes;
        S3D_CUDA(cudaEventRecord(me.ready, st));  // what I send is complete at this point of my stream
        S3D_TRY(bar());
        std::vector<int> cursor(world, 0);
        std::vector<char> waited(world, 0);
        for (const Xfer& r : recvs) {
            const std::vector<Xfer>& ps = g->slot[r.peer].sends;
            int k = cursor[r.peer];
            while (k < (int)ps.size() && ps[k].peer != rank) ++k;
            if (k >= (int)ps.size() || ps[k].bytes != r.bytes)
                return g->failed = true, fail(S3D_ERR_STATE, "local p2p: rank %d has no matching send for rank %d", r.peer, rank);
            cursor[r.peer] = k + 1;
            if (!waited[r.peer]) { S3D_CUDA(cudaStreamWaitEvent(st, g->slot[r.peer].ready, 0)); waited[r.peer] = 1; }
            S3D_TRY(copy_from(r.peer, r.ptr, ps[k].ptr, r.bytes, st));
        }
        return bar();  // everybody has read the send lists
    }
    int ensure_stage(size_t words) {
        if (words <= stage_words) return S3D_OK;
        if (d_stage) cudaFree(d_stage);
        stage_words = 0;
        S3D_CUDA(cudaMalloc((void**)&d_stage, words * sizeof(unsigned)));
        stage_words = words;
        return S3D_OK;
    }
    int allreduce_max_u32(unsigned* d, int n, cudaStream_t st) override {
        if (world == 1 || n <= 0) return S3D_OK;
        S3D_TRY(ensure_stage((size_t)world * n));
        LocalGroup::Slot& me = g->slot[rank];
        me.ptr = d;
        S3D_CUDA(cudaEventRecord(me.ready, st));
        S3D_TRY(bar());
        for (int r = 0; r < world; ++r) {
            if (r != rank) S3D_CUDA(cudaStreamWaitEvent(st, g->slot[r].ready, 0));
            S3D_TRY(copy_from(r, d_stage + (size_t)r * n, g->slot[r].ptr, sizeof(unsigned) * n, st));
        }
        S3D_CUDA(cudaEventRecord(me.done, st));  // I hold every peer's values
        S3D_TRY(bar());
        for (int r = 0; r < world; ++r)          // nobody still reads my d when I overwrite it
            if (r != rank) S3D_CUDA(cudaStreamWaitEvent(st, g->slot[r].done, 0));
        max_rows_u32_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_stage, world, n, d);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        sent += 4ull * n * (world - 1);
        return bar();
    }
    int allgather(const void* d_send, void* d_recv, size_t bytes_each, cudaStream_t st) override {
        LocalGroup::Slot& me = g->slot[rank];
        me.ptr = d_send;
        S3D_CUDA(cudaEventRecord(me.ready, st));
        S3D_TRY(bar());
        for (int r = 0; r < world; ++r) {
            if (r != rank) S3D_CUDA(cudaStreamWaitEvent(st, g->slot[r].ready, 0));
            S3D_TRY(copy_from(r, (unsigned char*)d_recv + (size_t)r * bytes_each, g->slot[r].ptr, bytes_each, st));
        }
        sent += bytes_each * (world - 1);
        return bar();
    }
    int allgather_host(const void* mine, void* all, size_t bytes_each, cudaStream_t) override {
        LocalGroup::Slot& me = g->slot[rank];
        me.host.assign((const unsigned char*)mine, (const unsigned char*)mine + bytes_each);
        S3D_TRY(bar());
        for (int r = 0; r < world; ++r) memcpy((unsigned char*)all + (size_t)r * bytes_each, g->slot[r].host.data(), bytes_each);
        return bar();
    }
    int finish(cudaStream_t st) override {
        LocalGroup::Slot& me = g->slot[rank];
        S3D_CUDA(cudaEventRecord(me.done, st));  // my reads of the peers' buffers are complete here
        S3D_TRY(bar());
        for (int r = 0; r < world; ++r)
            if (r != rank) S3D_CUDA(cudaStreamWaitEvent(st, g->slot[r].done, 0));
        return bar();
    }
};

LocalGroup* local_group_create(int world) { return new LocalGroup(world); }
void local_group_destroy(LocalGroup* g) { delete g; }
void local_group_fail(LocalGroup* g) { g->failed = true; }
Comm* local_comm_create(LocalGroup* g, int rank, int device) {
    LocalComm* cm = new LocalComm();
    cm->world = g->world; cm->rank = rank; cm->device = device; cm->g = g;
    if (cm->init() != S3D_OK) { delete cm; return nullptr; }
    return cm;
}

// =================================================================================================
// Plane bookkeeping (host logic; identical on every rank)
// =================================================================================================
// A sharded octave gives every shard at least this many planes (> hw_max + 1): a per-level halo then comes from the direct
// neighbours only.  Grouped exchanges and window halos can be deeper than a thin slab; halo_plan handles any depth by
// intersecting the needed planes with every peer's owned range.
constexpr int kMinOwnedPlanes = 10;

static void slab_bounds_of(int nz, int world, int rank, int* own0, int* own1) {
    auto b = [&](int r) { return r <= 0 ? 0 : (r >= world ? nz : std::min(nz, (int)(((long long)nz * r / world) & ~1LL))); };
    *own0 = b(rank);
    *own1 = b(rank + 1);
}

static void owned_of(int nz, int world, int rank, int o, int* p0, int* p1) {
    int own0, own1;
    slab_bounds_of(nz, world, rank, &own0, &own1);
    const int nzo = nz >> o, sh = 1 << o;
    *p0 = std::min(nzo, (own0 + sh - 1) >> o);
    *p1 = std::min(nzo, (own1 + sh - 1) >> o);
}

static int slab_group_size() {
    const char* e = getenv("S3D_SLAB_GROUP");  // levels per halo exchange (1..6), read per call: the tests switch it
    const int v = e ? atoi(e) : 3;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
}

static int min_sharded_nz() {
    const char* e = getenv("S3D_SLAB_MIN_NZ");  // read per call: the tests switch it
    return e ? atoi(e) : 96;
}

// First octave that is replicated: the first one where some shard would own fewer than kMinOwnedPlanes planes, or
// whose plane count is below S3D_SLAB_MIN_NZ (default 96: sharding a 64^3 octave buys nothing but collectives).
static int first_replicated(int nx, int ny, int nz, int world) {
    const int mn = std::min(nx, std::min(ny, nz));
    const int noct = (int)log2f((float)mn) - 3 + 1;
    if (world <= 1) return noct;
    for (int o = 0; o < noct; ++o) {
        const int nzo = nz >> o;
        bool ok = nzo >= min_sharded_nz();
        for (int r = 0; r < world && ok; ++r) {
            int p0, p1;
            owned_of(nz, world, r, o, &p0, &p1);
            ok = p1 - p0 >= kMinOwnedPlanes;
        }
        if (!ok) return o;
    }
    return noct;
}

struct Span { int peer, kind, k0, k1; };  // kind 0 = receive from peer, 1 = send to peer

// The transfers of one halo fill as `rank` sees them, in the pairing order of Comm::p2p: receives by ascending peer and
// (low, high) side of MY halo; sends by ascending peer and (low, high) side of THE PEER's halo.
static void halo_plan(int nz, int world, int rank, int o, int depth, std::vector<Span>& out) {
    const int nzo = nz >> o;
    auto need = [&](int r, int side, int* a, int* b) {
        int p0, p1;
        owned_of(nz, world, r, o, &p0, &p1);
        if (depth >= 0 && p1 <= p0) { *a = *b = 0; return; }
        if (side == 0) { *a = depth < 0 ? 0 : std::max(0, p0 - depth); *b = p0; }
        else { *a = p1; *b = depth < 0 ? nzo : std::min(nzo, p1 + depth); }
    };
    int m0, m1;
    owned_of(nz, world, rank, o, &m0, &m1);
    for (int s = 0; s < world; ++s) {
        if (s == rank) continue;
        int s0, s1;
        owned_of(nz, world, s, o, &s0, &s1);
        for (int side = 0; side < 2; ++side) {
            int a, b;
            need(rank, side, &a, &b);
            const int k0 = std::max(a, s0), k1 = std::min(b, s1);
            if (k1 > k0) out.push_back({s, 0, k0, k1});
        }
    }
    for (int d = 0; d < world; ++d) {
        if (d == rank) continue;
        for (int side = 0; side < 2; ++side) {
            int a, b;
            need(d, side, &a, &b);
            const int k0 = std::max(a, m0), k1 = std::min(b, m1);
            if (k1 > k0) out.push_back({d, 1, k0, k1});
        }
    }
}

// Append the transfers that fill `buf`'s halo (depth planes; < 0: everything not owned) to the send / recv lists.
static void add_halo(const s3d_ctx* c, const Comm& cm, float* buf, int o, int depth, std::vector<Xfer>& sends, std::vector<Xfer>& recvs) {
    std::vector<Span> plan;
    halo_plan(c->nz, cm.world, cm.rank, o, depth, plan);
    const size_t plane = c->plane(o);
    for (const Span& s : plan) {
        Xfer x;
        x.peer = s.peer;
        x.ptr = buf + (size_t)(s.k0 - c->za[o]) * plane;
        x.bytes = (size_t)(s.k1 - s.k0) * plane * sizeof(float);
        (s.kind == 0 ? recvs : sends).push_back(x);
    }
}

static int exchange(s3d_ctx* c, Comm& cm, float* buf, int o, int depth) {
    if (cm.world == 1) return S3D_OK;
    std::vector<Xfer> sends, recvs;
    add_halo(c, cm, buf, o, depth, sends, recvs);
    return cm.p2p(sends, recvs, c->stream);
}

// =================================================================================================
// The sharded run
// =================================================================================================
// CreateCSIFT3D for one shard: allocate, copy the owned planes, first sweep of data_scale (max|v| over them).  Only
// ENQUEUES work on the shard's own stream and involves no other rank, so the upload of the next volume can run under
// the extraction of the current one.
static int slab_create_impl(int world, int rank, int device, const float* vol_own, int on_device, int nx, int ny, int nz,
                            const s3d_params* p, s3d_ctx** out) {
    s3d_ctx* c = new s3d_ctx();
    *out = c;  // the caller destroys it, also on failure
    s3d_params prm;
    if (p) prm = *p; else s3d_default_params(&prm);
    prm.device = device;
    S3D_TRY(ctx_common_init(c, nx, ny, nz, &prm));
    c->slab = true;
    slab_bounds_of(nz, world, rank, &c->own0, &c->own1);
    c->first_full = first_replicated(nx, ny, nz, world);
    cudaStream_t st = c->stream;
    for (auto& e : c->ph_ev) S3D_CUDA(cudaEventCreate(&e));
    S3D_CUDA(cudaEventRecord(c->ph_ev[6], st));
    S3D_TRY(stage_init(c));
    const size_t plane0 = c->plane(0);
    // ---- CreateCSIFT3D: copy + data_scale (Src/cSIFT3D.cc:146-163, Src/cUtil.cc:536-564) ------------------
    S3D_CUDA(s3d::dev_alloc((void**)&c->d_input, std::max<size_t>(c->lvox(0), 4) * sizeof(float), st));
    if (getenv("S3D_SLAB_POISON")) S3D_CUDA(cudaMemsetAsync(c->d_input, 0xFF, c->lvox(0) * sizeof(float), st));
    const size_t nown = plane0 * (size_t)(c->own1 - c->own0);
    float* own_ptr = c->d_input + plane0 * (size_t)(c->own0 - c->za[0]);
    if (nown) {
        if (!vol_own) return fail(S3D_ERR_ARG, "null volume");
        S3D_CUDA(cudaMemcpyAsync(own_ptr, vol_own, nown * sizeof(float), on_device ? cudaMemcpyDefault : cudaMemcpyHostToDevice, st));
    }
    S3D_TRY(stage_input_max(c, own_ptr, nown));
    S3D_CUDA(cudaEventRecord(c->ph_ev[5], st));  // end of the upload phase proper
    return S3D_OK;
}

// KpSiftAlgorithm over the shards (collective): second sweep of data_scale with the global maximum, pyramid with halo
// exchanges, thresholds, sparse stages on the owned planes.  Only ENQUEUES (kernels and collectives) on the shard's
// stream: the host can go on to enqueue the next volume while this one runs; s3d_wait completes the run.
static int slab_run_impl(Comm& cm, s3d_ctx* c) {
    if (!c->slab || c->stage != 1 || c->ran) return fail(S3D_ERR_STATE, "not a freshly created slab handle");
    cudaStream_t st = c->stream;
    const int G = c->G, D = c->D, L = c->L;
    const size_t plane0 = c->plane(0);
    const size_t nown = plane0 * (size_t)(c->own1 - c->own0);
    float* own_ptr = c->d_input + plane0 * (size_t)(c->own0 - c->za[0]);
    S3D_CUDA(cudaEventRecord(c->ph_ev[0], st));
    S3D_TRY(cm.allreduce_max_u32(c->d_slots, 1, st));
    S3D_TRY(stage_input_normalize(c, own_ptr, own_ptr, nown));
    S3D_CUDA(cudaEventRecord(c->ph_ev[1], st));
    // ---- Build_Gaussian_Scale_Space + Build_DOG_Scale_Space (:268-360) --------------------------------------
    std::vector<Xfer> wsends, wrecvs;  // window halos of all sharded octaves, exchanged in one batch at the end
    for (int o = 0; o < c->noct; ++o) {
        const int nzo = c->dims[o][2];
        float* first_src = o == 0 ? c->d_input : c->gss[o * G];
        if (c->full[o]) {
            if (o == c->first_full) {  // the last sharded data this octave is built from: gather it once
                if (o > 0) S3D_TRY(stage_seed(c, o, c->p0[o], c->p1[o]));
                S3D_TRY(exchange(c, cm, first_src, o, -1));
            } else {
                S3D_TRY(stage_seed(c, o, 0, nzo));
            }
            for (int i = 0; i < G; ++i) S3D_TRY(stage_level(c, o, i, 0, nzo));
        } else {
            if (o > 0) S3D_TRY(stage_seed(c, o, c->p0[o], c->p1[o]));
            // Levels are built in GROUPS of `gsz` with one halo exchange per group: the group's first source level is
            // exchanged 1 + sum(hw) planes deep and level j is produced on owned +- (1 + the hw of the group's later
            // levels), so the later levels find their halo locally (the same arithmetic on the same inputs: the
            // redundant planes carry the owner's bits).  gsz = 1: an exchange of hw + 1 planes before every level.
            // Every exchange is a rendezvous of neighbouring ranks; fewer, deeper exchanges trade a few redundant
            // planes for fewer collective launches.
            const int gsz = slab_group_size();
            for (int a = (o == 0 ? 0 : 1); a < G; a += gsz) {
                const int b = std::min(G - 1, a + gsz - 1);
                int depth = 1;
                for (int j = a; j <= b; ++j) depth += c->taps[j].hw;
                float* src = (o == 0 && a == 0) ? c->d_input : c->gss[o * G + a - 1];
                S3D_TRY(exchange(c, cm, src, o, depth));
                for (int j = a; j <= b; ++j) {
                    int ext = 1;
                    for (int k = j + 1; k <= b; ++k) ext += c->taps[k].hw;
                    S3D_TRY(stage_level(c, o, j, std::max(0, c->p0[o] - ext), std::min(nzo, c->p1[o] + ext)));
                }
            }
            for (int l = 1; l <= L; ++l) add_halo(c, cm, c->gss[o * G + l], o, slab_window_halo(c, l), wsends, wrecvs);
        }
    }
    S3D_CUDA(cudaEventRecord(c->ph_ev[2], st));
    // planes of levels 1..L that the neighbours' orientation / descriptor windows reach (:939-955, :1182-1198)
    if (cm.world > 1) S3D_TRY(cm.p2p(wsends, wrecvs, st));
    S3D_CUDA(cudaEventRecord(c->ph_ev[3], st));
    // thresholds: global max|DoG| per level (:384-385)
    S3D_TRY(cm.allreduce_max_u32(c->d_slots + 1, c->noct * D, st));
    // from here on the levels are only read by this shard's own kernels: once every peer's copies out of them are
    // complete they may go back to the block cache in stream order (stage_sparse releases them)
    S3D_TRY(cm.finish(st));
    // ---- Detect_KeyPoints, Assign_Orientation, Extract_Description on the owned planes ---------------------------
    S3D_TRY(stage_sparse(c));
    S3D_CUDA(cudaEventRecord(c->ph_ev[4], st));
    return S3D_OK;  // everything is enqueued; s3d_wait(c) completes the run
}

// ---- result gather ------------------------------------------------------------------------------------------
constexpr int kUnits = kCtxMaxOct * kCtxMaxG;  // unit = octave * kCtxMaxG + level

__global__ void unit_count_kernel(const int* __restrict__ rec, int stride_words, int off_octave, int off_level, int n, int* __restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int o = rec[(size_t)i * stride_words + off_octave], l = rec[(size_t)i * stride_words + off_level];
    const int u = o * kCtxMaxG + l;
    if (u >= 0 && u < kUnits) atomicAdd(&cnt[u], 1);
}

struct SegTable {
    int nseg;
    int start[32 + kCtxMaxOct * 4];  // first source row of the segment
    int dest[32 + kCtxMaxOct * 4];   // its destination row
};

// One CTA per source row: rows of a source list are grouped by unit (in the reference's order inside a unit), so a
// row's destination is its unit's destination base plus its rank inside the unit.
__global__ void permute_rows_kernel(const unsigned* __restrict__ src, unsigned* __restrict__ dst, int row_words, int nrows, SegTable t) {
    const int j = blockIdx.x;
    if (j >= nrows) return;
    int sgm = 0;
    for (int k = 1; k < t.nseg; ++k)
        if (t.start[k] <= j) sgm = k;
    const size_t d = (size_t)(t.dest[sgm] + (j - t.start[sgm])) * row_words, s = (size_t)j * row_words;
    for (int w = threadIdx.x; w < row_words; w += blockDim.x) dst[d + w] = src[s + w];
}

static int slab_gather_impl(Comm& cm, s3d_ctx* c, int root, int with_extrema) {
    cudaStream_t st = c->stream;
    S3D_CUDA(cudaSetDevice(c->device));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    S3D_CUDA(cudaEventCreate(&e0));
    S3D_CUDA(cudaEventCreate(&e1));
    S3D_CUDA(cudaEventRecord(e0, st));
    // per-unit counts of this rank's two lists
    int* d_cnt = nullptr;
    S3D_CUDA(s3d::dev_alloc((void**)&d_cnt, sizeof(int) * 2 * kUnits, st));
    S3D_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int) * 2 * kUnits, st));
    if (c->n_kps > 0) {
        unit_count_kernel<<<(c->n_kps + 255) / 256, 256, 0, st>>>((const int*)c->d_kps, S3D_KP_BYTES / 4, 4, 5, c->n_kps, d_cnt);
        g_launches++;
    }
    if (with_extrema && c->n_extre > 0) {
        unit_count_kernel<<<(c->n_extre + 255) / 256, 256, 0, st>>>(c->d_xyz5, 5, 3, 4, c->n_extre, d_cnt + kUnits);
        g_launches++;
    }
    std::vector<int> mine(2 * kUnits + 2), all((size_t)cm.world * (2 * kUnits + 2));
    S3D_CUDA(cudaMemcpyAsync(mine.data(), d_cnt, sizeof(int) * 2 * kUnits, cudaMemcpyDeviceToHost, st));
    S3D_CUDA(cudaStreamSynchronize(st));
    s3d::dev_free(d_cnt, st);
    mine[2 * kUnits] = c->n_kps;
    mine[2 * kUnits + 1] = with_extrema ? c->n_extre : 0;
    S3D_TRY(cm.allgather_host(mine.data(), all.data(), sizeof(int) * mine.size(), st));
    auto cnt = [&](int r, int list, int u) { return all[(size_t)r * mine.size() + list * kUnits + u]; };
    auto total_of = [&](int r, int list) { return all[(size_t)r * mine.size() + 2 * kUnits + list]; };
    for (int r = 0; r < cm.world; ++r)
        for (int list = 0; list < 2; ++list) {
            int s = 0;
            for (int u = 0; u < kUnits; ++u) s += cnt(r, list, u);
            if (s != total_of(r, list)) return fail(S3D_ERR_STATE, "slab gather: rank %d reports %d rows in list %d but %d by unit", r, total_of(r, list), list, s);
        }
    const int K = [&] { int s = 0; for (int r = 0; r < cm.world; ++r) s += total_of(r, 0); return s; }();
    const int E = [&] { int s = 0; for (int r = 0; r < cm.world; ++r) s += total_of(r, 1); return s; }();
    struct Arr { void** slot; int list; int row_bytes; };
    void *m_kps = nullptr, *m_desc = nullptr, *m_extre = nullptr, *m_codes = nullptr, *m_xyz5 = nullptr;
    void* mine_ptr[5] = {c->d_kps, c->d_desc, c->d_extre, c->d_codes, c->d_xyz5};
    void** merged[5] = {&m_kps, &m_desc, &m_extre, &m_codes, &m_xyz5};
    const int list_of[5] = {0, 0, 1, 1, 1};
    const int row_bytes[5] = {S3D_KP_BYTES, S3D_DESC_LEN * 4, S3D_KP_BYTES, 4, 20};
    const int narr = with_extrema ? 5 : 2;
    std::vector<Xfer> sends, recvs;
    std::vector<void*> staging;
    if (cm.rank != root) {
        for (int a = 0; a < narr; ++a) {
            const int n = list_of[a] == 0 ? c->n_kps : c->n_extre;
            if (n > 0) sends.push_back({root, mine_ptr[a], (size_t)n * row_bytes[a]});
        }
        S3D_TRY(cm.p2p(sends, recvs, st));
    } else {
        // staging: the other ranks' lists, contiguous per (rank, array)
        std::vector<std::vector<void*>> stg(cm.world, std::vector<void*>(5, nullptr));
        for (int r = 0; r < cm.world; ++r) {
            if (r == root) continue;
            for (int a = 0; a < narr; ++a) {
                const int n = total_of(r, list_of[a]);
                if (n <= 0) continue;
                void* buf = nullptr;
                S3D_CUDA(s3d::dev_alloc(&buf, (size_t)n * row_bytes[a], st));
                staging.push_back(buf);
                stg[r][a] = buf;
                recvs.push_back({r, buf, (size_t)n * row_bytes[a]});
            }
        }
        S3D_TRY(cm.p2p(sends, recvs, st));
        for (int a = 0; a < narr; ++a) {
            const int tot = list_of[a] == 0 ? K : E;
            S3D_CUDA(s3d::dev_alloc(merged[a], (size_t)std::max(tot, 1) * row_bytes[a], st));
        }
        // destination of (rank r, unit u): units ascending, ranks ascending inside a unit
        for (int list = 0; list < (with_extrema ? 2 : 1); ++list) {
            std::vector<SegTable> tab(cm.world);
            for (auto& t : tab) t.nseg = 0;
            std::vector<int> srcpos(cm.world, 0);
            int dest = 0;
            for (int u = 0; u < kUnits; ++u)
                for (int r = 0; r < cm.world; ++r) {
                    const int n = cnt(r, list, u);
                    if (n <= 0) continue;
                    SegTable& t = tab[r];
                    if (t.nseg >= (int)(sizeof(t.start) / sizeof(int))) return fail(S3D_ERR_CAPACITY, "slab gather: too many (octave, level) units");
                    t.start[t.nseg] = srcpos[r]; t.dest[t.nseg] = dest; t.nseg++;
                    srcpos[r] += n; dest += n;
                }
            for (int a = 0; a < narr; ++a) {
                if (list_of[a] != list) continue;
                for (int r = 0; r < cm.world; ++r) {
                    const int n = total_of(r, list);
                    if (n <= 0) continue;
                    const void* src = r == root ? mine_ptr[a] : stg[r][a];
                    permute_rows_kernel<<<n, row_bytes[a] >= 1024 ? 256 : 64, 0, st>>>((const unsigned*)src, (unsigned*)*merged[a], row_bytes[a] / 4, n, tab[r]);
                    g_launches++;
                }
            }
        }
        S3D_CUDA(cudaGetLastError());
        for (void* b : staging) s3d::dev_free(b, st);
        for (int a = 0; a < narr; ++a) if (mine_ptr[a]) s3d::dev_free(mine_ptr[a], st);
        c->d_kps = (s3d_keypoint*)m_kps; c->d_desc = (float*)m_desc;
        c->n_kps = K;
        if (with_extrema) {
            c->d_extre = (s3d_keypoint*)m_extre; c->d_codes = (int*)m_codes; c->d_xyz5 = (int*)m_xyz5;
            c->n_extre = E;
        }
    }
    S3D_CUDA(cudaEventRecord(e1, st));
    S3D_TRY(cm.finish(st));
    S3D_CUDA(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    c->ph_ms[4] = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return S3D_OK;
}

}  // namespace s3d

using namespace s3d;

extern "C" {

int s3d_slab_bounds(int nz, int world, int rank, int* own0, int* own1) {
    clear_error();
    if (nz < 1 || world < 1 || rank < 0 || rank >= world || !own0 || !own1) return fail(S3D_ERR_ARG, "bad argument");
    slab_bounds_of(nz, world, rank, own0, own1);
    return S3D_OK;
}

int s3d_slab_first_replicated(int nx, int ny, int nz, int world, const s3d_params*, int* octave) {
    clear_error();
    if (nx < 8 || ny < 8 || nz < 8 || world < 1 || !octave) return fail(S3D_ERR_ARG, "bad argument");
    *octave = first_replicated(nx, ny, nz, world);
    return S3D_OK;
}

int s3d_slab_plan(int nz, int world, int rank, int octave, int depth, int* quads, int cap, int* n) {
    clear_error();
    if (nz < 1 || world < 1 || rank < 0 || rank >= world || octave < 0 || octave >= kCtxMaxOct || !n) return fail(S3D_ERR_ARG, "bad argument");
    std::vector<Span> plan;
    halo_plan(nz, world, rank, octave, depth, plan);
    *n = (int)plan.size();
    for (int i = 0; i < (int)plan.size() && i < cap && quads; ++i) {
        quads[4 * i] = plan[i].peer; quads[4 * i + 1] = plan[i].kind; quads[4 * i + 2] = plan[i].k0; quads[4 * i + 3] = plan[i].k1;
    }
    return S3D_OK;
}

int s3d_comm_unique_id(void* id128) {
    clear_error();
    if (!id128) return fail(S3D_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (!g_nccl.load()) return fail(S3D_ERR_CUDA, "NCCL is not available: %s", g_nccl.err.c_str());
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    S3D_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return S3D_OK;
}

int s3d_comm_create(const void* id128, int world, int rank, int device, s3d_comm_t* out) {
    clear_error();
    if (!id128 || !out || world < 1 || rank < 0 || rank >= world) return fail(S3D_ERR_ARG, "bad argument");
    {
        std::lock_guard<std::mutex> lk(g_nccl_mu);
        if (!g_nccl.load()) return fail(S3D_ERR_CUDA, "NCCL is not available: %s", g_nccl.err.c_str());
    }
    int dev;
    S3D_TRY(use_device(device, &dev));
    NcclComm* nc = new NcclComm();
    nc->world = world; nc->rank = rank; nc->device = dev;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&nc->comm, world, id, rank);
    if (r != ncclSuccess) {
        nc->comm = nullptr;
        delete nc;
        return fail(S3D_ERR_CUDA, "ncclCommInitRank(world %d, rank %d): %s", world, rank, g_nccl.GetErrorString(r));
    }
    s3d_comm* c = new s3d_comm();
    c->impl = nc;
    *out = c;
    return S3D_OK;
}

void s3d_comm_destroy(s3d_comm_t comm) {
    if (!comm) return;
    if (comm->impl) { cudaSetDevice(comm->impl->device); delete comm->impl; }
    delete comm;
}

int s3d_comm_info(s3d_comm_t comm, int* world, int* rank, int* device) {
    if (!comm || !comm->impl) return fail(S3D_ERR_ARG, "null communicator");
    if (world) *world = comm->impl->world;
    if (rank) *rank = comm->impl->rank;
    if (device) *device = comm->impl->device;
    return S3D_OK;
}

int s3d_comm_traffic(s3d_comm_t comm, unsigned long long* sent, unsigned long long* received) {
    if (!comm || !comm->impl) return fail(S3D_ERR_ARG, "null communicator");
    if (sent) *sent = comm->impl->sent;
    if (received) *received = comm->impl->recvd;
    return S3D_OK;
}

int s3d_slab_create(s3d_comm_t comm, const float* vol_own, int on_device, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out) {
    clear_error();
    if (!comm || !comm->impl || !out) return fail(S3D_ERR_ARG, "null argument");
    S3D_CUDA(cudaSetDevice(comm->impl->device));
    s3d_ctx* c = nullptr;
    const int r = slab_create_impl(comm->impl->world, comm->impl->rank, comm->impl->device, vol_own, on_device, nx, ny, nz, p, &c);
    if (r != S3D_OK) { if (c) s3d_destroy(c); *out = nullptr; return r; }
    *out = c;
    return S3D_OK;
}

int s3d_slab_execute_async(s3d_comm_t comm, s3d_handle h) {
    clear_error();
    if (!comm || !comm->impl || !h) return fail(S3D_ERR_ARG, "null argument");
    S3D_CUDA(cudaSetDevice(comm->impl->device));
    return slab_run_impl(*comm->impl, h);
}

int s3d_slab_execute(s3d_comm_t comm, s3d_handle h) {
    const int r = s3d_slab_execute_async(comm, h);
    return r != S3D_OK ? r : s3d_wait(h);
}

int s3d_slab_run(s3d_comm_t comm, const float* vol_own, int on_device, int nx, int ny, int nz, const s3d_params* p, s3d_handle* out) {
    int r = s3d_slab_create(comm, vol_own, on_device, nx, ny, nz, p, out);
    if (r != S3D_OK) return r;
    r = s3d_slab_execute(comm, *out);
    if (r != S3D_OK) { s3d_destroy(*out); *out = nullptr; }
    return r;
}

int s3d_slab_gather(s3d_comm_t comm, s3d_handle h, int root, int with_extrema) {
    clear_error();
    if (!comm || !comm->impl || !h) return fail(S3D_ERR_ARG, "null argument");
    if (!h->ran) return fail(S3D_ERR_STATE, "s3d_slab_gather before s3d_slab_run");
    if (root < 0 || root >= comm->impl->world) return fail(S3D_ERR_ARG, "root %d out of range", root);
    return slab_gather_impl(*comm->impl, h, root, with_extrema);
}

int s3d_slab_phases(s3d_handle h, double* ms8) {
    if (!h || !ms8) return fail(S3D_ERR_ARG, "null argument");
    if (h->slab && h->ran && h->ph_ev[0]) {
        for (int k = 0; k < 4; ++k) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, h->ph_ev[k], h->ph_ev[k + 1]) == cudaSuccess) h->ph_ms[k] = ms;
        }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->ph_ev[6], h->ph_ev[5]) == cudaSuccess) h->ph_ms[5] = ms;
        cudaGetLastError();
    }
    for (int k = 0; k < 8; ++k) ms8[k] = h->ph_ms[k];
    return S3D_OK;
}

int s3d_extract_multi(const float* vol, int nx, int ny, int nz, const s3d_params* p, const int* devices, int nshards,
                      int with_extrema, s3d_handle* handles) {
    clear_error();
    if (!vol || !handles || nshards < 1 || nshards > 64) return fail(S3D_ERR_ARG, "bad argument");
    int cur = 0;
    S3D_TRY(use_device(-1, &cur));
    LocalGroup group(nshards);
    std::vector<int> rc(nshards, S3D_OK);
    std::vector<std::string> msg(nshards);
    std::vector<s3d_ctx*> ctx(nshards, nullptr);
    auto work = [&](int g) {
        const int dev = devices ? devices[g] : cur;
        int rdev = 0;
        int r = use_device(dev, &rdev);
        if (r == S3D_OK && devices)  // direct peer copies between the shards' devices (else they stage through the host)
            for (int k = 0; k < nshards; ++k) {
                int can = 0;
                if (devices[k] != rdev && cudaDeviceCanAccessPeer(&can, rdev, devices[k]) == cudaSuccess && can)
                    if (cudaDeviceEnablePeerAccess(devices[k], 0) != cudaSuccess) cudaGetLastError();  // already enabled
            }
        s3d_params prm;
        if (p) prm = *p; else s3d_default_params(&prm);
        if (nshards > 1) prm.stream = nullptr;  // every shard needs a stream of its own
        LocalComm cm;
        cm.world = nshards; cm.rank = g; cm.device = rdev; cm.g = &group;
        if (r == S3D_OK) r = cm.init();
        if (r == S3D_OK) {
            int own0, own1;
            slab_bounds_of(nz, nshards, g, &own0, &own1);
            r = slab_create_impl(nshards, g, rdev, vol + (size_t)own0 * nx * ny, 0, nx, ny, nz, &prm, &ctx[g]);
        }
        if (r == S3D_OK) r = slab_run_impl(cm, ctx[g]);
        if (r == S3D_OK) r = s3d_wait(ctx[g]);
        if (r == S3D_OK) r = slab_gather_impl(cm, ctx[g], 0, with_extrema);
        if (r != S3D_OK) {
            group.failed = true;
            msg[g] = s3d_last_error();
        }
        if (ctx[g]) cudaStreamSynchronize(ctx[g]->stream);
        rc[g] = r;
    };
    if (nshards == 1) {
        work(0);
    } else {
        std::vector<std::thread> th;
        for (int g = 0; g < nshards; ++g) th.emplace_back(work, g);
        for (auto& t : th) t.join();
    }
    cudaSetDevice(cur);
    for (int g = 0; g < nshards; ++g)
        if (rc[g] != S3D_OK) {
            // report the first shard that failed for a reason of its own
            int first = g;
            for (int k = 0; k < nshards; ++k)
                if (rc[k] != S3D_OK && msg[k].find("another shard") == std::string::npos) { first = k; break; }
            for (auto*& c : ctx) if (c) { s3d_destroy(c); c = nullptr; }
            return fail(rc[first], "shard %d: %s", first, msg[first].c_str());
        }
    for (int g = 0; g < nshards; ++g) handles[g] = ctx[g];
    return S3D_OK;
}

}  // extern "C"
