// s3d_devcache.h — size-class block cache for the extraction handles' device buffers.
//
// Every volume of a job has the same shape, so a handle asks for the same ~90 buffer sizes as the one
// before it.  cudaMallocAsync/cudaFreeAsync serve that well while ONE handle is alive at a time; with
// two handles alive on different streams (the next volume uploading or running while the previous
// one's results go back to the host) the stream-ordered pool intermittently allocates fresh gigabytes
// (measured: steps of 45-700 ms among 19 ms ones).  This cache keeps freed blocks by size class and
// hands them out again with explicit ordering: a block freed on stream A and reused on stream B makes
// B wait for the event recorded at the free (nothing to wait for on the same stream, or when the
// event has completed).  Blocks come from cudaMalloc, only while the cache warms up.
// S3D_ALLOC=pool in the environment restores cudaMallocAsync/cudaFreeAsync.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <unordered_map>
#include <vector>

namespace s3d {

class DevCache {
public:
    static DevCache& get() { static DevCache c; return c; }

    cudaError_t alloc(void** out, size_t bytes, cudaStream_t st) {
        *out = nullptr;
        if (!enabled_) return cudaMallocAsync(out, bytes, st);
        int dev = 0;
        cudaGetDevice(&dev);
        const size_t cls = size_class(bytes);
        Blk b;
        bool hit = false, done = false;
        {
            std::lock_guard<std::mutex> lk(mu_);
            // Oldest first, and a block whose last use has COMPLETED before any other: the most recently freed
            // block of a class is typically still in use on its stream (the staging buffer of the volume that is
            // uploading has the size of an octave-0 level), and taking it would chain this stream behind that one.
            // (No short-cut for "same stream": a destroyed stream's handle value can be recycled by a new stream.)
            auto& fl = free_[dev & 63][cls];
            size_t pick = fl.size();
            for (size_t i = 0; i < fl.size(); ++i)
                if (cudaEventQuery(fl[i].ev) == cudaSuccess) { pick = i; done = true; break; }
            if (!done) cudaGetLastError();  // cudaErrorNotReady is not sticky, but clear it for the caller's checks
            // nothing idle in this class: a new block is better than queueing this stream behind another one, as long
            // as the device's total stays under the cap (the cache grows until no step has to wait)
            if (pick == fl.size() && !fl.empty() && total_[dev & 63] + cls > cap_) pick = 0;
            if (pick < fl.size()) { b = fl[pick]; fl.erase(fl.begin() + pick); cached_[dev & 63] -= cls; hit = true; }
        }
        if (hit) {
            if (!done) {
                cudaError_t e = cudaStreamWaitEvent(st, b.ev, 0);
                if (e != cudaSuccess) return e;
            }
            put_event(dev, b.ev);
            *out = b.p;
        } else {
            cudaError_t e = cudaMalloc(out, cls);
            if (e != cudaSuccess) {  // out of memory: give the cached blocks back and retry once
                cudaGetLastError();
                trim(dev);
                e = cudaMalloc(out, cls);
                if (e != cudaSuccess) return e;
            }
            std::lock_guard<std::mutex> lk(mu_);
            total_[dev & 63] += cls;
        }
        std::lock_guard<std::mutex> lk(mu_);
        live_[*out] = Live{cls, dev};
        return cudaSuccess;
    }

    void free(void* p, cudaStream_t st) {
        if (!p) return;
        if (!enabled_) { cudaFreeAsync(p, st); return; }
        int cur = 0, dev = 0;
        cudaGetDevice(&cur);
        size_t cls = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) { cudaFreeAsync(p, st); return; }  // not ours (allocated before a mode switch)
            cls = it->second.cls;
            dev = it->second.dev;
            live_.erase(it);
        }
        if (dev != cur) cudaSetDevice(dev);  // the block, the stream and the event belong to the block's device
        Blk b;
        b.p = p; b.st = st; b.cls = cls; b.ev = get_event(dev);
        cudaEventRecord(b.ev, st);
        bool keep = true;
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (total_[dev & 63] > cap_) keep = false;
            else { free_[dev & 63][cls].push_back(b); cached_[dev & 63] += cls; }
        }
        if (!keep) {
            cudaEventSynchronize(b.ev);
            put_event(dev, b.ev);
            cudaFree(p);
            std::lock_guard<std::mutex> lk(mu_);
            total_[dev & 63] -= cls;
        }
        if (dev != cur) cudaSetDevice(cur);
    }

    // give every cached block of the device back to the driver (after its last use has completed)
    void trim(int dev) {
        std::vector<Blk> all;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (auto& kv : free_[dev & 63]) { for (auto& b : kv.second) all.push_back(b); kv.second.clear(); }
            cached_[dev & 63] = 0;
        }
        size_t freed = 0;
        for (auto& b : all) { cudaEventSynchronize(b.ev); put_event(dev, b.ev); cudaFree(b.p); freed += b.cls; }
        std::lock_guard<std::mutex> lk(mu_);
        total_[dev & 63] -= freed;
    }
    size_t cached_bytes(int dev) { std::lock_guard<std::mutex> lk(mu_); return cached_[dev & 63]; }
    bool enabled() const { return enabled_; }

private:
    struct Blk { void* p; cudaStream_t st; cudaEvent_t ev; size_t cls; };
    struct Live { size_t cls; int dev; };
    DevCache() {
        const char* e = getenv("S3D_ALLOC");
        enabled_ = !(e && strcmp(e, "pool") == 0);
        const char* c = getenv("S3D_ALLOC_CAP_GB");
        cap_ = (size_t)((c ? atof(c) : 64.0) * (double)((size_t)1 << 30));  // fractional GB allowed (0.5 is 512 MiB, not 0)
        memset(cached_, 0, sizeof cached_);
        memset(total_, 0, sizeof total_);
    }
    // classes: multiples of the largest power of two <= bytes/8 (>= 512 B): at most 12.5 % slack, and the
    // count-dependent buffers (detections, keypoints) of similar volumes land in the same class
    static size_t size_class(size_t bytes) {
        if (bytes < 512) return 512;
        size_t g = 512;
        while (g * 16 <= bytes) g <<= 1;
        return (bytes + g - 1) / g * g;
    }
    cudaEvent_t get_event(int dev) {
        {
            std::lock_guard<std::mutex> lk(mu_);
            auto& f = events_[dev & 63];
            if (!f.empty()) { cudaEvent_t e = f.back(); f.pop_back(); return e; }
        }
        cudaEvent_t e = nullptr;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        return e;
    }
    void put_event(int dev, cudaEvent_t e) { std::lock_guard<std::mutex> lk(mu_); events_[dev & 63].push_back(e); }

    std::mutex mu_;
    bool enabled_ = true;
    size_t cap_ = (size_t)64 << 30;
    std::unordered_map<size_t, std::vector<Blk>> free_[64];
    std::unordered_map<void*, Live> live_;
    std::vector<cudaEvent_t> events_[64];
    size_t cached_[64];   // bytes sitting in the free lists
    size_t total_[64];    // bytes obtained from cudaMalloc and not yet returned (live + cached)
};

inline cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t st) { return DevCache::get().alloc(p, bytes, st); }
inline void dev_free(void* p, cudaStream_t st) { DevCache::get().free(p, st); }

}  // namespace s3d
