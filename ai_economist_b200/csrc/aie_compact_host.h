// aie_compact_host.h — host side of the compacted transfer: expanding compact env records into the caller's tensors
// with a small persistent thread pool (no CUDA here; shared by the CUDA library and the emulation build).
#pragma once
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/aie_b200.h"
#include "aie_compact.cuh"

namespace aie {

// Persistent workers: run(n, fn) calls fn(i) for i in [0, n) on the pool plus the calling thread and returns when all
// are done.  One job at a time (calls on a handle are serialised by contract).
class HostPool {
public:
    explicit HostPool(int n_threads) {
        for (int i = 0; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; epoch_++; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    void run(int n, const std::function<void(int)> &fn) {
        { std::lock_guard<std::mutex> g(m_); fn_ = &fn; n_ = n; next_.store(0); pending_ = (int)workers_.size(); epoch_++; }
        cv_.notify_all();
        work();
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void work() { for (int i; (i = next_.fetch_add(1)) < n_;) (*fn_)(i); }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> g(m_); cv_.wait(g, [&] { return epoch_ != seen; }); seen = epoch_; if (stop_) return; }
            work();
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::atomic<int> next_{0};
    int n_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

// bits -> 0.0f / 1.0f.  The expansion is a pure write stream (296 MB per step for c2 at 8 192 envs), so on x86 the
// 16-byte groups go out with non-temporal stores (no read-for-ownership of the destination lines); a 16-entry table maps
// four bits to four floats.  `src` must be readable up to the 8 bytes holding the last bit (true inside a compact record).
#if defined(__SSE2__)
#include <emmintrin.h>
struct NibbleTable {
    __m128 v[16];
    NibbleTable() { for (int k = 0; k < 16; k++) v[k] = _mm_set_ps((float)((k >> 3) & 1), (float)((k >> 2) & 1), (float)((k >> 1) & 1), (float)(k & 1)); }
};
inline void expand_bits(const uint32_t *src, int n, float *dst) {
    static const NibbleTable T;
    auto bit = [&](int i) { return (float)((src[i >> 5] >> (i & 31)) & 1u); };
    int i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = bit(i); i++; }
    for (; i + 4 <= n; i += 4) {
        uint64_t w;
        memcpy(&w, (const uint8_t *)src + 4 * (i >> 5), 8);     // the word holding bit i and the next one
        _mm_stream_ps(dst + i, T.v[(w >> (i & 31)) & 15u]);
    }
    for (; i < n; i++) dst[i] = bit(i);
}
inline void expand_fence() { _mm_sfence(); }
#else
inline void expand_bits(const uint32_t *src, int n, float *dst) {
    for (int i = 0; i < n; i++) dst[i] = (float)((src[i >> 5] >> (i & 31)) & 1u);
}
inline void expand_fence() {}
#endif

// one env: compact record -> the caller's (host) tensors; NULL outputs are skipped
inline void expand_env(const CompactLayout &L, const uint8_t *rec, size_t env, const aie_host_out &o) {
    if (o.obs_agent_map) expand_bits((const uint32_t *)(rec + L.off_a_map), L.n_a_map, o.obs_agent_map + env * L.n_a_map);
    if (o.mask_agent) expand_bits((const uint32_t *)(rec + L.off_a_mask), L.n_a_mask, o.mask_agent + env * L.n_a_mask);
    if (o.obs_planner_map && L.n_p_map) expand_bits((const uint32_t *)(rec + L.off_p_map), L.n_p_map, o.obs_planner_map + env * L.n_p_map);
    if (o.mask_planner) expand_bits((const uint32_t *)(rec + L.off_p_mask), L.n_p_mask, o.mask_planner + env * L.n_p_mask);
    if (o.obs_agent_idx) { int16_t *d = o.obs_agent_idx + env * L.n_a_idx; const uint8_t *s = rec + L.off_a_idx; for (int i = 0; i < L.n_a_idx; i++) d[i] = (int16_t)s[i]; }
    if (o.obs_planner_idx && L.n_p_idx) { int16_t *d = o.obs_planner_idx + env * L.n_p_idx; const uint8_t *s = rec + L.off_p_idx; for (int i = 0; i < L.n_p_idx; i++) d[i] = (int16_t)s[i]; }
    if (o.obs_agent_flat) memcpy(o.obs_agent_flat + env * L.n_a_flat, rec + L.off_a_flat, 4 * (size_t)L.n_a_flat);
    if (o.obs_planner_flat) memcpy(o.obs_planner_flat + env * L.n_p_flat, rec + L.off_p_flat, 4 * (size_t)L.n_p_flat);
    if (o.obs_planner_agents && L.n_p_agents) memcpy(o.obs_planner_agents + env * L.n_p_agents, rec + L.off_p_agents, 4 * (size_t)L.n_p_agents);
    if (o.obs_time) o.obs_time[env] = *(const float *)(rec + L.off_time);
    if (o.done) o.done[env] = *(const int32_t *)(rec + L.off_done);
    if (o.reward) memcpy(o.reward + env * L.n_rew, rec + L.off_rew, 8 * (size_t)L.n_rew);
    expand_fence();   // the non-temporal stores of this env are globally visible before the work item is reported done
}

}  // namespace aie
