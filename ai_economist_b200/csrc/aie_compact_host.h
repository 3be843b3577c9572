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

constexpr int AIE_MAX_SLICES = 16;         // transfer slices of the compacted D2H copy (one event each)
constexpr int AIE_MAX_HOST_THREADS = 128;  // expansion threads

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

// bits -> 0.0f / 1.0f with non-temporal stores, AVX-512 or SSE2 picked at run time (aie_expand_host.cpp, host compiler)
void expand_bits(const uint32_t *src, int n, float *dst);
void expand_fence();
const char *expand_isa();

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
