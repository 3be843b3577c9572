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

int numa_node_of(const void *addr);          // node of the page holding addr (-1: unknown)           (aie_expand_host.cpp)
int numa_node_count();
bool pin_current_thread_to_node(int node);
void flag_wait_zero(std::atomic<int> *flag);  // sleeps while *flag == 0 (futex)
void flag_wake_all(std::atomic<int> *flag);
int pci_numa_node(const char *bus_id);       // node a PCI device hangs off (-1: unknown)
void *alloc_on_node(size_t bytes, int node);  // anonymous pages bound to one node (nullptr: not possible here)
void free_on_node(void *p, size_t bytes);

// Persistent workers: run(items, fn) calls fn(i) for every item on the pool plus the calling thread and returns when all
// are done.  One job at a time (calls on a handle are serialised by contract).  numa_mode 0: one queue, unpinned threads.
// 1 / 2: worker k is pinned to node k % nodes and items are queued per node (the node of their destination memory); a
// worker drains its own node's queue first and then (mode 1) helps the other nodes, or (mode 2) stops.
class HostPool {
public:
    HostPool(int n_threads, int numa_mode) : mode_(numa_mode), nodes_(numa_mode ? numa_node_count() : 1) {
        if (nodes_ > MAXN) nodes_ = MAXN;
        for (int i = 0; i < n_threads; i++) workers_.emplace_back([this, i] { loop(i); });
    }
    ~HostPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; epoch_++; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    int size() const { return (int)workers_.size(); }
    int mode() const { return mode_; }
    int nodes() const { return nodes_; }
    // item_node: preferred node of every item (nullptr or mode 0: a single queue in item order)
    void run(int n, const signed char *item_node, const std::function<void(int)> &fn) {
        {
            std::lock_guard<std::mutex> g(m_);
            for (int k = 0; k < MAXN; k++) { q_[k].clear(); next_[k].store(0); }
            for (int i = 0; i < n; i++) {
                int k = (mode_ && item_node && item_node[i] >= 0 && item_node[i] < nodes_) ? item_node[i] : 0;
                q_[k].push_back(i);
            }
            fn_ = &fn; pending_ = (int)workers_.size(); epoch_++;
        }
        cv_.notify_all();
        work(-1);
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    static constexpr int MAXN = 8;
    void drain(int k) { for (size_t i; (i = next_[k].fetch_add(1)) < q_[k].size();) (*fn_)(q_[k][i]); }
    void work(int my_node) {   // my_node < 0: the calling thread, helps everywhere
        if (my_node >= 0) drain(my_node);
        if (my_node < 0 || mode_ != 2)
            for (int k = 0; k < nodes_; k++) if (k != my_node) drain(k);
    }
    void loop(int idx) {
        const int my_node = mode_ ? idx % nodes_ : 0;
        if (mode_) pin_current_thread_to_node(my_node);
        uint64_t seen = 0;
        for (;;) {
            { std::unique_lock<std::mutex> g(m_); cv_.wait(g, [&] { return epoch_ != seen; }); seen = epoch_; if (stop_) return; }
            work(mode_ ? my_node : -1);
            { std::lock_guard<std::mutex> g(m_); if (--pending_ == 0) done_.notify_one(); }
        }
    }
    int mode_, nodes_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(int)> *fn_ = nullptr;
    std::vector<int> q_[MAXN];
    std::atomic<size_t> next_[MAXN];
    int pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};

// bits -> 0.0f / 1.0f with non-temporal stores, AVX-512 or SSE2 picked at run time (aie_expand_host.cpp, host compiler)
void expand_bits(const uint32_t *src, int n, float *dst);
void expand_fence();
const char *expand_isa();

// non-zero bitmap + byte values -> int16 plane (zero-filled first)
inline void expand_sparse_i16(const uint32_t *mask, const uint8_t *vals, int n, int16_t *dst) {
    memset(dst, 0, 2 * (size_t)n);
    int k = 0;
    for (int w0 = 0; w0 < n; w0 += 32) {
        uint32_t m = mask[w0 >> 5];
        while (m) { const int j = __builtin_ctz(m); m &= m - 1; dst[w0 + j] = (int16_t)vals[k++]; }
    }
}
// one env: compact record -> the caller's (host) tensors; NULL outputs are skipped.  prog / cslot: the agents' flat
// program and the class slots of its entries (Tables).  Returns false when an index plane overflowed its capacity: the
// caller fetches that env's index planes directly.
inline bool expand_env(const CompactLayout &L, const uint8_t *rec, size_t env, const aie_host_out &o, const uint16_t *prog,
                       const uint16_t *cslot) {
    if (o.obs_agent_map) expand_bits((const uint32_t *)(rec + L.off_a_map), L.n_a_map, o.obs_agent_map + env * L.n_a_map);
    if (o.mask_agent) expand_bits((const uint32_t *)(rec + L.off_a_mask), L.n_a_mask, o.mask_agent + env * L.n_a_mask);
    if (o.obs_planner_map && L.n_p_map) expand_bits((const uint32_t *)(rec + L.off_p_map), L.n_p_map, o.obs_planner_map + env * L.n_p_map);
    if (o.mask_planner) expand_bits((const uint32_t *)(rec + L.off_p_mask), L.n_p_mask, o.mask_planner + env * L.n_p_mask);
    const int32_t *cnt = (const int32_t *)(rec + L.off_idx_cnt);
    bool ok = true;
    if (o.obs_agent_idx) {
        if (cnt[0] <= L.cap_a_idx) expand_sparse_i16((const uint32_t *)(rec + L.off_a_idx_mask), rec + L.off_a_idx_vals, L.n_a_idx, o.obs_agent_idx + env * L.n_a_idx);
        else ok = false;
    }
    if (o.obs_planner_idx && L.n_p_idx) {
        if (cnt[1] <= L.cap_p_idx) expand_sparse_i16((const uint32_t *)(rec + L.off_p_idx_mask), rec + L.off_p_idx_vals, L.n_p_idx, o.obs_planner_idx + env * L.n_p_idx);
        else ok = false;
    }
    if (o.obs_agent_flat) {
        const float *f_sh = (const float *)(rec + L.off_f_sh), *f_ag = (const float *)(rec + L.off_f_ag);
        const uint8_t *c8 = rec + L.off_f_cnt; const uint16_t *c16 = (const uint16_t *)(rec + L.off_f_cnt);
        float *d = o.obs_agent_flat + env * L.n_a_flat;
        for (int a = 0; a < L.A; a++, d += L.Fa) {
            const float *ag = f_ag + a * L.n_ag;
            for (int j = 0; j < L.Fa; j++) {
                const int kind = AIE_FLAT_KIND(prog[j]), slot = cslot[j];
                d[j] = kind == FK_SHARED ? f_sh[slot] : kind == FK_AGENT ? ag[slot]
                       : (float)(L.cnt_bytes == 1 ? (unsigned)c8[a * L.n_cnt + slot] : (unsigned)c16[a * L.n_cnt + slot]);
            }
        }
    }
    if (o.obs_planner_flat) memcpy(o.obs_planner_flat + env * L.n_p_flat, rec + L.off_p_flat, 4 * (size_t)L.n_p_flat);
    if (o.obs_planner_agents && L.n_p_agents) memcpy(o.obs_planner_agents + env * L.n_p_agents, rec + L.off_p_agents, 4 * (size_t)L.n_p_agents);
    if (o.obs_time) o.obs_time[env] = *(const float *)(rec + L.off_time);
    if (o.done) o.done[env] = *(const int32_t *)(rec + L.off_done);
    if (o.reward) memcpy(o.reward + env * L.n_rew, rec + L.off_rew, 8 * (size_t)L.n_rew);
    expand_fence();   // the non-temporal stores of this env are globally visible before the work item is reported done
    return ok;
}

}  // namespace aie
