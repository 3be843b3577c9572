// aie_abi.inl — the extern "C" entry points declared in include/aie_b200.h.
// Included by aie_abi.cu (CUDA backend = the product) and by tests/emu/aie_emu.cpp (1-lane host emulation
// of the same device source, CPU logic tests only).  The includer provides, in namespace aie::be:
//   struct State;  int init(aie_env*);  void destroy(aie_env*);
//   int upload(aie_env*, void *dst, const void *src, size_t n, void *stream);
//   int download(aie_env*, void *dst, const void *src, size_t n, void *stream);
//   int dev_copy(aie_env*, void *dst, const void *src, size_t n, void *stream);
//   int sync(aie_env*, void *stream);   int sync_all(aie_env*);
//   int launch_finish_reset(aie_env*, int lo, int n, void *stream);
//   int launch_step(aie_env*, int emit_obs, void *stream);
//   int launch_observe(aie_env*, int lo, int n, void *stream);
//   int launch_sample(aie_env*, uint64_t seed, void *stream);
//   int compact_buffers(aie_env*, size_t bytes, uint8_t **dev, uint8_t **host);   (library-owned, allocated once)
//   int launch_pack(aie_env*, const aie::CompactLayout&, uint8_t *dev, void *stream);
//   int download_slice(aie_env*, int k, void *host, const void *dev, size_t n, void *stream);   (async copy + event k)
//   int wait_slice(aie_env*, int k);                                                           (any thread)
//   double slice_device_ms(aie_env*, int k);   device-clock time of slice k's arrival since the call's first enqueue (-1: n/a)
//   void *const_upload(const void *host, size_t bytes);  void *dev_alloc(size_t bytes);  void const_free(void *dev);
//   struct DevScope { DevScope(int device); ~DevScope(); bool ok() const; };   makes `device` current for the scope of one
//       entry point and restores the caller's device on exit (a handle can be used while another device is current)
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "aie_compact_host.h"

static thread_local std::string g_last_error;

struct aie_env {
    aie::DevCfg cfg;
    aie::Tables tables;
    aie_config ucfg;
    int n_envs, device;
    aie::DevBufs bufs;
    bool bound, loaded;
    int64_t launches;
    uint64_t sample_calls;
    aie::be::State be;
    std::vector<aie_flat_field> flat_layout[3];
    void *dyn_prob_dev = nullptr, *dyn_work_dev = nullptr;   // dynamic layouts: probability maps, per-env work maps
    aie::HostPool *pool = nullptr;   // aie_step_host_compact: created on first use
    std::vector<signed char> item_node;   // NUMA node of every work item's destination (cached per output pointer)
    const void *item_node_key = nullptr; size_t item_node_n = 0;
    double host_timing[AIE_HOST_TIMING_WORDS] = {};   // last aie_step_host_compact call (aie_get_host_timing)
};

static int fail(int code, const std::string &msg) { g_last_error = msg; return code; }
// every entry point that touches the device runs with the handle's own device current (and puts the caller's back)
#define AIE_DEVICE_SCOPE(dev) aie::be::DevScope dev_scope_(dev); \
    if (!dev_scope_.ok()) return fail(AIE_ECUDA, "cannot make the handle's CUDA device current")

extern "C" {

const char *aie_last_error(void) { return g_last_error.c_str(); }
int aie_abi_version(void) { return AIE_ABI_VERSION; }

int aie_create(const aie_config *cfg, int32_t n_envs, int32_t device, aie_env **out) {
    if (!cfg || !out) return fail(AIE_EINVAL, "null argument");
    aie_env *env = new (std::nothrow) aie_env();
    if (!env) return fail(AIE_ENOMEM, "out of host memory");
    std::string err;
    int rc = aie::build_devcfg(*cfg, n_envs, env->cfg, env->tables, err, env->flat_layout);
    if (rc != AIE_OK) { delete env; return fail(rc, "aie_create: " + err); }
    env->ucfg = *cfg; env->n_envs = n_envs; env->device = device;
    if (getenv("AIE_VERBOSE")) {   // record / staging layout (host-side facts; the backend adds its launch geometry)
        const aie::DevCfg &c = env->cfg;
        fprintf(stderr, "[aie] A %d map %dx%d: record %d B (resident %d, obs prefix %d, mt @%d, price_hist @%d), step scratch %d, obs extra %d "
                        "(alias mt %d), chunk %d agents, mw %d, Fa %d Fp %d Fpa %d Na %d Np %d; ob:", c.A, c.H, c.W, c.rec_bytes, c.resident_bytes,
                c.obs_prefix_bytes, c.off_mt, c.off_price_hist, c.step_scratch_bytes, c.obs_extra_bytes, c.obs_alias_mt, c.ob_chunk, c.mw,
                c.Fa, c.Fp, c.Fpa, c.Na, c.Np);
        for (int i = 0; i < aie::OB_COUNT; i++) fprintf(stderr, " %d", c.ob[i]);
        fprintf(stderr, "\n");
    }
    memset(&env->bufs, 0, sizeof(env->bufs));
    env->bound = env->loaded = false; env->launches = 0; env->sample_calls = 0;
    aie::be::DevScope dev_scope_(device);   // init() validates the ordinal itself and reports the precise error
    rc = aie::be::init(env);
    if (rc != AIE_OK) { delete env; return rc; }
    if (env->cfg.dyn_layout) {   // library-owned device memory of the layout generator
        const size_t hw = (size_t)env->cfg.HW;
        // per env: the float64 work map + 16 doubles holding MultiZone's region -> zone-type bytes
        env->dyn_prob_dev = cfg->dyn_prob ? aie::be::const_upload(cfg->dyn_prob, 2 * hw * sizeof(double)) : aie::be::dev_alloc(2 * hw * sizeof(double));
        env->dyn_work_dev = aie::be::dev_alloc((size_t)n_envs * (hw + 16) * sizeof(double));
        if (!env->dyn_prob_dev || !env->dyn_work_dev) { aie_destroy(env); return fail(AIE_ENOMEM, "aie_create: device memory for the layout generator"); }
        env->bufs.dyn_prob = (const double *)env->dyn_prob_dev; env->bufs.dyn_work = (double *)env->dyn_work_dev;
    }
    env->ucfg.dyn_prob = nullptr;   // the caller's pointer is not kept
    *out = env;
    return AIE_OK;
}

int aie_destroy(aie_env *env) {
    if (!env) return AIE_OK;
    aie::be::DevScope dev_scope_(env->device);
    aie::be::destroy(env);
    if (env->dyn_prob_dev) aie::be::const_free(env->dyn_prob_dev);
    if (env->dyn_work_dev) aie::be::const_free(env->dyn_work_dev);
    delete env->pool;
    delete env;
    return AIE_OK;
}

int aie_get_dims(const aie_env *env, aie_dims *out) {
    if (!env || !out) return fail(AIE_EINVAL, "null argument");
    aie::fill_dims(env->cfg, *out);
    return AIE_OK;
}

int aie_get_field(const aie_env *env, const char *name, aie_field *out) {
    if (!env || !name || !out) return fail(AIE_EINVAL, "null argument");
    if (aie::lookup_field(env->cfg, name, out) != AIE_OK) return fail(AIE_EINVAL, std::string("unknown state field ") + name);
    return AIE_OK;
}

int aie_get_flat_layout(const aie_env *env, int32_t which, aie_flat_field *out, int32_t cap) {
    if (!env || which < 0 || which > 2 || (cap > 0 && !out)) return fail(AIE_EINVAL, "aie_get_flat_layout: bad argument");
    const std::vector<aie_flat_field> &v = env->flat_layout[which];
    for (int i = 0; i < (int)v.size() && i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

int aie_bind_buffers(aie_env *env, const aie_buffers *b) {
    if (!env || !b) return fail(AIE_EINVAL, "null argument");
    const aie::DevCfg &c = env->cfg;
    if (!b->state || !b->state0 || !b->actions_agent || (!c.no_spatial && (!b->obs_agent_map || !b->obs_agent_idx)) || !b->obs_agent_flat ||
        !b->mask_agent || !b->obs_planner_flat || (c.Fpa > 0 && !b->obs_planner_agents) || !b->mask_planner || !b->obs_time ||
        !b->reward || !b->done)
        return fail(AIE_EINVAL, "aie_bind_buffers: a required buffer is NULL");
    if (c.planner_spatial && (!b->obs_planner_map || !b->obs_planner_idx))
        return fail(AIE_EINVAL, "aie_bind_buffers: planner map buffers required when planner_gets_spatial_info");
    if (c.n_act_p > 0 && !b->actions_planner) return fail(AIE_EINVAL, "aie_bind_buffers: actions_planner required");
    if (((uintptr_t)b->state & 15) || ((uintptr_t)b->state0 & 15)) return fail(AIE_EINVAL, "state buffers must be 16-byte aligned");
    aie::DevBufs &d = env->bufs;
    d.state = (uint8_t *)b->state; d.state0 = (uint8_t *)b->state0; d.final = (uint8_t *)b->episode_final;
    if ((uintptr_t)b->episode_final & 15) return fail(AIE_EINVAL, "episode_final must be 16-byte aligned");
    d.act_a = b->actions_agent; d.act_p = b->actions_planner;
    d.a_map = b->obs_agent_map; d.a_idx = b->obs_agent_idx; d.a_flat = b->obs_agent_flat; d.a_mask = b->mask_agent;
    d.p_map = b->obs_planner_map; d.p_idx = b->obs_planner_idx; d.p_flat = b->obs_planner_flat;
    d.p_agents = b->obs_planner_agents; d.p_mask = b->mask_planner; d.time_obs = b->obs_time;
    d.events = b->events; d.event_envs = b->events ? b->event_envs : 0; d.event_cap = b->event_cap;
    if (b->events && (b->event_envs < 0 || b->event_envs > env->n_envs || b->event_cap < 1))
        return fail(AIE_EINVAL, "aie_bind_buffers: bad event log shape");
    d.rew = b->reward; d.done = b->done;  // d.tab was set by the backend at creation
    env->bound = true;
    return AIE_OK;
}

int aie_load_state(aie_env *env, const aie_host_state *hs, int32_t env_lo, void *stream) {
    if (!env || !hs) return fail(AIE_EINVAL, "null argument");
    if (!env->bound) return fail(AIE_ESTATE, "aie_load_state: buffers not bound");
    AIE_DEVICE_SCOPE(env->device);
    const aie::DevCfg &c = env->cfg;
    if (hs->n < 1 || env_lo < 0 || env_lo + hs->n > env->n_envs) return fail(AIE_EINVAL, "aie_load_state: env range out of bounds");
    if (!hs->stone || !hs->wood || !hs->stone_src || !hs->wood_src || !hs->loc || !hs->coin || !hs->build_payment ||
        !hs->build_skill || !hs->bonus_gather_prob || !hs->mt_key || !hs->mt_pos)
        return fail(AIE_EINVAL, "aie_load_state: a required host array is NULL");
    std::vector<uint8_t> host((size_t)hs->n * c.rec_bytes);
    std::string err;
    for (int i = 0; i < hs->n; i++) {
        int rc = aie::pack_record(c, *hs, i, host.data() + (size_t)i * c.rec_bytes, err);
        if (rc != AIE_OK) return fail(rc, "aie_load_state: " + err);
    }
    uint8_t *dst = env->bufs.state + (size_t)env_lo * c.rec_bytes;
    int rc = aie::be::upload(env, dst, host.data(), host.size(), stream);
    if (rc != AIE_OK) return rc;
    rc = aie::be::sync(env, stream);  // the staging vector goes out of scope below
    if (rc != AIE_OK) return rc;
    rc = aie::be::launch_finish_reset(env, env_lo, hs->n, stream);
    if (rc != AIE_OK) return rc;
    rc = aie::be::dev_copy(env, env->bufs.state0 + (size_t)env_lo * c.rec_bytes, dst, host.size(), stream);
    if (rc != AIE_OK) return rc;
    rc = aie::be::launch_observe(env, env_lo, hs->n, stream);
    if (rc != AIE_OK) return rc;
    env->loaded = true;
    return aie::be::sync(env, stream);
}

int aie_step(aie_env *env, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_step: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::launch_step(env, 1, stream);  // dynamics + observations fused in one launch
}

int aie_step_dynamics(aie_env *env, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_step_dynamics: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::launch_step(env, 0, stream);
}

int aie_sample_random_actions(aie_env *env, uint64_t seed, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_sample_random_actions: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::launch_sample(env, seed, stream);
}

int aie_set_fused_policy(aie_env *env, uint64_t seed, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_set_fused_policy: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    env->bufs.policy_seed = seed ? (aie::host_mix64(seed) | 1ull) : 0ull;
    return seed ? aie::be::launch_sample(env, seed, stream) : AIE_OK;   // actions for the very next step
}

int aie_observe(aie_env *env, void *stream) {
    if (!env) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_observe: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    return aie::be::launch_observe(env, 0, env->n_envs, stream);
}

int aie_step_host(aie_env *env, const int32_t *act_a, const int32_t *act_p, const aie_host_out *o, void *stream) {
    if (!env || !act_a || !o) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_step_host: bind buffers and load state first");
    AIE_DEVICE_SCOPE(env->device);
    const aie::DevCfg &c = env->cfg;
    const size_t E = env->n_envs, A = c.A;
    int rc = aie::be::upload(env, (void *)env->bufs.act_a, act_a, E * A * c.n_act_a * 4, stream);
    if (rc != AIE_OK) return rc;
    if (c.n_act_p > 0) {
        if (!act_p) return fail(AIE_EINVAL, "aie_step_host: planner actions required");
        rc = aie::be::upload(env, (void *)env->bufs.act_p, act_p, E * c.n_act_p * 4, stream);
        if (rc != AIE_OK) return rc;
    }
    rc = aie_step(env, stream);
    if (rc != AIE_OK) return rc;
    const aie::DevBufs &d = env->bufs;
    struct { void *h; const void *dev; size_t n; } cp[] = {
        {o->obs_agent_map, d.a_map, E * A * (size_t)c.a_map_elems * 4}, {o->obs_agent_idx, d.a_idx, E * A * (size_t)c.a_idx_elems * 2},
        {o->obs_agent_flat, d.a_flat, E * A * c.Fa * 4}, {o->mask_agent, d.a_mask, E * A * c.Na * 4},
        {o->obs_planner_map, c.planner_spatial ? d.p_map : nullptr, E * c.M * c.HW * 4},
        {o->obs_planner_idx, c.planner_spatial ? d.p_idx : nullptr, E * 2 * c.HW * 2},
        {o->obs_planner_flat, d.p_flat, E * c.Fp * 4}, {o->obs_planner_agents, d.p_agents, E * A * c.Fpa * 4},
        {o->mask_planner, d.p_mask, E * c.Np * 4}, {o->obs_time, d.time_obs, E * 4},
        {o->reward, d.rew, E * (A + 1) * 8}, {o->done, d.done, E * 4},
    };
    for (auto &x : cp)
        if (x.h && x.dev) {
            rc = aie::be::download(env, x.h, x.dev, x.n, stream);
            if (rc != AIE_OK) return rc;
        }
    return aie::be::sync(env, stream);
}

int32_t aie_compact_bytes_per_env(const aie_env *env) { return env ? aie::compact_layout(env->cfg).bytes : 0; }

int aie_step_host_compact(aie_env *env, const int32_t *act_a, const int32_t *act_p, const aie_host_out *o, int32_t n_threads,
                          void *stream) {
    if (!env || !act_a || !o) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->loaded) return fail(AIE_ESTATE, "aie_step_host_compact: bind buffers and load state first");
    const std::chrono::steady_clock::time_point t_call = std::chrono::steady_clock::now();
    AIE_DEVICE_SCOPE(env->device);
    const aie::DevCfg &c = env->cfg;
    const size_t E = env->n_envs;
    int rc = aie::be::mark_call_start(env, stream);
    if (rc != AIE_OK) return rc;
    rc = aie::be::upload(env, (void *)env->bufs.act_a, act_a, E * c.A * c.n_act_a * 4, stream);
    if (rc != AIE_OK) return rc;
    if (c.n_act_p > 0) {
        if (!act_p) return fail(AIE_EINVAL, "aie_step_host_compact: planner actions required");
        rc = aie::be::upload(env, (void *)env->bufs.act_p, act_p, E * c.n_act_p * 4, stream);
        if (rc != AIE_OK) return rc;
    }
    const aie::CompactLayout L = aie::compact_layout(c);
    uint8_t *dev = nullptr, *host = nullptr;
    rc = aie::be::compact_buffers(env, E * (size_t)L.bytes, &dev, &host);
    if (rc != AIE_OK) return rc;
    // The batch is stepped in up to AIE_E2E_CHUNKS launches (env replicas never interact); the compact records of a chunk go
    // down, on the library's copy stream, in slices (AIE_MAX_SLICES in total), each followed by an event, while the next chunk
    // still steps.  A work item first waits for the slice holding its envs, so the expansion of the early slices overlaps
    // the later chunks' kernels and transfers.
    using clk = std::chrono::steady_clock;
    const clk::time_point t0 = clk::now();
    auto ms_since = [&](clk::time_point t) { return std::chrono::duration<double, std::milli>(t - t0).count(); };
    int chunk = 16;   // envs per work item (AIE_E2E_ITEM_ENVS: 8 .. 64); short items keep the tail after the last slice short
    if (const char *v = getenv("AIE_E2E_ITEM_ENVS")) { const int k = atoi(v); if (k >= 8 && k <= 64) chunk = k; }
    const int n_items = (int)((E + chunk - 1) / chunk);
    int n_slices = n_items < aie::AIE_MAX_SLICES ? n_items : aie::AIE_MAX_SLICES;
    const int items_per_slice = (n_items + n_slices - 1) / n_slices;
    n_slices = (n_items + items_per_slice - 1) / items_per_slice;
    // small records (one warp per env): 4 launches; large records (one CTA per env, a few hundred resident per wave): 2, more
    // would lose to wave quantisation what the earlier first slice gains (profiles/r02z_e2e_transfer_knobs.txt section 8)
    int n_chunks = c.mw > 1 ? 2 : 4;
    if (const char *v = getenv("AIE_E2E_CHUNKS")) n_chunks = atoi(v);
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > n_slices) n_chunks = n_slices;
    const int slices_per_chunk = (n_slices + n_chunks - 1) / n_chunks;
    auto slice_lo = [&](int k) { const size_t v = (size_t)k * items_per_slice * chunk; return v < E ? v : E; };
    for (int k = 0; k < n_slices; k++) {
        if (k % slices_per_chunk == 0) {
            const int k_end = k + slices_per_chunk < n_slices ? k + slices_per_chunk : n_slices;
            const int lo = (int)slice_lo(k), hi = (int)slice_lo(k_end);
            rc = aie::be::launch_step_range(env, 1, lo, hi, stream);
            if (rc != AIE_OK) return rc;
            rc = aie::be::launch_pack_range(env, L, dev, lo, hi, stream);
            if (rc != AIE_OK) return rc;
            rc = aie::be::chunk_ready(env, stream);
            if (rc != AIE_OK) return rc;
        }
        const size_t lo = slice_lo(k), hi = slice_lo(k + 1);
        rc = aie::be::download_slice(env, k, host + lo * (size_t)L.bytes, dev + lo * (size_t)L.bytes, (hi - lo) * (size_t)L.bytes, stream);
        if (rc != AIE_OK) return rc;
    }
    rc = aie::be::copies_done(env, stream);
    if (rc != AIE_OK) return rc;
    const double t_enqueued = ms_since(clk::now());
    // <= 0: one thread per physical core of an SMT-2 host - the expansion is bound by the memory controllers, 64 threads were
    // as fast as 96 or 128 on the 2 x 32-core B200 host and burn half the CPU time (profiles/r02z_e2e_transfer_knobs.txt)
    int want = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency() / 2;
    if (want < 1) want = 1;
    if (want > aie::AIE_MAX_HOST_THREADS) want = aie::AIE_MAX_HOST_THREADS;
    // AIE_E2E_NUMA: 0 unpinned threads, one queue; 1 node-local items first, then help the other nodes; 2 node-local only
    // (default: measured fastest on a two-socket host, profiles/r02g_e2e_probe_c2.log - remote non-temporal stores cost
    // more than the idle threads would add)
    int numa_mode = 2;
    if (const char *v = getenv("AIE_E2E_NUMA")) numa_mode = atoi(v);
    if (numa_mode < 0 || numa_mode > 2) numa_mode = 2;
    if (!env->pool || env->pool->size() != want - 1 || env->pool->mode() != numa_mode) {
        delete env->pool; env->pool = new aie::HostPool(want - 1, numa_mode); env->item_node_key = nullptr;
    }
    if (numa_mode && env->pool->nodes() > 1 && o->obs_agent_map &&
        (env->item_node_key != (const void *)o->obs_agent_map || env->item_node_n != (size_t)n_items)) {
        env->item_node.assign((size_t)n_items, -1);   // where the (dominant) agent-map rows of every item live
        for (int i = 0; i < n_items; i++) {
            size_t mid = (size_t)i * chunk + chunk / 2;
            if (mid >= E) mid = E - 1;
            env->item_node[i] = (signed char)aie::numa_node_of(o->obs_agent_map + mid * (size_t)L.n_a_map);
        }
        env->item_node_key = (const void *)o->obs_agent_map; env->item_node_n = (size_t)n_items;
    }
    const signed char *item_node = (numa_mode && env->item_node_n == (size_t)n_items && env->item_node_key == (const void *)o->obs_agent_map)
                                       ? env->item_node.data() : nullptr;
    const aie_host_out out = *o;
    std::atomic<int> failed{0};
    std::atomic<int64_t> first_slice_us{-1}, last_slice_us{-1}, wait_us{0}, busy_us{0};
    std::mutex overflow_m;
    std::vector<size_t> overflow;   // envs whose index planes hold more non-zero elements than the compact record carries
    // One thread per slice waits for its event in the driver; the others watch a flag (hundreds of concurrent
    // cudaEventSynchronize calls serialise on the driver's lock and cost more than the transfer they wait for).  The watchers
    // spin, yielding their CPU on every round: measured on the two-socket B200 host, sleeping waiters (futex) let the sockets'
    // uncore clock down and the DMA itself takes twice as long (profiles/r02z_e2e_transfer_knobs.txt).  AIE_E2E_SPIN_US=n
    // makes them sleep after n microseconds - for hosts where the process runs under a tight CPU quota.
    std::atomic<int> claimed[aie::AIE_MAX_SLICES], arrived[aie::AIE_MAX_SLICES];
    for (int k = 0; k < aie::AIE_MAX_SLICES; k++) { claimed[k].store(0); arrived[k].store(0); }
    int spin_us = -1;
    if (const char *v = getenv("AIE_E2E_SPIN_US")) spin_us = atoi(v);
    auto slice_arrived = [&](int k) {
        int v = arrived[k].load(std::memory_order_acquire);
        if (v == 0) {
            int expected = 0;
            if (claimed[k].compare_exchange_strong(expected, 1)) {
                v = aie::be::wait_slice(env, k) == AIE_OK ? 1 : -1;
                arrived[k].store(v, std::memory_order_release);
                aie::flag_wake_all(&arrived[k]);
            } else {
                const clk::time_point s0 = clk::now();
                while ((v = arrived[k].load(std::memory_order_acquire)) == 0) {
                    if (spin_us >= 0 && std::chrono::duration<double, std::micro>(clk::now() - s0).count() >= spin_us) {
                        aie::flag_wait_zero(&arrived[k]);
                        v = arrived[k].load(std::memory_order_acquire);
                        break;
                    }
#if defined(__x86_64__) || defined(__i386__)
                    for (int i = 0; i < 32; i++) __builtin_ia32_pause();
#endif
                    std::this_thread::yield();   // never hold a CPU against the thread that waits in the driver
                }
            }
        }
        return v == 1;
    };
    auto job = [&](int item) {
        const int k = item / items_per_slice;
        const clk::time_point w0 = clk::now();
        if (!slice_arrived(k)) { failed.store(1); return; }
        const clk::time_point w1 = clk::now();
        if (item % items_per_slice == 0 && (k == 0 || k == n_slices - 1))
            (k == 0 ? first_slice_us : last_slice_us).store((int64_t)(1e3 * ms_since(w1)));
        const size_t hi = (size_t)(item + 1) * chunk < E ? (size_t)(item + 1) * chunk : E;
        for (size_t e = (size_t)item * chunk; e < hi; e++)
            if (!aie::expand_env(L, host + e * (size_t)L.bytes, e, out, env->tables.w, env->tables.w + c.tab_cslot)) {
                std::lock_guard<std::mutex> g(overflow_m);
                overflow.push_back(e);
            }
        const clk::time_point w2 = clk::now();
        wait_us.fetch_add(std::chrono::duration_cast<std::chrono::microseconds>(w1 - w0).count());
        busy_us.fetch_add(std::chrono::duration_cast<std::chrono::microseconds>(w2 - w1).count());
    };
    env->pool->run(n_items, item_node, job);
    const double t_expanded = ms_since(clk::now());
    double *ht = env->host_timing;
    ht[0] = t_enqueued; ht[1] = 1e-3 * first_slice_us.load(); ht[2] = 1e-3 * (n_slices > 1 ? last_slice_us.load() : first_slice_us.load());
    ht[3] = t_expanded; ht[4] = (double)n_slices; ht[14] = (double)n_chunks; ht[15] = (double)aie::be::staging_node(env); ht[5] = (double)want; ht[6] = (double)(E * (size_t)L.bytes);
    ht[7] = std::chrono::duration<double, std::milli>(t0 - t_call).count();
    ht[8] = 1e-3 * wait_us.load(); ht[9] = 1e-3 * busy_us.load();   // summed over the threads
    ht[10] = aie::be::slice_device_ms(env, 0); ht[11] = aie::be::slice_device_ms(env, n_slices - 1);
    if (const char *rep = getenv("AIE_E2E_REPEAT_EXPAND")) {   // tuning aid: the expansion alone, every slice already on the host
        const int n = atoi(rep);
        const clk::time_point r0 = clk::now();
        for (int i = 0; i < n; i++) env->pool->run(n_items, item_node, job);
        ht[12] = n > 0 ? std::chrono::duration<double, std::milli>(clk::now() - r0).count() / n : 0.0;
    }
    if (failed.load()) return fail(AIE_ECUDA, "aie_step_host_compact: waiting for a transfer slice failed");
    for (size_t e : overflow) {   // rare: fetch those envs' index planes as they are
        if (out.obs_agent_idx) { rc = aie::be::download(env, out.obs_agent_idx + e * L.n_a_idx, env->bufs.a_idx + e * L.n_a_idx, 2 * (size_t)L.n_a_idx, stream); if (rc != AIE_OK) return rc; }
        if (out.obs_planner_idx && L.n_p_idx) { rc = aie::be::download(env, out.obs_planner_idx + e * L.n_p_idx, env->bufs.p_idx + e * L.n_p_idx, 2 * (size_t)L.n_p_idx, stream); if (rc != AIE_OK) return rc; }
    }
    env->host_timing[13] = (double)overflow.size();
    rc = aie::be::sync(env, stream);
    if (rc != AIE_OK) return rc;
    return AIE_OK;
}

int aie_get_host_timing(const aie_env *env, double *out, int32_t cap) {
    if (!env || (cap > 0 && !out)) return fail(AIE_EINVAL, "null argument");
    for (int i = 0; i < cap && i < AIE_HOST_TIMING_WORDS; i++) out[i] = env->host_timing[i];
    return AIE_HOST_TIMING_WORDS;
}

int aie_read_state(aie_env *env, int32_t e, const aie_state_dump *out) {
    if (!env || !out) return fail(AIE_EINVAL, "null argument");
    if (!env->bound) return fail(AIE_ESTATE, "aie_read_state: buffers not bound");
    if (e < 0 || e >= env->n_envs) return fail(AIE_EINVAL, "aie_read_state: env index out of range");
    AIE_DEVICE_SCOPE(env->device);
    const aie::DevCfg &c = env->cfg;
    std::vector<uint8_t> rec(c.rec_bytes);
    int rc = aie::be::sync_all(env);  // debug path: order against work on any stream
    if (rc != AIE_OK) return rc;
    rc = aie::be::download(env, rec.data(), env->bufs.state + (size_t)e * c.rec_bytes, rec.size(), nullptr);
    if (rc != AIE_OK) return rc;
    rc = aie::be::sync(env, nullptr);
    if (rc != AIE_OK) return rc;
    aie::unpack_record(c, rec.data(), *out);
    return AIE_OK;
}

int aie_read_episode_final(aie_env *env, int32_t e, const aie_state_dump *out) {
    if (!env || !out) return fail(AIE_EINVAL, "null argument");
    if (!env->bound || !env->bufs.final) return fail(AIE_ESTATE, "aie_read_episode_final: no episode_final buffer bound");
    if (e < 0 || e >= env->n_envs) return fail(AIE_EINVAL, "aie_read_episode_final: env index out of range");
    AIE_DEVICE_SCOPE(env->device);
    const aie::DevCfg &c = env->cfg;
    std::vector<uint8_t> rec(c.rec_bytes);
    int rc = aie::be::sync_all(env);
    if (rc != AIE_OK) return rc;
    rc = aie::be::download(env, rec.data(), env->bufs.final + (size_t)e * c.rec_bytes, rec.size(), nullptr);
    if (rc != AIE_OK) return rc;
    rc = aie::be::sync(env, nullptr);
    if (rc != AIE_OK) return rc;
    aie_state_dump d = *out;
    d.book_rows = nullptr; d.book_count = nullptr;  // the snapshot holds the resident sections only
    aie::unpack_record(c, rec.data(), d);
    return AIE_OK;
}

int64_t aie_launch_count(const aie_env *env) { return env ? env->launches : 0; }

}  // extern "C"
