// aie_compact.cuh — compacted device->host transfer of one step's outputs (aie_step_host_compact).
//
// The observation tensors are mostly 0/1-valued float32 planes (maps, masks) and small int16 indices; copied as they
// are they make the end-to-end step PCIe-bound (c2: 36 KB per env-step).  Here a pack pass rewrites each env's outputs
// as a compact record - planes and masks as bits, index planes as bytes, everything else verbatim - one D2H copy moves
// the compact records, and host threads expand them into the caller's tensors, which end up holding exactly the bytes
// the plain path (aie_step_host) delivers.  No simulation work happens on the host: this is a transfer format.
#pragma once
#include <stdint.h>
#include <string.h>

#include "aie_core.cuh"

namespace aie {

struct CompactLayout {
    // element counts per env
    int32_t n_a_map, n_a_mask, n_p_map, n_p_mask, n_a_idx, n_p_idx, n_a_flat, n_p_flat, n_p_agents, n_rew;
    // byte offsets inside an env's compact record (bit sections are uint32 words)
    int32_t off_a_map, off_a_mask, off_p_map, off_p_mask, off_a_idx, off_p_idx, off_a_flat, off_p_flat, off_p_agents,
        off_time, off_done, off_rew, bytes;
};

inline CompactLayout compact_layout(const DevCfg &c) {
    CompactLayout L;
    memset(&L, 0, sizeof(L));
    L.n_a_map = c.A * c.a_map_elems; L.n_a_mask = c.A * c.Na;
    L.n_p_map = c.planner_spatial ? c.M * c.HW : 0; L.n_p_mask = c.Np;
    L.n_a_idx = c.A * c.a_idx_elems; L.n_p_idx = c.planner_spatial ? 2 * c.HW : 0;
    L.n_a_flat = c.A * c.Fa; L.n_p_flat = c.Fp; L.n_p_agents = c.A * c.Fpa; L.n_rew = c.A + 1;
    int off = 0;
    auto words = [](int bits) { return 4 * ((bits + 31) / 32); };
    auto take = [&](int bytes) { int o = off; off += (bytes + 3) & ~3; return o; };
    L.off_a_map = take(words(L.n_a_map)); L.off_a_mask = take(words(L.n_a_mask));
    L.off_p_map = take(words(L.n_p_map)); L.off_p_mask = take(words(L.n_p_mask));
    L.off_a_idx = take(L.n_a_idx); L.off_p_idx = take(L.n_p_idx);
    L.off_a_flat = take(4 * L.n_a_flat); L.off_p_flat = take(4 * L.n_p_flat); L.off_p_agents = take(4 * L.n_p_agents);
    L.off_time = take(4); L.off_done = take(4);
    off = (off + 7) & ~7;
    L.off_rew = take(8 * L.n_rew);
    L.bytes = (off + 15) & ~15;
    return L;
}

// ---- device side: one warp packs one env -----------------------------------------------------------------------
AIE_DEV void pack_bits(const float *src, int n, uint32_t *dst, int lane) {
#if AIE_ON_DEVICE
    for (int w0 = 0; w0 < n; w0 += 32) {
        const int i = w0 + lane;
        const uint32_t b = wballot(i < n && src[i] != 0.0f);
        if (lane == 0) dst[w0 >> 5] = b;
    }
#else
    for (int w0 = lane * 32; w0 < n; w0 += 32 * NL) {
        uint32_t b = 0;
        for (int j = 0; j < 32 && w0 + j < n; j++) b |= (src[w0 + j] != 0.0f ? 1u : 0u) << j;
        dst[w0 >> 5] = b;
    }
#endif
}
AIE_DEV void pack_env(const DevCfg &c, const DevBufs &b, const CompactLayout &L, size_t env, uint8_t *dst, int lane) {
    pack_bits(b.a_map + env * L.n_a_map, L.n_a_map, (uint32_t *)(dst + L.off_a_map), lane);
    pack_bits(b.a_mask + env * L.n_a_mask, L.n_a_mask, (uint32_t *)(dst + L.off_a_mask), lane);
    if (L.n_p_map) pack_bits(b.p_map + env * L.n_p_map, L.n_p_map, (uint32_t *)(dst + L.off_p_map), lane);
    pack_bits(b.p_mask + env * L.n_p_mask, L.n_p_mask, (uint32_t *)(dst + L.off_p_mask), lane);
    {
        const int16_t *s = b.a_idx + env * L.n_a_idx;
        for (int i = lane; i < L.n_a_idx; i += NL) dst[L.off_a_idx + i] = (uint8_t)s[i];   // indices are 0 .. A + 1 <= 65
    }
    if (L.n_p_idx) {
        const int16_t *s = b.p_idx + env * L.n_p_idx;
        for (int i = lane; i < L.n_p_idx; i += NL) dst[L.off_p_idx + i] = (uint8_t)s[i];
    }
    {
        const float *s = b.a_flat + env * L.n_a_flat; float *d = (float *)(dst + L.off_a_flat);
        for (int i = lane; i < L.n_a_flat; i += NL) d[i] = s[i];
    }
    {
        const float *s = b.p_flat + env * L.n_p_flat; float *d = (float *)(dst + L.off_p_flat);
        for (int i = lane; i < L.n_p_flat; i += NL) d[i] = s[i];
    }
    if (L.n_p_agents) {
        const float *s = b.p_agents + env * L.n_p_agents; float *d = (float *)(dst + L.off_p_agents);
        for (int i = lane; i < L.n_p_agents; i += NL) d[i] = s[i];
    }
    {
        const double *s = b.rew + env * L.n_rew; double *d = (double *)(dst + L.off_rew);
        for (int i = lane; i < L.n_rew; i += NL) d[i] = s[i];
    }
    if (lane == 0) {
        *(float *)(dst + L.off_time) = b.time_obs[env];
        *(int32_t *)(dst + L.off_done) = b.done[env];
    }
}

}  // namespace aie
