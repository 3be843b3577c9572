// aie_compact.cuh — compacted device->host transfer of one step's outputs (aie_step_host_compact).
//
// The observation tensors are mostly 0/1-valued float32 planes (maps, masks) and small int16 indices; copied as they
// are they make the end-to-end step PCIe-bound (c2: 36 KB per env-step).  Here a pack pass rewrites each env's outputs
// as a compact record - planes and masks as bits; the (mostly zero) index planes as a bitmap of their non-zero elements
// plus those values as bytes; the agents' flat vectors de-duplicated by entry class (entries every agent of an env shares
// once, per-agent scalars, open-order counts as small integers); the rest verbatim - one D2H copy moves the compact
// records, and host threads expand them into the caller's tensors, which end up holding exactly the bytes the plain path
// (aie_step_host) delivers.  No simulation work happens on the host: this is a transfer format (c2: 2.6 KB per env-step).
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "aie_core.cuh"

namespace aie {

struct CompactLayout {
    // element counts per env
    int32_t n_a_map, n_a_mask, n_p_map, n_p_mask, n_a_idx, n_p_idx, n_a_flat, n_p_flat, n_p_agents, n_rew;
    // Index planes (int16, mostly 0): a bitmap of the non-zero elements + their values as bytes, at most cap_* of them;
    // the count is stored too - an env with more non-zero elements than the capacity is fetched directly (rare).
    int32_t cap_a_idx, cap_p_idx;
    // The agents' flat vectors [A][Fa] by class of their program entries (aie_layout.h: FK_*): entries shared by all
    // agents of an env once (n_sh floats), per-agent scalars (A x n_ag floats), open-order counts (A x n_cnt small
    // integers as uint8, or uint16 when (A - 1) * K can exceed 255)
    int32_t A, Fa, n_sh, n_ag, n_cnt, cnt_bytes;
    // byte offsets inside an env's compact record (bit sections are uint32 words)
    int32_t off_a_map, off_a_mask, off_p_map, off_p_mask, off_a_idx_mask, off_a_idx_vals, off_p_idx_mask, off_p_idx_vals,
        off_idx_cnt /* int32[2] */, off_f_sh, off_f_ag, off_f_cnt, off_p_flat, off_p_agents, off_time, off_done, off_rew, bytes;
};

inline CompactLayout compact_layout(const DevCfg &c) {
    CompactLayout L;
    memset(&L, 0, sizeof(L));
    L.n_a_map = c.A * c.a_map_elems; L.n_a_mask = c.A * c.Na;
    L.n_p_map = c.planner_spatial ? c.M * c.HW : 0; L.n_p_mask = c.Np;
    L.n_a_idx = c.A * c.a_idx_elems; L.n_p_idx = c.planner_spatial ? 2 * c.HW : 0;
    L.n_a_flat = c.A * c.Fa; L.n_p_flat = c.Fp; L.n_p_agents = c.A * c.Fpa; L.n_rew = c.A + 1;
    L.cap_a_idx = (L.n_a_idx / 4 + 31) & ~3; L.cap_p_idx = L.n_p_idx ? ((L.n_p_idx / 8 + 31) & ~3) : 0;
    if (getenv("AIE_COMPACT_TINY_CAPS")) { L.cap_a_idx = 4; L.cap_p_idx = L.n_p_idx ? 4 : 0; }   // test knob: force the overflow path
    L.A = c.A; L.Fa = c.Fa; L.n_sh = c.cf_n_sh; L.n_ag = c.cf_n_ag; L.n_cnt = c.cf_n_cnt;
    L.cnt_bytes = ((c.A - 1) * c.K > 255) ? 2 : 1;
    int off = 0;
    auto words = [](int bits) { return 4 * ((bits + 31) / 32); };
    auto take = [&](int bytes) { int o = off; off += (bytes + 3) & ~3; return o; };
    L.off_a_map = take(words(L.n_a_map)); L.off_a_mask = take(words(L.n_a_mask));
    L.off_p_map = take(words(L.n_p_map)); L.off_p_mask = take(words(L.n_p_mask));
    L.off_a_idx_mask = take(words(L.n_a_idx)); L.off_a_idx_vals = take(L.cap_a_idx);
    L.off_p_idx_mask = take(words(L.n_p_idx)); L.off_p_idx_vals = take(L.cap_p_idx);
    L.off_idx_cnt = take(8);
    L.off_f_sh = take(4 * L.n_sh); L.off_f_ag = take(4 * c.A * L.n_ag); L.off_f_cnt = take(L.cnt_bytes * c.A * L.n_cnt);
    L.off_p_flat = take(4 * L.n_p_flat); L.off_p_agents = take(4 * L.n_p_agents);
    L.off_time = take(4); L.off_done = take(4);
    off = (off + 7) & ~7;
    L.off_rew = take(8 * L.n_rew);
    L.bytes = (off + 15) & ~15;
    return L;
}

// ---- device side: one warp packs one env -----------------------------------------------------------------------
AIE_DEV void pack_bits(const float *src, int n, uint32_t *dst, int lane) {
#if AIE_ON_DEVICE
    // eight independent loads in flight per lane (the pass is latency-bound otherwise: one 128-byte row per round trip)
    int w0 = 0;
    for (; w0 + 256 <= n; w0 += 256) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = src[w0 + 32 * j + lane];
        uint32_t mine = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { const uint32_t b = wballot(v[j] != 0.0f); if (lane == j) mine = b; }
        if (lane < 8) dst[(w0 >> 5) + lane] = mine;
    }
    for (; w0 < n; w0 += 32) {
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
// non-zero bitmap + byte values (in element order, the first `cap` of them); returns the number of non-zero elements
AIE_DEV int pack_sparse_u8(const int16_t *src, int n, uint32_t *mask, uint8_t *vals, int cap, int lane) {
    int running = 0;
#if AIE_ON_DEVICE
    int w0 = 0;
    for (; w0 + 128 <= n; w0 += 128) {   // four rows in flight
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = (int)src[w0 + 32 * j + lane];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t b = wballot(v[j] != 0);
            if (lane == 0) mask[(w0 >> 5) + j] = b;
            const int pos = running + __popc(b & ((1u << lane) - 1u));
            if (v[j] != 0 && pos < cap) vals[pos] = (uint8_t)v[j];
            running += __popc(b);
        }
    }
    for (; w0 < n; w0 += 32) {
        const int i = w0 + lane;
        const int v = i < n ? (int)src[i] : 0;
        const uint32_t b = wballot(v != 0);
        if (lane == 0) mask[w0 >> 5] = b;
        const int pos = running + __popc(b & ((1u << lane) - 1u));
        if (v != 0 && pos < cap) vals[pos] = (uint8_t)v;   // indices are 0 .. A + 1 <= 65
        running += __popc(b);
    }
#else
    (void)lane;
    for (int w0 = 0; w0 < n; w0 += 32) {
        uint32_t b = 0;
        for (int j = 0; j < 32 && w0 + j < n; j++)
            if (src[w0 + j] != 0) { b |= 1u << j; if (running < cap) vals[running] = (uint8_t)src[w0 + j]; running++; }
        mask[w0 >> 5] = b;
    }
#endif
    return running;
}
AIE_DEV void pack_env(const DevCfg &c, const DevBufs &b, const CompactLayout &L, size_t env, uint8_t *dst, int lane) {
    pack_bits(b.a_map + env * L.n_a_map, L.n_a_map, (uint32_t *)(dst + L.off_a_map), lane);
    pack_bits(b.a_mask + env * L.n_a_mask, L.n_a_mask, (uint32_t *)(dst + L.off_a_mask), lane);
    if (L.n_p_map) pack_bits(b.p_map + env * L.n_p_map, L.n_p_map, (uint32_t *)(dst + L.off_p_map), lane);
    pack_bits(b.p_mask + env * L.n_p_mask, L.n_p_mask, (uint32_t *)(dst + L.off_p_mask), lane);
    {
        const int na = pack_sparse_u8(b.a_idx + env * L.n_a_idx, L.n_a_idx, (uint32_t *)(dst + L.off_a_idx_mask), dst + L.off_a_idx_vals,
                                      L.cap_a_idx, lane);
        const int np = L.n_p_idx ? pack_sparse_u8(b.p_idx + env * L.n_p_idx, L.n_p_idx, (uint32_t *)(dst + L.off_p_idx_mask),
                                                  dst + L.off_p_idx_vals, L.cap_p_idx, lane) : 0;
        if (lane == 0) { ((int32_t *)(dst + L.off_idx_cnt))[0] = na; ((int32_t *)(dst + L.off_idx_cnt))[1] = np; }
    }
    {   // agents' flat vectors by entry class
        const float *s = b.a_flat + env * L.n_a_flat;
        const uint16_t *prog = b.tab, *cslot = b.tab + c.tab_cslot;
        float *f_sh = (float *)(dst + L.off_f_sh), *f_ag = (float *)(dst + L.off_f_ag);
        uint8_t *cnt8 = dst + L.off_f_cnt; uint16_t *cnt16 = (uint16_t *)(dst + L.off_f_cnt);
        for (int x = lane; x < L.n_a_flat; x += NL) {
            const int a = x / L.Fa, j = x - a * L.Fa;
            const int kind = AIE_FLAT_KIND(prog[j]), slot = cslot[j];
            const float v = s[x];
            if (kind == FK_SHARED) { if (a == 0) f_sh[slot] = v; }
            else if (kind == FK_AGENT) f_ag[a * L.n_ag + slot] = v;
            else if (L.cnt_bytes == 1) cnt8[a * L.n_cnt + slot] = (uint8_t)v;
            else cnt16[a * L.n_cnt + slot] = (uint16_t)v;
        }
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
