// aie_expand_host.cpp — host side of the compacted transfer (aie_step_host_compact): bits -> 0.0f / 1.0f.
// Plain C++ (no CUDA): compiled by the host compiler so that the AVX-512 path can be built with a function-level target
// attribute and picked at run time.  The expansion is a pure write stream (c2: 296 MB per step at 8 192 envs), so the
// groups go out with non-temporal stores (no read-for-ownership of the destination lines): whole 64-byte lines from a
// 16-bit mask with AVX-512, 16-byte groups from a 16-entry nibble table with SSE2.
#include <stdint.h>
#include <string.h>
#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#endif

namespace aie {

static inline float bit_at(const uint32_t *src, int i) { return (float)((src[i >> 5] >> (i & 31)) & 1u); }

#if defined(__x86_64__) || defined(__i386__)

static void expand_bits_sse2(const uint32_t *src, int n, float *dst) {
    static const struct Table {
        __m128 v[16];
        Table() { for (int k = 0; k < 16; k++) v[k] = _mm_set_ps((float)((k >> 3) & 1), (float)((k >> 2) & 1), (float)((k >> 1) & 1), (float)(k & 1)); }
    } T;
    int i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 15)) { dst[i] = bit_at(src, i); i++; }
    for (; i + 4 <= n; i += 4) {
        uint64_t w;
        memcpy(&w, (const uint8_t *)src + 4 * (i >> 5), 8);     // the word holding bit i and the next one
        _mm_stream_ps(dst + i, T.v[(w >> (i & 31)) & 15u]);
    }
    for (; i < n; i++) dst[i] = bit_at(src, i);
}

__attribute__((target("avx512f"))) static void expand_bits_avx512(const uint32_t *src, int n, float *dst) {
    int i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63)) { dst[i] = bit_at(src, i); i++; }
    const __m512 one = _mm512_set1_ps(1.0f);
    for (; i + 16 <= n; i += 16) {
        uint64_t w;
        memcpy(&w, (const uint8_t *)src + 4 * (i >> 5), 8);     // bits i .. i + 15 sit inside these 64 bits
        _mm512_stream_ps(dst + i, _mm512_maskz_mov_ps((__mmask16)(w >> (i & 31)), one));
    }
    for (; i < n; i++) dst[i] = bit_at(src, i);
}

typedef void (*expand_fn)(const uint32_t *, int, float *);
static expand_fn pick() {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") ? expand_bits_avx512 : expand_bits_sse2;
}
// `src` must be readable up to the 8 bytes holding the last bit (true inside a compact record)
void expand_bits(const uint32_t *src, int n, float *dst) {
    static const expand_fn fn = pick();
    fn(src, n, dst);
}
void expand_fence() { _mm_sfence(); }
const char *expand_isa() { __builtin_cpu_init(); return __builtin_cpu_supports("avx512f") ? "avx512f" : "sse2"; }

#else
void expand_bits(const uint32_t *src, int n, float *dst) { for (int i = 0; i < n; i++) dst[i] = bit_at(src, i); }
void expand_fence() {}
const char *expand_isa() { return "scalar"; }
#endif

}  // namespace aie
