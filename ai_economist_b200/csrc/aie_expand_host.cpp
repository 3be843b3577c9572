// aie_expand_host.cpp — host side of the compacted transfer (aie_step_host_compact): bits -> 0.0f / 1.0f.
// Plain C++ (no CUDA): compiled by the host compiler so that the AVX-512 path can be built with a function-level target
// attribute and picked at run time.  The expansion is a pure write stream (c2: 296 MB per step at 8 192 envs), so the
// groups go out with non-temporal stores (no read-for-ownership of the destination lines): whole 64-byte lines from a
// 16-bit mask with AVX-512, 16-byte groups from a 16-entry nibble table with SSE2.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#if defined(__linux__)
#include <linux/futex.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#endif
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


// ---- NUMA placement of the expansion (Linux; everything degrades to "unknown node" elsewhere) ------------------------
// The expansion is bound by the host's memory controllers: a thread writing to the other socket's memory goes through the
// inter-socket link.  The pool therefore pins its workers to nodes and hands each work item to the node its destination
// pages live on (get_mempolicy(MPOL_F_NODE | MPOL_F_ADDR), no libnuma needed).
int numa_node_of(const void *addr) {
#if defined(__linux__) && defined(SYS_get_mempolicy)
    int node = -1;
    if (syscall(SYS_get_mempolicy, &node, nullptr, 0UL, const_cast<void *>(addr), 3UL /* MPOL_F_NODE | MPOL_F_ADDR */) == 0) return node;
#endif
    (void)addr;
    return -1;
}
int numa_node_count() {
#if defined(__linux__)
    int n = 0;
    for (; n < 64; n++) {
        char path[96];
        snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", n);
        if (access(path, R_OK) != 0) break;
    }
    return n > 0 ? n : 1;
#else
    return 1;
#endif
}
bool pin_current_thread_to_node(int node) {
#if defined(__linux__)
    char path[96];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[1024];
    const bool ok = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    int any = 0;
    for (char *p = buf; *p;) {   // "0-31,64-95"
        char *end;
        long a = strtol(p, &end, 10);
        if (end == p) break;
        long b = a;
        if (*end == '-') { p = end + 1; b = strtol(p, &end, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, &set); any = 1; }
        p = (*end == ',') ? end + 1 : end;
        if (*end != ',') break;
    }
    return any && sched_setaffinity(0, sizeof(set), &set) == 0;
#else
    (void)node;
    return false;
#endif
}

// Block while *flag == 0 / wake every thread blocked on it (futex(2): sleeping waiters cost no CPU time, which matters when
// the process runs under a CPU quota; elsewhere: yield loop).
void flag_wait_zero(std::atomic<int> *flag) {
    static_assert(sizeof(std::atomic<int>) == sizeof(int), "futex word");
    while (flag->load(std::memory_order_acquire) == 0) {
#if defined(__linux__) && defined(SYS_futex)
        syscall(SYS_futex, reinterpret_cast<int *>(flag), FUTEX_WAIT_PRIVATE, 0, nullptr, nullptr, 0);
#else
        std::this_thread::yield();
#endif
    }
}
void flag_wake_all(std::atomic<int> *flag) {
#if defined(__linux__) && defined(SYS_futex)
    syscall(SYS_futex, reinterpret_cast<int *>(flag), FUTEX_WAKE_PRIVATE, 0x7fffffff, nullptr, nullptr, 0);
#else
    (void)flag;
#endif
}

// NUMA node a PCI device hangs off ("0000:3b:00.0" as cudaDeviceGetPCIBusId prints it), -1 when the kernel does not say.
int pci_numa_node(const char *bus_id) {
#if defined(__linux__)
    char path[160], low[64];
    size_t n = 0;
    for (; bus_id[n] && n + 1 < sizeof(low); n++) low[n] = (char)((bus_id[n] >= 'A' && bus_id[n] <= 'F') ? bus_id[n] + 32 : bus_id[n]);
    low[n] = 0;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", low);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
#else
    (void)bus_id;
    return -1;
#endif
}
// Page-aligned anonymous memory bound to one NUMA node (mbind(2) before first touch); nullptr when that is not possible.
void *alloc_on_node(size_t bytes, int node) {
#if defined(__linux__) && defined(SYS_mbind)
    if (node < 0 || node >= 64 || bytes == 0) return nullptr;
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) return nullptr;
    unsigned long mask = 1UL << node;
    if (syscall(SYS_mbind, p, (unsigned long)bytes, 2 /* MPOL_BIND */, &mask, 65UL, 0U) != 0) { munmap(p, bytes); return nullptr; }
    memset(p, 0, bytes);   // first touch
    return p;
#else
    (void)bytes; (void)node;
    return nullptr;
#endif
}
void free_on_node(void *p, size_t bytes) {
#if defined(__linux__)
    if (p) munmap(p, bytes);
#else
    (void)p; (void)bytes;
#endif
}

}  // namespace aie
