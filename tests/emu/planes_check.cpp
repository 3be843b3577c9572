// TEST INFRASTRUCTURE: walks the per-lane store loops of csrc/aie_core.cuh (store_bitplanes_f32, store_bytes_i16,
// store_run_f32, store_rows_f32) with NL = 32, lane by lane on the host, and compares every output run with a plain
// element-wise computation - for the default build and for the tuning variants (-DAIE_PLANES_V2=1).  These functions
// use no warp collectives, so running the 32 lanes one after the other is exact.
#define AIE_EMU 1
#define AIE_EMU_NL 32
#include "../../ai_economist_b200/csrc/aie_core.cuh"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace aie;

static uint32_t magic(int n) { return n > 1 ? (uint32_t)((1ull << 32) / (uint64_t)n) + 1u : 0u; }

int main() {
    int fails = 0, cases = 0;
    uint8_t pbits[8] = {1, 2, 32, 16, 4, 8, 0x40, 0};
    srand(7);
    const int ns[] = {1, 3, 4, 9, 25, 49, 64, 65, 121, 127, 128, 129, 225, 361, 444, 625, 1600, 4096};
    for (int n : ns)
        for (int np = 1; np <= 7; np++)
            for (int off = 0; off < 4; off++) {   // alignment of the run inside its 16-byte group
                std::vector<uint8_t> bytes((n + 3) / 4 * 4 + 16);
                for (auto &b : bytes) b = (uint8_t)(rand() & 0x7F);
                std::vector<float> out(np * n + 16, -1.0f), ref(np * n + 16, -1.0f);
                float *dst = out.data();
                while (((uintptr_t)dst & 15) != 0) dst++;
                dst += off;
                float *rdst = ref.data() + (dst - out.data());
                for (int lane = 0; lane < NL; lane++) store_bitplanes_f32(dst, np, n, magic(n), bytes.data(), pbits, lane);
                for (int m = 0; m < np; m++)
                    for (int i = 0; i < n; i++) rdst[m * n + i] = (bytes[i] & pbits[m]) ? 1.0f : 0.0f;
                cases++;
                if (out != ref) { fails++; printf("bitplanes n=%d np=%d off=%d differ\n", n, np, off); }
                // int16 index planes
                std::vector<int16_t> o16(n + 16, -7), r16(n + 16, -7);
                int16_t *d16 = o16.data();
                while (((uintptr_t)d16 & 15) != 0) d16++;
                d16 += 2 * off + 1;
                int16_t *rd16 = r16.data() + (d16 - o16.data());
                for (int lane = 0; lane < NL; lane++) store_bytes_i16<false>(d16, n, bytes.data(), lane);
                for (int i = 0; i < n; i++) rd16[i] = bytes[i];
                cases++;
                if (o16 != r16) { fails++; printf("bytes_i16 n=%d off=%d differ\n", n, off); }
                // flat runs / row matrices
                std::vector<float> of(np * n + 16, -1.0f), rf(np * n + 16, -1.0f);
                float *df = of.data();
                while (((uintptr_t)df & 15) != 0) df++;
                df += off;
                float *rdf = rf.data() + (df - of.data());
                for (int lane = 0; lane < NL; lane++)
                    store_rows_f32(df, np, n, magic(n), lane, [&](int a, int j) { return (float)(a * 1000 + j); });
                for (int a = 0; a < np; a++) for (int j = 0; j < n; j++) rdf[a * n + j] = (float)(a * 1000 + j);
                cases++;
                if (of != rf) { fails++; printf("rows n=%d np=%d off=%d differ\n", n, np, off); }
            }
    printf("%d cases, %d failures (AIE_PLANES_V2=%d)\n", cases, fails, AIE_PLANES_V2);
    return fails ? 1 : 0;
}
