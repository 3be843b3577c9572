"""Build the tuning variants of the CUDA library next to the default one (csrc/variants/*.so; nvcc cross-compiles here,
so no GPU time is spent compiling).  Each is the default source with one macro flipped; `AIE_LIB_PATH=<so>` makes the
Python binding load it (bench.py, pytest).  Usage: python tools/build_variants.py [name ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_economist_b200.build import CSRC, build_library  # noqa: E402

VARIANTS = {
    # leaner bit-plane writer: carried (plane, offset) and a running output pointer (lane-by-lane checked on the host by
    # tests/test_store_loops.py; ~30 instead of ~38 SASS instructions per 16-byte group)
    "planes_v2": ["-DAIE_PLANES_V2=1"],
    # MT19937 twist unroll depth (code size vs. ILP); the default 8 is the loop's full trip count
    "twist4": ["-DAIE_TWIST_UNROLL=4"],
    "planes_v2_twist4": ["-DAIE_PLANES_V2=1", "-DAIE_TWIST_UNROLL=4"],
    # the bench's random policy drawn inside the step kernel's observation pass (from the mask limits it has just staged)
    # instead of a separate 16 us sampler launch per step; aie_sample_random_actions only refreshes the seed
    "fused_policy": ["-DAIE_FUSED_POLICY=1"],
    "planes_v2_fused_policy": ["-DAIE_PLANES_V2=1", "-DAIE_FUSED_POLICY=1"],
}

if __name__ == "__main__":
    os.makedirs(os.path.join(CSRC, "variants"), exist_ok=True)
    for name in (sys.argv[1:] or list(VARIANTS)):
        out = build_library(extra_flags=VARIANTS[name], out=os.path.join("variants", "libaie_%s.so" % name))
        print(name, "->", os.path.join(CSRC, out))
