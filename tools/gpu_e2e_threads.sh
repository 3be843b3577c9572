for t in 16 24 32 48 64 96 32 48; do BENCH_ARGS="--e2e-threads $t" bash tools/gpu_e2e_chunks.sh c2 T=$t | grep "^c2"; done
