OUT=gpurun_out/r03l; mkdir -p $OUT
timeout 900 python tools/fuzz_knn.py --cases 400 > $OUT/fuzz_knn.json 2> $OUT/fuzz_knn.err; tail -3 $OUT/fuzz_knn.err
timeout 600 python tools/stress_fused.py 60 2>&1 | tail -5
