OUT=gpurun_out/r03o; mkdir -p $OUT
timeout 1500 python tools/fuzz_misc.py --cases 80 --seed 1 > $OUT/fuzz_misc.json 2> $OUT/fuzz_misc.err; tail -25 $OUT/fuzz_misc.err | cut -c1-600
