bash tools/gpu_cmd.sh r03f \
 'timeout 900 python tools/fuzz_parity.py --cases 1500 --seed 1 > $OUT/fuzz.json'
