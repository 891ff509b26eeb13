bash tools/gpu_cmd.sh r03g \
 'timeout 900 python tools/fuzz_check.py tools/fuzz_r03f.json > $OUT/fuzz_check.json'
