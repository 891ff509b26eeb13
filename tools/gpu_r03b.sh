bash tools/gpu_cmd.sh r03b \
 'timeout 900 python -m pytest tests/test_knn_gpu.py tests/test_dsm_gpu.py -x -q' \
 'timeout 300 python tools/bench_knn.py > $OUT/knn.json'
