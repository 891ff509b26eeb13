bash tools/gpu_cmd.sh r03a \
 'python tests/golden/make_golden_knn.py' \
 'timeout 600 python -m pytest tests/test_knn_gpu.py -x -q' \
 'timeout 300 python tools/bench_knn.py > $OUT/knn.json' \
 'timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_knn_gpu.py' \
 'timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench_ours.json' \
 'timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/bench_reference.json'
