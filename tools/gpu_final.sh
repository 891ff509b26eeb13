#!/bin/bash
# Final sanity visit of a round: GPU tests, smoke(), and the randomised sweeps at the current HEAD.
OUT=gpurun_out/${1:-final}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python tools/fuzz_parity.py --cases 4000 --seed 31 > $OUT/fuzz_parity.json 2> $OUT/fuzz_parity.err; tail -1 $OUT/fuzz_parity.err
timeout 600 python tools/fuzz_misc.py --cases 40 --seed 3 > $OUT/fuzz_misc.json 2> $OUT/fuzz_misc.err; tail -1 $OUT/fuzz_misc.err
timeout 600 python tools/fuzz_knn.py --cases 150 --seed 2 > $OUT/fuzz_knn.json 2> $OUT/fuzz_knn.err; tail -1 $OUT/fuzz_knn.err
