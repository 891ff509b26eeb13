bash tools/gpu_cmd.sh r03e \
 'timeout 600 python -m pytest tests/test_densify_gpu.py tests/test_optim_gpu.py -x -q'
