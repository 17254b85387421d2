mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core_filter" 2>&1 | tail -3
timeout 120 python tools/profile_step.py --steps 20 2>&1 | grep -E "filter|total"
timeout 120 python tools/profile_step.py --steps 20 --variant tiny 2>&1 | grep -E "filter|total"
timeout 120 python tools/profile_step.py --steps 20 --variant ultra_tiny 2>&1 | grep -E "filter|total"
(cd mlff_distiller_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -DMLFFD_FILTER_TIMING -o libmlffd.so mlffd.cu)
timeout 120 python tools/profile_filter.py tc 2>&1 | tail -3
