set -x
mkdir -p gpurun_out
python scripts/e2e_breakdown.py > gpurun_out/r02h_e2e.log 2>&1
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "step_host or collector or checkpoint" 2>&1 | tail -3 > gpurun_out/r02h_tests.log
bash scripts/prof_kernel.sh r02h_D step_kernel 13 --workload D
