set -x
mkdir -p gpurun_out
for i in 1 2; do
(cd variants/old_tree && python bench.py --workload C_both --steps 128 --warmup 16 --no-cpu) > gpurun_out/r02n_old_C_both_$i.json 2>/dev/null
python bench.py --workload C_both --steps 128 --warmup 16 --no-cpu --no-secondary > gpurun_out/r02n_new_C_both_$i.json 2>/dev/null
done
(cd variants/old_tree && python bench.py --workload C_beam --steps 128 --warmup 16 --no-cpu) > gpurun_out/r02n_old_C_beam.json 2>/dev/null
python bench.py --workload C_beam --steps 128 --warmup 16 --no-cpu --no-secondary > gpurun_out/r02n_new_C_beam.json 2>/dev/null
