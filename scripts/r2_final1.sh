# final one-GPU measurement pass of round 2 (every output under gpurun_out/, copied to profiles/ afterwards)
set -x
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/r2z_pytest_gpu.log 2>&1; tail -2 gpurun_out/r2z_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2z_C3.json 2>gpurun_out/r2z_C3.err; tail -c 400 gpurun_out/r2z_C3.err
for c in C1 C2 C3S C5; do timeout 400 python bench.py --config $c > gpurun_out/r2z_$c.json 2>gpurun_out/r2z_$c.err; done
timeout 400 python bench.py --config C5 --tree-reuse 4 --no-cpu > gpurun_out/r2z_C5_reuse4.json 2>>gpurun_out/r2z_C5.err
timeout 600 python bench.py --config C4 --steps 3 --warmup 3 > gpurun_out/r2z_C4.json 2>gpurun_out/r2z_C4.err
for f in gpurun_out/r2z_C*.json; do echo $f; grep -o "\"ms_per_step\": [0-9.]*" $f | head -2 | tr '\n' ' '; echo; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2z_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 4 -c 1 -f -o gpurun_out/walk_r2z python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2z_ncu_full.log 2>&1
ls -la gpurun_out/walk_r2z.ncu-rep
