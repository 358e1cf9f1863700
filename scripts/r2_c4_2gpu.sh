# periodic box (C4 at 128^3) on one and on two GPUs: the state_crc must agree (sharded warp-shared search with the periodic fallback)
set -x
timeout 300 python bench.py --config C4 --n 2097152 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2w_C4_128_1gpu.json 2>gpurun_out/r2w_C4_128_1gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --config C4 --n 2097152 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2w_C4_128_2gpu.json 2>gpurun_out/r2w_C4_128_2gpu.err
B200_PERIODIC_SEARCH_FROM_ROOT=1 timeout 300 python bench.py --config C4 --n 2097152 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2w_C4_128_1gpu_root.json 2>gpurun_out/r2w_C4_128_1gpu_root.err
for f in gpurun_out/r2w_C4_128_*.json; do echo $f; grep -o "\"ms_per_step\": [0-9.]*\|state_crc[^,}]*\|\"sidm_ms\": [0-9.]*\|\"ensure_ms\": [0-9.]*" $f | tr '\n' ' '; echo; done
tail -3 gpurun_out/r2w_C4_128_2gpu.err
