# two-GPU record of the final binary: the NCCL variants of the slab tests (+ the C-ABI driver) and the N = 2 bench line
set -x
python -m pytest tests/test_slab_gpu.py -m gpu -q -rP -k "nccl or c_abi" 2>&1 | grep -v "^$" | tail -15 | tee gpurun_out/r2f_nccl_2gpu.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_r2f_n2.err | tee gpurun_out/bench_r2f_n2.json | cut -c1-400
tail -5 gpurun_out/bench_r2f_n2.err
