set -x
python tools/time_build.py 100 200 2>&1 | tee gpurun_out/time_build_r2f.txt
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -5) 2>&1 | tee gpurun_out/pytest_r2f.txt
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r2f.err | tee gpurun_out/bench_r2f.json | cut -c1-300
