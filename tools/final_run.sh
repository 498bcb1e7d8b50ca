# final single-GPU record of a round: smoke, GPU tests, ncu launch list, ncu --set full capture of one step (-> DRAM traffic per
# kernel, keyed by the source hash), then the reference arm and the bench, so that the bench line quotes the traffic of THIS binary
set -x
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
(time python -m pytest tests -m gpu -x -q 2>&1 | tail -5) 2>&1 | tee gpurun_out/pytest_r2_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline --no-f64 > gpurun_out/bench_under_ncu_r2_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bin|k_rows|k_twin|k_sweep_n3|k_force_finish|k_zero" -s 14 -c 8 -f -o gpurun_out/prof_r2_step python tools/prof_c2.py 100 f32 4 > gpurun_out/prof_r2_step.log 2>&1
tail -2 gpurun_out/prof_r2_step.log
python tools/traffic_from_ncu.py gpurun_out/prof_r2_step.ncu-rep > /dev/null && cp profiles/r2_traffic.json gpurun_out/r2_traffic.json
python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_r2_final_ref.err | tee gpurun_out/bench_r2_final_ref.json | cut -c1-400
python bench.py --steps 20 --warmup 5 2>gpurun_out/bench_r2_final.err | tee gpurun_out/bench_r2_final.json | cut -c1-300
