set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r1b.log 2>&1; tail -3 gpurun_out/pytest_gpu_r1b.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1b.log 2>&1; tail -2 gpurun_out/smoke_r1b.log
python bench.py > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; cut -c1-400 gpurun_out/bench_r1b.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1b.json 2> gpurun_out/bench_ref_r1b.err; cut -c1-300 gpurun_out/bench_ref_r1b.json
CMD="python bench.py --trees 200000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_r1b.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_launches.csv $CMD > gpurun_out/ncu_launch_r1b.log 2>&1
$CMD > gpurun_out/plain_r1b2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 3 -c 1 -o gpurun_out/r1b_validate_full $CMD > gpurun_out/ncu_full_r1b.log 2>&1
tail -2 gpurun_out/ncu_full_r1b.log
