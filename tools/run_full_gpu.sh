# Round-2 evidence run on one B200 (gpurun):  bash tools/run_full_gpu.sh [tag]
# tests, smoke, both bench arms (each plain), then the ncu launch list and one --set full capture of the top kernel
# (only after the same command has exited 0 without ncu).  Outputs: gpurun_out/*_<tag>.*
TAG=${1:-r2}
set -x
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; cut -c1-300 gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --trees 200000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --ref-wall 0"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain_${TAG}b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:validate_kernel -s 6 -c 2 -o gpurun_out/${TAG}_validate_full $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
