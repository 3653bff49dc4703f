#!/bin/bash
# Round artifacts on the GPU box (run through gpurun from the repo root): full GPU test suite, smoke, bench lines,
# ncu launch list of the bench command and one --set full capture of the dominant kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
T=${1:-final}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
timeout 300 python tools/bench_aux.py > gpurun_out/bench_aux_$T.json 2> gpurun_out/bench_aux_$T.err
timeout 300 python tools/bench_fit.py --steps 300 > gpurun_out/fit_$T.json 2>/dev/null
timeout 300 python tools/bench_joint.py --steps 40 > gpurun_out/joint_$T.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensemble_tc_kernel -c 1 -f -o gpurun_out/prof_tc_$T \
    python bench.py --steps 1 --warmup 0 --res 256 --no-cpu-baseline > gpurun_out/prof_tc_$T.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_fit_$T.csv \
    python tools/bench_fit.py --steps 4 > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu_$T.log; cat gpurun_out/bench_$T.json; cat gpurun_out/bench_aux_$T.json gpurun_out/fit_$T.json gpurun_out/joint_$T.json
