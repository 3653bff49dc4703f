#!/bin/bash
# Round artifacts on the GPU box (run through gpurun from the repo root, 1 GPU): GPU test suite, smoke, bench lines, aux benches,
# ncu launch lists and --set full captures of the dominant kernels.  Everything lands in gpurun_out/ (tag = $1).
mkdir -p gpurun_out
T=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu_$T.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 6 --warmup 2 > gpurun_out/bench_ref_$T.json 2> gpurun_out/bench_ref_$T.err; echo "bench ref rc=$?"
timeout 300 python tools/bench_aux.py > gpurun_out/bench_aux_$T.json 2> gpurun_out/bench_aux_$T.err
timeout 300 python tools/bench_fit.py --steps 300 > gpurun_out/fit_$T.json 2>/dev/null
timeout 300 python tools/bench_joint.py --steps 100 > gpurun_out/joint_$T.json 2>/dev/null
timeout 100 python tools/mc_time.py > gpurun_out/mc_time_$T.txt 2>&1
# launch lists (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-stock-gpu > gpurun_out/launches_$T.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_joint_$T.csv \
    python tools/bench_joint.py --steps 3 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_fit_$T.csv \
    python tools/bench_fit.py --steps 4 > /dev/null 2>&1
# full captures: dense ensemble kernel (one launch of the bench workload), the three marching-cubes kernels, the generic linear layer
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ensemble_tc_kernel_v8 -s 3 -c 1 -f -o gpurun_out/prof_tc_$T \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-stock-gpu > gpurun_out/prof_tc_$T.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:mc_ -s 3 -c 3 -f -o gpurun_out/prof_mc_$T python tools/mc_time.py > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:linear_tc_kernel -s 40 -c 3 -f -o gpurun_out/prof_lin_$T \
    python tools/bench_joint.py --steps 2 > /dev/null 2>&1
tail -3 gpurun_out/pytest_gpu_$T.log; cat gpurun_out/bench_$T.json | cut -c1-600
