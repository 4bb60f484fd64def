#!/bin/bash
# One gpurun call's worth of evidence (1 GPU): the GPU suite, the default bench line, the ncu
# launch list of the same command, deformMesh through the C++ shim.  Everything lands in
# gpurun_out/; copy what should be judged into profiles/.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round_check.sh'
set -u
mkdir -p gpurun_out build
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline \
  --no-compression --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "ncu rc=$?"
g++ -std=c++17 -O2 -pthread -Iinclude scripts/bench_shim.cpp -Lkimera_pgmo_b200/lib -lpgmo_b200 \
  -Wl,-rpath,"$PWD/kimera_pgmo_b200/lib" -o build/bench_shim &&
  timeout 60 build/bench_shim > gpurun_out/bench_shim.json 2> gpurun_out/bench_shim.err
echo "bench_shim rc=$?"
cat gpurun_out/bench_shim.json
tail -c 600 gpurun_out/bench_default.json
