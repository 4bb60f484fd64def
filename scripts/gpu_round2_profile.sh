#!/bin/bash
# Round-2 evidence in one gpurun call (1 GPU): GPU suite, default bench line, ncu launch list of
# the deformation step, `ncu --set full` of the blend-only kernel (K3b) and of the search kernel
# (K3a), raw pages exported on the box.
#   gpurun --timeout 1200 -- 'bash scripts/gpu_round2_profile.sh'
set -u
mkdir -p gpurun_out build
timeout 500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench rc=$?"
B="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-compression --no-e2e --no-tie-sweep --no-config3 --no-config4"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/r02_launches.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deformBlendCached \
  -s 6 -c 2 -f -o gpurun_out/r02_blend $B > gpurun_out/ncu_blend.log 2>&1
echo "ncu blend rc=$?"
ncu -i gpurun_out/r02_blend.ncu-rep --page raw --csv > gpurun_out/r02_blend_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:deformSearchSorted \
  -s 4 -c 2 -f -o gpurun_out/r02_search $B > gpurun_out/ncu_search.log 2>&1
echo "ncu search rc=$?"
ncu -i gpurun_out/r02_search.ncu-rep --page raw --csv > gpurun_out/r02_search_raw.csv 2>/dev/null
tail -c 400 gpurun_out/bench_default.json
