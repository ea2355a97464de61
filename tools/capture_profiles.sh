#!/bin/bash
# Re-creates the evidence under profiles/ on a B200 box:  gpurun --timeout 1800 -- bash tools/capture_profiles.sh
# (outputs land in gpurun_out/cap/; summarise with tools/ncu_summary.py / ncu_lines.py / launch_summary.py)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/cap
# 1. launch list of the bench command (short)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/cap/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/cap/ncu_launches.log 2>&1
# 2. full capture of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:sweep_packed -s 4 -c 1 -o gpurun_out/cap/r02_sweep_packed python bench.py --steps 2 --warmup 3 --no-side --no-cpu-baseline > gpurun_out/cap/ncu_full.log 2>&1
# 3. MSLR-shaped launch
N=3771125 Q=31531 STEPS=1 ncu --set full --clock-control none --import-source on -k regex:sweep_packed -s 2 -c 1 -o gpurun_out/cap/r02_sweep_packed_mslr python tools/bench_sweep.py "" > gpurun_out/cap/ncu_mslr.log 2>&1
# 4. micro-benchmarks
for m in cmp_throughput:cmp read_pattern:read_pattern; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/micro/${m#*:} tools/micro/${m%:*}.cu
done
./tools/micro/cmp > gpurun_out/cap/r02_micro_compare_throughput.txt 2>&1
./tools/micro/read_pattern > gpurun_out/cap/r02_micro_read_pattern.txt 2>&1
# 5. evaluate bench, both load paths
CANDS=1,8,26 python tools/bench_evaluate.py "" "FASTRANK_TMA_EVAL=1" > gpurun_out/cap/r02_evaluate_bench.json 2>/dev/null
# 6. the bench itself
python bench.py > gpurun_out/cap/r02_bench.json 2>gpurun_out/cap/bench_err.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/cap/r02_bench_reference_arm.json 2>/dev/null
ls -la gpurun_out/cap
