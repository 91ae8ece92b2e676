#!/bin/bash
# Profiling pass of a round (run on the GPU box through gpurun): launch lists of the bench and of the full per-frame path,
# one ncu --set full capture of the dominant kernel (k_brox_sor, finest level).  Outputs land in gpurun_out/.
# NOTE: ncu costs ~170 ms per launch on this pool (save/restore of the resident buffers), so the bench list is cut after
# the first 2500 launches (= the first five steps of the run, ~7 GPU-minutes); shares, not absolutes, are what it is for.
set -u
TAG=${1:-r1l}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 3 --warmup 3 --headline-only > $OUT/${TAG}_bench_short.log 2>&1 || { echo "bench failed"; tail -5 $OUT/${TAG}_bench_short.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --headline-only > $OUT/${TAG}_ncu_bench.log 2>&1
echo "bench launch list rc=$?"
python tools/profile_detect.py 2 > $OUT/${TAG}_detect.log 2>&1 || { echo "profile_detect failed"; tail -5 $OUT/${TAG}_detect.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/${TAG}_launches_detect.csv \
    python tools/profile_detect.py 2 > $OUT/${TAG}_ncu_detect.log 2>&1
echo "detect launch list rc=$?"
python tools/run_brox_once.py 1 > /dev/null 2>&1 || { echo "run_brox_once failed"; exit 1; }
# 180 k_brox_sor launches per solve, the last 20 are the 384x288 level
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_brox_sor --launch-skip 175 --launch-count 1 \
    -o $OUT/${TAG}_brox_sor python tools/run_brox_once.py 1 > $OUT/${TAG}_ncu_sor.log 2>&1
echo "sor capture rc=$?"
ncu -i $OUT/${TAG}_brox_sor.ncu-rep --page details > $OUT/${TAG}_brox_sor_ncu_details.txt 2>&1
ncu -i $OUT/${TAG}_brox_sor.ncu-rep --page raw --csv > $OUT/${TAG}_brox_sor_ncu_raw.csv 2>&1
ls -la $OUT | tail -12
