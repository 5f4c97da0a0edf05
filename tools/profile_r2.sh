#!/bin/bash
# ncu evidence of round 2 (run under gpurun, one GPU): full-set captures of the dominant kernels + the launch list of a bench step.
set -x
mkdir -p gpurun_out
# the filtered tcgen05 scan (2-CTA): launches alternate sample <false> / filtered <true>; -s 3 picks a filtered one
ncu --set full --clock-control none --import-source on -k regex:stage1_umma_kernel -s 3 -c 1 -f -o gpurun_out/r2_prof_umma python tools/prof_targets.py knn > gpurun_out/r2_prof_umma.log 2>&1
# SHA-256 over the chunk table of a 4 GiB segment (second call: pooled workspaces warm)
ncu --set full --clock-control none --import-source on -k regex:sha256_chunks_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_sha python tools/prof_targets.py sha > gpurun_out/r2_prof_sha.log 2>&1
# PQ ADC scan
ncu --set full --clock-control none --import-source on -k regex:pq_adc_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_pq python tools/prof_targets.py pq > gpurun_out/r2_prof_pq.log 2>&1
# launch list of a default bench step (knn + 16 GiB ingest, side measurements off)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-side --no-parity --no-cpu-baseline --ingest-gib 16 --e2e-ingest-gib 1 > gpurun_out/r2_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
