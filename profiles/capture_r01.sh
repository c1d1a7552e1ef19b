#!/bin/bash
# Evidence capture for profiles/ (run on a B200 box through gpurun; outputs land in gpurun_out/).
#   bash profiles/capture_r01.sh
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r01b_bench_n1.json 2> gpurun_out/r01b_bench_n1.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01b_bench_reference.json 2> gpurun_out/r01b_bench_reference.log
# launch list of one warm-up + timed run (cold-cache, serialised times: compare SHARES with the bench line's table)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01b_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01b_ncu_launches.log 2>&1
# full sections of the hot kernels of the first step: -s 5 skips the five weight-pyramid launches of the first call
ncu --set full --clock-control none --import-source on -k regex:"k_final|k_psnr_ssim|k_down_u8|k_up_levels|k_down_f32s" -s 5 -c 12 -f \
    -o gpurun_out/r01b_ncu_full python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r01b_ncu_full.log 2>&1
ls -la gpurun_out/r01b_*
