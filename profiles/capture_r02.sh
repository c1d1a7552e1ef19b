#!/bin/bash
# Round-2 evidence capture (run on a 1-GPU B200 box through gpurun; outputs land in gpurun_out/):   bash profiles/capture_r02.sh
set -x
python bench.py --steps 10 --warmup 3 --verify > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.log
python bench.py --workload cfg5_iqa --steps 10 --warmup 3 > gpurun_out/r02_bench_cfg5_n1.json 2> gpurun_out/r02_bench_cfg5_n1.log
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.log
python profiles/aux_bench.py > gpurun_out/r02_aux_kernels.json 2> gpurun_out/r02_aux_kernels.log
# launch list of warm-up + 2 timed steps (cold-cache, serialised times: compare SHARES with the bench line's table)
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_ncu_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1
# full sections of every launch of ONE step (the first call's weight-pyramid launches are skipped by the -k filter)
ncu --set full --clock-control none --import-source on -k regex:"k_gather|k_final|k_psnr_ssim|k_down_u8|k_down_u16|k_up_level|k_down_f32s" -c 30 -f \
    -o gpurun_out/r02_ncu_full python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_full.log 2>&1
ls -la gpurun_out/r02_*
