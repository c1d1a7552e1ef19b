#!/bin/bash
# Round-2 closing capture after the SSIM kernel change (run on a 1-GPU B200 box through gpurun; outputs land in gpurun_out/):
#   bash profiles/capture_r02b.sh
# The ncu-wrapped commands time the serial loop (--serial; the committed files were captured before the two-deep loop became
# bench.py's default, i.e. with exactly that loop): a profiler serialises the launches anyway, the kernels are the same.
# Only srf_iqa.cu moved since profiles/capture_r02.sh ran, so the full-section capture is of k_psnr_ssim alone; the other
# kernels keep their entries in profiles/ncu_traffic.json (per-kernel source digests, see profiles/summarize_ncu.py).
set -x
# 1. plain run first: ncu only ever wraps a command that has exited 0 without it
python bench.py --serial --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_plain.log 2>&1 || exit 1
# 2. full sections of ONE launch of the SSIM kernel, summarised on the box so that the bench line below carries its traffic
ncu --set full --clock-control none --import-source on -k regex:"k_psnr_ssim" -c 1 -f -o gpurun_out/r02b_ncu_ssim \
    python bench.py --serial --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02b_ncu_ssim.log 2>&1
python profiles/summarize_ncu.py gpurun_out/r02b_ncu_ssim.ncu-rep gpurun_out/r02b_ncu_ssim_summary.txt --traffic profiles/ncu_traffic.json \
    > gpurun_out/r02b_summarize.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
# 3. the bench lines
python bench.py --steps 10 --warmup 3 --verify > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.log
python bench.py --workload cfg5_iqa --steps 10 --warmup 3 > gpurun_out/r02b_bench_cfg5_n1.json 2> gpurun_out/r02b_bench_cfg5_n1.log
# 4. launch list of warm-up + 2 timed steps (cold-cache, serialised times: compare SHARES with the bench line's table)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02b_ncu_launches.csv \
    python bench.py --serial --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_ncu_launches.log 2>&1
ls -la gpurun_out/r02b_*
