#!/bin/bash
# Everything profiles/ needs from one GPU box: bash scripts/gpu_profile_round.sh r02   (run under gpurun; ~4 minutes)
#  1. the bench line (N=1) and the reference arm                              -> gpurun_out/<tag>_bench_line.json, _bench_reference.json
#  2. the ncu launch list of the same bench command (shares, not absolutes)   -> gpurun_out/<tag>_launches.csv
#  3. ncu --set full of the dominant kernels (kernel 1, kernel 2, fused)      -> gpurun_out/<tag>_k1k2.ncu-rep, _fused.ncu-rep, _otsu.ncu-rep
#  4. all BASELINE configs at full size                                        -> gpurun_out/<tag>_configs.json
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/${tag}_bench_line.json 2> gpurun_out/${tag}_bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/${tag}_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"integral_sq_kernel|threshold_tma_kernel" -s 2 -c 2 \
    -o gpurun_out/${tag}_k1k2 python scripts/prof_step.py pages=256 steps=2 enable_fused=0 > gpurun_out/${tag}_ncu_k1k2.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"local_fused_kernel|fixup_kernel" -s 2 -c 2 \
    -o gpurun_out/${tag}_fused python scripts/prof_step.py pages=256 steps=2 > gpurun_out/${tag}_ncu_fused.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"hist_units_kernel|otsu_apply_kernel|otsu_tiles" -s 3 -c 3 \
    -o gpurun_out/${tag}_otsu python scripts/profile_ops.py 256 tiles otsu > gpurun_out/${tag}_ncu_otsu.log 2>&1
timeout 600 python scripts/bench_configs.py > gpurun_out/${tag}_configs.json 2>> gpurun_out/${tag}_bench.err
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${tag}_gpu.txt
ls -la gpurun_out | tail -15
