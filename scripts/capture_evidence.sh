#!/bin/sh
# Round evidence in one GPU call: bench lines, launch list, ncu captures of the hot kernels.
# usage (on the GPU box, from the repo root): sh scripts/capture_evidence.sh r01
R=${1:-r01}
O=gpurun_out
mkdir -p $O
python bench.py --steps 200 --warmup 5 > $O/${R}_bench_native.json 2> $O/${R}_bench_native.err
python bench.py --eager --steps 200 --warmup 5 --no-extras --no-cpu-baseline > $O/${R}_bench_native_eager.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2>/dev/null
# launch list (only after the same command has passed without ncu, above)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_bench_launches.csv \
    python bench.py --eager --steps 3 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1
# full captures, one launch per kernel
python scripts/prof_case.py fwdbwd level2 iid canon 2 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"warpcorr_fwd_tma|corr_bwd_seq|warp_bwd_tile|zero2|deinterleave8" -c 6 \
    -o $O/${R}_level2_fwdbwd python scripts/prof_case.py fwdbwd level2 iid canon 1 > /dev/null 2>&1
python scripts/prof_case.py fwdbwd level6 iid canon 2 > /dev/null 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"small_kernel" -c 2 \
    -o $O/${R}_level6_fwdbwd python scripts/prof_case.py fwdbwd level6 iid canon 1 > /dev/null 2>&1
ls -la $O/${R}_*
