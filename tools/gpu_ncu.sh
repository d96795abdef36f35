#!/bin/bash
# One GPU-box visit for the profiler evidence of the DEFAULT kernels at HEAD (B200_PROFILING.md recipe):
#   1. ncu launch list (gpu__time_duration.sum, --clock-control none) of one training view in the trainer's kernel order
#      (tools/native/view_check: plain C++ over the C ABI, so ncu's per-launch serialisation costs seconds, not minutes);
#   2. ncu --set full of the hot kernels of that view (source-level, -lineinfo);
# then, back on the CPU box:  python tools/ncu_summary.py gpurun_out/prof_<tag>.ncu-rep --json profiles/ncu_dram_bytes.json > profiles/<tag>_ncu_full.csv
# usage: gpurun --timeout 900 -- 'bash tools/gpu_ncu.sh <tag>'
TAG=${1:-rX}
mkdir -p gpurun_out
make -s -C tools/native >/dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
    tools/native/view_check 1000000 1280 720 0.01 2 > gpurun_out/ncu_launches_$TAG.log 2>&1; echo "launch list rc=$?"
KRE=${2:-'deform_mlp|composite_|rs_onesweep|rs_histogram|emit_instances|preprocess_|hexplane_time|hexplane_bwd|hexplane_fwd|tile_ranges|adam_'}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s ${3:-39} -c ${4:-21} -f -o gpurun_out/prof_$TAG \
    tools/native/view_check 1000000 1280 720 0.01 2 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "full rc=$?"
tail -3 gpurun_out/ncu_full_$TAG.log
# summaries are extracted HERE (the report itself can exceed what gpurun copies back: it is dropped if larger than 45 MB)
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep --json gpurun_out/ncu_dram_bytes_$TAG.json > gpurun_out/ncu_full_$TAG.csv 2> gpurun_out/ncu_summary_$TAG.err
for k in deform_mlp_bwd deform_mlp_fwd composite_bwd2 composite_fwd rs_onesweep_pass hexplane_time_bwd2 hexplane_bwd_kernel; do
    python tools/ncu_stalls.py gpurun_out/prof_$TAG.ncu-rep $k > gpurun_out/stalls_${k}_$TAG.txt 2>&1
done
# 3. the launch list of bench.py itself (the command whose `kernels` shares it must agree with): training steps only
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --render-frames 0 --no-raster-only --no-c5 --no-launcher-path --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.json 2> gpurun_out/ncu_bench_$TAG.err; echo "bench launch list rc=$?"
python tools/ncu_launch_shares.py gpurun_out/launches_bench_$TAG.csv > gpurun_out/launch_shares_bench_$TAG.txt 2>&1
python tools/ncu_launch_shares.py gpurun_out/launches_$TAG.csv > gpurun_out/launch_shares_view_$TAG.txt 2>&1
gzip -f gpurun_out/launches_bench_$TAG.csv
ls -la gpurun_out/prof_$TAG.ncu-rep
[ $(stat -c %s gpurun_out/prof_$TAG.ncu-rep) -gt 45000000 ] && rm -f gpurun_out/prof_$TAG.ncu-rep
du -sh gpurun_out
