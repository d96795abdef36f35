#!/bin/bash
# One GPU-box visit that reproduces what the driver runs at round end: the GPU test suite, both bench arms at N = 1, smoke().
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_final.sh <tag>'
TAG=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; tail -2 gpurun_out/pytest_$TAG.log
python bench.py --impl reference > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "reference arm rc=$?"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
