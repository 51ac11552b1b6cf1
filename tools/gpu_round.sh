#!/bin/bash
# One gpurun call: GPU tests, the default bench line, a launch list.  Logs -> gpurun_out/.
# usage: tools/gpu_round.sh <tag> [pytest-args...]
tag=${1:-run}; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
nproc >> gpurun_out/${tag}_smi.txt

timeout 1500 python -m pytest tests -m gpu -q --timeout 600 "$@" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 50 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -5 gpurun_out/${tag}_bench.err; head -c 1500 gpurun_out/${tag}_bench.json
# every launch of two steps with its device time (cold-cache, serialised: compare shares)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_step.py 2 > gpurun_out/${tag}_onestep.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${tag}_onestep.log
