#!/bin/bash
# One gpurun call: ncu --set full with source of k_descriptor / k_orientation (one launch each).
tag=${1:-profd}
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_descriptor|k_orientation" --launch-skip 2 -c 2 -o gpurun_out/${tag}_desc python tools/one_step.py 2 > gpurun_out/${tag}_desc.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/${tag}_desc.log
ls -la gpurun_out/${tag}_*.ncu-rep
