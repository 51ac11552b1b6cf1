#!/bin/bash
# One gpurun call of ncu captures (1 GPU): the step's top kernels and the 100k sweep GEMM.
tag=${1:-prof}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_descriptor|k_orientation|k_blur_dog_fast|k_extrema_scan|k_linear_blend|k_tc_pass|k_working_resize|k_octave_grey" --launch-skip 10 -c 10 -o gpurun_out/${tag}_step python tools/one_step.py 2 > gpurun_out/${tag}_step.log 2>&1
echo "ncu step exit $?"; tail -2 gpurun_out/${tag}_step.log
timeout 900 ncu --set full --import-source on --clock-control none -k "regex:k_tc_pass" --launch-skip 7 -c 2 -o gpurun_out/${tag}_sweep python tools/sweep_step.py 100000 2 > gpurun_out/${tag}_sweep.log 2>&1
echo "ncu sweep exit $?"; tail -2 gpurun_out/${tag}_sweep.log
ls -la gpurun_out/${tag}_*.ncu-rep
