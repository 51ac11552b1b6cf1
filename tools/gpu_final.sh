#!/bin/bash
# The closing 1-GPU call of a round: A/B of build variants, the whole GPU suite, smoke, the default bench
# line, the launch list and the --set full capture of the step's kernels.  Logs -> gpurun_out/.
tag=${1:-final}
mkdir -p gpurun_out
{
  python tools/ab_value.py 40
  for v in openpano_b200/_variants/*.so; do [ -f "$v" ] && PANO_B200_LIB=$v python tools/ab_value.py 40; done
  python tools/ab_value.py 40
} > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 50 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json; echo
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_step.py 2 > gpurun_out/${tag}_onestep.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:k_descriptor|k_orientation|k_blur_dog_fast|k_extrema_scan|k_linear_blend|k_tc_pass|k_working_resize|k_octave_grey|k_rank_sort|k_refine" --launch-skip 13 -c 13 -o gpurun_out/${tag}_step python tools/one_step.py 2 > gpurun_out/${tag}_step.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/${tag}_step.log
