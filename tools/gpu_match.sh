#!/bin/bash
# Short GPU call for matcher work: matcher tests + the match-heavy bench legs (+ the e2e pause probe).
tag=${1:-m}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x -k "match or config3 or config4 or golden or adaptors" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -6 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 30 --configs 3,4 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -3 gpurun_out/${tag}_bench.err
python tools/show_bench.py gpurun_out/${tag}_bench.json 2>&1 | grep -E "^value|k_tc_top2|k_tc_nominate|^[0-9]+ \{|ms_device"
if [ -n "$2" ]; then timeout 300 python tools/e2e_pause_probe.py 600 > gpurun_out/${tag}_pause_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/${tag}_pause_probe.log | tail -30; fi
