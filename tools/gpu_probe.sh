#!/bin/bash
# matcher shard test + the e2e pause probe (twice: the pauses depend on the box)
tag=${1:-p}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "match or config3 or config4 or golden" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
for i in 1 2; do timeout 200 python tools/e2e_pause_probe.py 900 > gpurun_out/${tag}_pause_probe_$i.log 2>&1; echo "probe exit $?"; done
cat gpurun_out/${tag}_pause_probe_1.log gpurun_out/${tag}_pause_probe_2.log | tail -150
