#!/bin/bash
# matcher shard test + the e2e pause probe (twice: the pauses depend on the box)
tag=${1:-p}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x -k "row_shards or match_pairs_bit_exact or config4_sweep_vs" > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
for i in 1 2; do timeout 200 python tools/e2e_pause_probe.py 900 > gpurun_out/${tag}_pause_probe_$i.log 2>&1; echo "probe exit $?"; done
cat gpurun_out/${tag}_pause_probe_1.log gpurun_out/${tag}_pause_probe_2.log | tail -150
