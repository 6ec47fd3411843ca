timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 100 python tools/rollout_ab.py 2>&1 | tail -1; done
ALINE_QUERY_TC4=1 timeout 100 python tools/rollout_ab.py 2>&1 | tail -1
timeout 300 python tools/bench_configs.py > gpurun_out/r2f_configs.json 2> gpurun_out/r2f_configs.err; echo rc=$?
