timeout 300 python -m pytest tests/test_query_parity_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_configs.py > gpurun_out/r2f_configs.json 2> gpurun_out/r2f_configs.err; echo rc=$?
ALINE_QUERY_FOLD=0 timeout 300 python tools/bench_configs.py > gpurun_out/r2f_configs_nofold.json 2> /dev/null; echo rc=$?
python - <<EOP
import json
a=json.load(open("gpurun_out/r2f_configs.json")); b=json.load(open("gpurun_out/r2f_configs_nofold.json"))
for k,v in a.items():
    if isinstance(v,dict) and "rollout_ms" in v: print(k, "fold", round(v["rollout_ms"],3), "nofold", round(b[k]["rollout_ms"],3))
EOP
