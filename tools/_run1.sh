timeout 300 python -m pytest tests/test_query_parity_gpu.py tests/test_forward_gpu.py tests/test_tc_gpu.py -x -q -m gpu 2>&1 | tail -4
timeout 100 python tools/bench_fold.py > gpurun_out/fold_ctx.json 2>gpurun_out/fold_ctx.err
python - <<EOP
import json
a=json.load(open("gpurun_out/fold_ctx.json"))
for k in a: print(k, a[k]["keys"], "unfolded", round(a[k]["fold0_us"],1), "fold", round(a[k]["fold1_us"],1), a[k]["fold1_max_abs_vs_fp32"])
EOP
for f in 0 1 0 1; do echo fold=$f; ALINE_QUERY_FOLD=$f timeout 100 python tools/rollout_ab.py 2>&1 | tail -2; done
