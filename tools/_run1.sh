timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do timeout 100 python tools/rollout_ab.py 2>&1 | tail -1; done
timeout 100 python tools/bench_fold.py > gpurun_out/fold_nkf.json 2>gpurun_out/fold_nkf.err
python - <<EOP
import json
a=json.load(open("gpurun_out/fold_nkf.json"))
for k in a: print(k, a[k]["keys"], "unfolded", round(a[k]["fold0_us"],1), "fold", round(a[k]["fold1_us"],1), "tc4", a[k].get("tc4_us"), a[k]["fold0_max_abs_vs_fp32"], a[k]["fold1_max_abs_vs_fp32"])
EOP
