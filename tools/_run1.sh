timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for i in 1 2; do
ALINE_B200_LIB=$PWD/aline_b200/lib/libaline_b200_exp.so timeout 100 python tools/bench_fold.py > gpurun_out/fold_twice.json 2>/dev/null
timeout 100 python tools/bench_fold.py > gpurun_out/fold_once.json 2>/dev/null
python - <<EOP
import json
a=json.load(open("gpurun_out/fold_once.json")); b=json.load(open("gpurun_out/fold_twice.json"))
for k in a: print(k, a[k]["keys"], "unfolded", round(a[k]["fold0_us"],1), "fold", round(a[k]["fold1_us"],1), "fold twice", round(b[k]["fold1_us"],1))
EOP
done
