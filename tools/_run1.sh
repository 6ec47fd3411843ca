timeout 300 python -m pytest tests/test_query_parity_gpu.py tests/test_forward_gpu.py tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -3
for i in 1 2; do
echo emit; timeout 100 python tools/rollout_ab.py 2>&1 | tail -2
done
