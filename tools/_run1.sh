for i in 1 2; do
echo emit; timeout 100 python tools/rollout_ab.py 2>&1 | tail -2
echo skip; ALINE_B200_LIB=$PWD/aline_b200/lib/libaline_b200_exp.so timeout 100 python tools/rollout_ab.py 2>&1 | tail -2
done
