timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo bench rc=$?
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-ref-cuda > gpurun_out/r2g_ncu_bench.log 2>&1; echo ncu1 rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:query_tc3 -c 2 -o gpurun_out/r2g_q3_nc18 -f python tools/one_query.py 18 tc3 > gpurun_out/r2g_ncu_q3.log 2>&1; echo ncu2 rc=$?
