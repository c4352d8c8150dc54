mkdir -p gpurun_out
NCU_ONCE=1 timeout 700 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -o gpurun_out/gemm_r2 -f python tools/bench_gemm.py > gpurun_out/ncu_gemm_r2.log 2>&1; tail -2 gpurun_out/ncu_gemm_r2.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1; tail -2 gpurun_out/ncu_b.log | cut -c1-200
timeout 120 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cat gpurun_out/bench_gemm.txt | cut -c1-200
ls -la gpurun_out/gemm_r2.ncu-rep gpurun_out/launches.csv
