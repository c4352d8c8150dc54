mkdir -p gpurun_out
timeout 200 python tools/bench_gemm.py > gpurun_out/bench_gemm.txt 2>&1; cut -c1-170 gpurun_out/bench_gemm.txt
timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -12 gpurun_out/profile_step.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-120
NCU_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_skin|k_gemm_h16" -o gpurun_out/k7_h16_r2 -f python tools/bench_skin.py 4096 > gpurun_out/ncu_k7_r2.log 2>&1; tail -2 gpurun_out/ncu_k7_r2.log
NCU_ONCE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_h16 -o gpurun_out/h16_r2 -f python tools/bench_gemm.py > gpurun_out/ncu_h16_r2.log 2>&1; tail -2 gpurun_out/ncu_h16_r2.log
ls -la gpurun_out/*.ncu-rep
