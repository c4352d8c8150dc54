# final single-GPU evidence of round 2 (fp16-pair forward blend, side-stream scheduling, pipelined read-back)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-160 gpurun_out/bench_final.json
for c in 0 2 3 4; do
  timeout 200 python bench.py --config $c --steps 10 > gpurun_out/bench_final_c$c.json 2> gpurun_out/bench_final_c$c.err; cut -c1-160 gpurun_out/bench_final_c$c.json; tail -1 gpurun_out/bench_final_c$c.err | cut -c1-160
done
timeout 200 python tools/profile_step.py > gpurun_out/profile_step.txt 2>&1; head -8 gpurun_out/profile_step.txt
timeout 200 python tools/bench_skin.py > gpurun_out/bench_skin.txt 2>&1; cat gpurun_out/bench_skin.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_b.log 2>&1; tail -1 gpurun_out/ncu_b.log | cut -c1-120
NCU_ONCE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_skin" -o gpurun_out/k7_final2 -f python tools/bench_skin.py 4096 > gpurun_out/ncu_k7_final2.log 2>&1; tail -2 gpurun_out/ncu_k7_final2.log
ls -la gpurun_out/*.ncu-rep | tail -3
