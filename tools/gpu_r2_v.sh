mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python tools/bench_skin.py 4096 2>&1 | head -4
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench.json; tail -2 gpurun_out/bench.err | cut -c1-200
