python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02n_pytest_gpu.log; cat gpurun_out/r02n_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02n_smoke.log 2>&1; tail -1 gpurun_out/r02n_smoke.log
python bench.py > gpurun_out/r02n_bench_16M.json 2> gpurun_out/r02n_bench.err; tail -c 300 gpurun_out/r02n_bench.err
