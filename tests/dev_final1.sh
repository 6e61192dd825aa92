set -x
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r02i_pytest_gpu.log
tail -3 gpurun_out/r02i_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02i_smoke.log 2>&1; tail -1 gpurun_out/r02i_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_bench_16M.json 2> gpurun_out/r02i_bench_16M.err
tail -c 400 gpurun_out/r02i_bench_16M.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02i_launches_16M.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-traffic > gpurun_out/r02i_b_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_pre_interaction|k_fluid_force|k_gravity' -s 3 -c 3 -o gpurun_out/r02i_full_16M -f python tests/dev_profile.py 312 2 > gpurun_out/r02i_full_ncu.log 2>&1
ls -la gpurun_out/r02i_*
