python -m pytest tests/test_gpu_fullsize.py -q -m gpu -k "energy_history" -s 2>&1 | grep -E "steps;|passed|failed" > gpurun_out/r02j_energy.log; cat gpurun_out/r02j_energy.log
for c in c1 c2 c3 c4 c5_64m; do python bench.py --config $c --steps 5 --warmup 3 --no-traffic > gpurun_out/r02j_bench_$c.json 2> gpurun_out/r02j_bench_$c.err; tail -c 200 gpurun_out/r02j_bench_$c.err; done
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02j_bench_reference_arm.json 2>gpurun_out/r02j_ref.err
ls -la gpurun_out/r02j_*
