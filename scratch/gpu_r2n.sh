python -m pytest tests -m gpu -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log; cp gpurun_out/parity_errors_gpu.json gpurun_out/r2n_parity_errors_gpu.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2n_bench_ref.json 2> gpurun_out/r2n_bench_ref.err; echo "ref rc=$?"
for wl in real_like pvc t1 spatial; do python bench.py --workload $wl --steps 20 --no-cpu-baseline > gpurun_out/r2n_$wl.json 2> gpurun_out/r2n_$wl.err; done
python bench.py --workload nn --voxels 4000000 --steps 10 --no-cpu-baseline > gpurun_out/r2n_nn4m.json 2> gpurun_out/r2n_nn4m.err
python bench.py --workload disp --voxels 500000 --steps 10 --no-cpu-baseline > gpurun_out/r2n_disp500k.json 2> gpurun_out/r2n_disp500k.err
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5 --sustained-seconds 0.1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5 --sustained-seconds 0.1 > gpurun_out/r2n_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r2n_headline python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5 --sustained-seconds 0.1 > gpurun_out/r2n_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r2n_headline_1it python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-c5 --sustained-seconds 0.1 --iters-per-launch 1 > gpurun_out/r2n_ncu3.log 2>&1
