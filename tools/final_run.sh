set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r3z_gpu_tests.txt
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3z_bench.json 2> gpurun_out/r3z_bench.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r3z_bench_reference.json 2> gpurun_out/r3z_bench_reference.err
python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r3z_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r3z_launches.csv python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r3z_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gp_fit_kernel|predict_ei_fused_kernel' -s 8 -c 4 -o gpurun_out/r3z_full python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r3z_ncu_full.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3z_smoke.txt 2>&1
ls -la gpurun_out/ | tail -10
