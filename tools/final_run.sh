set -x
python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2v_bench_reference.json 2> gpurun_out/r2v_bench_reference.err
python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r2v_plain.json 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2v_launches.csv python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r2v_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gp_fit_kernel|predict_ei_fused_kernel' -s 8 -c 4 -o gpurun_out/r2v_full python bench.py --steps 2 --warmup 3 --sub none --no-cpu-baseline --passes 1 > gpurun_out/r2v_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
