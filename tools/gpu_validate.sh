# Full GPU validation of a build (run through gpurun): pytest -m gpu, stage times, bench.py, ncu launch list, ncu --set full of the top kernels.
# Everything lands in gpurun_out/P_*; copy what is to be kept into profiles/.
set -x
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/P_pytest.log 2>&1
tail -4 gpurun_out/P_pytest.log
timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/P_stage.log 2>&1
cat gpurun_out/P_stage.log
timeout -s KILL 400 python bench.py > gpurun_out/P_bench.json 2> gpurun_out/P_bench.err
cat gpurun_out/P_bench.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/P_launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/P_ncu_bench.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_lz_find|k_spec_resolve|k_find_blocks|k_checksum" -c 8 -f -o gpurun_out/prof_P python tools/stage_times.py 64 A > gpurun_out/P_ncu.log 2>&1
tail -2 gpurun_out/P_ncu.log
