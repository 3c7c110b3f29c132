set -x
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/Q_pytest.log 2>&1
tail -4 gpurun_out/Q_pytest.log
timeout -s KILL 120 python tools/stage_times.py 265 A > gpurun_out/Q_stage.log 2>&1
cat gpurun_out/Q_stage.log
timeout -s KILL 400 python bench.py > gpurun_out/Q_bench.json 2> gpurun_out/Q_bench.err
cat gpurun_out/Q_bench.json | cut -c1-1200
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/Q_launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/Q_ncu_bench.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_spec_resolve" -c 1 -f -o gpurun_out/prof_Q python tools/stage_times.py 64 A > gpurun_out/Q_ncu.log 2>&1
tail -2 gpurun_out/Q_ncu.log
