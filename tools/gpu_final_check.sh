# Final GPU check (gpurun): pytest -m gpu, smoke(), resident/e2e split, bench.py.
set -x
T=${1:-S}
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout -s KILL 200 python tools/resident_times.py > gpurun_out/${T}_resident.log 2>&1; grep -v Warn gpurun_out/${T}_resident.log | tail -6
timeout -s KILL 200 python tools/stage_times.py 265 A > gpurun_out/${T}_stage.log 2>&1; tail -3 gpurun_out/${T}_stage.log
timeout -s KILL 200 python tools/e2e_times.py > gpurun_out/${T}_e2e.log 2>&1; head -8 gpurun_out/${T}_e2e.log
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-300 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
