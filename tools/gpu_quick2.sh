# quick GPU pass: all gpu tests (fullsize last), e2e split, headline bench
set -x
T=$1
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout -s KILL 200 python tools/stage_times.py 265 A > gpurun_out/${T}_stage.log 2>&1; tail -40 gpurun_out/${T}_stage.log
timeout -s KILL 200 python tools/e2e_times.py > gpurun_out/${T}_e2e.log 2>&1; tail -30 gpurun_out/${T}_e2e.log
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-900 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
