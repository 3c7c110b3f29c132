# quick GPU pass: all gpu tests, stage times, e2e split, headline bench + decode-foreign
set -x
T=$1
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout -s KILL 200 python tools/stage_times.py 265 A > gpurun_out/${T}_stage.log 2>&1; tail -3 gpurun_out/${T}_stage.log
timeout -s KILL 600 python bench.py --workload decode-foreign --skip-cpu > gpurun_out/${T}_foreign.json 2> gpurun_out/${T}_foreign.err; cut -c1-1500 gpurun_out/${T}_foreign.json; tail -3 gpurun_out/${T}_foreign.err
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; cut -c1-600 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
