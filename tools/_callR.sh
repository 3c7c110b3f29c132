set -x
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/R_pytest.log 2>&1
tail -4 gpurun_out/R_pytest.log
timeout -s KILL 200 python bench.py --skip-cpu > gpurun_out/R_bench.json 2> gpurun_out/R_bench.err
cut -c1-700 gpurun_out/R_bench.json
