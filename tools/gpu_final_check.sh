# Quick GPU check (gpurun): pytest -m gpu, smoke(), bench.py.
set -x
mkdir -p gpurun_out
( time timeout -s KILL 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/S_pytest.log 2>&1
tail -4 gpurun_out/S_pytest.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/S_smoke.log 2>&1; tail -1 gpurun_out/S_smoke.log
timeout -s KILL 400 python bench.py > gpurun_out/S_bench.json 2> gpurun_out/S_bench.err
cut -c1-400 gpurun_out/S_bench.json
