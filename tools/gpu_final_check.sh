# Final GPU check (gpurun): pytest -m gpu, smoke(), bench.py, resident split.
set -x
T=${1:-S}
mkdir -p gpurun_out
( time timeout -s KILL 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/${T}_pytest.log 2>&1
tail -4 gpurun_out/${T}_pytest.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cut -c1-300 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err
timeout -s KILL 200 python tools/resident_times.py > gpurun_out/${T}_resident.log 2>&1; grep -v Warn gpurun_out/${T}_resident.log | tail -6
for so in libflate_b200/libb2f_*.so; do echo $so; B2F_LIB=$so timeout -s KILL 100 python tools/resident_times.py 2>&1 | grep -E "overlap=True|overlap=False|enc:|dec:" | head -6 | cut -c1-260; done
