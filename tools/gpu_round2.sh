# One GPU validation pass (run through gpurun): parity + full-size tests, stage times, bench line, ncu of the round-2 kernels.
# usage: bash tools/gpu_round2.sh <tag> [quick]
set -x
T=$1
mkdir -p gpurun_out
( time timeout -s KILL 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi_order.py tests/test_gpu_split.py -x -q ) > gpurun_out/${T}_parity.log 2>&1
tail -3 gpurun_out/${T}_parity.log
timeout -s KILL 200 python tools/stage_times.py 265 A > gpurun_out/${T}_stage.log 2>&1
tail -3 gpurun_out/${T}_stage.log
timeout -s KILL 200 python tools/e2e_times.py 265 > gpurun_out/${T}_e2e.log 2>&1
cat gpurun_out/${T}_e2e.log
timeout -s KILL 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
cat gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
if [ "$2" != "quick" ]; then
( time timeout -s KILL 1200 python -m pytest tests/test_gpu_fullsize.py -x -q ) > gpurun_out/${T}_full.log 2>&1
tail -3 gpurun_out/${T}_full.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:"${NCU_K:-k_seg_resolve|k_seg_subst|k_checksum<}" -c ${NCU_C:-4} -f -o gpurun_out/prof_${T} python tools/stage_times.py 64 A > gpurun_out/${T}_ncu.log 2>&1
tail -2 gpurun_out/${T}_ncu.log
fi
