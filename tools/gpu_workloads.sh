# bench.py over the BASELINE workloads on ONE GPU (run through gpurun); lines land in gpurun_out/<tag>_<workload>.json
set -x
T=$1
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -x -q -k "decoder_read or split or abi" > gpurun_out/${T}_fix.log 2>&1; tail -2 gpurun_out/${T}_fix.log
timeout -s KILL 300 python -m pytest tests/test_gpu_split.py tests/test_gpu_abi_order.py -x -q > gpurun_out/${T}_split.log 2>&1; tail -3 gpurun_out/${T}_split.log
for W in config2 config4 decode-foreign; do
  timeout -s KILL 900 python bench.py --workload $W > gpurun_out/${T}_$W.json 2> gpurun_out/${T}_$W.err
  cat gpurun_out/${T}_$W.json | cut -c1-900; tail -2 gpurun_out/${T}_$W.err
done
free -g | head -2; nproc
timeout -s KILL 1500 python bench.py --workload config5 --steps 3 > gpurun_out/${T}_config5.json 2> gpurun_out/${T}_config5.err
cat gpurun_out/${T}_config5.json | cut -c1-900; tail -3 gpurun_out/${T}_config5.err
timeout -s KILL 600 python bench.py --impl reference > gpurun_out/${T}_reference.json 2> gpurun_out/${T}_reference.err
cat gpurun_out/${T}_reference.json | cut -c1-600
