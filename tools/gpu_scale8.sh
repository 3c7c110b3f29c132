# 8-GPU runs of the bench workloads (run through `gpurun --gpus 8`): config3 (weak), config4 (strong, 1024 streams), config5 (weak, --gib 4)
set -x
T=$1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1; free -g | head -2; nproc
run() { W=$1; shift; timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload $W "$@" > gpurun_out/${T}_${W}_n8.json 2> gpurun_out/${T}_${W}_n8.err; cat gpurun_out/${T}_${W}_n8.json | cut -c1-700; tail -2 gpurun_out/${T}_${W}_n8.err; }
run config3 --skip-cpu
run config4 --skip-cpu
run config5 --gib 4 --steps 3 --skip-cpu
