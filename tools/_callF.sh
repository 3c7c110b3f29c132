set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/F_gpu.log
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/F_pytest.log 2>&1
tail -5 gpurun_out/F_pytest.log
python bench.py > gpurun_out/F_bench.json 2> gpurun_out/F_bench.err
cat gpurun_out/F_bench.json
python tools/stage_times.py 265 A > gpurun_out/F_stage.log 2>&1
cat gpurun_out/F_stage.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/F_launches.csv python bench.py --steps 1 --warmup 3 > gpurun_out/F_ncu_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_lz_match|k_lz_chain|k_spec_resolve|k_spec_round|k_find_blocks|k_parse_exits|k_spec_tokens|k_checksum" -c 16 -f -o gpurun_out/prof_F python tools/stage_times.py 64 A > gpurun_out/F_ncu.log 2>&1
tail -3 gpurun_out/F_ncu.log
