# ncu evidence for profiles/: launch list of the bench command + one --set full capture of every kernel of the path.
# The .ncu-rep stays on the box (too large to bring back whole); summaries and per-source-line tables are written as text.
set -x
T=$1; TT=$1
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TT}_launches.csv python bench.py --steps 1 --warmup 3 --skip-cpu > gpurun_out/${TT}_launches.log 2>&1
tail -2 gpurun_out/${TT}_launches.log
B2F_DECODE_PARTS=1 timeout -s KILL 1200 ncu --set full --clock-control none --import-source on -k regex:"k_" -c ${NCU_C:-100} -f -o /tmp/prof_${TT} python tools/stage_times.py ${NCU_MIB:-64} P > gpurun_out/${TT}_ncu.log 2>&1
tail -2 gpurun_out/${TT}_ncu.log
python tools/ncu_summary.py /tmp/prof_${TT}.ncu-rep > gpurun_out/${TT}_ncu_summary.txt 2> gpurun_out/${TT}_ncu_summary.err
wc -l gpurun_out/${TT}_ncu_summary.txt
for spec in "k_lz_find encode_kernels" "k_spec_round spec_kernels" "k_find_blocks decode_kernels" "k_seg_resolve spec_kernels" "k_spec_tokens spec_kernels" "k_parse_exits encode_kernels"; do
  set -- $spec
  python tools/ncu_lines.py /tmp/prof_${TT}.ncu-rep "$1" libflate_b200/libb2f.so $2 60 > gpurun_out/${TT}_lines_$(echo $1 | tr -cd 'a-z_0-9').txt 2>&1
done
ls -la gpurun_out | tail -12
