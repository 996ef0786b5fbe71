# Round profile: GPU tests, the default bench line, the ncu launch list, one --set full capture of k_expand and the integer metrics.
# Usage (from the repo root, under gpurun): bash tools/profile_round.sh r2
R=${1:-r2}
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gputests.log 2>&1; echo tests=$?
tail -3 gpurun_out/${R}_gputests.log
python bench.py > gpurun_out/${R}_bench_cfg2_n1.json 2> gpurun_out/bench.err; echo bench=$?
CMD="python bench.py --steps 5 --warmup 3 --launches-per-step 1 --no-cpu --verify 0 --no-witness-d2h --no-north-star"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_cfg2.csv $CMD > gpurun_out/ncu1.log 2>&1; echo ncu1=$?
ncu --set full --clock-control none --import-source on -k regex:k_expand -s 4 -c 1 -f -o gpurun_out/${R}_k_expand_full $CMD > gpurun_out/ncu2.log 2>&1; echo ncu2=$?
ncu --metrics smsp__sass_thread_inst_executed_op_integer_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:k_expand -s 4 -c 1 --csv --log-file gpurun_out/${R}_k_expand_int.csv $CMD > gpurun_out/ncu3.log 2>&1; echo ncu3=$?
tail -2 gpurun_out/ncu2.log; ls -la gpurun_out | tail -12
