# ncu evidence of round 2 (run under gpurun, one part per call: the .ncu-rep files are large)
#   bash tools/ncu_round2.sh gates | train_fwd | bwd | lists
mkdir -p gpurun_out
case "$1" in
gates)
  python tools/profile_target.py 1024 parity 3 8 auto > gpurun_out/r2p_plain_eval.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 12 -c 6 -f -o gpurun_out/r2p_full_gates_eval_B1024 python tools/profile_target.py 1024 parity 3 8 auto > gpurun_out/r2p_ncu_eval.log 2>&1
  echo "eval full rc=$?" ;;
train_fwd)
  python tools/profile_train_fwd_target.py 1024 parity 2 > gpurun_out/r2f_plain_train_fwd.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_gemm -c 12 -f -o gpurun_out/r2f_full_gates_train_fwd_B1024 python tools/profile_train_fwd_target.py 1024 parity 2 > gpurun_out/r2f_ncu_train_fwd.log 2>&1
  echo "train fwd full rc=$?" ;;
bwd)
  python tools/profile_train_target.py 1024 parity 3 > gpurun_out/r2p_plain_train.log 2>&1 && \
  ncu --set full --clock-control none -k regex:'conv_gemm|wgrad_kernel|tail_|head_|lstm_bwd|bottleneck' -s 158 -c 27 -f -o gpurun_out/r2p_full_bwd_iter_B1024 python tools/profile_train_target.py 1024 parity 3 > gpurun_out/r2p_ncu_train.log 2>&1
  echo "train full rc=$?" ;;
lists)
  python bench.py --mode train --steps 1 --warmup 3 --no-cpu --no-fast > gpurun_out/r2p_plain_bench_train.json 2> gpurun_out/r2p_plain_bench_train.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 900 --csv --log-file gpurun_out/r2p_launches_bench_train.csv python bench.py --mode train --steps 1 --warmup 3 --no-cpu --no-fast > gpurun_out/r2p_ncu_bench_train.log 2>&1
  echo "launch list train rc=$?"
  python bench.py --mode eval --steps 1 --warmup 3 --no-cpu --no-fast > gpurun_out/r2p_plain_bench_eval.json 2> gpurun_out/r2p_plain_bench_eval.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 300 --csv --log-file gpurun_out/r2p_launches_bench_eval.csv python bench.py --mode eval --steps 1 --warmup 3 --no-cpu --no-fast > gpurun_out/r2p_ncu_bench_eval.log 2>&1
  echo "launch list eval rc=$?" ;;
esac
