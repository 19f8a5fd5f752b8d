set -x
python bench.py > gpurun_out/r2q_bench_ours.json 2> gpurun_out/r2q_bench.err || echo BENCH_FAILED
python bench.py --impl reference > gpurun_out/r2q_bench_ref.json 2>> gpurun_out/r2q_bench.err || echo REF_FAILED
python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/r2q_short.json 2>> gpurun_out/r2q_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-sweep > gpurun_out/r2q_ncu_list.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"nmpc_qp_kernel|nmpc_linearize|sdf_eval_tc" --launch-skip 9 --launch-count 3 -o gpurun_out/rti_r2q -f python bench.py --steps 3 --warmup 3 --no-cpu --no-sweep > gpurun_out/r2q_ncu_full.log 2>&1
ls -la gpurun_out/rti_r2q.ncu-rep
tail -c 300 gpurun_out/r2q_bench.err
