mkdir -p gpurun_out/r01c; O=gpurun_out/r01c
python bench.py > $O/bench_default.log 2>&1; tail -1 $O/bench_default.log | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.log 2>&1
python bench.py --workload fib4g --steps 20 --warmup 3 --no-cpu > $O/bench_fib4g.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"hb_fsm_sync|hb_emitw" -c 2 -o $O/prof_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $O/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"hb_fsm_sync|hb_emitw" -c 2 -o $O/prof_full_fib -f python bench.py --workload fib4g --steps 1 --warmup 3 --no-cpu --no-e2e > $O/ncu_full_fib.log 2>&1
make -C huffmandecoderongpus_b200 host/HuffFrameworkBaselines > /dev/null 2>&1
(cd huffmandecoderongpus_b200/host && B200_REPEATS=25 ./HuffFrameworkBaselines all ../../oracle/_ref/files) > $O/harness_all.log 2>&1
(cd huffmandecoderongpus_b200/host && ./HuffFramework synth1g; ./HuffFramework synthfib) > $O/harness_synth.log 2>&1
python tools/compare_reference_gpu.py 10 > $O/compare_ref_gpu.log 2>&1
python tools/graph_sweep.py > $O/graph_sweep_kjv.log 2>&1
python tools/slow_path_probe.py > $O/slow_path.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/smi.txt
tail -3 $O/compare_ref_gpu.log; tail -3 $O/slow_path.log
