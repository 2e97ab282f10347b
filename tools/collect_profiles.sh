# Round-2 evidence, one gpurun call on one B200:  bash tools/collect_profiles.sh [tag]   (default tag r02j)
T=${1:-r02l}; mkdir -p gpurun_out/$T; O=gpurun_out/$T
git rev-parse HEAD > $O/commit.txt 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/smi.txt
python bench.py > $O/bench_default.log 2>&1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-secondary"
$CMD > $O/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_english1g.csv $CMD > $O/ncu_list.log 2>&1
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary"
$CMD1 > $O/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"hb_fsm_sync|hb_emit32|hb_emitw" -s 2 -c 2 -o $O/prof_full_english1g -f $CMD1 > $O/ncu_full.log 2>&1
CMD2="python bench.py --workload fib4g --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary"
$CMD2 > $O/plain2.log 2>&1 && ncu --set full --clock-control none -k regex:"hb_fsm_sync|hb_emit32|hb_emitw" -s 2 -c 2 -o $O/prof_full_fib4g -f $CMD2 > $O/ncu_full_fib.log 2>&1
make -C huffmandecoderongpus_b200 host/HuffFrameworkBaselines > /dev/null 2>&1
(cd huffmandecoderongpus_b200/host && B200_REPEATS=25 ./HuffFrameworkBaselines all ../../oracle/_ref/files) > $O/harness_all.log 2>&1
(cd huffmandecoderongpus_b200/host && ./HuffFramework synth1g; ./HuffFramework synthfib; ./HuffFramework synth16g) > $O/harness_synth.log 2>&1
python tools/compare_reference_gpu.py 10 > $O/compare_ref_gpu.log 2>&1
python tools/graph_sweep.py > $O/graph_sweep_kjv.log 2>&1
python tools/latency_probe.py > $O/latency_probe.log 2>&1
python tools/slow_path_probe.py > $O/slow_paths.log 2>&1
tail -2 $O/ncu_full.log; tail -3 $O/compare_ref_gpu.log; grep '^{' $O/bench_default.log | cut -c1-300
