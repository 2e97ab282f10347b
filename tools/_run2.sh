mkdir -p gpurun_out/r02a
for cfg in "10 4" "12 2"; do set -- $cfg
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --emit-path flat --ep-wf $1 --ep-copies-log2 $2"
$CMD > gpurun_out/r02a/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:hb_emitf -s 2 -c 1 -o gpurun_out/r02a/prof_emitf_$1_$2 -f $CMD > gpurun_out/r02a/ncu.log 2>&1; tail -1 gpurun_out/r02a/ncu.log
done
