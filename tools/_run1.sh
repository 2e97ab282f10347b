mkdir -p gpurun_out/r02a
for cfg in "words 0 -1" "flat 10 4" "flat 11 3" "flat 12 2" "flat 9 4"; do
  set -- $cfg
  timeout 300 python bench.py --workload fib4g --steps 10 --warmup 3 --no-cpu --no-e2e --emit-path $1 --ep-wf $2 --ep-copies-log2 $3 > gpurun_out/r02a/fib_$1_$2_$3.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02a/fib_$1_$2_$3.log) $(grep -o '"kernel_ms": {[^}]*}' gpurun_out/r02a/fib_$1_$2_$3.log)"
done
