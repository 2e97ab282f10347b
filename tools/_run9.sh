O=gpurun_out/r03l; mkdir -p $O
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary"
for v in "--emit-path words --ep-wf 10 --ep-copies-log2 3" "--emit-path words --ep-wf 10 --ep-copies-log2 2" "--emit-path words --ep-wf 9 --ep-copies-log2 3"; do
  echo "== fib4g $v" >> $O/ab.log; $B --workload fib4g $v >> $O/ab.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r03l/ab.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GB/s %.1f launches %s kernels %s' % (d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_ms']))
    else: print('   '+l.strip()[:200])
PY
