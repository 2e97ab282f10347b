O=gpurun_out/r04e; mkdir -p $O
B="python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-secondary"
rm -f $O/ab.log
for v in "--emit-path words64w --ep-wf 14 --ep-copies-log2 0" "--emit-path words64w --ep-wf 13 --ep-copies-log2 1" "--emit-path words64w --ep-wf 13 --ep-copies-log2 0"; do
echo "== english1g $v" >> $O/ab.log; $B $v >> $O/ab.log 2>&1
done
echo "== fib4g" >> $O/ab.log; $B --workload fib4g >> $O/ab.log 2>&1
echo "== fib16g" >> $O/ab.log; $B --workload fib16g --steps 10 >> $O/ab.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r04e/ab.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GB/s %.1f launches %s kernels %s frac %.3f' % (d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['decode_frac']))
    else: print('   '+l.strip()[:200])
PY
