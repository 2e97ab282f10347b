O=gpurun_out/r04d; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
B="python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e --no-secondary"
rm -f $O/ab.log
for v in "" "--emit-path words64w" "--emit-path words64w --ep-wf 12 --ep-copies-log2 1" "--emit-path words64w --ep-wf 10 --ep-copies-log2 3" "--emit-path words64w --ep-wf 11 --ep-copies-log2 1"; do
echo "== fib4g $v" >> $O/ab.log; $B --workload fib4g $v >> $O/ab.log 2>&1
done
echo "== english1g --emit-path words64w" >> $O/ab.log; $B --emit-path words64w >> $O/ab.log 2>&1
echo "== english1g --emit-path words64w 12 1" >> $O/ab.log; $B --emit-path words64w --ep-wf 12 --ep-copies-log2 1 >> $O/ab.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r04d/ab.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GB/s %.1f launches %s kernels %s frac %.3f' % (d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['decode_frac']))
    else: print('   '+l.strip()[:200])
PY
