O=gpurun_out/r03u; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary"
echo "== english1g" >> $O/ab.log; $B >> $O/ab.log 2>&1
echo "== fib4g" >> $O/ab.log; $B --workload fib4g >> $O/ab.log 2>&1
echo "== fib16g" >> $O/ab.log; $B --workload fib16g --steps 5 >> $O/ab.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r03u/ab.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GB/s %.1f launches %s kernels %s frac %.3f' % (d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_ms'], d['roofline']['decode_frac']))
    else: print('   '+l.strip()[:200])
PY
