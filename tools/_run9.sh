O=gpurun_out/r03g; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary"
for v in "" ; do
  echo "== english1g $v" >> $O/ab.log; $B $v >> $O/ab.log 2>&1
  echo "== fib4g $v" >> $O/ab.log; $B --workload fib4g $v >> $O/ab.log 2>&1
done
python tools/latency_probe.py > $O/latency.log 2>&1; tail -12 $O/latency.log
python - <<'PY'
import json
for l in open('gpurun_out/r03g/ab.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   ms/step %.4f  GB/s %.1f launches %s kernels %s' % (d['ms_per_step'], d['value'], d['gpu_launches'], d['roofline']['kernel_ms']))
    else: print('   '+l.strip()[:200])
PY
