O=gpurun_out/r03m; mkdir -p $O
for c in 4 8 16 32 64; do
  echo "== chunk $c MiB" >> $O/e2e.log
  python bench.py --steps 5 --warmup 3 --no-cpu --no-secondary --host-chunk-mib $c >> $O/e2e.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r03m/e2e.log'):
    if l.startswith('=='): print(l.strip()); continue
    if l.startswith('{'):
        d=json.loads(l); print('   e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    else: print('   '+l.strip()[:200])
PY
