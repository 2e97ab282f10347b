mkdir -p gpurun_out/r02g
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02g/pytest_gpu.log 2>&1; tail -3 gpurun_out/r02g/pytest_gpu.log
timeout 600 python bench.py --steps 50 > gpurun_out/r02g/bench_default.log 2>&1; grep '^{' gpurun_out/r02g/bench_default.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('value',round(d['value'],1),'ms',round(d['ms_per_step'],4),d['roofline']['kernel_ms'],'frac',round(d['roofline']['decode_frac'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), {k:(round(v['value'],1),round(v['ms_per_step'],4),round(v['decode_frac_rank0'],3)) for k,v in d['secondary'].items() if 'value' in v})
"
