O=gpurun_out/r04f; mkdir -p $O
HB_STRESS_SEEDS=16 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -3 $O/pytest.log
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3) > $O/bench_n2.log 2>&1; echo "rc $?" >> $O/bench_n2.log
grep '^{' $O/bench_n2.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('N',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), {k:(round(v['value'],1),round(v['ms_per_step'],4)) for k,v in d['secondary'].items() if 'value' in v})
"
tail -3 $O/bench_n2.log
cd huffmandecoderongpus_b200/host
(B200_DEVICES=2 timeout 600 ./HuffFramework synth1g; B200_DEVICES=2 timeout 600 ./HuffFramework synth16g) > ../../$O/harness_synth_2.log 2>&1; grep "b200" ../../$O/harness_synth_2.log
