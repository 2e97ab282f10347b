O=gpurun_out/r03k; mkdir -p $O
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
n=2
(time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3) > $O/bench_n$n.log 2>&1; echo "rc $?" >> $O/bench_n$n.log
grep '^{' $O/bench_n$n.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('N',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), {k:(round(v['value'],1),round(v['ms_per_step'],4)) for k,v in d['secondary'].items() if 'value' in v})
"
tail -3 $O/bench_n$n.log
cd huffmandecoderongpus_b200/host
(B200_DEVICES=2 timeout 600 ./HuffFramework synth1g; echo "rc $?") > ../../$O/harness_synth1g_2.log 2>&1; head -3 ../../$O/harness_synth1g_2.log
cd ../..; timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_ref_harness.py -x -q -m gpu > $O/pytest_multi.log 2>&1; tail -2 $O/pytest_multi.log
