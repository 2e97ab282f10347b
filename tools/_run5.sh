mkdir -p gpurun_out/r02d
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
nvidia-smi topo -m > gpurun_out/r02d/topo.txt 2>&1
for n in 2 4 8; do
  [ $n -le $N ] || continue
  (time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 3) > gpurun_out/r02d/bench_n$n.log 2>&1; echo "rc $?" >> gpurun_out/r02d/bench_n$n.log
  grep '^{' gpurun_out/r02d/bench_n$n.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('N',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],4),'e2e',d['e2e'] and round(d['e2e']['value'],1), {k:(round(v['value'],1),round(v['ms_per_step'],4)) for k,v in d['secondary'].items() if 'value' in v})
"
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe.py > gpurun_out/r02d/pcie_n$n.log 2>&1; grep '^{' gpurun_out/r02d/pcie_n$n.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('pcie ranks',d['ranks'],{k:(round(v['ms'],1),round(v['aggregate_GBps'],1)) for k,v in d['results'].items()})
"
done
cd huffmandecoderongpus_b200/host
for n in 1 2 4 8; do [ $n -le $N ] || continue; (B200_DEVICES=$n timeout 600 ./HuffFramework synth16g; echo "rc $?") > ../../gpurun_out/r02d/harness_synth16g_$n.log 2>&1; head -2 ../../gpurun_out/r02d/harness_synth16g_$n.log | tail -1; done
for n in 2 4 8; do [ $n -le $N ] || continue; (B200_DEVICES=$n timeout 600 ./HuffFramework synth1g; echo "rc $?") > ../../gpurun_out/r02d/harness_synth1g_$n.log 2>&1; head -2 ../../gpurun_out/r02d/harness_synth1g_$n.log | tail -1; done
cd ../..; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02d/pytest_multi.log 2>&1; tail -2 gpurun_out/r02d/pytest_multi.log
