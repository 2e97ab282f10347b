mkdir -p gpurun_out/r02f
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r02f/pytest_gpu.log 2>&1; tail -3 gpurun_out/r02f/pytest_gpu.log
timeout 900 python tools/slow_path_probe.py > gpurun_out/r02f/slow_paths.txt 2>&1; cat gpurun_out/r02f/slow_paths.txt
cd huffmandecoderongpus_b200/host
for n in 1 2; do (B200_DEVICES=$n timeout 300 ./HuffFramework synth1g; echo "rc $?") 2>&1 | sed -n 2p; done
(HB_MULTI_THREADS=0 B200_DEVICES=2 timeout 300 ./HuffFramework synth1g) 2>&1 | sed -n 2p
(timeout 300 ./HuffFramework multi) 2>&1 | tail -3
