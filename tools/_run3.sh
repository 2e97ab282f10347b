mkdir -p gpurun_out/r02b
nvidia-smi -L > gpurun_out/r02b/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r02b/pytest_multi.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b/pytest_multi.log; tail -15 gpurun_out/r02b/pytest_multi.log
cd huffmandecoderongpus_b200/host
(timeout 300 ./HuffFramework multi; echo "rc $?") > ../../gpurun_out/r02b/harness_multi.log 2>&1; tail -12 ../../gpurun_out/r02b/harness_multi.log
(timeout 300 ./HuffFramework onethread; echo "rc $?") > ../../gpurun_out/r02b/harness_onethread.log 2>&1; tail -4 ../../gpurun_out/r02b/harness_onethread.log
(B200_DEVICES=2 timeout 300 ./HuffFramework synth1g; echo "rc $?") > ../../gpurun_out/r02b/harness_synth1g_2.log 2>&1; tail -5 ../../gpurun_out/r02b/harness_synth1g_2.log
(B200_DEVICES=1 timeout 300 ./HuffFramework synth1g; echo "rc $?") > ../../gpurun_out/r02b/harness_synth1g_1.log 2>&1; tail -3 ../../gpurun_out/r02b/harness_synth1g_1.log
(timeout 300 ./HuffFramework bigtable; echo "rc $?") > ../../gpurun_out/r02b/harness_bigtable.log 2>&1; tail -7 ../../gpurun_out/r02b/harness_bigtable.log
