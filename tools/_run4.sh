mkdir -p gpurun_out/r02c
(time timeout 1500 python -m pytest tests -x -q -m gpu) > gpurun_out/r02c/pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c/pytest_gpu.log; tail -6 gpurun_out/r02c/pytest_gpu.log
cd huffmandecoderongpus_b200/host && (timeout 300 ./HuffFramework bigtable; echo "rc $?") > ../../gpurun_out/r02c/harness_bigtable.log 2>&1; tail -6 ../../gpurun_out/r02c/harness_bigtable.log
(B200_PIN=0 timeout 300 ./HuffFramework kjv; echo "rc $?") 2>&1 | tail -2
