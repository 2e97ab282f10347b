mkdir -p gpurun_out/r02e
timeout 300 python -m pytest tests/test_gpu_ref_harness.py -x -q -m gpu > gpurun_out/r02e/pytest_refharness.log 2>&1; tail -3 gpurun_out/r02e/pytest_refharness.log
oracle/_ref/ref_harness oracle/_ref/files > gpurun_out/r02e/ref_harness.log 2>&1; cat gpurun_out/r02e/ref_harness.log
timeout 120 python tools/sanitize_cases.py > gpurun_out/r02e/sanitize_plain.log 2>&1; echo "plain rc $?"; tail -3 gpurun_out/r02e/sanitize_plain.log
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/r02e/memcheck.log python tools/sanitize_cases.py > gpurun_out/r02e/memcheck_stdout.log 2>&1; echo "memcheck rc $?"; tail -5 gpurun_out/r02e/memcheck.log
