O=gpurun_out/r03s; mkdir -p $O
HB_STRESS_SEEDS=30 HB_STRESS_CASES=16 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -4 $O/pytest.log
