O=gpurun_out/r03j; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; tail -5 $O/pytest.log
python tools/slow_path_probe.py > $O/slow_paths.txt 2>&1; cat $O/slow_paths.txt
