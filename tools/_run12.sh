O=gpurun_out/r04c; mkdir -p $O
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary"
$CMD1 > $O/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"hb_scan" -s 6 -c 2 -o $O/prof_scan -f $CMD1 > $O/ncu_full.log 2>&1
tail -3 $O/ncu_full.log
