O=gpurun_out/r03o; mkdir -p $O
CMD1="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary --emit-path words32w"
$CMD1 > $O/plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"hb_emit32w" -s 1 -c 1 -o $O/prof_e32w -f $CMD1 > $O/ncu_full.log 2>&1
tail -2 $O/ncu_full.log
