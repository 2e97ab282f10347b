# usage: bash tools/multi_push_ab.sh N   -- the C driver's map exchange: hb_shard_exchange (default), push kernel + events (1), peer copies (0)
n=${1:-2}; cd huffmandecoderongpus_b200/host
for p in 2 1 0; do
echo "== HB_MULTI_PUSH=$p, $n GPUs"
(HB_MULTI_PUSH=$p B200_DEVICES=$n timeout 600 ./HuffFramework synth1g; HB_MULTI_PUSH=$p B200_DEVICES=$n timeout 600 ./HuffFramework synth16g) 2>&1 | grep b200
done
