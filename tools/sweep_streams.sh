for S in 2 3 4; do for CH in 65536 131072 262144 524288; do
  V=$(MLKEM_B200_STREAMS=$S MLKEM_B200_CHUNK=$CH python bench.py --steps 3 --no-extras --no-cpu-baseline --e2e-log2-items 16 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.3f M pairs/s  %.2f ms'%(d['value']/1e6, d['ms_per_step']))")
  echo "streams=$S chunk=$CH : $V"
done; done
