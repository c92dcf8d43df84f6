for S in 2 3 4; do for CH in 32768 65536 131072 262144; do
  V=$(MLKEM_B200_HOST_SLOTS=$S MLKEM_B200_HOST_CHUNK=$CH python bench.py --steps 2 --log2-items 20 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e %.3f M pairs/s  %.2f ms  (H2D %.1f GB/s)'%(d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/d['e2e']['ms_per_step']/1e6))")
  echo "host_slots=$S host_chunk=$CH : $V"
done; done
