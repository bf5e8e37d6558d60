# A/B of the SM partition on the headline bench: bash tools/ab_island.sh
for v in "" "VO_B200_ISLAND=8" "VO_B200_ISLAND=16" "" "VO_B200_ISLAND=8" "VO_B200_ISLAND=16"; do
  env $v VO_B200_DEBUG_FALLBACK=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-matrix 2>gpurun_out/island.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', d['value'], d['e2e']['value'], d['oracle_parity'] if 'oracle_parity' in d else '')"
  grep -m1 "partition" gpurun_out/island.err; tail -2 gpurun_out/island.err | cut -c1-200
done
