for v in "" "VO_B200_LK_ORDER=1" "VO_B200_SEQ_HOST=1" "VO_B200_STEREO_FIRST=1"; do
  for r in 1 2; do
    env $v python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-matrix 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['e2e']['value'])"
  done
done
