# A/B of an environment switch on the headline bench: bash tools/ab_env.sh VAR=1 [steps]
V="$1"; S="${2:-20}"
for v in "" "$V" "" "$V" "" "$V"; do
  env $v python bench.py --steps $S --warmup 5 --no-cpu-baseline --no-matrix 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v]', d['value'], d['e2e']['value'])"
done
