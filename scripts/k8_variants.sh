# K8 (plan_rewire_kernel) variants: each librrtk_<tag>.so built by rrtplanner_b200.build with extra -D flags
for L in rrtplanner_b200/librrtk.so rrtplanner_b200/librrtk_*.so; do
  RRTK_LIB=$PWD/$L timeout 300 python bench.py --dubins-only --steps 3 --no-cpu --dubins-plans ${PLANS:-1024} 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$L', round(d['dubins_rrtstar']['plans_per_s']), round(d['euclid_rrtstar_with_rewire']['plans_per_s']), d['dubins_rrtstar']['overflow'], d['dubins_rrtstar']['blocks_per_sm'])"
done
