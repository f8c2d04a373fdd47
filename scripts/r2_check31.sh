bash scripts/variants.sh 0 main 2>&1 | tail -1
for by in 3 4 5; do echo bsy $by; RRTK_GRID_BSY=$by bash scripts/variants.sh 0 main 2>&1 | tail -1; done
for cap in 192 160; do echo cap $cap; RRTK_PLAN_CAP=$cap bash scripts/variants.sh 0 main 2>&1 | tail -1; done
echo cap 192 bsy 4; RRTK_PLAN_CAP=192 RRTK_GRID_BSY=4 bash scripts/variants.sh 0 main 2>&1 | tail -1
