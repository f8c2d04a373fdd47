RRTK_GRID_BSY=4 RRTK_PLAN_CAP=192 bash scripts/variants.sh 0 g8 main 2>&1 | tail -2
RRTK_GRID_BSY=5 RRTK_PLAN_CAP=160 bash scripts/variants.sh 0 g8 2>&1 | tail -1
