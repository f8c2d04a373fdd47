mkdir -p gpurun_out
for P in 512 1036; do for T in 128 256; do
RRTK_LIB=$PWD/exp_clk.so timeout 300 python scripts/phase_clocks.py $P $T 2>&1 | tail -1
done; done
./scripts/micro/scatter
