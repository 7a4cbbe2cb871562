#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/make_profiles.sh TAG'): the ncu evidence behind bench.py's numbers.
#  1. bench.py without ncu (the numbers) ; 2. launch list of the same command (gpu__time_duration per launch) ;
#  3. one --set full capture of every kernel of one td3_hopper / sac_hopper iteration (cold-cache, serialised).
TAG=${1:-r1}
set -x
python bench.py > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
  python bench.py --steps 30 --warmup 3 --no-legs --no-torch-baseline --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
#    (the config-4 / config-5 legs of the same line have their own launch lists below and in ${TAG}_launches_pop1024*:
#     1024 agents' initialisation under ncu takes minutes)
for wl in td3_hopper sac_hopper; do
  python tools/profile_target.py $wl 12 > gpurun_out/${TAG}_plain_$wl.log 2>&1 || exit 1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'critic_fused|actor_fused|wgrad|adam_polyak|gather|alpha_kernel' -s 30 -c 12 \
    -o gpurun_out/${TAG}_full_$wl -f python tools/profile_target.py $wl 12 > gpurun_out/${TAG}_ncu_$wl.log 2>&1
done
ls -la gpurun_out | tail -12
# 4. the wide (tensor-core) path at batch 65 536: launch list of two DP iterations and one --set full capture of the
#    tcgen05 kernels (tc_linear forward / backward, tc_wgrad)
PYTHONPATH=. python tools/bench_dp.py 65536 20 3xtf32 > gpurun_out/${TAG}_plain_wide.log 2>&1 || exit 1
PYTHONPATH=. timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_wide.csv \
  python tools/bench_dp.py 65536 2 3xtf32 > gpurun_out/${TAG}_ncu_wide.log 2>&1
PYTHONPATH=. timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_linear_kernel|tc_wgrad_kernel' --launch-skip 40 -c 10 \
  -o gpurun_out/${TAG}_full_wide -f python tools/bench_dp.py 65536 2 3xtf32 > gpurun_out/${TAG}_ncu_wide_full.log 2>&1
ls -la gpurun_out | tail -8
# 5. the replay sampler alone at 2^20 rows per launch (roofline_gather.traffic)
PYTHONPATH=. timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' --launch-skip 3 -c 2 \
  -o gpurun_out/${TAG}_full_gather -f python tools/profile_target.py gather > gpurun_out/${TAG}_ncu_gather.log 2>&1
ls -la gpurun_out | tail -4
