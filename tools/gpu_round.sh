#!/bin/bash
# One GPU round on a fresh box: parity tests, the bench (both arms), a launch list and one
# `ncu --set full` capture of a steady-state tick.  Usage: tools/gpu_round.sh TAG [notest]
# Everything lands in gpurun_out/TAG_*.
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "notest" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  tail -3 $OUT/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2>> $OUT/${TAG}_bench.err
# launch list: 400 steady-state launches
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 3000 -c 400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --kernel-only --steps 1 --warmup 1 > $OUT/${TAG}_ncu1.log 2>&1
# full capture of one steady-state tick at 4096 trials
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 3000 -c 12 \
   -o $OUT/${TAG}_prof -f python bench.py --kernel-only --trials 4096 --steps 1 --warmup 1 > $OUT/${TAG}_ncu2.log 2>&1
ls -la $OUT | grep ${TAG}
# launch lists of the secondary workloads (wide LU path): 1000-junction line and the 100k-MOSFET deck
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 400 --csv \
   --log-file $OUT/${TAG}_rsfq_launches.csv python -c "import bench; bench.bench_rsfq(0)" > $OUT/${TAG}_ncu3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 -c 300 --csv \
   --log-file $OUT/${TAG}_large_launches.csv python tools/t_large.py 50001 1 1 > $OUT/${TAG}_ncu4.log 2>&1
python tools/launch_summary.py $OUT/${TAG}_rsfq_launches.csv > $OUT/${TAG}_rsfq_launch_summary.txt 2>&1
python tools/launch_summary.py $OUT/${TAG}_large_launches.csv > $OUT/${TAG}_large_launch_summary.txt 2>&1
ls -la $OUT | grep ${TAG}
