#!/bin/bash
# One measurement session on the GPU box (run through gpurun from the repo root):
#   tests -> bench (ours + reference arm) -> ncu launch list -> ncu --set full of one launch of every hot kernel.
# Outputs land in gpurun_out/ with the tag given as $1 (e.g. r02); tools/refresh_profiles.py <tag> turns them into profiles/.
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gputest.log 2>&1
echo "pytest exit $?" >> $OUT/${TAG}_gputest.log
tail -3 $OUT/${TAG}_gputest.log
python bench.py --steps 20 --warmup 3 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err || { echo bench failed; tail -20 $OUT/bench_${TAG}.err; exit 1; }
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_${TAG}_reference.json 2>> $OUT/bench_${TAG}.err
python tools/bench_configs.py > $OUT/configs_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
# one launch of each kernel of the device-resident step: the step is pack, leg_syn, lon_syn, lon_ana, leg_ana, unpack;
# skip the first step (cold), capture the second
ncu --set full --clock-control none --import-source on --launch-skip 6 --launch-count 6 -f -o $OUT/prof_${TAG} \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT
