#!/bin/bash
# Round-2 evidence run on one B200 (called through gpurun): the bench line, the reference arm, and -- only after the same
# command has exited 0 without ncu -- the launch list of ONE timed-shape step and full captures of the dominant kernels.
# Everything lands in gpurun_out/; scripts/refresh_profiles.py turns it into the files under profiles/.
set -u
O=gpurun_out
mkdir -p $O
python bench.py --impl reference --steps 2 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err; echo "reference arm rc=$?"
python bench.py > $O/r02_bench_1gpu.json 2> $O/r02_bench_1gpu.err; echo "bench rc=$?"
python bench.py --steps 1 --warmup 2 --profile-step > /dev/null 2> $O/r02_profile_step.err; rc=$?; echo "profile-step plain rc=$rc"
[ $rc -eq 0 ] || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file $O/r02_launches_bench_step.csv python bench.py --steps 1 --warmup 2 --profile-step > $O/r02_ncu_launches.log 2>&1; echo "ncu launch list rc=$?"
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:ppp_prefilter2 --launch-skip 15 --launch-count 3 \
    -f -o $O/r02_prefilter2_bench python bench.py --steps 1 --warmup 2 --profile-step > $O/r02_ncu_prefilter2.log 2>&1; echo "ncu prefilter2 rc=$?"
ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:ppp_refine --launch-skip 8 --launch-count 1 \
    -f -o $O/r02_refine_bench python bench.py --steps 1 --warmup 2 --profile-step > $O/r02_ncu_refine.log 2>&1; echo "ncu refine rc=$?"
ls -la $O | tail -20
