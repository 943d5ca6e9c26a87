#!/bin/bash
# end-of-round-2 ncu captures: the step's two kernels with their fused summaries, the packed-format conversion kernels, and
# the launch list of the bench command.  Each capture follows a plain run of the same command that exited 0.
set -u
OUT=gpurun_out; mkdir -p $OUT
python scripts/r02p_profile_target.py > $OUT/r02p_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/r02p_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'prove_f32_tma_kernel|verify_tma_kernel|unpack_witness_kernel|pack_proof_kernel|unpack_proof_kernel' -c 10 -f -o /tmp/r02p python scripts/r02p_profile_target.py > $OUT/r02p_ncu.log 2>&1
echo "ncu rc=$?"
python scripts/ncu_dynamic_mix.py /tmp/r02p.ncu-rep 1048576 2 > $OUT/r02p_prove_digest_dynamic_mix.txt 2>> $OUT/r02p_ncu.log
python scripts/ncu_dynamic_mix.py /tmp/r02p.ncu-rep 1048576 6 > $OUT/r02p_verify_bitmap_dynamic_mix.txt 2>> $OUT/r02p_ncu.log
python scripts/summarize_sweeps_ncu.py /tmp/r02p.ncu-rep 1048576 "end of round 2: prove + fused digest, verify + fused bitmap, packed-format conversions; 2^20 items; rows are the LAST launch of each kernel name" > $OUT/r02p_summary.txt 2>> $OUT/r02p_ncu.log
ncu -i /tmp/r02p.ncu-rep --page raw --csv > $OUT/r02p_raw.csv 2>> $OUT/r02p_ncu.log
B="--steps 3 --warmup 3 --no-cpu --no-graph --ring 2 --e2e-steps 0 --no-arith --no-fs --no-uniform --no-sweeps --no-config2 --no-config4 --reps 1"
python bench.py $B > $OUT/r02p_bench_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02p_launches.csv python bench.py $B > $OUT/r02p_launches_ncu.log 2>&1
echo "launch list rc=$?"
cat $OUT/r02p_summary.txt
