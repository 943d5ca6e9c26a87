#!/bin/bash
# round-2 ncu captures on the GPU box: the fused kernels (r02g) and the sweep kernels (r02h), each after a plain run.
# The reports are summarised ON THE BOX (text files under gpurun_out/) and then deleted: together they exceed what a call
# may bring back.
set -u
OUT=gpurun_out; mkdir -p $OUT
python scripts/r02_profile_target.py > $OUT/r02g_plain.log 2>&1 || { echo plain failed; tail -5 $OUT/r02g_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'prove_f32_tma_kernel|verify_tma_kernel|prove_fs_kernel|verify_fs_kernel' -c 14 -f -o /tmp/r02g_fused python scripts/r02_profile_target.py > $OUT/r02g_ncu.log 2>&1
echo "fused ncu rc=$?"
# the source page holds two sections per captured launch; pair i of the target = launches 2i, 2i+1 = sections 4i .. 4i+3
i=0
for name in prove_table verify_table prove_arith verify_arith prove_fs verify_fs prove_table_uniform; do
  python scripts/ncu_dynamic_mix.py /tmp/r02g_fused.ncu-rep 1048576 $((4 * i + 2)) > $OUT/r02g_${name}_dynamic_mix.txt 2>> $OUT/r02g_ncu.log
  i=$((i + 1))
done
python scripts/summarize_sweeps_ncu.py /tmp/r02g_fused.ncu-rep 1048576 "fused kernels, 2^20 items; rows are the LAST launch of each kernel name (prove TABLE: the D_uniform launch)" > $OUT/r02g_fused_summary.txt 2>> $OUT/r02g_ncu.log
ncu -i /tmp/r02g_fused.ncu-rep --page raw --csv > $OUT/r02g_fused_raw.csv 2>> $OUT/r02g_ncu.log
python scripts/sweep_profile_target.py 22 > $OUT/r02h_plain.log 2>&1 || { echo sweeps plain failed; tail -5 $OUT/r02h_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'g1_smul|pairing_kernel|poly_mul|kzg_commit_kernel|poly_unary|ntt4_kernel' -c 60 -f -o /tmp/r02h_sweeps python scripts/sweep_profile_target.py 22 > $OUT/r02h_ncu.log 2>&1
echo "sweeps ncu rc=$?"
python scripts/summarize_sweeps_ncu.py /tmp/r02h_sweeps.ncu-rep 4194304 "round-2 sweep kernels (2^22 items: inputs fit L2)" > $OUT/r02h_sweeps_after.txt 2>> $OUT/r02h_ncu.log
# launch list of the bench command (shares of the step's kernels)
python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --ring 2 --e2e-steps 0 --no-arith --no-fs --no-uniform --no-sweeps --no-config2 --no-config4 --reps 1 > $OUT/r02i_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02i_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --ring 2 --e2e-steps 0 --no-arith --no-fs --no-uniform --no-sweeps --no-config2 --no-config4 --reps 1 > $OUT/r02i_ncu.log 2>&1
echo "launch list rc=$?"
ls -la $OUT | grep r02
