import sys; sys.path.insert(0,'plonk-by-fingers_b200/python')
import pbh_b200
ctx=pbh_b200.Context()
names=["imad","lop3_iadd3","half_imad_half_alu","ffma","hfma2_instr","dp4a_instr","imad_hi_iadd","half_ffma_half_imad","ffma_3reg","imad_3reg","ffma2_3pair_instr","ffma2_bcast_instr","shf","imad_wide","lop3","iadd3","prmt"]
only=[int(a) for a in sys.argv[1:]]
for i,n in enumerate(names):
    if only and i not in only: continue
    v=ctx.measure_int32_peak(i); print(f"{n:24s} {v/1e12:7.2f} T/s  {v/148/1.965e9:6.1f} per clk per SM")
