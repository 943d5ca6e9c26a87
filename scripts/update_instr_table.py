#!/usr/bin/env python
"""Maintains profiles/instr_per_item.json: executed thread instructions per item of a kernel, split by the pipe that
issues them, from an executed-instruction mix produced by scripts/ncu_dynamic_mix.py (itself read from an ncu --set full
--import-source on report).  bench.py combines these STATIC counts with times, clocks and pipe peaks measured in its run.
usage: python scripts/update_instr_table.py <key> <profiles/xxx_dynamic_mix.txt> [<key> <file> ...]"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATH = os.path.join(ROOT, "profiles", "instr_per_item.json")
FP = ("FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "FFMA2", "FADD2", "FMUL2")
IMAD = ("IMAD", "IDP")
ALU = ("LOP3", "SHF", "SEL", "FSEL", "IADD3", "VIADD", "LEA", "ISETP", "FSETP", "PLOP3", "PRMT", "VIMNMX", "IMNMX", "FMNMX", "MOV", "P2R", "R2P",
       "BMSK", "SGXT", "IABS", "HSETP", "HSET2", "HMNMX2", "FSET", "I2FP", "F2FP", "VABSDIFF", "FCHK")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "RED", "LDL", "STL", "UTMA", "SYNCS")

def classify(op):
    base = op.split(".")[0]
    if base in FP: return "fp"
    if base in IMAD: return "imad"
    if base in ALU: return "alu"
    if base in LSU or base.startswith("UTMA"): return "lsu"
    return "other"

def parse(path):
    tot, per = None, {"fp": 0.0, "imad": 0.0, "alu": 0.0, "lsu": 0.0, "other": 0.0}
    listed = 0.0
    for line in open(path):
        m = re.search(r"= ([0-9.]+) thread instructions per item", line)
        if m: tot = float(m.group(1))
        m = re.match(r"^([A-Z][A-Z0-9_.x]*)\s+([0-9.]+)\s+[0-9.]+\s*$", line)
        if m:
            per[classify(m.group(1))] += float(m.group(2)); listed += float(m.group(2))
    per["other"] += max(0.0, (tot or listed) - listed)     # opcodes beyond the 30 listed
    return tot or listed, per

tab = json.load(open(PATH)) if os.path.exists(PATH) else {}
args = sys.argv[1:]
for key, path in zip(args[0::2], args[1::2]):
    tot, per = parse(path)
    tab[key] = {"total": round(tot, 1), "fma": round(per["fp"] + per["imad"], 1), "fp32": round(per["fp"], 1), "imad": round(per["imad"], 1),
                "alu": round(per["alu"], 1), "lsu": round(per["lsu"], 1), "source": os.path.relpath(path, ROOT)}
    print(key, tab[key])
json.dump(tab, open(PATH, "w"), indent=1, sort_keys=True)
