#!/usr/bin/env python
"""Write the SASS of the kernels the round's claims rest on into profiles/sass/ (one file per kernel, encodings stripped),
with an index of their instruction mix and the loop tables of tools/sass_loops.py.  Run after `make -C nvspeechplayer_b200/csrc`."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "nvspeechplayer_b200", "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = {
    "klatt_f32_sched": ["klatt_f32_sched_kernel"],
    "klatt_f32": ["klatt_f32_hold_kernel", "klatt_f32_general_pair_kernel", "klatt_plan_kernel"],
    "klatt_f32_block": ["klatt_f32_block_kernel"],
    "klatt_pull": ["klatt_pull_kernel"],
    "klatt_long": ["klatt_long_stage_kernelILi2ELi2", "klatt_long_stage_kernelILi0ELi2", "klatt_long_stage_kernelILi3ELi2",
                   "klatt_long_spec_kernel", "klatt_long_timeline_kernel"],
    "klatt_f64": ["klatt_batch_f64_kernel"],
}
os.makedirs(OUT, exist_ok=True)
index = []
for obj, names in WANT.items():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj + ".o")], capture_output=True, text=True).stdout
    loops = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loops.py"), ""], input=txt, capture_output=True, text=True).stdout
    for part in re.split(r"(?=\t\tFunction : )", txt):
        m = re.match(r"\t\tFunction : (\S+)", part)
        if not m:
            continue
        fn = m.group(1)
        for nm in names:
            if nm not in fn:
                continue
            base = "%s_%s" % (TAG, nm.replace("ILi", "_").replace("ELi", "_"))
            lines = [l.rstrip() for l in part.splitlines() if not re.match(r"^\s*/\* 0x[0-9a-f]+ \*/\s*$", l)]
            lines = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in lines]
            open(os.path.join(OUT, base + ".sass"), "w").write("\n".join(lines) + "\n")
            ops = {}
            n = 0
            for l in lines:
                mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
                if mm:
                    n += 1
                    op = mm.group(1).split(".")[0]
                    ops[op] = ops.get(op, 0) + 1
            lt = [b for b in loops.split("== ") if b.startswith(fn)]
            if lt:
                open(os.path.join(OUT, base + ".loops.txt"), "w").write("== " + lt[0])
            index.append((base + ".sass", fn, n, ops))
with open(os.path.join(OUT, "INDEX.md"), "w") as f:
    f.write("# profiles/sass -- `cuobjdump -sass` of the built objects (sm_100a), one file per kernel, encodings stripped\n\n"
            "Regenerate with `python tools/dump_sass.py` after `make -C nvspeechplayer_b200/csrc`; `*.loops.txt` = `tools/sass_loops.py` "
            "(every backward branch closes a loop: address range, size, instruction mix).\n"
            "No `UTCMMA` / `UTMA*` is expected anywhere: the path is not a contraction (the north star rules tensor cores out); what "
            "to look for is `FFMA2 / FMUL2 / FADD2` (packed FP32 pairs), the FP64 oscillator chain (`DADD / DMUL / DFMA`) and the absence "
            "of local-memory traffic (`LDL / STL`) in the loops.\n\n"
            "| file | kernel | instr | FFMA2 | FMUL2 | FADD2 | FFMA | FADD | FMUL | DFMA+DADD+DMUL | IMAD | LDL+STL | bytes |\n|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for name, fn, n, ops in index:
        g = lambda k: ops.get(k, 0)
        f.write("| `%s` | `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n" % (
            name, fn[:48], n, g("FFMA2"), g("FMUL2"), g("FADD2"), g("FFMA"), g("FADD"), g("FMUL"), g("DFMA") + g("DADD") + g("DMUL"), g("IMAD"),
            g("LDL") + g("STL"), n * 16))
print(open(os.path.join(OUT, "INDEX.md")).read())
