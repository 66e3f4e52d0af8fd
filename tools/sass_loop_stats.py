"""Static per-warp cost of the loops of a kernel from its SASS (no GPU needed): for every backward branch, the number
of instructions, FP64 instructions, register moves and the sum of the stall counts encoded in the control bits (the
cycles ONE warp needs to issue the loop body once, before any scoreboard wait).  With two warps per scheduler the
on-chip CG kernels are bound by this in-order stream, not by the FP64 pipe (profiles/README.md, r02).

    python tools/sass_loop_stats.py thirring2d_b200/csrc/build/tb_resident.o resident_wt_kernelILi64ELi64ELb1ELb0ELb0ELb1E
"""
import re
import subprocess
import sys


def instructions(obj, pattern):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.split("\n")
    out, on, i = [], False, 0
    while i < len(sass):
        ln = sass[i]
        if "Function :" in ln:
            on = pattern in ln
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", ln) if on else None
        if m and i + 1 < len(sass):
            m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", sass[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                out.append((int(m.group(1), 16), m.group(2).strip(), (hi >> 41) & 0xF))
                i += 2
                continue
        i += 1
    return out


def main():
    obj, pattern = sys.argv[1], sys.argv[2]
    ins = instructions(obj, pattern)
    print(f"{pattern}: {len(ins)} instructions")
    for addr, text, _ in ins:
        if "BRA" in text:
            t = re.search(r"0x([0-9a-f]+)", text)
            if t and int(t.group(1), 16) < addr:
                tgt = int(t.group(1), 16)
                seg = [y for y in ins if tgt <= y[0] <= addr]
                if len(seg) < 200:
                    continue
                cnt = lambda rx: sum(1 for y in seg if re.search(rx, y[1]))
                print(f"  loop {tgt:#x}..{addr:#x}: {len(seg)} instr, FP64 {cnt(r'(DFMA|DMUL|DADD)')}, moves "
                      f"{cnt(r'(IMAD.MOV|^MOV)')}, LDS/STS {cnt(r'(LDS|STS)')}, LDTM/STTM {cnt(r'(LDTM|STTM)')}, "
                      f"local LDL/STL {cnt(r'(LDL|STL)')}, sum of stall counts {sum(y[2] for y in seg)}")


if __name__ == "__main__":
    main()
