"""Attribute the warp-stall samples and executed instructions of an ncu capture to CUDA source lines.

    ncu -i capture.ncu-rep --page source --csv > sass.csv            # per-SASS-instruction samples / executed counts
    nvcc <the flags of build.py> -cubin -o k.cubin csrc/vine_b200.cu   # the same source revision as the capture
    nvdisasm -g -c k.cubin > k.sass                                   # SASS with '//## File "...", line N' markers
    python tools/ncu_source_lines.py sass.csv k.sass KERNEL_SUBSTRING [top]

The two instruction streams must come from the same source and compiler (the script checks that the opcode sequences agree).
Prints, per source line: share of the stall samples, share of the executed warp-instructions, number of SASS instructions.
"""
import csv
import re
import sys
from collections import defaultdict


def disassembly(path, kernel):
    lines = open(path).read().split("\n")
    starts = [i for i, l in enumerate(lines) if l.startswith(".text.")]
    begin = [i for i in starts if kernel in lines[i]][0]
    end = min([i for i in starts if i > begin] + [len(lines)])
    cur, out = None, []
    for l in lines[begin:end]:
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out.append((m.group(2).strip(), cur))
    return out


def opcode(text):
    text = re.sub(r"^@!?U?P\d+\s+", "", text.strip())
    return text.split()[0].split(".")[0]


def main():
    sass_csv, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = [r for r in csv.reader(open(sass_csv)) if len(r) > 5 and r[0].startswith("0x")]
    ins = disassembly(dis, kernel)
    assert len(rows) == len(ins), (len(rows), len(ins))
    assert all(opcode(a[1]) == opcode(b[0]) for a, b in zip(rows, ins)), "different builds"
    samples, executed, count = defaultdict(int), defaultdict(int), defaultdict(int)
    for r, (_, loc) in zip(rows, ins):
        samples[loc] += int(r[2]); executed[loc] += int(r[5]); count[loc] += 1
    ts, te = sum(samples.values()), sum(executed.values())
    print(f"{kernel}: {len(rows)} SASS instructions, {ts} stall samples, {te} executed warp-instructions")
    cache = {}
    for loc, s in sorted(samples.items(), key=lambda kv: -kv[1])[:top]:
        text = ""
        if loc:
            try:
                src = cache.setdefault(loc[0], open(loc[0]).read().split("\n"))
                text = src[loc[1] - 1].strip()[:100]
            except OSError:
                pass
        name = f"{loc[0].split('/')[-1]}:{loc[1]}" if loc else "?"
        print(f"{100 * s / ts:5.1f} % samples {100 * executed[loc] / te:5.1f} % executed  {count[loc]:4d} SASS  {name:24s} {text}")


if __name__ == "__main__":
    main()
