"""Attribute an ncu source-page CSV (SASS rows with 'Instructions Executed' / '# Samples') to CUDA
source lines, using the line markers of `nvdisasm -g -c <cubin>` for the same kernel (instruction
order is identical).  usage: sass_by_line.py <ncu_source.csv> <nvdisasm.sass> <kernel substring>"""
import collections
import csv
import re
import sys


def main(src_csv, sass, kernel):
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    # nvdisasm: find the function, then collect (line, opcode) in order
    lines = open(sass).read().split("\n")
    start = next(i for i, l in enumerate(lines) if ".text." in l and kernel in l and l.strip().startswith(".section"))
    cur_line, seq = None, []
    for l in lines[start + 1:]:
        if l.strip().startswith(".section") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
        if m:
            seq.append((cur_line, m.group(1).strip()))
    print("ncu rows %d, nvdisasm instructions %d" % (len(data), len(seq)))
    n = min(len(data), len(seq))
    inst = collections.Counter()
    samp = collections.Counter()
    for k in range(n):
        line = seq[k][0]
        inst[line] += int(data[k][ix["Instructions Executed"]])
        samp[line] += int(data[k][ix["# Samples"]])
    ti, ts = sum(inst.values()), sum(samp.values())
    print("%-28s %14s %7s %10s %7s" % ("file:line", "warp-inst", "share", "samples", "share"))
    for line, v in sorted(inst.items(), key=lambda x: -x[1])[:45]:
        print("%-28s %14d %6.1f%% %10d %6.1f%%" % ("%s:%s" % line if line else "?", v, 100.0 * v / ti, samp[line], 100.0 * samp[line] / ts))


if __name__ == "__main__":
    main(*sys.argv[1:4])
