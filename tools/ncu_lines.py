"""Join an ncu SASS-level source page (csv) with nvdisasm -g line info: per source line totals of
instructions executed and stall samples.  usage: ncu_lines.py sass.csv dis.txt kernel_substr [file_filter]"""
import csv, re, sys, collections
sass_csv, dis, ksub = sys.argv[1:4]
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ia, isrc, isamp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0)) for r in rows[2:] if len(r) > iex]
# parse the disassembly of the kernel
cur_fn = None; line = None; fname = None; dis_ins = []
inl = None
for l in open(dis):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur_fn = m.group(1); continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        fname, line = m.group(1).split('/')[-1], int(m.group(2)); inl = m.group(3)
        # attribute to the outermost call site (last "inlined at")
        mm = re.findall(r'inlined at "([^"]+)", line (\d+)', inl)
        if mm: fname, line = mm[-1][0].split('/')[-1], int(mm[-1][1])
        continue
    if cur_fn and ksub in cur_fn:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m: dis_ins.append((m.group(2).strip(), fname, line, inl))
print(len(ins), len(dis_ins), file=sys.stderr)
n = min(len(ins), len(dis_ins))
agg = collections.defaultdict(lambda: [0, 0])
for k in range(n):
    key = (dis_ins[k][1], dis_ins[k][2])
    agg[key][0] += ins[k][2]; agg[key][1] += ins[k][1]
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print("total inst", tot_i, "samples", tot_s)
for key in sorted(agg):
    v = agg[key]
    if v[0] > tot_i * 0.002 or v[1] > tot_s * 0.002:
        print(f"{key[0]}:{key[1]:4d}  inst {v[0]:10d} {100*v[0]/tot_i:5.1f}%   samples {v[1]:7d} {100*v[1]/tot_s:5.1f}%")
