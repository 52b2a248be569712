"""Static instruction mix of the largest loop body of one kernel in an object file (cuobjdump -sass): a GPU-free proxy for the
executed warp instructions per iteration.  usage: python scripts/sass_loop_count.py <obj> <kernel-name-substring>"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    loops = []
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr:
                loops.append((tgt, addr))
    loops.sort(key=lambda l: l[0] - l[1])
    print("== %s: %d instructions" % (name, len(ins)))
    for lo, hi in loops[:3]:                      # the three largest loops (outer frame loop, window loop, ...)
        body = [t for a, t in ins if lo <= a <= hi]
        ops = collections.Counter()
        for t in body:
            m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", t); ops[m.group(2) if m else "?"] += 1
        print("   loop [%#x, %#x]: %d instructions" % (lo, hi, len(body)))
        print("      " + ", ".join("%s %d" % kv for kv in ops.most_common(20)))
