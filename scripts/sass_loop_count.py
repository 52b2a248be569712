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
    best = None
    for addr, text in ins:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= addr and (best is None or addr - tgt > best[1] - best[0]):
                best = (tgt, addr)
    print("== %s: %d instructions, largest loop [%#x, %#x]" % (name, len(ins), best[0], best[1]))
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    ops = collections.Counter()
    for t in body:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", t); ops[m.group(2) if m else "?"] += 1
    print("   loop body: %d instructions" % len(body))
    print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(20)))
