"""Per-source-line view of one kernel in an .ncu-rep (read with `ncu -i`, no GPU needed): joins the SASS page of the report
(instructions executed, stall samples, shared-memory wavefronts per SASS instruction) with the line table of the same kernel in the
built object (`cuobjdump -xelf` + `nvdisasm -g`), then sums per source line and per phase of bands_fused_kernel.

usage: python scripts/ncu_by_line.py gpurun_out/prof_extract.ncu-rep lbaudiodetective_b200/build/lbad_extract.cu.o bands_fused_kernelILi32ELb1ELb1 24320000
       (report, object file, mangled-name fragment of the kernel, units per launch — here windows)"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, frag, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith("_ZN") and frag in l)
cur, off2line = None, {}
for l in dis[start:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        off2line[int(m.group(1), 16)] = cur
    elif l.startswith("//-----") and off2line:
        break
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ia, iex, ism, iwf = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
base = int(data[0][ia], 16)
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in data:
    a = agg[off2line.get(int(r[ia], 16) - base)]
    a[0] += int(r[iex]); a[1] += int(r[ism]); a[2] += int(r[iwf] or 0)
tot = [sum(a[i] for a in agg.values()) for i in range(3)]
print("per %s: %.1f warp instructions, %.1f shared-memory wavefronts (LDS/STS only); %d stall samples" % ("unit", tot[0] / units, tot[2] / units, tot[1]))

# phases of bands_fused_kernel by source line (lbad_extract.cu / lbad_math.cuh); packed-FMA intrinsics are inlined from the CUDA
# headers and cannot be attributed to a phase, so they are listed on their own
PHASES = [("pass 1: sample loads (half transform)", "lbad_extract.cu", 715, 715), ("pass 1: twiddle table loads", "lbad_extract.cu", 719, 719),
          ("pass 1: twiddle products (carried half, odd outputs)", "lbad_extract.cu", 720, 742), ("transposition: stores", "lbad_extract.cu", 756, 759),
          ("transposition: stores", "lbad_extract.cu", 766, 768), ("transposition: 128-bit loads", "lbad_extract.cu", 760, 765), ("transposition: 128-bit loads", "lbad_extract.cu", 769, 773),
          ("scalar butterfly arithmetic (w = 1, -i, t = +-1, tangent products)", "lbad_math.cuh", 80, 101), ("pass 2: first stage from the 128-bit loads", "lbad_math.cuh", 154, 162),
          ("real split: twiddle loads", "lbad_extract.cu", 786, 787), ("real split: partner shuffles + lane-0 select", "lbad_extract.cu", 788, 793),
          ("real split: partner shuffles + lane-0 select", "sm_30_intrinsics.hpp", 400, 460),
          ("warp syncs + the 32 register moves that hand the carried half transform to the next window", "sm_30_intrinsics.hpp", 100, 120), ("real split: arithmetic", "lbad_math.cuh", 177, 196),
          ("bin energies (Q4 scaling, squares)", "lbad_math.cuh", 197, 222), ("bin energies: stores", "lbad_extract.cu", 794, 808),
          ("band sums: loads + adds", "lbad_extract.cu", 603, 621), ("band sums: combine, divide, image store", "lbad_extract.cu", 834, 851),
          ("packed FFMA2 / FADD2 (butterflies, real split, energies)", "sm_100_rt.hpp", 0, 10 ** 6)]
ph = collections.OrderedDict()
for k, a in agg.items():
    name = "other (addressing, loop, frame staging)"
    if k:
        for n, f, lo, hi_ in PHASES:
            if k[0] == f and lo <= k[1] <= hi_:
                name = n; break
    p = ph.setdefault(name, [0, 0, 0]); p[0] += a[0]; p[1] += a[1]; p[2] += a[2]
print("%-72s %12s %12s %10s" % ("phase", "instr/unit", "smem wf/unit", "samples %"))
order = [n for n, *_ in PHASES]
for name in sorted(ph, key=lambda n: order.index(n) if n in order else 99):
    a = ph[name]
    print("%-72s %12.1f %12.1f %9.1f%%" % (name, a[0] / units, a[2] / units, 100.0 * a[1] / max(tot[1], 1)))
if "--lines" in sys.argv:
    for k, a in sorted(agg.items(), key=lambda kv: kv[0] or ("", 0)):
        if a[0] / units >= 0.5 or a[1] >= tot[1] * 0.002:
            print("%-22s %5s  instr %8.2f  wavefronts %7.2f  samples %5.2f%%" % (k[0] if k else None, k[1] if k else "", a[0] / units, a[2] / units, 100.0 * a[1] / tot[1]))
