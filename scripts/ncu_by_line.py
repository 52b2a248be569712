"""Per-phase view of bands_fused_kernel in an .ncu-rep (read with `ncu -i`, no GPU needed): joins the SASS page of the report
(instructions executed, stall samples, shared-memory wavefronts per SASS instruction) with the inline-aware line table of the same
kernel in the built object (`cuobjdump -xelf` + `nvdisasm -gi`: every instruction carries its chain of inlined-at source lines), and
attributes each instruction to a phase by the TEXT of the kernel-body lines in its chain — no line numbers are hard-coded, so the
table survives edits of the kernel.

usage: python scripts/ncu_by_line.py gpurun_out/prof_extract.ncu-rep lbaudiodetective_b200/build/lbad_extract.cu.o bands_fused_kernelILi32ELb1ELb1ELi1ELb1ELi2E 24320000 [--lines]
       (report, object file, mangled-name fragment of the kernel, units per launch — here windows)"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, obj, frag, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text._ZN") and frag in l)
chain, off2chain, fresh = [], {}, True
for l in dis[start + 1:]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        if fresh:
            chain, fresh = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        if m.group(3):
            chain.append((os.path.basename(m.group(3)), int(m.group(4))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", l)
    if m:
        off2chain[int(m.group(1), 16)] = tuple(chain); fresh = True
    elif l.startswith("//-----") and off2chain:
        break
src_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lbaudiodetective_b200", "csrc")
SRC = {f: open(os.path.join(src_dir, f)).read().split("\n") for f in ("lbad_extract.cu", "lbad_math.cuh", "lbad_common.cuh")}
def text(fl):
    f, n = fl
    return SRC[f][n - 1] if f in SRC and 0 < n <= len(SRC[f]) else ""

# phase <- first anchor (in this order) found in the text of any source line of the instruction's inline chain
ANCHORS = [
    ("pass 1: sample loads", ["h[m] = *reinterpret_cast<const float2*>(w + 2"]),
    ("pass 1: 16-point transform of the new half", ["dft16(h)"]),
    ("pass 1: twiddle loads and products (carried half)", ["h[q] = cmul(h[q], tw1h"]),
    ("pass 1: combining stage (carried + new half)", ["dit32_combine("]),
    ("pass 1: odd outputs x the lane's exp(-2 pi i 16 n2 / M)", ["const float2 om = tw1h", "cmul(z[2 * q + 1], om)", "carry[q] = odd[q]"]),
    ("transposition: 32 STS.64", ["scr2[bitrev5(p) * SCR_LD2 + lane] = z[p]"]),
    ("transposition: 16 LDS.128", ["scr2[lane * SCR_LD2 + 2 * q]", "z[2 * q] = make_float2(t.x"]),
    ("pass 2: 32-point transform", ["fft32_tail<R>(z)"]),
    ("real split: partner shuffles + lane-0 select", ["lane == 0 ? z[p0] : z[pp]", "__shfl_sync(0xffffffffu, offer"]),
    ("real split: arithmetic (+ the one twiddle load)", ["real_split_pair_rows(", "real_split_pair_2x(", "wl = reinterpret_cast<const float2*>(tw2)[lane]", "lo.x = 2.0f * (z[p].x"]),
    ("bin energies and their stores", ["vbuf[k] = bin_energy_raw", "vbuf[1024 - k] = bin_energy_raw_conj", "vbuf[512] = bin_energy_raw"]),
    ("band sums: loads and adds", ["seg_sum_static<", "seg_sum(v, "]),
    ("band sums: combine, divide, image store", ["__shfl_xor_sync(0xffffffffu, sa", "__shfl_xor_sync(0xffffffffu, sb", "float tot = (lane & 1)", "if (!(tot <=", "images[((size_t)f * UNIT_ROWS"]),
]
OTHER = "other (addressing, loop control, warp syncs, frame staging, barriers)"
def phase(ch):
    texts = [text(fl) for fl in ch]
    for name, pats in ANCHORS:
        if any(p in t for t in texts for p in pats):
            return name
    return OTHER

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h, data = rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]
ia, iex, ism, iwf, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared"), h.index("Source")
base = int(data[0][ia], 16)
ph, lines = collections.OrderedDict(), collections.defaultdict(lambda: [0, 0, 0])
pipes = collections.defaultdict(lambda: collections.Counter())
for r in data:
    ch = off2chain.get(int(r[ia], 16) - base, ())
    name = phase(ch)
    a = ph.setdefault(name, [0, 0, 0]); ex, sm_, wf = int(r[iex]), int(r[ism]), int(r[iwf] or 0)
    a[0] += ex; a[1] += sm_; a[2] += wf
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[isrc]); op = m.group(2) if m else "?"
    pipes[name]["packed" if op in ("FFMA2", "FADD2", "FMUL2") else "scalar fp" if op in ("FFMA", "FADD", "FMUL") else "other"] += ex
    body = next((fl for fl in reversed(ch) if fl[0] == "lbad_extract.cu"), ch[-1] if ch else None)
    b = lines[body]; b[0] += ex; b[1] += sm_; b[2] += wf
tot = [sum(a[i] for a in ph.values()) for i in range(3)]
print("per %s: %.1f warp instructions, %.1f shared-memory wavefronts of LDS/STS (shuffles and the bulk copy's writes are counted by the l1tex counter, not here); %d stall samples"
      % ("window", tot[0] / units, tot[2] / units, tot[1]))
print("%-74s %10s %10s %10s %12s %10s" % ("phase", "instr", "packed fp", "scalar fp", "smem wavefr.", "samples %"))
order = [n for n, _ in ANCHORS] + [OTHER]
for name in sorted(ph, key=order.index):
    a = ph[name]
    print("%-74s %10.1f %10.1f %10.1f %12.1f %9.1f%%" % (name, a[0] / units, pipes[name]["packed"] / units, pipes[name]["scalar fp"] / units, a[2] / units, 100.0 * a[1] / max(tot[1], 1)))
print("%-74s %10.1f %10.1f %10.1f %12.1f" % ("total", tot[0] / units, sum(p["packed"] for p in pipes.values()) / units, sum(p["scalar fp"] for p in pipes.values()) / units, tot[2] / units))
if "--lines" in sys.argv:
    for k, a in sorted(lines.items(), key=lambda kv: kv[0] or ("", 0)):
        if a[0] / units >= 0.5 or a[1] >= tot[1] * 0.002:
            print("%-22s %5s  instr %8.2f  wavefronts %7.2f  samples %5.2f%%  %s" % (k[0] if k else None, k[1] if k else "", a[0] / units, a[2] / units, 100.0 * a[1] / tot[1], text(k).strip()[:90] if k else ""))
if "--debug" in sys.argv:
    want = sys.argv[sys.argv.index("--debug") + 1]
    cnt = collections.Counter()
    for r in data:
        ch = off2chain.get(int(r[ia], 16) - base, ())
        if phase(ch).startswith(want):
            cnt[ch] += int(r[iex])
    for ch, c in cnt.most_common(12):
        print("%8.1f  %s" % (c / units, " <- ".join("%s:%d" % fl for fl in ch)))
