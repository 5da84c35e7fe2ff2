"""Per-source-line and per-opcode breakdown of an ncu report (needs -lineinfo): instructions, stall samples and shared-memory
wavefronts per dof row.  usage: ncu_lines.py report.ncu-rep kernel_substring rows [top] [source stem, default apply_mma]"""
import collections, csv, re, subprocess, sys, os, tempfile
rep, ksub, nrows = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
stem = sys.argv[5] if len(sys.argv) > 5 else "apply_mma"
so = os.path.join(os.path.dirname(__file__), "..", "extendableasgfem.jl_b200", "libasgfem_cuda.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
dis = ""
for f in os.listdir(tmp):
    if f.startswith(stem) and f.endswith(".cubin"):
        dis = subprocess.run(["nvdisasm", "-g", f], cwd=tmp, capture_output=True, text=True).stdout
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kname = rows[0][1]
mangled = None
lines = dis.split("\n")
# find the function whose demangled template args match the kernel name of the report
m = re.search(r"(k_apply_\w+)<([^>]*)>", kname)
tag = m.group(1) + "I" + "".join(("Li%sE" if t == "int" else "Lb%sE") % a for t, a in re.findall(r"\((int|bool)\)(\d+)", m.group(2))) + "E" if m else ksub
start = None
for i, l in enumerate(lines):
    if l.startswith("_ZN") and tag in l and l.rstrip().endswith(":"):
        start = i
        break
off2line, cur = {}, None
for l in lines[start + 1:]:
    mm = re.search(r'//## File ".*?([a-z_0-9]+\.cuh?)", line (\d+)', l)
    if mm:
        cur = int(mm.group(2))
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if mm:
        off2line[int(mm.group(1), 16)] = cur
    if l.startswith("//-----"):
        break
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][0], 16)
per, samp, wf = collections.Counter(), collections.Counter(), collections.Counter()
op, ops, opw = collections.Counter(), collections.Counter(), collections.Counter()
stall = collections.Counter()
lstall = collections.defaultdict(collections.Counter)
for r in data:
    ln = off2line.get(int(r[0], 16) - base)
    n, s, w = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]), int(r[ix["L1 Wavefronts Shared"]] or 0)
    per[ln] += n; samp[ln] += s; wf[ln] += w
    t = r[ix["Source"]].split()
    o = t[0] if not t[0].startswith("@") else t[1]
    op[o] += n; ops[o] += s; opw[o] += w
    for h in hdr:
        if h.startswith("stall_") and "Not Issued" not in h:
            stall[h] += int(r[ix[h]] or 0)
            lstall[ln][h[6:]] += int(r[ix[h]] or 0)
tot = sum(samp.values())
print(kname)
print("warp-instructions per row %.0f, smem wavefronts per row %.0f" % (sum(per.values()) / nrows, sum(wf.values()) / nrows))
print("stalls: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in stall.most_common(8)))
src = open(os.path.join(os.path.dirname(__file__), "..", "extendableasgfem.jl_b200", "csrc", stem + ".cu")).read().split("\n")
for ln, c in sorted(per.items(), key=lambda x: -x[1])[:top]:
    print("%4s: instr/row %7.1f samples %5.1f%% wf/row %7.1f | %s" % (ln, c / nrows, 100 * samp[ln] / tot, wf[ln] / nrows, src[ln - 1].strip()[:90] if ln else ""))
print("\nlines by stall samples:")
for ln, c in sorted(samp.items(), key=lambda x: -x[1])[:14]:
    print("%4s: samples %5.1f%% %s | %s" % (ln, 100 * c / tot, ", ".join("%s %.0f%%" % (k, 100 * v / max(c, 1)) for k, v in lstall[ln].most_common(3)), src[ln - 1].strip()[:70] if ln else ""))
print()
for o, c in op.most_common(22):
    print("%-26s instr/row %8.1f samples %5.1f%% wf/row %7.1f" % (o, c / nrows, 100 * ops[o] / tot, opw[o] / nrows))
