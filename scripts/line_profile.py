"""Per-source-line instruction and stall-sample shares of one kernel from an ncu report captured with
`--set full --import-source on` (SASS page) joined with `nvdisasm -g` line info of the same build.

    python scripts/line_profile.py gpurun_out/aug_r01_final.ncu-rep build/obj/aug_tile.cu.o aug_tile_kernelILb0E
"""
import collections, csv, os, re, subprocess, sys, tempfile

rep, obj, needle = sys.argv[1], sys.argv[2], sys.argv[3]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text.") and needle in l][0]
end = next((i for i in range(start + 1, len(dis)) if dis[i].startswith(".text.") and needle not in dis[i]), len(dis))
cur, lines = None, []
for l in dis[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[1]
iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
data = [(int(r[iE]), int(r[iS])) for r in rows[2:]]
assert len(data) == len(lines), (len(data), len(lines), "report and object file are different builds")
tot, ts = sum(d[0] for d in data), sum(d[1] for d in data)
agg = collections.defaultdict(lambda: [0, 0])
for k, (e, s) in zip(lines, data):
    agg[k][0] += e
    agg[k][1] += s
src = {}
print(f"{tot} warp instructions, {ts} samples, {len(data)} SASS instructions")
for k, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if e / tot > 0.004 or s / ts > 0.006:
        text = ""
        if k:
            path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "..", "medical_image_segmentation_b200", "csrc", k[0])
            if k[0] not in src:
                src[k[0]] = open(path).read().split("\n") if os.path.exists(path) else None
            if src[k[0]]:
                text = src[k[0]][k[1] - 1].strip()[:90]
        print(f"{(k[0] if k else '?'):>16s}:{(k[1] if k else 0):4d}  instr {100 * e / tot:5.2f}%  samples {100 * s / ts:5.2f}%  {text}")
