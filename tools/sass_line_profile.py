"""Joins an `ncu --page source --csv` SASS listing (per-instruction counters) with `nvdisasm --print-line-info` of the same
cubin, and sums the counters per CUDA source line / per source range. Development tool.
usage: sass_line_profile.py ncu_source.csv dis.txt kernel_substring [callee_substring ...]"""
import csv, re, sys, collections, os

def load_ncu(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = rows[hdr]
    ia, isrc, iinst, ithr = cols.index("Address"), cols.index("Source"), cols.index("Instructions Executed"), cols.index("Thread Instructions Executed")
    isamp = cols.index("# Samples")
    out = []
    for r in rows[hdr + 1:]:
        if len(r) <= ithr or not r[ia]:
            continue
        out.append((r[isrc].strip(), float(r[iinst] or 0), float(r[ithr] or 0), float(r[isamp] or 0)))
    return out

def load_dis(path, names):
    """list of (opcode text, line) in order for the functions whose section name contains any of `names`, in file order of names"""
    funcs = {}
    cur, line = None, 0
    for l in open(path):
        m = re.match(r"\s*\.section\s+\.text\.(\S+),", l)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.search(r'//## File ".*?", line (\d+)', l)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m and cur:
            funcs[cur].append((m.group(2).strip(), line))
    out = []
    for n in names:
        for k, v in funcs.items():
            if n in k:
                out.append((k, v))
    return out

ncu = load_ncu(sys.argv[1])
dis = load_dis(sys.argv[2], sys.argv[3:])
flat = [x for _, v in dis for x in v]
print("ncu rows", len(ncu), "dis rows", len(flat), [(k[-40:], len(v)) for k, v in dis])
per_line = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
n = min(len(ncu), len(flat))
mismatch = 0
for i in range(n):
    op_n = ncu[i][0].split()[0] if ncu[i][0] else ""
    op_d = flat[i][0].split()[0] if flat[i][0] else ""
    if op_n.lstrip("@!P0123456789 ") != op_d.lstrip("@!P0123456789 ") and op_n != op_d:
        mismatch += 1
    a = per_line[flat[i][1]]
    a[0] += ncu[i][1]; a[1] += ncu[i][2]; a[2] += ncu[i][3]
tot = sum(v[0] for v in per_line.values())
tots = sum(v[2] for v in per_line.values())
print("opcode mismatches", mismatch, "total warp instructions %.3e" % tot, "samples", tots)
src = open("/root/repo/sqeazy_b200/csrc/device/" + (os.environ.get("SRC") or "lz4_encode.cu")).read().split("\n") if len(sys.argv) > 3 else []
for ln in sorted(per_line):
    v = per_line[ln]
    if v[0] / tot > 0.004 or v[2] / max(tots, 1) > 0.004:
        print("%5d %6.2f%% inst %6.2f%% samp  thr/inst %4.1f  %s" % (ln, 100 * v[0] / tot, 100 * v[2] / max(tots, 1), v[1] / max(v[0], 1), src[ln - 1].strip()[:110] if 0 < ln <= len(src) else ""))
# ranges
if len(sys.argv) > 3:
    import bisect
    # phase boundaries of lz4_encode.cu, found by their header comments
    pats = [("ext_bytes", "__device__ __forceinline__ int ext_bytes"), ("load4", "uint32_t load4("), ("eq4_shift_in (A0)", "uint32_t eq4_shift_in("),
            ("warp_totals", "void warp_totals("), ("put_ext", "int put_ext("), ("run_ones (B)", "int run_ones("), ("prologue", "int encode_general("),
            ("A0 masks", "auto analyse = "), ("A1 short cand", "phase A1: fixed-offset candidates"), ("sampling + votes", "Four warps analyse their sub-blocks first"),
            ("A2 list+hash", "phase A2: hash candidates"), ("lead table", "ones at the start of every segment and beyond"), ("B parse", "phase B: every thread parses"),
            ("C compaction", "phase C: the selected matches"), ("D sizes", "sequences [s0, s1) of this thread"), ("D emit", "phase D: emission"),
            ("kernel: load", "lz4_encode_kernel(const uint8_t*"), ("closed form", "closed form: 1 literal"), ("hand-off", "hand-off: size word"), ("other", "Compaction (replaces remove_blanks")]
    marks = [(1, "helpers")]
    for name, pat in pats:
        for i, l in enumerate(src):
            if pat in l:
                marks.append((i + 1, name)); break
    marks.sort()
    agg = collections.defaultdict(lambda: [0.0, 0.0])
    for ln, v in per_line.items():
        k = bisect.bisect_right([m[0] for m in marks], ln) - 1
        agg[marks[k][1]][0] += v[0]; agg[marks[k][1]][1] += v[2]
    for _, name in marks:
        if name in agg:
            print("%-16s %6.2f%% inst %6.2f%% samples" % (name, 100 * agg[name][0] / tot, 100 * agg[name][1] / max(tots, 1)))
