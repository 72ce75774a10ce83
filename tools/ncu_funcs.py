"""Per-device-function and per-instruction stall breakdown from an ncu SASS source page.
    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_funcs.py sass.csv <lib.so> <cubin-prefix> <kernel-substring> [function-substring-for-instruction-listing]
"""
import collections, csv, os, re, subprocess, sys, tempfile

sass_csv, lib, cubpre, kern = sys.argv[1:5]
detail = sys.argv[5] if len(sys.argv) > 5 else None
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith(cubpre)][0]
txt = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cub)], capture_output=True, text=True).stdout
active, cur, funcs, loc = False, 'entry', {}, ('?', 0)
for line in txt.splitlines():
    m = re.match(r'\s*\.section\s+\.text\.(\S+?),', line)
    if m:
        active = kern in m.group(1); cur = 'entry'; continue
    if not active:
        continue
    m = re.match(r'^\$_ZN3cpz\S*?\$(_ZN3cpz\S+?):', line)
    if m:
        cur = m.group(1); continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        funcs[int(m.group(1), 16)] = (cur, m.group(2).strip(), loc)
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address'); hdr = rows[hi]
body = [r for r in rows[hi + 1:] if r and r[0].startswith('0x')]
col = {h: i for i, h in enumerate(hdr)}; base = int(body[0][0], 16)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()]); sel = []
for r in body:
    f, op, loc = funcs.get(int(r[0], 16) - base, ('?', '?', ('?', 0)))
    short = re.sub(r'_ZN3cpz\d+', '', f)[:50]
    a = agg[short]; a[0] += int(r[col['Instructions Executed']]); a[1] += int(r[col['# Samples']])
    st = {h[6:]: int(r[col[h]]) for h in stalls if r[col[h]] not in ('', '0')}
    for k, v in st.items():
        a[2][k] += v
    if detail and detail in f:
        sel.append((int(r[col['# Samples']]), int(r[col['Instructions Executed']]), op[:64], loc, st))
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{k:52s} inst {100*a[0]/ti:5.1f}% samples {100*a[1]/ts:5.1f}%  " + ", ".join(f"{s}:{100*v/max(a[1],1):.0f}%" for s, v in a[2].most_common(6)))
if detail:
    print('--- top instructions in', detail, 'total samples', sum(s[0] for s in sel))
    for s in sorted(sel, key=lambda x: -x[0])[:40]:
        print(f"{s[0]:5d} {s[1]:8d} {s[2]:64s} {s[3][0]}:{s[3][1]} {dict(sorted(s[4].items(), key=lambda kv: -kv[1])[:3])}")
