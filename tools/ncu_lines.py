"""Join an ncu SASS source page (CSV) with nvdisasm line info to get per-source-line instruction counts and
stall samples.

    ncu -i rep.ncu-rep --page source --csv > sass.csv
    python tools/ncu_lines.py sass.csv <lib.so> <kernel-substring> [top]
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(lib, kern_sub):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    out = []
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur, loc, active = None, ("?", 0), False
        for line in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
            if m:
                active = kern_sub in m.group(1)
                cur = m.group(1)
                continue
            if not active:
                continue
            m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
            if m:
                loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip(), loc))
        if out:
            break
    return out


def main():
    sass_csv, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = disasm_lines(lib, kern)
    rows = list(csv.reader(open(sass_csv)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if r and r[0].startswith("0x")]
    col = {h: i for i, h in enumerate(hdr)}
    base = int(body[0][0], 16)
    by_off = {off: (txt, loc) for off, txt, loc in dis}
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    tot_inst = tot_samp = 0
    miss = 0
    for r in body:
        off = int(r[0], 16) - base
        txt, loc = by_off.get(off, (None, ("?", 0)))
        if txt is None:
            miss += 1
        inst = int(float(r[col["Instructions Executed"]] or 0))
        samp = int(float(r[col["# Samples"]] or 0))
        a = agg[loc]
        a[0] += inst
        a[1] += samp
        for h in stall_cols:
            v = r[col[h]]
            if v and v != "0":
                a[2][h[6:]] += int(float(v))
        tot_inst += inst
        tot_samp += samp
    print(f"instructions {tot_inst}, samples {tot_samp}, sass rows {len(body)}, unmatched {miss}")
    for loc, (inst, samp, st) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        stalls = ", ".join(f"{k}:{v}" for k, v in st.most_common(4))
        print(f"{loc[0]}:{loc[1]:<5d} inst {100*inst/tot_inst:5.1f}%  samples {100*samp/max(tot_samp,1):5.1f}%  [{stalls}]")


if __name__ == "__main__":
    main()
