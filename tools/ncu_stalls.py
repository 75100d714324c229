"""Warp-stall samples of an `ncu --set full --import-source on` capture, attributed to source lines.

  python tools/ncu_stalls.py <report.ncu-rep> <object with -lineinfo, e.g. molecular-vae_b200/_build/gru_rec2.o> [top]

ncu's CSV source page lists SASS instructions with their stall samples but no line numbers; nvdisasm -g of the same build
lists the same instructions with `//## File "...", line N` markers.  The two are joined by instruction index per kernel."""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os


def sass_sections(obj):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    out = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout.splitlines()
        name, cur_file, cur_line = None, None, None
        for l in txt:
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
            if m:
                name = m.group(1)
                out[name] = []
                continue
            m = re.search(r'//## File "(.*?)", line (\d+)', l)
            if m:
                cur_file, cur_line = os.path.basename(m.group(1)), int(m.group(2))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m and name:
                out[name].append((cur_file, cur_line, m.group(2).strip()))
    return out


def main():
    rep, obj = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    sass = sass_sections(obj)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for r in csv.reader(io.StringIO(raw)):
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    src_cache = {}
    seen = set()
    for s in secs:
        if s["name"] in seen:
            continue
        seen.add(s["name"])
        # template arguments of the demangled name -> the mangled section with the same instruction count
        cands = [k for k, v in sass.items() if len(v) == len(s["rows"])]
        if not cands:
            print("no SASS section with", len(s["rows"]), "instructions for", s["name"][:80])
            continue
        dis = sass[cands[0]]
        h = s["hdr"]
        isamp = h.index("# Samples")
        stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        agg, why = {}, {}
        for i, r in enumerate(s["rows"]):
            key = (dis[i][0], dis[i][1])
            n = int(r[isamp] or 0)
            agg[key] = agg.get(key, 0) + n
            w = why.setdefault(key, {})
            for c in stall_cols:
                v = int(r[c] or 0)
                if v:
                    w[h[c][6:]] = w.get(h[c][6:], 0) + v
        tot = sum(agg.values())
        print(f"kernel: {s['name'][:110]}\n  {tot} warp samples over {len(s['rows'])} SASS instructions ({cands[0][:60]}...)")
        for (f, ln), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
            text = "?"
            if f:
                path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
                if path not in src_cache:
                    src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
                if src_cache[path] and ln and ln <= len(src_cache[path]):
                    text = src_cache[path][ln - 1].strip()[:90]
            reasons = ", ".join(f"{k} {v}" for k, v in sorted(why[(f, ln)].items(), key=lambda kv: -kv[1])[:2])
            print(f"  {n:7d} {100.0 * n / tot:5.1f}%  {f}:{ln}  {text}   [{reasons}]")


if __name__ == "__main__":
    main()
