"""Where a kernel's instructions and stall samples go, by CUDA source function (read here, without a GPU).
usage: python tools/ncu_source_breakdown.py <report.ncu-rep> [out.json]
Reads `ncu -i <rep> --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on), attributes every
SASS instruction's executed count and stall samples to the source line ncu maps it to, and sums them per enclosing function
(found by scanning the source files for the RZK_VM / __device__ / template function headers)."""
import csv
import json
import re
import subprocess
import sys
from collections import defaultdict


_REPORT_SOURCES = {}


def load_report_sources(rep):
    """The CUDA sources as imported into the report (--import-source on), so that line numbers match the profiled build."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
    cur = None
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] in ("File Name", "File Path"):
            cur = r[1]; _REPORT_SOURCES[cur] = {}
        elif cur is not None and len(r) >= 2 and r[0].isdigit():
            _REPORT_SOURCES[cur][int(r[0])] = r[1]


def function_ranges(path):
    """[(first_line, name)] of function definitions in a source file (good enough for this code base's style)."""
    out = []
    if path in _REPORT_SOURCES and _REPORT_SOURCES[path]:
        src = _REPORT_SOURCES[path]
        lines = [src.get(i, "") for i in range(1, max(src) + 1)]
    else:
        try:
            lines = open(path).read().splitlines()
        except OSError:
            return out
    pat = re.compile(r'^\s*(?:RZK_VM|RZK_HD|RZK_D|__device__|__global__|static|inline|constexpr)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(')
    for i, ln in enumerate(lines, 1):
        if ln.startswith((' ', '\t')) and not ln.lstrip().startswith(('RZK_', '__device__', '__global__')):
            continue
        m = pat.match(ln)
        if m and not ln.rstrip().endswith(';'):
            out.append((i, m.group(1)))
    return out


def main():
    rep = sys.argv[1]
    load_report_sources(rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    per_line = defaultdict(lambda: defaultdict(float))
    fpath, hdr = None, None
    kernel = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1]; hdr = None; continue
        if r[0] == "Function Name":
            kernel = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or fpath is None:
            continue
        d = dict(zip(hdr, r))
        try:
            line = int(d["Line No"])
        except (KeyError, ValueError):
            continue
        k = (fpath, line)
        for name in ("Instructions Executed", "# Samples"):
            try:
                per_line[k][name] += float(d.get(name) or 0)
            except ValueError:
                pass
        for name, v in d.items():
            if name.startswith("stall_") and "Not Issued" not in name:
                try:
                    per_line[k][name] += float(v or 0)
                except ValueError:
                    pass
    ranges = {}
    per_fn = defaultdict(lambda: defaultdict(float))
    for (f, line), m in per_line.items():
        if f not in ranges:
            ranges[f] = function_ranges(f)
        name = "?"
        for first, fn in ranges[f]:
            if first <= line:
                name = fn
            else:
                break
        key = f.split("/")[-1] + ":" + name
        for a, b in m.items():
            per_fn[key][a] += b
    tot_i = sum(m["Instructions Executed"] for m in per_fn.values()) or 1
    tot_s = sum(m["# Samples"] for m in per_fn.values()) or 1
    res = {"kernel": kernel, "instructions_executed": tot_i, "samples": tot_s, "functions": []}
    print(f"{kernel}\n  warp instructions executed {tot_i:.0f}, stall samples {tot_s:.0f}")
    print(f"  {'function':44s} {'inst %':>7s} {'samples %':>9s}   top stall reasons (share of the function's samples)")
    for key, m in sorted(per_fn.items(), key=lambda kv: -kv[1]["# Samples"]):
        st = sorted(((a[6:], b) for a, b in m.items() if a.startswith("stall_") and b > 0), key=lambda ab: -ab[1])
        s = m["# Samples"] or 1
        top = ", ".join(f"{a} {100 * b / s:.0f}" for a, b in st[:5])
        print(f"  {key:44s} {100 * m['Instructions Executed'] / tot_i:7.2f} {100 * m['# Samples'] / tot_s:9.2f}   {top}")
        res["functions"].append({"function": key, "inst_pct": round(100 * m["Instructions Executed"] / tot_i, 2),
                                 "samples_pct": round(100 * m["# Samples"] / tot_s, 2),
                                 "stalls_pct_of_function": {a: round(100 * b / s, 1) for a, b in st[:8]}})
    if len(sys.argv) > 2:
        json.dump(res, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
