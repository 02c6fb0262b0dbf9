#!/usr/bin/env python3
"""Where a kernel loses lanes: joins ncu's per-SASS-instruction counters (warp-level and thread-level executions, stall
samples) with the source lines of a local rebuild and lists, per source line, the warp-instructions that perfect 32-lane
execution of the same thread work would not have needed ("wasted"), per unit of work.

Usage: python tools/ncu_divergence.py source_page.csv UNITS [kernel-substring]
  source_page.csv = `ncu -i rep.ncu-rep --page source --csv -k regex:<kernel> -c 1`
  UNITS           = units of work in the profiled launch (e.g. env-frames), the divisor of every count
(the working tree must be the source the report was taken from)
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "neuro_genetic_pong_self_play_b200", "csrc")


def main():
    page, units = sys.argv[1], float(sys.argv[2])
    kernel = sys.argv[3] if len(sys.argv) > 3 else "rollout_kernel"
    cubin = os.path.join(tempfile.mkdtemp(), "core.cubin")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin", "-o", cubin,
                           os.path.join(CSRC, "ngp_core.cu")], stderr=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    rows = list(csv.reader(open(page)))
    hdr, data = rows[1], rows[2:]
    iE, iS, iT = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
    best = None
    for st in [i for i, l in enumerate(sass) if l.startswith(".text.") and kernel in l]:
        end = next(i for i in range(st + 1, len(sass)) if sass[i].startswith("//---------------------") or i == len(sass) - 1)
        cur, seq = ("?", 0), []
        for l in sass[st:end]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
                seq.append(cur)
        if len(seq) == len(data):
            best = seq
    if best is None:
        raise SystemExit(f"no section of the local build matches the report's {len(data)} SASS instructions; is the working tree the profiled source?")
    W, T, S = collections.Counter(), collections.Counter(), collections.Counter()
    for loc, row in zip(best, data):
        W[loc] += int(row[iE]); T[loc] += int(row[iT]); S[loc] += int(row[iS])
    tw, tt, ts = sum(W.values()), sum(T.values()), max(1, sum(S.values()))
    print(f"warp instructions per unit {tw / units:.1f}, average active threads {tt / tw:.2f}")
    waste = {k: W[k] - T[k] / 32 for k in W}
    print(f"wasted warp instructions per unit {sum(waste.values()) / units:.1f}")
    byfile = collections.Counter()
    for k, v in waste.items():
        byfile[k[0]] += v
    print({k: round(v / units) for k, v in byfile.most_common(8)})
    cache = {}

    def text(f, ln):
        path = os.path.join(CSRC, f) if os.path.exists(os.path.join(CSRC, f)) else os.path.join(CSRC, "generated", f)
        if path not in cache:
            cache[path] = open(path).read().split("\n") if os.path.exists(path) else []
        return cache[path][ln - 1].strip()[:80] if 0 < ln <= len(cache[path]) else ""

    for (f, ln), v in sorted(waste.items(), key=lambda kv: -kv[1])[:45]:
        print(f"{v / units:7.1f} wasted  {W[(f, ln)] / units:7.1f} warp-inst  active {T[(f, ln)] / max(1, W[(f, ln)]):5.1f}  samples {100 * S[(f, ln)] / ts:4.1f}%  {f}:{ln} : {text(f, ln)}")


if __name__ == "__main__":
    main()
