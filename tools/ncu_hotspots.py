#!/usr/bin/env python3
"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.

ncu's source page of a report taken on the GPU box lists SASS only; this script rebuilds the same
translation unit here with -lineinfo, disassembles it with nvdisasm -g (which annotates every SASS
instruction with file:line) and joins the two listings by instruction index.

Usage: python tools/ncu_hotspots.py gpurun_out/prof.ncu-rep [kernel-substring] [--top N]
(the working tree must be the source the report was taken from)
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CSRC = os.path.join(ROOT, "neuro_genetic_pong_self_play_b200", "csrc")


def main():
    rep = sys.argv[1]
    kernel = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else "rollout_kernel"
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
    global CSRC
    if "--src" in sys.argv:      # csrc directory of the profiled source (e.g. from `git archive <commit>`)
        CSRC = sys.argv[sys.argv.index("--src") + 1]
    tu = sys.argv[sys.argv.index("--tu") + 1] if "--tu" in sys.argv else "ngp_core.cu"
    tmp = tempfile.mkdtemp()
    cubin = os.path.join(tmp, "core.cubin")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-cubin", "-o", cubin,
                           os.path.join(CSRC, tu)], stderr=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kernel, "-c", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    kname = rows[0][1]
    hdr, data = rows[1], rows[2:]
    iE, iS, iSrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
    # candidate sections: kernels whose name contains the substring; pick the one with matching length
    starts = [i for i, l in enumerate(sass) if l.startswith(".text.") and kernel in l]
    best = None
    for st in starts:
        end = next(i for i in range(st + 1, len(sass)) if sass[i].startswith("//---------------------") or i == len(sass) - 1)
        cur, seq = ("?", 0), []
        for l in sass[st:end]:
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                seq.append((cur, m.group(2)))
        if len(seq) == len(data):
            best = seq
    if best is None:
        raise SystemExit(f"no section of the local build matches the report's {len(data)} SASS instructions (kernel {kname}); "
                         "is the working tree the profiled source?")
    byline, samp, byfile = collections.Counter(), collections.Counter(), collections.Counter()
    tot = 0
    for (loc, _), row in zip(best, data):
        e, s_ = int(row[iE]), int(row[iS])
        byline[loc] += e; samp[loc] += s_; byfile[loc[0]] += e; tot += e
    ssum = max(1, sum(samp.values()))
    print(f"kernel: {kname}\nwarp instructions executed: {tot}")
    for f, c in byfile.most_common():
        print("%6.2f%%  %s" % (100 * c / tot, f))
    print("\n  inst%  samples%  location : source")
    cache = {}
    for (f, ln), c in byline.most_common(top):
        path = os.path.join(CSRC, f) if os.path.exists(os.path.join(CSRC, f)) else os.path.join(CSRC, "generated", f)
        if path not in cache:
            cache[path] = open(path).read().split("\n") if os.path.exists(path) else []
        text = cache[path][ln - 1].strip()[:100] if 0 < ln <= len(cache[path]) else ""
        print("%6.2f%% %7.2f%%  %s:%d : %s" % (100 * c / tot, 100 * samp[(f, ln)] / ssum, f, ln, text))


if __name__ == "__main__":
    main()
