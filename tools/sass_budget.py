#!/usr/bin/env python3
"""Instruction budget of one kernel by SOURCE LINE: static SASS count (nvdisasm -g line table of the -lineinfo build) and,
optionally, the warp-stall samples of an `ncu --set full --import-source on` capture of the SAME build mapped onto the lines.

    python tools/sass_budget.py build/obj/physics_kernels.<flags hash>.o _Z9k_physics [gpurun_out/prof_physics.ncu-rep] [--top 40]

This is how the instruction diet of k_physics / k_post was driven: both kernels are instruction-fetch
bound, so their time follows the executed instruction stream; the table shows where the stream goes (inlined helpers are
attributed to the helper's own line, e.g. cross3 / sincos) and which lines the warps wait on.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def line_table(obj, kernel_prefix):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
    start = next(i for i, l in enumerate(dis) if l.startswith("\t.section\t.text." + kernel_prefix))
    ends = [i for i in range(start + 1, len(dis)) if dis[i].startswith("//--------------------- .text.")]
    end = ends[0] if ends else len(dis)
    cur, table = None, []
    for l in dis[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        mm = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
        if mm:
            table.append((cur, mm.group(1).split(".")[0]))
    return table


def ncu_samples(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel_regex], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hdr = next(r for r in rows if "# Samples" in r)
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    return [(int(r[isamp] or 0), int(r[iex] or 0)) for r in rows if len(r) > iex and r[ia].startswith("0x")]


def main():
    argv = sys.argv[1:]
    top = 40
    if "--top" in argv:
        i = argv.index("--top")
        top = int(argv[i + 1])
        del argv[i:i + 2]
    args = argv
    obj, kernel = args[0], args[1]
    table = line_table(obj, kernel)
    samples = None
    if len(args) > 2:
        samples = ncu_samples(args[2], re.sub(r"^_Z\d+", "", kernel))
        if len(samples) != len(table):
            print(f"# the capture has {len(samples)} instructions, this build {len(table)}: not the same build, samples ignored")
            samples = None
    cnt, ops, smp, exe = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter(), collections.Counter()
    for i, (loc, op) in enumerate(table):
        cnt[loc] += 1
        ops[loc][op] += 1
        if samples:
            smp[loc] += samples[i][0]
            exe[loc] += samples[i][1]
    total, stot = len(table), sum(smp.values()) or 1
    print(f"# {kernel}: {total} instructions" + (f", {stot} stall samples" if samples else ""))
    key = (lambda l: -smp[l]) if samples else (lambda l: -cnt[l])
    src_cache = {}
    for loc in sorted(cnt, key=key)[:top]:
        f, ln = loc if loc else ("?", 0)
        if f not in src_cache:
            try:
                src_cache[f] = open(f).read().split("\n")
            except OSError:
                src_cache[f] = []
        text = src_cache[f][ln - 1].strip()[:90] if 0 < ln <= len(src_cache[f]) else ""
        head = f"{smp[loc]:5d} {100 * smp[loc] / stot:4.1f}% exec {exe[loc]:7d} " if samples else ""
        print(f"{head}{cnt[loc]:5d} instr  {os.path.basename(f)}:{ln}  {dict(ops[loc].most_common(3))}  | {text}")


if __name__ == "__main__":
    main()
