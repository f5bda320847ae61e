#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into tracked text files under profiles/.

    python scripts/prof_summary.py launches gpurun_out/r1_launches.csv profiles/r1_launches.md [step_marker]
    python scripts/prof_summary.py full gpurun_out/r1_prof_gemm.ncu-rep profiles/r1_ncu_gemm.md

`launches`: per-kernel-class device time of ONE steady-state step (between two occurrences of the step marker
kernel, default k_count = the first kernel of the CSR build), from the `--metrics gpu__time_duration.sum` pass.
`full`: the headline counters of every profiled launch of a `--set full` report (read with `ncu -i ... --page raw`).
`sass`: per-kernel counts of the Blackwell-specific SASS mnemonics (tcgen05 MMA, tensor-memory loads / stores, TMA / bulk
copies, mbarrier waits) in the built library, from `cuobjdump -sass` -- the proof that the hot kernels are hand-written
tcgen05 / TMEM / TMA code and not recompiled mma.sync:

    python scripts/prof_summary.py sass swarm_ode_b200/libgnode_b200.so profiles/sass_summary.txt
"""
import collections
import csv
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.sum.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg",
]


def clean(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)[:90]


def launches(src, dst, marker="k_count", which=2):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    data = [(x[ki], float(x[vi].replace(",", "")) / (1000.0 if x[ui] == "ns" else 1.0)) for x in r if len(x) > vi]
    idx = [i for i, d in enumerate(data) if marker in d[0]]
    a, b = (idx[which], idx[which + 1]) if len(idx) > which + 1 else (0, len(data))
    agg, tot = collections.OrderedDict(), 0.0
    for name, t in data[a:b]:
        e = agg.setdefault(clean(name), [0, 0.0])
        e[0] += 1
        e[1] += t
        tot += t
    with open(dst, "w") as out:
        out.write(f"# ncu launch list ({src}): one steady-state step, launches {a}..{b} of {len(data)}\n\n")
        out.write("Per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes.\n\n")
        out.write(f"launches in step: {b - a}, summed device time {tot / 1000:.3f} ms\n\n")
        out.write("| us | launches | share | kernel |\n|---:|---:|---:|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f"| {v[1]:.1f} | {v[0]} | {100 * v[1] / tot:.1f}% | `{k}` |\n")
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, units, body = rows[0], rows[1], rows[2:]
    cols = [hdr.index(m) for m in FULL_METRICS if m in hdr]
    kn = hdr.index("Kernel Name")
    with open(dst, "w") as out:
        out.write(f"# ncu --set full ({src}), headline counters per profiled launch\n\n")
        for r in body:
            out.write(f"## `{clean(r[kn])}`  (ID {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for c in cols:
                out.write(f"| {hdr[c]} | {r[c]} | {units[c]} |\n")
            out.write("\n")
    print(open(dst).read())


SASS_MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS",
                  "HMMA", "FFMA", "LDS", "STS", "LDG", "STG", "BAR"]


def sass(lib, dst):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for mn in SASS_MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    kernels[cur][mn] += 1
                    break
            else:
                kernels[cur][op.split(".")[0]] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    cols = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "HMMA", "FFMA"]
    with open(dst, "w") as out:
        out.write(f"# cuobjdump -sass {lib}: instruction counts per kernel (sm_100a)\n")
        out.write("# UTCHMMA / UTCQMMA = tcgen05.mma (kind::tf32 / f16), LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit,\n")
        out.write("# UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier ops, HMMA = mma.sync (none expected)\n\n")
        out.write(f"{'kernel':70s} " + " ".join(f"{c:>8s}" for c in cols) + f" {'total':>8s}\n")
        tot = collections.Counter()
        for k, cnt in sorted(kernels.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCQMMA"] + kv[1]["LDTM"] + kv[1]["UBLKCP"] + kv[1]["UTMALDG"])):
            nm = clean(names[k])[:70]
            out.write(f"{nm:70s} " + " ".join(f"{cnt[c]:8d}" for c in cols) + f" {cnt['_total']:8d}\n")
            tot.update(cnt)
        out.write(f"\n{'ALL KERNELS (' + str(len(kernels)) + ')':70s} " + " ".join(f"{tot[c]:8d}" for c in cols) + f" {tot['_total']:8d}\n")
    print(open(dst).read())


if __name__ == "__main__":
    if sys.argv[1] == "sass":
        sass(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], *(sys.argv[4:5] or ["k_count"]))
    else:
        full(sys.argv[2], sys.argv[3])
