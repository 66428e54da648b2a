"""Rebuild the tracked summaries under profiles/ from one measurement run's files in gpurun_out/.
usage: python scripts/refresh_profiles.py PREFIX COMMIT [PREFILTER_LAUNCHES_PER_STEP]   (files gpurun_out/PREFIX_*, see the gpurun command in profiles/README)"""
import collections, csv, io, re, shutil, subprocess, sys
pre, commit = sys.argv[1], sys.argv[2]
n_step_launches = int(sys.argv[3]) if len(sys.argv) > 3 else None  # pre-filter launches of ONE step (the capture may run into the next)
G = "gpurun_out/" + pre + "_"


def rows_of(path):
    return list(csv.reader(l for l in open(path) if not l.startswith("==")))


# ---- launch list of the bench command
rows = rows_of(G + "launches_bench.csv")
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
with open("profiles/r01_launches_bench_config3_1gpu.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 1 --warmup 1 (config 3 per-GPU share, top-k 100;\n"
            f"# 2 resident-input runs + 2 host-buffer runs of the scan + the config-2 side measurement + the CPU baseline); commit {commit}\n"
            "# per-launch times are cold-cache and serialised: compare SHARES. kernel, launches, total ms, share\n")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:<62s} {n:4d} {ms:10.3f} ms {100 * ms / tot:6.2f}%\n")
shutil.copy(G + "launches_bench.csv", "profiles/r01_launches_bench_config3_1gpu.csv")
for src, dst in (("bench.json", "r01_bench_1gpu.json"), ("bench_ref.json", "r01_bench_reference_arm.json"), ("other_configs.json", "r01_other_configs.json")):
    shutil.copy(G + src, "profiles/" + dst)

# ---- ncu --set full summaries of the probe run (scripts/pf_ab.py 2)
txt = subprocess.run([sys.executable, "scripts/ncu_summary.py", G + "prof.ncu-rep"], capture_output=True, text=True).stdout.split("----\n")[1:]
src = ("# source: ncu --set full --clock-control none --import-source on (B200, sm_100a), command: python scripts/pf_ab.py 2 "
       f"(2048 queries x 125000 targets x G=4096, top-100); summary by scripts/ncu_summary.py; commit {commit}\n")
for body in txt:
    name = body.split("\n", 1)[0]
    if "refine" in name:
        fn, head = "r01_ncu_refine.txt", "# r01 ppp_refine_kernel<4> (warp per surviving pair: full tier A/B evaluation), speculative chunk\n"
    elif "<4, 1>" in name:
        fn, head = "r01_ncu_prefilter2_all_ones.txt", "# r01 ppp_prefilter2_kernel<4, true> (queries with an all-ones presence plane), speculative chunk\n"
    else:
        fn, head = "r01_ncu_prefilter2_general.txt", "# r01 ppp_prefilter2_kernel<4, false> (general presence planes), speculative chunk\n"
    open("profiles/" + fn, "w").write(head + src + body)

# ---- DRAM traffic of the bench step's pre-filter launches
rows = rows_of(G + "traffic_bench_pf.csv")
hdr = rows[0]
ii, ki, mi, ui, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "")})[r[mi]] = (float(r[vi].replace(",", "")), r[ui])
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
total = 0.0
with open("profiles/r01_ncu_bench_prefilter2_traffic.txt", "w") as f:
    f.write("# r01 ppp_prefilter2_kernel launches of ONE bench step (config 3 per-GPU share, top-k 100): ncu --metrics dram__bytes_read.sum,\n"
            "# dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:ppp_prefilter2 -c N python bench.py\n"
            f"# --steps 1 --warmup 1 (first run of the command = one step); commit {commit}.  The step's total is bench.py's roofline.traffic.\n"
            "# launch, kernel, ms, DRAM read MB, DRAM write MB, L2 hit %\n")
    ids = list(per)[:n_step_launches]
    n_step = len(ids) // 1
    # one step = launches up to (not including) the first repetition of the first launch's duration pattern: the command runs 4 scans,
    # each with the same launch sequence, so one step = len/4 launches when the capture covers whole scans; otherwise all captured
    for i, k in enumerate(ids):
        d = per[k]
        rd = d["dram__bytes_read.sum"][0] * scale[d["dram__bytes_read.sum"][1]]
        wr = d["dram__bytes_write.sum"][0] * scale[d["dram__bytes_write.sum"][1]]
        ms = d["gpu__time_duration.sum"][0] / (1e6 if d["gpu__time_duration.sum"][1] in ("ns", "nsecond") else 1e3 if d["gpu__time_duration.sum"][1] in ("us", "usecond") else 1)
        total += rd + wr
        f.write(f"{i:3d} {d['name']:<34s} {ms:8.3f} {rd / 1e6:9.3f} {wr / 1e6:8.3f} {d['lts__t_sector_hit_rate.pct'][0]:6.1f}\n")
    f.write(f"# total DRAM bytes over these {len(ids)} launches: {int(total)}\n")
print("traffic total", int(total), "launches", len(per))
