"""Rebuild the round-2 summaries under profiles/ from the files scripts/r02_capture.sh left in gpurun_out/.
usage: python scripts/refresh_profiles_r02.py COMMIT"""
import collections, csv, json, shutil, subprocess, sys

commit = sys.argv[1]
G = "gpurun_out/r02_"
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}

# ---- launch list of ONE step of the bench command (cudaProfilerStart/Stop around the step: bench.py --profile-step)
rows = list(csv.reader(l for l in open(G + "launches_bench_step.csv") if not l.startswith("==")))
hdr = rows[0]
ii, ki, mi, ui, vi = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
per = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    per.setdefault(r[ii], {"name": r[ki].split("(")[0].replace("void ", "")})[r[mi]] = float(r[vi].replace(",", "")) * scale[r[ui]]
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["name"], [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d["gpu__time_duration.sum"]
    a[2] += d["dram__bytes_read.sum"]
    a[3] += d["dram__bytes_write.sum"]
tot = sum(a[1] for a in agg.values())
bench = json.loads(open(G + "bench_1gpu.json").read().strip().splitlines()[-1])
with open("profiles/r02_launches_bench_step_1gpu.txt", "w") as f:
    f.write("# ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none:\n"
            "#   python bench.py --steps 1 --warmup 2 --profile-step     (ONE step of the headline run: full config 3, 10,000 queries x\n"
            f"#   1,000,000 targets x G=4096, top-100, one B200; the same command exited 0 without ncu first); commit {commit}\n"
            "# per-launch times are cold-cache and serialised: compare SHARES with bench.py's kernel_share_of_step\n"
            f"#   (bench, events, same commit: pre-filter {bench['roofline']['kernel_ms_per_step']:.1f} ms of {bench['ms_per_step']:.1f} ms = "
            f"{100 * bench['roofline']['kernel_share_of_step']:.1f} %)\n"
            "# kernel, launches, total ms, share of the step's kernel time, DRAM read MB, DRAM write MB\n")
    for k, (n, ms, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:<58s} {n:4d} {ms:10.3f} ms {100 * ms / tot:6.2f}% {rd / 1e6:10.1f} {wr / 1e6:10.1f}\n")
    f.write(f"{'total':<58s} {sum(a[0] for a in agg.values()):4d} {tot:10.3f} ms\n")
shutil.copy(G + "launches_bench_step.csv", "profiles/r02_launches_bench_step_1gpu.csv")
pf = [a for k, a in agg.items() if k.startswith("ppp_prefilter2_kernel<")]  # all-ones / general / P == Y instantiations
traffic = sum(a[2] + a[3] for a in pf)
cfg = bench["config"]
json.dump({"workload": [cfg["genomes"], cfg["queries"], cfg["targets"], 100, 1], "dram_bytes_per_step": traffic,
           "launches_per_step": sum(a[0] for a in pf), "kernel": "ppp_prefilter2_kernel (all three instantiations)",
           "source": f"profiles/r02_launches_bench_step_1gpu.csv: dram__bytes_read.sum + dram__bytes_write.sum over the {sum(a[0] for a in pf)} "
                     f"ppp_prefilter2_kernel launches of one step of `python bench.py --steps 1 --warmup 2 --profile-step` under ncu, commit {commit}"},
          open("profiles/r02_ncu_traffic.json", "w"), indent=1)
print("pre-filter DRAM traffic per step:", traffic / 1e6, "MB in", sum(a[0] for a in pf), "launches; share of kernel time",
      sum(a[1] for a in pf) / tot)

# ---- bench lines
for src, dst in (("bench_1gpu.json", "r02_bench_1gpu.json"), ("bench_reference_arm.json", "r02_bench_reference_arm.json")):
    shutil.copy(G + src, "profiles/" + dst)

# ---- ncu --set full summaries (bulk launches of the same step)
for rep, head in (("prefilter2_bench.ncu-rep", "ppp_prefilter2_kernel, the three launches of one bulk target range (all-ones presence planes / general / P == Y)"),
                  ("refine_bench.ncu-rep", "ppp_refine_kernel, one bulk launch")):
    txt = subprocess.run([sys.executable, "scripts/ncu_summary.py", G + rep], capture_output=True, text=True).stdout
    open("profiles/r02_ncu_" + rep.replace("_bench.ncu-rep", ".txt"), "w").write(
        f"# r02 {head}\n# source: ncu --profile-from-start off --set full --import-source on --clock-control none (B200, sm_100a) of\n"
        f"#   python bench.py --steps 1 --warmup 2 --profile-step   (full config 3 on one GPU); summary by scripts/ncu_summary.py; commit {commit}\n" + txt)
