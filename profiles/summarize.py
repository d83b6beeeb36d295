#!/usr/bin/env python3
"""Turn the csv exports of profiles/capture.sh (gpurun_out/<tag>_*.csv) into the tracked summaries:
   profiles/<tag>_launches.csv (+ _summary.md), profiles/<tag>_ncu_full_raw.csv, profiles/<tag>_ncu_full_table.md and
   profiles/ncu_traffic.json (what bench.py reads for roofline.traffic).   python profiles/summarize.py <tag> [bench.json]"""
import collections
import csv
import json
import os
import re
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")


def short(name):
    m = re.match(r"(?:void )?(k_\w+)(<[^>]*>)?", name)
    if not m:
        return name[:40]
    base, targs = m.group(1), m.group(2) or ""
    if base in ("k_first_march", "k_correct", "k_final", "k_ppm_slope", "k_ppm_states", "k_ppm_flux"):
        return f"{base}<{targs[1:].split(',')[0].strip()}>"
    if base == "k_ion_rates":        # <UPDATE[, MEMO, TASKS]> (the last two since r2h: sparse sub-cycle)
        a = [x.strip() for x in targs[1:-1].split(",")] if targs else []
        if a and a[0] in ("1", "true", "(bool)1"):
            return "k_ion_rates<fused update>"
        if len(a) >= 3 and a[2] in ("1", "true", "(bool)1"):
            return "k_ion_rates[memo]"           # over k_ion_walk's task list
        if len(a) == 4 and a[3] in ("1", "true", "(bool)1"):
            return "k_ion_rates[first]" if a[1] in ("1", "true", "(bool)1") else "k_ion_rates[first]<no memo>"
        if len(a) >= 3 and a[1] in ("0", "false", "(bool)0"):
            return "k_ion_rates<no memo>"
        return "k_ion_rates"                     # full pass (writes the chunk records when the memo is on)
    if base == "k_ion_update":       # <MEMO, TASKS>
        a = [x.strip() for x in targs[1:-1].split(",")] if targs else []
        if len(a) == 2 and a[1] in ("1", "true", "(bool)1"):
            return "k_ion_update[memo]"
        if len(a) == 2 and a[0] in ("0", "false", "(bool)0"):
            return "k_ion_update<no memo>"
        return "k_ion_update"
    return base


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def launches(tag, bench):
    src = os.path.join(OUT, f"{tag}_launches.csv")
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    ni, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        ms = fnum(r[vi]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
        k = re.sub(r"<.*", "", short(r[ni]))
        tot[k] += ms; cnt[k] += 1
    shutil.copy(src, os.path.join(HERE, f"{tag}_launches.csv"))
    live = {}
    if bench:
        d = json.loads(open(bench).read().strip().splitlines()[-1])
        for k, v in d["kernels"].items():
            live[re.sub(r"<.*", "", k)] = live.get(re.sub(r"<.*", "", k), 0.0) + (v["share"] or 0.0)
    allms = sum(tot.values())
    with open(os.path.join(HERE, f"{tag}_launches_summary.md"), "w") as f:
        f.write(f"# ncu launch list, {tag}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv` on `python bench.py --steps 2 "
                "--warmup 3 --no-cpu-baseline --e2e-steps 1` (256^3 ioniz_sphere), after the same command exited 0 without ncu "
                f"(`profiles/capture.sh`). Raw list: `{tag}_launches.csv`.\n\nPer-launch times under ncu are cold-cache and serialised: compare "
                "SHARES with bench.py's live CUDA-event shares (last column" + (f"; `{os.path.basename(bench)}`" if bench else "") +
                "). The ncu run covers set-up, 3 warm-up steps with more ion iterations per step, and the e2e transposes, so the ion "
                "kernels and transposes weigh more there.\n\n| kernel | launches | total ms | share (ncu) | share (bench.py events) |\n|---|---|---|---|---|\n")
        for k, ms in tot.most_common():
            f.write(f"| {k} | {cnt[k]} | {ms:.3f} | {ms / allms:.3f} | {live.get(k, float('nan')):.4f} |\n".replace("nan", "-"))


def full(tag):
    src = os.path.join(OUT, f"{tag}_ncu_full_raw.csv")
    rows = list(csv.reader(open(src)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    col = {h: i for i, h in enumerate(hdr)}
    for h, i in list(col.items()):          # section-prefixed duplicates ("FBSP.TriageCompute.dram__throughput...")
        col.setdefault(h.split(".TriageCompute.")[-1], i)

    def val(r, name, scale_bytes=False):
        i = col[name]
        x = fnum(r[i])
        if scale_bytes:
            x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
        return x

    def ms(r):
        i = col["gpu__time_duration.sum"]
        return fnum(r[i]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1e-6)

    agg = collections.OrderedDict()
    for r in rows[2:]:
        agg.setdefault(short(r[ni]), []).append(r)
    shutil.copy(src, os.path.join(HERE, f"{tag}_ncu_full_raw.csv"))
    traffic = {}
    with open(os.path.join(HERE, f"{tag}_ncu_full_table.md"), "w") as f:
        f.write(f"# ncu --set full, {tag}: one main-loop step of the default bench workload (256^3 ioniz_sphere)\n\n"
                "Averages over the captured launches of each kernel (`profiles/capture.sh`; raw: "
                f"`{tag}_ncu_full_raw.csv`).  `--clock-control none`; times are cold-cache, serialised.\n\n"
                "| kernel | launches captured | time/launch (ms) | DRAM read (GB) | DRAM write (GB) | DRAM % of peak | FP64 pipe % | issue active % | warps active % | regs | warp-inst (M) |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, rs in agg.items():
            n = len(rs)
            avg = lambda name, **kw: sum(val(r, name, **kw) for r in rs) / n
            t = sum(ms(r) for r in rs) / n
            rd, wr = avg("dram__bytes_read.sum", scale_bytes=True), avg("dram__bytes_write.sum", scale_bytes=True)
            dpct = avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
            fp64 = avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") if "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active" in col else float("nan")
            issue = avg("smsp__issue_active.avg.pct_of_peak_sustained_active")
            wa = avg("sm__warps_active.avg.pct_of_peak_sustained_active")
            regs = avg("launch__registers_per_thread")
            inst = avg("smsp__inst_executed.sum") / 1e6
            f.write(f"| {k} | {n} | {t:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {dpct:.1f} | {fp64:.1f} | {issue:.1f} | {wa:.1f} | {regs:.0f} | {inst:.0f} |\n")
            if t > 0.05:
                traffic[k] = {"dram_bytes_per_launch": rd + wr, "time_ms": t, "dram_pct": dpct, "fp64_pipe_pct": fp64, "regs": regs}
    json.dump(traffic, open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)


def full_multi(tag, names):
    """Round-2 captures (profiles/capture2.sh): one --set full csv per workload, gpurun_out/<tag>_<name>_ncu_full_raw.csv ->
    profiles/<tag>_ncu_full_table.md (a section per workload) + profiles/<tag>_<name>_ncu_full_raw.csv.gz + ncu_traffic.json
    (first capture that has a kernel wins: `default` is the bench workload)."""
    import gzip
    titles = {"smr_ops": "--workload ioniz_smr: the SMR operators, k_correct3 and k_update on the five levels of the shipped deck",
              "pair": "bench.py default workload: k_correct3 / k_final / k_update (the default passes C, D, E since r2)",
              "default": "bench.py default workload (256^3 ioniz_sphere), unfused correct/final pair, separate ion kernels",
              "fused": "the same with AB2_HYDRO_FUSE=1 (k_cfinal) and AB2_ION_FUSE=1 (k_ion_rates<fused update>)",
              "ion": "bench.py default workload, the ion kernels of the sparse sub-cycle: full first-iteration passes (k_ion_rates / k_ion_update, "
                     "writing the chunk records), k_ion_walk, and the passes over its task list ([memo])",
              "first": "bench.py default workload since the call's prologue rides on the first rates pass (k_ion_rates[first]); profiles/capture5.sh",
              "ppm": "--workload blast --order 3 (PPM pass A)", "smr": "--workload ioniz_smr (shipped 5-level deck)"}
    traffic = {}
    with open(os.path.join(HERE, f"{tag}_ncu_full_table.md"), "w") as f:
        f.write(f"# ncu --set full, {tag}\n\n`profiles/capture2.sh {tag}` (r2h: `profiles/capture4.sh`) on one B200: every command first ran plain (exit 0), then its launch list, "
                "then ONE `ncu --set full --clock-control none --import-source on -k regex:^k_` capture.  Averages over the captured launches of "
                "each kernel; times are cold-cache and serialised (compare shares, not absolutes, with bench.py's CUDA-event times).\n")
        for name in names:
            src = os.path.join(OUT, f"{tag}_{name}_ncu_full_raw.csv")
            if not os.path.exists(src) or os.path.getsize(src) < 1000:
                f.write(f"\n## {name}: capture missing\n")
                continue
            rows = list(csv.reader(open(src)))
            hdr, units = rows[0], rows[1]
            ni = hdr.index("Kernel Name")
            col = {h: i for i, h in enumerate(hdr)}
            for h, i in list(col.items()):
                col.setdefault(h.split(".TriageCompute.")[-1], i)

            def val(r, nm, scale_bytes=False):
                if nm not in col:
                    return float("nan")
                i = col[nm]
                x = fnum(r[i])
                if scale_bytes:
                    x *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
                return x

            def ms(r):
                i = col["gpu__time_duration.sum"]
                return fnum(r[i]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units[i], 1e-6)

            agg = collections.OrderedDict()
            for r in rows[2:]:
                agg.setdefault(short(r[ni]), []).append(r)
            with gzip.open(os.path.join(HERE, f"{tag}_{name}_ncu_full_raw.csv.gz"), "wt") as gz:
                gz.write(open(src).read())
            f.write(f"\n## {name}: {titles.get(name, name)}\n\n"
                    "| kernel | launches | time/launch (ms) | DRAM read (GB) | DRAM write (GB) | DRAM % | FP64 pipe % | issue % | LSU pipe % | L1 hit % | L2 hit % | warps active % | regs | smem/block (KB) | warp-inst (M) |\n"
                    "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
            for k, rs in agg.items():
                n = len(rs)
                avg = lambda nm, **kw: sum(val(r, nm, **kw) for r in rs) / n
                t = sum(ms(r) for r in rs) / n
                rd, wr = avg("dram__bytes_read.sum", scale_bytes=True), avg("dram__bytes_write.sum", scale_bytes=True)
                dpct = avg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
                fp64 = avg("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
                issue = avg("smsp__issue_active.avg.pct_of_peak_sustained_active")
                lsu = avg("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active")
                l1 = avg("l1tex__t_sector_hit_rate.pct")
                l2 = avg("lts__t_sector_hit_rate.pct")
                wa = avg("sm__warps_active.avg.pct_of_peak_sustained_active")
                regs = avg("launch__registers_per_thread")
                smem = (avg("launch__shared_mem_per_block_static", scale_bytes=True) + avg("launch__shared_mem_per_block_dynamic", scale_bytes=True)) / 1e3
                inst = avg("smsp__inst_executed.sum") / 1e6
                f.write(f"| {k} | {n} | {t:.3f} | {rd / 1e9:.3f} | {wr / 1e9:.3f} | {dpct:.1f} | {fp64:.1f} | {issue:.1f} | {lsu:.1f} | {l1:.1f} | {l2:.1f} | {wa:.1f} | {regs:.0f} | {smem:.1f} | {inst:.0f} |\n")
                if k not in traffic and t > 0.02:
                    traffic[k] = {"dram_bytes_per_launch": rd + wr, "time_ms": t, "dram_pct": dpct, "fp64_pipe_pct": fp64, "issue_pct": issue,
                                  "warps_active_pct": wa, "regs": regs, "capture": f"{tag}_{name}"}
    path = os.path.join(HERE, "ncu_traffic.json")
    if os.environ.get("NCU_TRAFFIC_MERGE") and os.path.exists(path):      # a supplementary capture adds kernels, keeps the rest
        old = json.load(open(path))
        old.update({k: v for k, v in traffic.items() if k not in old or os.environ["NCU_TRAFFIC_MERGE"] == "override"})
        traffic = old
    json.dump(traffic, open(path, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "multi":
        full_multi(sys.argv[1], sys.argv[3:] or ["default", "fused", "ppm", "smr"])
        print("wrote profiles/%s_ncu_full_table.md" % sys.argv[1])
        sys.exit(0)
    tag = sys.argv[1]
    bench = sys.argv[2] if len(sys.argv) > 2 else None
    launches(tag, bench)
    full(tag)
    print("wrote profiles/%s_*" % tag)
