#!/usr/bin/env python
"""Turn gpurun_out ncu artefacts into the small text summaries committed under profiles/.
  summarize_profile.py launches <launches.csv> <out.md> "<title>"      (csv may also hold dram__bytes_{read,write}.sum: then
                                                                        profiles/ncu_traffic.json is written next to <out.md>)
  summarize_profile.py kernel <report.ncu-rep> <mangled-substring> <out.md> "<title>"
"""
import csv, collections, os, re, subprocess, sys

def launches(path, out, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot = collections.defaultdict(float); cnt = collections.Counter(); dram = collections.defaultdict(float)
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        try: v = float(row["Metric Value"].replace(",", ""))
        except Exception: continue
        u = row["Metric Unit"]
        if row["Metric Name"].startswith("dram__bytes"):
            dram[name] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            continue
        if row["Metric Name"] != "gpu__time_duration.sum": continue
        v = v / 1e6 if u == "ns" else v / 1e3 if u.startswith("us") else v * 1e3 if u in ("s", "second") else v
        tot[name] += v; cnt[name] += 1
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches: compare SHARES, not absolutes)\n\n")
        f.write(f"launches captured: {sum(cnt.values())}, total {T:.1f} ms\n\n| kernel | launches | total ms | share | avg ms |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{k.strip()}` | {cnt[k]} | {v:.3f} | {v / T:.4f} | {v / cnt[k]:.4f} |\n")
        if dram:
            import json
            f.write("\nDRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch:\n\n| kernel | launches | total GB | avg MB / launch |\n|---|---:|---:|---:|\n")
            for k, v in sorted(dram.items(), key=lambda x: -x[1]):
                f.write(f"| `{k.strip()}` | {cnt[k]} | {v / 1e9:.2f} | {v / max(cnt[k], 1) / 1e6:.1f} |\n")
            bn = sum(c for k, c in cnt.items() if "bounce_kernel" in k); bb = sum(v for k, v in dram.items() if "bounce_kernel" in k)
            json.dump({"bounce_dram_bytes_per_launch": bb / max(bn, 1), "bounce_launches": bn, "source": os.path.basename(out),
                       "note": "average over every bounce_kernel launch of one bench step (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum)"},
                      open(os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json"), "w"), indent=1)

def kernel(rep, fn, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]; idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
            "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nncu --set full --clock-control none --import-source on; report read with `ncu -i ... --page raw --csv` / `--page source --csv`\n\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"## launch {n}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in want:
                if w in idx: f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            st = {h: float(r[idx[h]]) for h in hdr if "average_warps_issue_stalled" in h and r[idx[h]] not in ("", "n/a")}
            f.write("\nstall reasons (warps per issue-active cycle): " + ", ".join(f"{k.split('stalled_')[1].replace('_per_issue_active.ratio', '')} {v:.2f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]) + "\n\n")
        by = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_by_line.py"), rep, "bounce" if "bounce" in fn else "commit", fn], capture_output=True, text=True, env=dict(os.environ, TOP="30")).stdout
        f.write("## hottest CUDA source lines of launch 0 (stall samples / warp instructions / active lanes per instruction)\n\n```\n" + by + "```\n")

if __name__ == "__main__":
    if sys.argv[1] == "launches": launches(*sys.argv[2:5])
    else: kernel(*sys.argv[2:6])
