"""Turn gpurun_out/ artefacts of tools/gpu_full.sh into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

# 1. launch list (ncu --metrics gpu__time_duration.sum, cold-cache, serialised): keep as CSV + share table
rows = []
with open(os.path.join(G, "launches.csv")) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    rows.append((int(r["ID"]), r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3))
with open(os.path.join(OUT, f"{tag}_launch_list.csv"), "w") as fh:
    fh.write("id,kernel,grid,block,us\n")
    for r in rows:
        fh.write(f'{r[0]},"{r[1]}","{r[2]}","{r[3]}",{r[4]:.3f}\n')
tot = sum(r[4] for r in rows)
agg = {}
for r in rows:
    agg.setdefault(r[1], [0, 0.0]); agg[r[1]][0] += 1; agg[r[1]][1] += r[4]
md = [f"# {tag}: launch list of one HiFi-GAN V1 forward (fp16, B=16, T=759), ncu gpu__time_duration.sum\n",
      "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -k regex:\"conv_tc|conv_post|cf_to_cl|lr_\" -s 79 -c 80 --csv python tools/ncu_forward.py`",
      "(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n",
      "| kernel | launches | total us | share |", "|---|---|---|---|"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% |")
md.append(f"\nTotal {tot:.1f} us over {len(rows)} launches.\n")

# 2. full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for rep, what in (("prof_stage1_k11", "stage 1 (128 ch) k=11: conv1 (operand copy only) then conv2 (residual epilogue) -- tensor-bound layers"),
                  ("prof_stage3_k3", "stage 3 (32 ch) k=3: conv1 then conv2 -- HBM-bound layers")):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    md.append(f"## ncu --set full: {what}\n")
    md.append(f"Report: `{rep}.ncu-rep` (scratch, not committed); command: `ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s <n> -c 2 python tools/ncu_forward.py`\n")
    md.append("| metric | unit | launch 1 | launch 2 |"); md.append("|---|---|---|---|")
    for w in want:
        if w in idx:
            vals = [r[idx[w]] for r in rr[2:4]]
            md.append(f"| {w} | {units[idx[w]]} | " + " | ".join(vals) + " |")
    md.append("")
open(os.path.join(OUT, f"{tag}_ncu_summary.md"), "w").write("\n".join(md))

# 3. bench lines
for name in ("bench_default.log", "bench_reference.log"):
    p = os.path.join(G, name)
    if os.path.exists(p):
        last = [l for l in open(p) if l.startswith("{")]
        if last:
            open(os.path.join(OUT, f"{tag}_{name.replace('.log', '.json')}"), "w").write(last[-1])
print(open(os.path.join(OUT, f"{tag}_ncu_summary.md")).read()[:6000])
