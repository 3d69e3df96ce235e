"""Turn gpurun_out/ artefacts of tools/gpu_full.sh into the tracked summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("VTTS_PROFILE_OUT", os.path.join(ROOT, "profiles")); G = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs(OUT, exist_ok=True)

# 1. launch list (ncu --metrics gpu__time_duration.sum,dram bytes; cold-cache, serialised): CSV + share table
per = {}
with open(os.path.join(G, "launches.csv")) as fh:
    lines = [l for l in fh if l.startswith('"')]
for r in csv.DictReader(lines):
    d = per.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"].split("(")[0], "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    else:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        d[r["Metric Name"]] = v * mult
rows = [(i, d["kernel"], d["grid"], d["block"], d.get("us", 0.0), d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0))
        for i, d in sorted(per.items())]
with open(os.path.join(OUT, f"{tag}_launch_list.csv"), "w") as fh:
    fh.write("id,kernel,grid,block,us,dram_read_bytes,dram_write_bytes\n")
    for r in rows:
        fh.write(f'{r[0]},"{r[1]}","{r[2]}","{r[3]}",{r[4]:.3f},{r[5]:.0f},{r[6]:.0f}\n')
tot = sum(r[4] for r in rows)
traffic = sum(r[5] + r[6] for r in rows)
agg = {}
for r in rows:
    a = agg.setdefault(r[1], [0, 0.0, 0.0]); a[0] += 1; a[1] += r[4]; a[2] += r[5] + r[6]
n_l = len(rows)
md = [f"# {tag}: launch list of one HiFi-GAN V1 forward (fp16, B=16, T=759, no padding trim), ncu\n",
      "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
      f"-k 'regex:conv_tc|unit_tc|unit64_tc|chain_tc|conv_post|cf_to_cl' -s {n_l} -c {n_l} --csv python tools/ncu_forward.py`",
      "(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n",
      "| kernel | launches | total us | share | DRAM read+write MB |", "|---|---|---|---|---|"]
for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    md.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / tot:.1f}% | {by / 1e6:.1f} |")
md.append(f"\nTotal {tot:.1f} us over {len(rows)} launches; DRAM traffic of the launch set {traffic / 1e9:.3f} GB.\n")
json.dump({"launches": len(rows), "dram_bytes_per_forward": traffic, "ncu_us_per_forward": tot,
           "workload": "HiFi-GAN V1 forward, fp16 operands, B=16, T=759 frames, no padding trim"},
          open(os.path.join(OUT, f"{tag}_traffic.json"), "w"))

# 2. full captures
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]
REPORTS = {
    "r01": (("prof_unit_c128_k11", "stage 1 (128 ch) fused unit k=11, d=1 (`unit_tc_kernel`) -- tensor / L2-weight-stream bound"),
            ("prof_unit64_c64_k11", "stage 2 (64 ch) fused unit k=11, d=1 (`unit64_tc_kernel`, M=64, resident weights)"),
            ("prof_unit64_c32_k3", "stage 3 (32 ch) fused unit k=3, d=1 (`unit64_tc_kernel`) -- HBM-bound"),
            ("prof_conv_c256_k11", "stage 0 (256 ch) k=11: conv1 then conv2 (`conv_tc_kernel`)")),
    "r02": (("prof_chain_c32_k11", "stage 3 (32 ch) ResidualBlock k=11, all 6 convs in one launch (`chain_tc_kernel<32>`, streamed weights)"),
            ("prof_chain_c32_k3", "stage 3 (32 ch) ResidualBlock k=3 (`chain_tc_kernel<32>`, resident weights)"),
            ("prof_chain_c64_k7", "stage 2 (64 ch) ResidualBlock k=7 (`chain_tc_kernel<64>`, streamed weights)"),
            ("prof_unit_c128_k11", "stage 1 (128 ch) fused unit k=11, d=1 (`unit_tc_kernel`)"),
            ("prof_conv_c256_k11", "stage 0 (256 ch) k=11: conv1 then conv2 (`conv_tc_kernel`)"),
            ("prof_up1_256_128", "upsample 256 -> 128, x8 polyphase (`conv_tc_kernel`, fp32 time-packed + 16-bit copy out)"),
            ("prof_up3_64_32", "upsample 64 -> 32, x2 polyphase (`conv_tc_kernel`, fp32 channels-last out for the chain kernel)"),
            ("prof_conv_post", "output conv 32 -> 1, k=7 + tanh on the channels-last stream (`conv_post_cl_kernel`)"),
            ("prof_lr_gather", "LengthRegulator gather at (256, 330 -> ~2200, 256) (`lr_gather_kernel`) -- HBM-bound"),
            ("prof_gauss", "GaussianUpsampling (16, 120 -> ~760, 256) (`gauss_upsample_kernel`)"),
            ("prof_path", "vits2 generate_path / expansion (16, 192, 120 -> ~760) (`path_generate_kernel`, `path_expand_kernel`)")),
}
for rep, what in REPORTS.get(tag, REPORTS["r02"]):
    path = os.path.join(G, rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    idx = {h: i for i, h in enumerate(hdr)}
    md.append(f"## ncu --set full: {what}\n")
    md.append(f"Report: `{rep}.ncu-rep` (scratch, not committed); command in `tools/gpu_full.sh` (r01) / `tools/gpu_profile_r02.sh` (r02)\n")
    nl = len(rr) - 2
    md.append("| metric | unit | " + " | ".join(f"launch {i + 1}" for i in range(nl)) + " |"); md.append("|---|---|" + "---|" * nl)
    for w in want:
        if w in idx:
            vals = [r[idx[w]] for r in rr[2:2 + nl]]
            md.append(f"| {w} | {units[idx[w]]} | " + " | ".join(vals) + " |")
    # warp-stall distribution from the source page
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    hdr2, tots, samples = None, {}, 0
    for r in csv.reader(src.splitlines()):
        if r and r[0] == "Address":
            if hdr2 is not None: break          # first kernel only
            hdr2 = r; ix = {h: i for i, h in enumerate(hdr2)}; continue
        if hdr2 is None or len(r) < len(hdr2): continue
        try: n = int(r[ix["# Samples"]])
        except ValueError: continue
        samples += n
        for h in hdr2:
            if h.startswith("stall_") and "Not Issued" not in h:
                try: tots[h] = tots.get(h, 0) + int(r[ix[h]])
                except ValueError: pass
    if samples:
        top = sorted(tots.items(), key=lambda kv: -kv[1])[:6]
        md.append("\nWarp-state samples (launch 1): " + ", ".join(f"{h} {100 * v / samples:.1f}%" for h, v in top) + "\n")
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
