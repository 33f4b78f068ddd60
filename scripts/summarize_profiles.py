"""Turns the ncu captures under gpurun_out/ into the tracked summaries under profiles/ (run here, no GPU needed):
   python scripts/summarize_profiles.py <round-tag> <launches.csv> <full.ncu-rep>"""
import csv, io, json, re, subprocess, sys, collections
tag, launches, rep = sys.argv[1:4]

# ---- launch list: per-kernel share of one step
rows = [r for r in csv.reader(open(launches)) if r]
hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[hdr_i]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
mu = hdr.index("Metric Unit")
per = collections.OrderedDict()
seq = []
for r in rows[hdr_i + 1:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("resep::", "").replace("void ", "")
    t = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1e-3)
    seq.append((name, t))
# one forward pass = from one k_encoder_chunked to the next
starts = [i for i, (n, _) in enumerate(seq) if n.startswith("k_encoder_chunked")]
a, b = (starts[0], starts[1]) if len(starts) > 1 else (0, len(seq))
for n, t in seq[a:b]:
    d = per.setdefault(n, [0, 0.0])
    d[0] += 1; d[1] += t
tot = sum(v[1] for v in per.values())
out = {"source": launches, "note": "ncu --metrics gpu__time_duration.sum --clock-control none; one forward pass (encoder..decoder) of bench.py config 2; "
       "per-launch times are cold-cache and serialised, compare SHARES not absolutes", "total_us": round(tot, 1),
       "kernels": [{"kernel": n, "launches": v[0], "us": round(v[1], 1), "share": round(v[1] / tot, 4)} for n, v in sorted(per.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(f"profiles/{tag}_launch_shares.json", "w"), indent=1)
print(json.dumps(out["kernels"][:6]))

# ---- full capture: key metrics per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, u = rr[0], rr[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size", "smsp__inst_executed.sum"]
summ = []
for r in rr[2:]:
    d = {"kernel": re.sub(r"\(.*", "", r[h.index("Kernel Name")]).replace("resep::", "").replace("void ", "")}
    for w in want:
        if w in h:
            d[w] = f"{r[h.index(w)]} {u[h.index(w)]}".strip()
    summ.append(d)
json.dump({"source": rep, "note": "ncu --set full --clock-control none --import-source on, one 64,800-row layer (scripts/gpu_post_time.py); cold caches",
           "kernels": summ}, open(f"profiles/{tag}_ncu_layer_kernels.json", "w"), indent=1)
for d in summ:
    print(d["kernel"], d.get("gpu__time_duration.sum"), "tensor", d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
          "dram rd", d.get("dram__bytes_read.sum"), "wr", d.get("dram__bytes_write.sum"))
