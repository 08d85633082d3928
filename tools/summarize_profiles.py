"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (run in the dev container).
usage: python tools/summarize_profiles.py <round tag, e.g. r01>"""
import collections, csv, json, os, re, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []

def launches(path, title=None, keep=None):
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v * 1e6 if u == "s" else v
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        if keep and not keep(name, row["Grid Size"]):
            continue
        agg[name + " grid=" + row["Grid Size"]].append(v)
    tot = sum(sum(v) for v in agg.values())
    out.append(f"## {title or 'Launch list'} ({os.path.basename(path)}; `ncu --metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised)\n")
    out.append("| kernel | launches | avg us | share |\n|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"| `{k}` | {len(v)} | {sum(v)/len(v):.2f} | {100*sum(v)/tot:.1f} % |")
    out.append(f"\ntotal {tot/1000:.2f} ms over {sum(len(v) for v in agg.values())} launches\n")

def full(path, title):
    if not os.path.exists(path):
        return
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
            "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
    out.append(f"## {title} (`{os.path.basename(path)}`, `ncu --set full --clock-control none --import-source on`)\n")
    for line in rows[2:]:
        name = line[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        out.append(f"kernel `{name[:80]}`\n\n| metric | value |\n|---|---|")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                out.append(f"| {w} | {line[i]} {units[i]} |")
        out.append("")

launches(f"gpurun_out/launches_{tag}.csv")
# the batched decode step of config 3 (batch 64, sampling): only the kernels of the per-step graph (64-row launches)
launches(f"gpurun_out/launches_{tag}_cfg3.csv", "Launch list of the batched decode step (bench.py --config 3 --faces 16: contexts 258..402)",
         keep=lambda n, g: any(k in n for k in ("gemm_ws_kernel", "attention_stream_kernel", "sample_kernel", "embed_tokens", "set_flag"))
         or ("layernorm_kernel" in n and g.replace(" ", "").startswith("(64,")))
full(f"gpurun_out/prof_mega_{tag}.ncu-rep", "Persistent decode kernel")
full(f"gpurun_out/prof_gemm_tc_{tag}.ncu-rep", "tcgen05 + TMA GEMM of the encoder (batch 8: M = 2056 / 32768 rows)")
full(f"gpurun_out/prof_attn_tc_{tag}.ncu-rep", "tcgen05 flash attention of the encoder (batch 8: cross-attention 257 x 4096 keys, self-attention 257 x 257)")
full(f"gpurun_out/prof_gemm_ws_{tag}.ncu-rep", "Weight-streaming tcgen05 GEMM of the batched decode step (M = 64; K slices = one thread-block cluster)")
full(f"gpurun_out/prof_attn_stream_{tag}.ncu-rep", "Persistent pipelined decode attention of a batch (batch 64, 7459 / 4096 keys)")
full(f"gpurun_out/prof_batched_{tag}.ncu-rep", "Canonical CUDA-core GEMM and attention kernels (batch 8 decode / prefill)")
def traffic(path, tokens):
    if not os.path.exists(path):
        return
    r = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    hdr, units, line = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name)
        v = float(line[i].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[i], 1)
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    alg = 621205504 + 98304 * ((258 + 258 + tokens - 1) / 2 + 1)
    json.dump({"kernel": "decode_mega_kernel",
               "capture": f"profiles/ncu_summary_{tag}.md ({os.path.basename(path)}: bench.py --faces 16, one launch = {tokens} decode tokens at contexts 258..{257 + tokens})",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "tokens_in_launch": tokens,
               "traffic_bytes_per_token": (rd + wr) / tokens, "algorithmic_bytes_per_token": alg},
              open(f"profiles/traffic_{tag}.json", "w"), indent=1)

traffic(f"gpurun_out/prof_mega_{tag}.ncu-rep", 145)
os.makedirs("profiles", exist_ok=True)
open(f"profiles/ncu_summary_{tag}.md", "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])

# ---- profiles/README.md from the bench line
bj = f"gpurun_out/bench_{tag}.json"
if os.path.exists(bj) and os.path.getsize(bj) > 0:
    d = json.load(open(bj))
    import shutil
    shutil.copy(bj, f"profiles/bench_{tag}.json")
    r = d["roofline"]
    md = [f"# profiles — round {tag}\n",
          "All numbers from `gpurun` on one B200 of the pool (SM clock %s MHz during the timed region, reasons %s).\n" % (
              d["clocks"]["sm_mhz"], d["clocks"]["reasons"]),
          "## bench.py (default: batch 1, 800 faces, greedy; `bench_%s.json`)\n" % tag,
          "| quantity | value |\n|---|---|",
          f"| face-tokens/s, inputs resident (`value`) | {d['value']:.1f} |",
          f"| face-tokens/s, host in / host out (`e2e`) | {d['e2e']['value']:.1f} |",
          f"| ms per step (one 800-face mesh) | {d['ms_per_step']:.1f} |",
          f"| stage split (ms) | encoder {d['config']['stage_ms']['encoder']:.2f}, generate {d['config']['stage_ms']['generate']:.1f}, detokenize+rest {d['config']['stage_ms']['detokenize_and_rest']:.2f} |",
          f"| decode loop: us per token (average over contexts 258..7458) | {r['us_per_step_avg']:.1f} |",
          f"| decode loop: achieved HBM GB/s (algorithmic bytes) / measured peak | {r['achieved']:.0f} / {r['peak']:.1f} = **{r['frac']:.3f}** |",
          f"| short context (~357..557): us per token, fraction of peak | {r['short_context']['us_per_step']:.1f}, {r['short_context']['frac']:.3f} |",
          f"| prefill (257 prefix tokens) | {r['prefill_ms']:.2f} ms |",
          f"| CPU baseline (oracle, {d['cpu_baseline']['cores'] if d.get('cpu_baseline') else '?'} cores) | {d['cpu_baseline']['value'] if d.get('cpu_baseline') else float('nan'):.1f} tokens/s |",
          f"| kernels launched in the timed region | {d['gpu_launches']} |",
          "",
          ""]
    def other(fn, label):
        path = f"profiles/{fn}"
        if not os.path.exists(path):
            return
        o = json.load(open(path))
        rr = o["roofline"]
        md.append(f"| {label} (`{fn}`) | {o['value']:.0f} tokens/s (e2e {o['e2e']['value']:.0f}), {o['ms_per_step'] / 1e3:.1f} s per step of "
                  f"{o['config']['global_batch']} meshes on {o['n_gpus']} GPU(s), {rr['us_per_step_avg'] / 1e3:.2f} ms per decode step, "
                  f"{rr['frac']:.3f} of the HBM roofline |")
    md += ["## Other BASELINE configurations (full runs, `bench.py --config N --lean`)\n", "| run | result |\n|---|---|"]
    other(f"bench_{tag}_cfg3.json", "config 3: batch 64, F = 800, top-k / top-p")
    other(f"bench_{tag}_cfg3_before_stream_attention.json", "config 3 before attention_stream_kernel / cluster GEMM / PDL")
    other(f"bench_{tag}_cfg4_n8.json", "config 4: 512 shapes on 8 GPUs (64 per GPU)")
    other(f"bench_{tag}_cfg5.json", "config 5: batch 32 per GPU, F = 1600")
    other(f"bench_{tag}_cfg5_before_stream_attention.json", "config 5 before attention_stream_kernel / cluster GEMM / PDL")
    md += ["",
           "Bounded decode-step measurements of configs 3 and 5 at three contexts each are in the `extra` block of `bench_%s.json`." % tag,
           "",
           "## Files\n",
           "* `ncu_summary_%s.md` — launch list of a bench run + `ncu --set full` metrics of `decode_mega_kernel`, `gemm_tc_kernel`, "
           "`attention_tc_kernel`, `gemm_ws_kernel` (cluster split-K), `attention_stream_kernel`; `traffic_%s.json` = DRAM bytes per token of the "
           "persistent kernel from that capture." % (tag, tag),
           "* `batched_kernels_%s.json` — per-kernel A/B timings of the batched decode step (attention old / streaming at batch 8, 32, 64; every "
           "decoder GEMM on the canonical, tiled tcgen05 and weight-streaming tcgen05 kernels with both K-slice reductions)." % tag,
           "* `microbench_cluster_%s.txt` — cluster co-residency, DSMEM and flagged-word exchange micro-benchmarks behind DESIGN 4.1.2." % tag,
           "* `mega_trace_%s*.txt` — phase timelines of the persistent kernel from its own globaltimer stamps (the row-split kernel that ships "
           "and the abandoned column-split design)." % tag,
           "* `sass_histogram_%s.md` — SASS mnemonics per kernel (UTCHMMA / UTMALDG / FHFMA / SYNCS evidence)." % tag,
           ""]
    open("profiles/README.md", "w").write("\n".join(md))
    print("\n".join(md))
if os.path.exists(f"gpurun_out/mega_trace_{tag}.txt"):
    import shutil
    shutil.copy(f"gpurun_out/mega_trace_{tag}.txt", f"profiles/mega_trace_{tag}.txt")
