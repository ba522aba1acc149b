"""Turn the csv files a GPU session left in gpurun_out/ into the committed summaries under profiles/ (run on the CPU box).

  python tools/r2_make_profiles.py <session prefix, e.g. r2s4> <tag, e.g. r02>
"""
import csv, gzip, io, json, os, re, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")


def read_csv_after_banner(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.reader(io.StringIO("".join(lines))))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("lb200::", "")[:120]


def metrics_long(path):
    """--metrics capture in long csv format -> [ {kernel, id, metric: value} ] in launch order"""
    rows = read_csv_after_banner(path)
    h = {n: i for i, n in enumerate(rows[0])}
    out = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) < len(h):
            continue
        d = out.setdefault(r[h["ID"]], {"kernel": short(r[h["Kernel Name"]]), "grid": r[h["Grid Size"]], "block": r[h["Block Size"]]})
        d[r[h["Metric Name"]]] = (r[h["Metric Value"]], r[h["Metric Unit"]])
    return list(out.values())


def num(v):
    return float(v.replace(",", ""))


def launches(pre, tag):
    src = os.path.join(G, pre + "_launches_bench.csv")
    if not os.path.exists(src):
        return
    L = metrics_long(src)
    dst = os.path.join(P, tag + "_launches_bench.csv")
    tot = collections.OrderedDict()
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 (launch order; cold-cache, serialised: compare SHARES)\n")
        f.write("id,kernel,grid,block,duration_ns\n")
        for i, d in enumerate(L):
            v, u = d["gpu__time_duration.sum"]
            ns = num(v) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
            f.write("%d,\"%s\",\"%s\",\"%s\",%.0f\n" % (i, d["kernel"], d["grid"], d["block"], ns))
            t = tot.setdefault(d["kernel"], [0, 0.0]); t[0] += 1; t[1] += ns
    with open(os.path.join(P, tag + "_launches_bench_summary.md"), "w") as f:
        f.write("# per-kernel totals of %s_launches_bench.csv (first 400 launches of `python bench.py --steps 2 --warmup 3` under ncu)\n\n| kernel | launches | total ms | mean us |\n|---|---|---|---|\n" % tag)
        for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f |\n" % (k, n, ns / 1e6, ns / n / 1e3))
    print("wrote", dst)


def single_pass(pre, tag):
    src = os.path.join(G, pre + "_metrics.csv")
    if not os.path.exists(src):
        return
    L = metrics_long(src)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second"]
    dst = os.path.join(P, tag + "_ncu_single_pass_f16x3_8192.md")
    traffic = None
    with open(dst, "w") as f:
        f.write("# ncu --metrics (one pass per launch, --clock-control none): default fp32 path (F16X3) at 8192^3, `tools/r2_ncu_f16_target.py`\n\n")
        f.write("| # | kernel | " + " | ".join(k.split(".")[0].replace("__", " ") + "<br>" + ".".join(k.split(".")[1:]) for k in keys) + " |\n|" + "---|" * (len(keys) + 2) + "\n")
        for i, d in enumerate(L):
            f.write("| %d | `%s` | " % (i, d["kernel"][:60]) + " | ".join("%s %s" % d.get(k, ("-", "")) for k in keys) + " |\n")
            if "gemm_tc_kernel" in d["kernel"]:
                traffic = d       # the last (warm) GEMM launch
        f.write("\nAlgorithmic bytes of the GEMM launch: 805 MB (A + B + C in fp32); the kernel itself reads the prepared fp16 pieces (537 MB) and writes C (268 MB).\n")
    if traffic:
        t = {"dram_bytes_read": num(traffic["dram__bytes_read.sum"][0]), "dram_bytes_write": num(traffic["dram__bytes_write.sum"][0]),
             "unit": "byte", "kernel": traffic["kernel"], "l2_hit_rate_pct": num(traffic["lts__t_sector_hit_rate.pct"][0]),
             "source": "profiles/%s_ncu_single_pass_f16x3_8192.md (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one pass, --clock-control none, last gemm_tc_kernel launch of tools/r2_ncu_f16_target.py)" % tag}
        json.dump(t, open(os.path.join(P, tag + "_traffic.json"), "w"), indent=1)
        print("traffic", t["dram_bytes_read"] / 1e9, t["dram_bytes_write"] / 1e9)
    print("wrote", dst)


FULL_KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size", "launch__cluster_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__cycles_active.avg", "sm__cycles_active.avg",
]


def full(pre, tag, which, title):
    src = os.path.join(G, "%s_%s_raw.csv" % (pre, which))
    if not os.path.exists(src):
        return
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    dst = os.path.join(P, "%s_ncu_full_%s.md" % (tag, which))
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: %s\n\n" % title)
        for n, row in enumerate(data):
            f.write("## launch %d: `%s`\n\n| metric | value |\n|---|---|\n" % (n, short(row[idx["Kernel Name"]])))
            for k in FULL_KEYS:
                if k in idx:
                    f.write("| `%s` | %s %s |\n" % (k, row[idx[k]], units[idx[k]]))
            # every tensor / tmem / stall-ish metric the capture holds
            for k in hdr:
                if "peak_sustained" in k and "pct_of_peak" not in k:
                    continue          # the hardware's own peak constants, not measurements
                if re.search(r"\.(min|max)(\.|$)", k):
                    continue
                if re.search(r"pipe_tensor|tmem|utc|smsp__average_warps_issue_stalled.*_per_issue_active", k) and k not in FULL_KEYS:
                    v = row[idx[k]]
                    if v not in ("", "0", "n/a"):
                        f.write("| `%s` | %s %s |\n" % (k, v, units[idx[k]]))
            f.write("\n")
    print("wrote", dst)


def source_hot(pre, tag):
    src = os.path.join(G, pre + "_full_source.csv.gz")
    if not os.path.exists(src):
        return
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(src), newline="")))
    # find the header row
    hi = next(i for i, r in enumerate(rows) if "Source" in r and any("Sampl" in c for c in r))
    hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
    samp = next(c for c in hdr if c.startswith("# Samples") or c.startswith("Warp Stall Sampling (All"))
    data = []
    for r in rows[hi + 1:]:
        try:
            data.append((float(r[idx[samp]] or 0), r))
        except (ValueError, IndexError):
            pass
    tot = sum(d[0] for d in data) or 1.0
    dst = os.path.join(P, tag + "_ncu_source_hot_gemm_tc.md")
    with open(dst, "w") as f:
        f.write("# hottest SASS lines of gemm_tc_kernel (F16X3, CTA pair) by warp-stall samples (`%s`, total %.0f)\n\n| share | samples | SASS |\n|---|---|---|\n" % (samp, tot))
        for s, r in sorted(data, key=lambda d: -d[0])[:40]:
            f.write("| %.1f %% | %.0f | `%s` |\n" % (100 * s / tot, s, r[idx["Source"]][:110]))
    print("wrote", dst)


def copies(pre, tag):
    """plain copies of the small text artefacts of the session"""
    import shutil
    for src, dst in (("bench_n1.json", "bench_n1.json"), ("bench_ref.json", "bench_reference.json"), ("layouts.log", "layouts.log"),
                     ("probes.jsonl", "probes.jsonl"), ("pytest.log", "pytest.log"), ("smi.txt", "smi.txt"), ("f64.jsonl", "f64_dmma.jsonl"),
                     ("layers_bench.txt", "layers_bench.txt"), ("large_shapes.txt", "large_shapes.txt")):
        a = os.path.join(G, "%s_%s" % (pre, src))
        if os.path.exists(a):
            shutil.copy(a, os.path.join(P, "%s_%s" % (tag, dst)))
    with open(os.path.join(P, tag + "_compute_sanitizer.txt"), "w") as f:
        f.write("# compute-sanitizer on tools/sanitizer_target.py (every kernel family once, incl. the kernels of round 2)\n")
        for tool in ("memcheck", "racecheck"):
            a = os.path.join(G, "%s_sanitizer_%s.log" % (pre, tool))
            if os.path.exists(a):
                f.write("## --tool %s\n" % tool)
                for ln in open(a, errors="replace"):
                    if re.search(r"SUMMARY|Error|error|hazard|done|ring prep|tail split|dmma|few rows|simt|f16x3|tf32", ln):
                        f.write(ln)


if __name__ == "__main__":
    pre, tag = sys.argv[1], sys.argv[2]
    launches(pre, tag); single_pass(pre, tag)
    full(pre, tag, "full", "gemm_tc_kernel (F16X3 default fp32 mode, CTA pair) at 8192^3, second launch of tools/r2_ncu_f16_target.py")
    full(pre, tag, "prep", "operand preparation kernels of the F16X3 mode at 8192^3")
    full(pre, tag, "aux", "fp64 DMMA kernel, 4096^3 with split-K of the last wave (+ tail reduce), im2col + few-rows kernel (conv2d 16x3x224x224 -> 20), transpose 8192^2: tools/ncu_r2_aux_target.py")
    source_hot(pre, tag)
    copies(pre, tag)
