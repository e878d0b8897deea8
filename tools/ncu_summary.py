#!/usr/bin/env python
"""Summaries of `ncu --set full` reports for profiles/: per report the kernel, launch duration and the counters the
roofline discussion needs (DRAM bytes, L2 -> SM bytes, tensor-pipe / shared-memory-pipe / L2 / DRAM utilisation, issue
rate), plus the top stall sites of the source page. Runs where ncu is installed (no GPU needed to read a report).

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [...] [--md profiles/x.md] [--json profiles/traffic_r02.json --key name]
"""
import argparse
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read bytes"),
    ("l1tex__m_l1tex2xbar_write_bytes.sum", "SM -> L2 write bytes"),
    ("lts__t_bytes.sum", "L2 bytes (all)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__inst_executed_pipe_tensor_op_utchmma.avg.pct_of_peak_sustained_active", "tensor pipe (inst, % of peak, active)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor subpipe hmma cycles active %"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor ops tf32 % of peak (elapsed)"),
    ("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.peak_sustained", "tensor ops tf32 peak per SM"),
    ("sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "tensor ops bf16 % of peak (elapsed)"),
    ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe cycles active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: LSU shared wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed", "smem pipe: LSU shared loads %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem pipe: tensor-core operand wavefronts %"),
    ("l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed", "smem/L1 data bank reads %"),
    ("l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "smem/L1 data bank writes %"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem pipe: LSU wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.avg.pct_of_peak_sustained_elapsed", "smem pipe: LSU ld %"),
    ("l1tex__data_pipe_tensor_wavefronts.avg.pct_of_peak_sustained_elapsed", "smem pipe: tensor operand wavefronts %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1 data pipe LSU wavefronts %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("launch__grid_size", "grid"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
]


def raw_page(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def source_top(path, n=12):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]

    def f(r, k):
        try:
            return float(r[ix[k]])
        except (ValueError, KeyError):
            return 0.0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(f(r, "# Samples") for r in data) or 1.0
    agg = sorted(((s, sum(f(r, s) for r in data)) for s in stalls), key=lambda x: -x[1])
    top = sorted(data, key=lambda r: -f(r, "# Samples"))[:n]
    lines = []
    for r in top:
        st = sorted(((s, f(r, s)) for s in stalls), key=lambda x: -x[1])[0]
        lines.append((100 * f(r, "# Samples") / tot, " ".join(r[ix["Source"]].split())[:72], st[0][6:]))
    lds = [r for r in data if "LDS" in r[ix["Source"]]]
    wf = sum(f(r, "L1 Wavefronts Shared") for r in lds)
    wfi = sum(f(r, "L1 Wavefronts Shared Ideal") for r in lds)
    return tot, [(s[6:], 100 * v / tot) for s, v in agg[:8]], lines, (wf, wfi)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reports", nargs="+")
    ap.add_argument("--md")
    ap.add_argument("--json")
    ap.add_argument("--keys", default="")
    ap.add_argument("--no-source", action="store_true")
    args = ap.parse_args()
    md = []
    js = json.load(open(args.json)) if args.json and __import__("os").path.exists(args.json) else {}
    keys = args.keys.split(",") if args.keys else []
    for i, path in enumerate(args.reports):
        m = raw_page(path)
        name = m.get("Kernel Name", ("?", ""))[0]
        md.append(f"### `{path.split('/')[-1]}` -- `{name[:110]}`\n")
        md.append("| metric | value |\n|---|---|")
        vals = {}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
        for key, label in METRICS:
            if key in m:
                v, u = m[key]
                md.append(f"| {label} (`{key}`) | {v} {u} |")
                try:
                    vals[key] = float(v.replace(",", "")) * scale.get(u, 1.0)   # bytes / microseconds / plain numbers
                except ValueError:
                    pass
        if not args.no_source:
            tot, agg, lines, (wf, wfi) = source_top(path)
            md.append(f"\nWarp-state samples ({tot:.0f}): " + ", ".join(f"{s} {p:.1f} %" for s, p in agg))
            if wfi:
                md.append(f"\nLDS wavefronts / ideal: {wf:.0f} / {wfi:.0f} = {wf / wfi:.2f}")
            md.append("\nTop stall sites (share of samples, SASS, dominant reason):\n")
            md.append("| % | instruction | reason |\n|---|---|---|")
            for p, src, why in lines:
                md.append(f"| {p:.1f} | `{src}` | {why} |")
        md.append("")
        if args.json and i < len(keys):
            def g(k):
                return vals.get(k)
            js[keys[i]] = {
                "report": path.split("/")[-1], "kernel": name[:120],
                "duration_us_under_ncu": g("gpu__time_duration.sum"),
                "dram_bytes_per_launch": (g("dram__bytes_read.sum") or 0) + (g("dram__bytes_write.sum") or 0),
                "dram_read_bytes": g("dram__bytes_read.sum"), "dram_write_bytes": g("dram__bytes_write.sum"),
                "l2_to_sm_read_bytes": g("l1tex__m_xbar2l1tex_read_bytes.sum"),
                "sm_to_l2_write_bytes": g("l1tex__m_l1tex2xbar_write_bytes.sum"),
                "how": "ncu --set full --clock-control none, one launch after warm-up, rotating input/output sets; "
                       "tools/gpu_round.sh + tools/ncu_summary.py",
                "tensor_tf32_pct_of_peak_elapsed": g("sm__ops_path_tensor_op_utchmma_src_tf32_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed"),
                "lts_throughput_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                "dram_throughput_pct": g("dram__throughput.avg.pct_of_peak_sustained_elapsed"),
            }
    text = "\n".join(md)
    if args.md:
        open(args.md, "w").write(text + "\n")
    else:
        print(text)
    if args.json:
        json.dump(js, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
