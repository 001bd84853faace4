#!/usr/bin/env python
"""Round-2 profile summary: turns what tools/r2_measure.sh left in gpurun_out/ into the tracked files under profiles/.

    python profiles/summarise_r2.py            (run here, no GPU: `ncu -i` only reads the reports)

Writes  profiles/r2_final_bench*.json       the bench lines
        profiles/r2_final_launches.csv      the ncu launch list (gpu__time_duration.sum per launch)
        profiles/r2_ncu_full_<kernel>.csv   raw page of every `ncu --set full` capture
        profiles/traffic.json               dram bytes per launch of each stage's dominant kernel
        profiles/r2_summary.md              the tables
"""
import collections
import csv
import json
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "gpurun_out")

KEYS = [("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("smsp__inst_executed.sum", "warp inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefront %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(val.replace(",", "")) * mult


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None, out
    return (rows[0], rows[1], rows[2]), out


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[4].split("(")[0].replace("void ", "").split("<")[0], []).append((float(r[-1]) / 1000, r[8], r[7]))
    return agg


def main():
    lines = ["# Round 2, final state - one B200 (`tools/r2_measure.sh`, summarised by `profiles/summarise_r2.py`)", ""]
    # ---- bench lines
    for name in ("r2_final_bench", "r2_final_bench_c3", "r2_final_bench_c5", "r2_final_bench_reference", "r2_bench_n2", "r2_bench_c4_n2",
                 "r2_bench_n4", "r2_bench_n8", "r2_bench_c4_n8"):
        src = os.path.join(OUT, name + ".json")
        if os.path.exists(src):
            txt = [l for l in open(src).read().splitlines() if l.startswith("{")]
            if txt:
                open(os.path.join(HERE, name + ".json"), "w").write(txt[-1] + "\n")
    b = json.load(open(os.path.join(HERE, "r2_final_bench.json")))
    lines += ["## Config 2 (B=8, 10x272x480, K=1500, 256-channel features; %d RoIs per step)" % b["config"]["rois_per_step"], "",
              "| | |", "|---|---|",
              "| device-resident, two batches in flight (`value`) | **%.0f images/s**, %.3f ms per step |" % (b["value"], b["ms_per_step"]),
              "| one batch at a time (`single_batch`) | %.0f images/s, %.3f ms per step |" % (b["single_batch"]["value"], b["single_batch"]["ms_per_step"]),
              "| through host buffers (`e2e`, %.3f GB H2D per step) | %.1f images/s |" % (b["e2e"]["h2d_bytes_per_step"] / 1e9, b["e2e"]["value"]),
              "| reference's own CUDA/torch sequence, same batch (`reference_cuda`) | %.0f images/s (%.1f ms per step) -> %.1fx |" % (
                  b["reference_cuda"]["value"], b["reference_cuda"]["ms_per_step"], b["value"] / b["reference_cuda"]["value"]),
              "| reference CPU sequence, %d cores (`cpu_baseline`) | %.2f images/s |" % (b["cpu_baseline"]["cores"], b["cpu_baseline"]["value"]),
              "| clocks during the timed region | %s MHz of %s, reasons %s |" % (b["clocks"]["sm_mhz"], b["clocks"]["sm_max_mhz"], b["clocks"]["reasons"]),
              "| `variants.feat_is_relu` (features already ReLU-ed, as the host mirror passes them) | %s |" % json.dumps(b.get("variants", {}).get("feat_is_relu")),
              ""]
    lines += ["Stage times (CUDA events between the stages, eager): " + ", ".join("%s %.3f ms" % kv for kv in b["stages_ms"].items()), "",
              "| kernel (device time from `rr_kernel_trace`, eager, queue kept full) | us | designed bytes | GB/s | of 6546 | of 8000 | bound |",
              "|---|---|---|---|---|---|---|"]
    for r in b["rooflines"]["kernels"]:
        lines.append("| %s | %.1f | %.1f MB | %.0f | %.3f | %.3f | %s |" % (r["kernel"], r["ms"] * 1e3, r["bytes"] / 1e6, r["achieved"],
                                                                           r["frac"], r["frac_nominal_8tbs"], r["bound"]))
    lines += ["", "| stage (SURVEY 8d algorithmic bytes) | ms | bytes | GB/s | of 6546 | of 8000 |", "|---|---|---|---|---|---|"]
    for r in b["rooflines"]["stages"]:
        lines.append("| %s | %.4f | %.1f MB | %.0f | %.3f | %.3f |" % (r["kernel"], r["ms"], r["bytes"] / 1e6, r["achieved"], r["frac"], r["frac_nominal_8tbs"]))
    lines += ["", "`aux` (other scope rows, cold L2): " + ", ".join("%s %.1f us" % (k, v["ms"] * 1e3) for k, v in b["aux"].items() if "ms" in v), "",
              "`aux.hm_tail_fusion_c2`: " + json.dumps(b["aux"].get("hm_tail_fusion_c2")), "",
              "`nms_baselines` (ms, host arrays in / keep list out unless `device`): ", "", "```", json.dumps(b.get("nms_baselines"), indent=1), "```", ""]
    for name, title in (("r2_final_bench_c3", "Config 3 (training loss path, B=32 at 512x512)"), ("r2_final_bench_c5", "Config 5 (dense scene, B=16, K=5000)")):
        p = os.path.join(HERE, name + ".json")
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        lines += ["## " + title, "", "%.0f images/s, %.4f ms per step; e2e %.0f images/s; roofline %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], json.dumps({k: d["roofline"][k] for k in ("kernel", "achieved", "frac")})), "",
            "| kernel | us | bytes | GB/s | of 6546 | of 8000 |", "|---|---|---|---|---|---|"]
        for r in d["rooflines"]["kernels"]:
            lines.append("| %s | %.1f | %.1f MB | %.0f | %.3f | %.3f |" % (r["kernel"], r["ms"] * 1e3, r["bytes"] / 1e6, r["achieved"], r["frac"], r["frac_nominal_8tbs"]))
        lines.append("")
    for name in ("r2_bench_n2", "r2_bench_c4_n2", "r2_bench_n4", "r2_bench_n8", "r2_bench_c4_n8"):
        p = os.path.join(HERE, name + ".json")
        if os.path.exists(p):
            d = json.load(open(p))
            lines.append("* `%s`: %d GPUs, %s scaling, B=%d per GPU: %.0f images/s, %.3f ms per step (one batch at a time %.3f ms); e2e %.0f images/s; %s" % (
                name, d["n_gpus"], d["scaling"], d["config"]["B"], d["value"], d["ms_per_step"],
                (d.get("single_batch") or {}).get("ms_per_step", float("nan")), d["e2e"]["value"], d["config"]["parallelism"]))
    lines.append("")
    # ---- launch list
    lp = os.path.join(OUT, "r2_final_launches.csv")
    if os.path.exists(lp):
        shutil.copy(lp, os.path.join(HERE, "r2_final_launches.csv"))
        agg = launches(lp)
        tot = sum(sum(x[0] for x in v) / len(v) for v in agg.values())
        lines += ["## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold cache, serialised: shares, not absolutes)", "",
                  "| kernel | launches | mean us | share | grid | block |", "|---|---|---|---|---|---|"]
        for k, v in agg.items():
            m = sum(x[0] for x in v) / len(v)
            lines.append("| %s | %d | %.1f | %.1f %% | %s | %s |" % (k, len(v), m, 100 * m / tot, v[0][1], v[0][2]))
        lines += ["| sum of means | | %.1f | | | |" % tot, ""]
    # ---- full captures
    traffic = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch, from `ncu --set full --clock-control none` "
                           "(profiles/r2_ncu_full_<kernel>.csv; tools/r2_measure.sh)"}
    stage_of = {"roi_tile_tma_kernel": "roi_align", "head_tc_kernel": "head", "decode_collect_kernel": "decode"}
    lines += ["## Full captures (`ncu --set full --clock-control none --import-source on`, one launch each)", "",
              "| kernel | " + " | ".join(t for _, t in KEYS) + " |", "|---|" + "---|" * len(KEYS)]
    for f in sorted(os.listdir(OUT)):
        if not (f.startswith("r2_prof_") and f.endswith(".ncu-rep")):
            continue
        kern = f[len("r2_prof_"):-len(".ncu-rep")]
        page, text = raw_page(os.path.join(OUT, f))
        if page is None:
            continue
        open(os.path.join(HERE, "r2_ncu_full_%s.csv" % kern), "w").write(text)
        h, u, v = page
        cells = []
        for key, _ in KEYS:
            if key in h:
                i = h.index(key)
                cells.append(("%s %s" % (v[i], u[i])).strip())
            else:
                cells.append("")
        lines.append("| %s | %s |" % (kern, " | ".join(cells)))
        if "dram__bytes_read.sum" in h:
            ir, iw = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
            tot = to_bytes(v[ir], u[ir]) + to_bytes(v[iw], u[iw])
            traffic[kern] = int(tot)
            if kern in stage_of:
                traffic[stage_of[kern]] = int(tot)
    lines.append("")
    json.dump(traffic, open(os.path.join(HERE, "traffic.json"), "w"), indent=1)
    open(os.path.join(HERE, "r2_summary.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
