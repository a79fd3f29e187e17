"""profiles/<tag>_<kernel>_ncu.json for every kernel of the two workloads, from the raw-page CSV exports that
tools/gpu_ncu_all.sh leaves in gpurun_out/ (ncu --set full of one forward of the 2-layer, full-width models).

Each file: per launch shape (grid, block) of that kernel, the launch count, mean duration, DRAM bytes read+written per
launch, achieved DRAM GB/s and its fraction of MEASURED_PEAKS.json's copy bandwidth, tensor-pipe / XU / FMA / ALU /
issue utilisation, registers, dynamic shared memory. ncu durations are cold-cache and serialised (and the clock is not
locked): use them for shares and for the counters, not as bench numbers.

    python tools/summarize_ncu_all.py [tag]      # default r02
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

M = {
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "xu_pct": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "alu_pct": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lsu_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "smem": "launch__shared_mem_per_block_dynamic",
    "sm_mhz": "sm__cycles_elapsed.avg.per_second",
}
UNIT_SCALE = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
              "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short_name(full: str) -> str:
    m = re.search(r"(\w+?)(_kernel)?(<[^>]*>)?\(", full)
    name = full.split("(")[0].split("::")[-1]
    name = re.sub(r"^void ", "", name).strip()
    return re.sub(r"[^\w<>,]", "", name)


def fnum(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        rec = {"name": r[ix["Kernel Name"]]}
        for key, metric in M.items():
            if metric not in ix:
                rec[key] = None
                continue
            v = fnum(r[ix[metric]])
            u = units[ix[metric]]
            if v is not None and key in ("dur", "rd", "wr"):
                v *= UNIT_SCALE.get(u, 1.0)
            if v is not None and key == "sm_mhz":
                v = v / 1e6 if u in ("hz", "Hz", "cycle/second") else v * {"Mhz": 1.0, "Ghz": 1e3}.get(u, 1.0)
            rec[key] = v
        out.append(rec)
    return out


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    index = {}
    for wl in ("wavlm", "whisper", "whisper_full_length"):
        path = os.path.join(OUT, f"{tag}_all_{wl}.raw.csv")
        if not os.path.exists(path):
            continue
        groups = collections.OrderedDict()
        for rec in load(path):
            key = (short_name(rec["name"]), rec["grid"], rec["block"])
            groups.setdefault(key, []).append(rec)
        total_us = sum(r["dur"] for g in groups.values() for r in g)
        per_kernel = collections.OrderedDict()
        for (name, grid, block), recs in groups.items():
            n = len(recs)

            def mean(k):
                v = [r[k] for r in recs if r[k] is not None]
                return sum(v) / len(v) if v else None

            dur = mean("dur")
            byt = (mean("rd") or 0.0) + (mean("wr") or 0.0)
            gbs = byt / (dur * 1e-6) / 1e9 if dur else None
            per_kernel.setdefault(name, []).append({
                "workload": wl, "grid": int(grid), "block": int(block), "launches": n,
                "duration_us_mean": round(dur, 2), "share_of_forward": round(n * dur / total_us, 4),
                "dram_bytes_per_launch": int(byt), "dram_gbs": round(gbs, 1) if gbs else None,
                "dram_frac_of_measured_peak": round(gbs / hbm, 4) if gbs else None,
                "dram_throughput_pct": mean("dram_pct"), "l2_throughput_pct": mean("l2_pct"),
                "tensor_pipe_pct": mean("tensor_pct"), "xu_pipe_pct": mean("xu_pct"), "fma_pipe_pct": mean("fma_pct"),
                "alu_pipe_pct": mean("alu_pct"), "issue_active_pct": mean("issue_pct"),
                "lsu_wavefronts_pct": mean("lsu_pct"), "warps_active_pct": mean("warps_pct"),
                "registers": mean("regs"), "dyn_smem_bytes": mean("smem"), "sm_mhz": mean("sm_mhz")})
        for name, entries in per_kernel.items():
            fn = re.sub(r"[<>,]", "_", name).strip("_")
            path_out = os.path.join(PROF, f"{tag}_{fn}_ncu.json")
            prev = json.load(open(path_out)) if os.path.exists(path_out) and fn in index else \
                {"kernel": name, "hbm_peak_gbs_measured": hbm,
                 "note": "ncu --set full --clock-control none, one forward of the 2-layer full-width model "
                         "(tools/ncu_all.py); cold-cache serialised launches", "shapes": []}
            prev["shapes"].extend(entries)
            json.dump(prev, open(path_out, "w"), indent=1)
            index.setdefault(fn, []).extend(entries)
    rows = []
    for fn, entries in index.items():
        for e in entries:
            rows.append((e["workload"], fn, e["grid"], e["launches"], e["duration_us_mean"], e["share_of_forward"],
                         e["dram_gbs"], e["tensor_pipe_pct"], e["xu_pipe_pct"], e["fma_pipe_pct"], e["issue_active_pct"]))
    rows.sort(key=lambda r: (r[0], -r[3] * r[4]))
    print("%-20s %-34s %7s %3s %10s %6s %8s %6s %6s %6s %6s" % ("workload", "kernel", "grid", "n", "us", "share", "GB/s",
                                                               "tens%", "xu%", "fma%", "iss%"))
    for r in rows:
        print("%-20s %-34s %7d %3d %10.1f %6.3f %8s %6s %6s %6s %6s" % (
            r[0], r[1][:34], r[2], r[3], r[4], r[5], r[6], *[("%.1f" % v) if v is not None else "-" for v in r[7:]]))


if __name__ == "__main__":
    main()
