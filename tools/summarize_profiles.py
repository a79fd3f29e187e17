"""Turn the raw outputs of tools/gpu_ncu_bench.sh + a bench.py run (gpurun_out/) into the tracked summaries under
profiles/: the ncu launch list, kernel-group shares (ncu vs in-run CUDA events), the --set full GEMM capture summary.

    python tools/summarize_profiles.py [round_tag]        # default r02
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def group_of(name: str) -> str:
    if "gemm_tc" in name or "gemm_ln" in name or "posconv_tc" in name or name.startswith("gemm"):
        return "gemm (tcgen05)"
    if "attention" in name:
        return "attention"
    if "layernorm" in name:
        return "layernorm"
    if "conv0" in name:
        return "conv0"
    return "other"


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    shutil.copy(os.path.join(OUT, "launches.csv"), os.path.join(PROF, f"{tag}_ncu_launches.csv"))
    shutil.copy(os.path.join(OUT, "bench_final.log"), os.path.join(PROF, f"{tag}_bench_final.log"))
    rows = list(csv.reader(open(os.path.join(OUT, "launches.csv"))))
    hdr, d = None, collections.defaultdict(list)
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            rec = dict(zip(hdr, r))
            if rec.get("Metric Name") == "gpu__time_duration.sum":
                v, u = float(rec["Metric Value"].replace(",", "")), rec["Metric Unit"]
                v = v / 1e3 if u in ("ns", "nsecond") else v if u in ("us", "usecond") else v * 1e3
                d[group_of(rec["Kernel Name"])].append(v)
    tot = sum(sum(v) for v in d.values())
    ncu = {k: {"launches": len(v), "us": round(sum(v), 1), "share": round(sum(v) / tot, 4)} for k, v in d.items()}
    bench = None
    for ln in open(os.path.join(OUT, "bench_final.log")):
        if ln.startswith("{"):
            bench = json.loads(ln)
    grp = collections.defaultdict(float)
    for k, v in bench["kernels_ms_per_step"].items():
        grp[group_of(k if not k.startswith("wavlm_") else k[6:])] += v["ms"]
    t = sum(grp.values())
    shares = {"note": "share of one WavLM-L step per kernel group: ncu launch list (cold cache, serialised, 2 steps "
                      "captured) vs in-run CUDA events (bench.py, profile option)",
              "ncu_launch_list": ncu,
              "in_run_cuda_events": {k: {"ms": round(v, 3), "share": round(v / t, 4)} for k, v in grp.items()}}
    json.dump(shares, open(os.path.join(PROF, f"{tag}_launch_shares.json"), "w"), indent=1)
    print(json.dumps(shares["ncu_launch_list"]))
    print(json.dumps(shares["in_run_cuda_events"]))

    rep = os.path.join(OUT, "prof_gemm_tc2.ncu-rep")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        h, units = rr[0], rr[1]
        want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                "lts__throughput.avg.pct_of_peak_sustained_elapsed",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__cycles_active.avg", "launch__shared_mem_per_block_dynamic"]
        ix = {k: i for i, k in enumerate(h)}
        summ = [{k: (r[ix[k]][:40] + " " + units[ix[k]]).strip() for k in want if k in ix} for r in rr[2:]]
        json.dump(summ, open(os.path.join(PROF, f"{tag}_gemm_tc2_ncu_full_summary.json"), "w"), indent=1)

        def num(sv):
            v, u = sv.rsplit(" ", 1) if " " in sv else (sv, "")
            return float(v.replace(",", "")) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)

        per = [num(x["dram__bytes_read.sum"]) + num(x["dram__bytes_write.sum"]) for x in summ]
        traffic = {"kernel": "gemm_tc2_kernel",
                   "source": "ncu --set full --clock-control none -k regex:gemm_tc2 -s 40 -c 4 on the bench command "
                             "(one encoder layer's four GEMM launches); tools/gpu_ncu_bench.sh",
                   "launch_traffic_bytes": [int(v) for v in per],
                   "traffic_bytes_per_launch_avg": int(sum(per) / max(len(per), 1)),
                   "algorithmic_bytes_per_launch_avg": 407000000}
        json.dump(traffic, open(os.path.join(PROF, f"{tag}_gemm_traffic.json"), "w"), indent=1)
        for s in summ:
            print(s["gpu__time_duration.sum"], s["dram__bytes_read.sum"], s["dram__bytes_write.sum"],
                  s["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"])


if __name__ == "__main__":
    main()
