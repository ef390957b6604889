"""Turn the raw outputs of scripts/r2_final_profile.sh (gpurun_out/r02_*) into the tracked summaries under profiles/:
r02_bench_1gpu.json, r02_launches_one_step.csv + r02_launches_summary.txt, r02_ncu_igemm_step.csv, fprop_traffic.json."""
import collections
import csv
import io
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
SUSTAINED = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}


def rows_of(path):
    txt = open(path).read()
    return list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))


def us_of(r):
    v = float(r["Metric Value"].replace(",", ""))
    return v / 1000 if r["Metric Unit"] in ("ns", "nsecond") else v


def main():
    shutil.copy(os.path.join(G, "r02_bench_1gpu.json"), os.path.join(P, "r02_bench_1gpu.json"))
    shutil.copy(os.path.join(G, "r02_launches.csv"), os.path.join(P, "r02_launches_one_step.csv"))
    for f in ("r02_fprop_g128to64.ncu-rep", "r02_gpu_tests.log"):
        if os.path.isfile(os.path.join(G, f)):
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
    bench = json.load(open(os.path.join(G, "r02_bench_1gpu.json")))
    # ---- launch list per kernel
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows_of(os.path.join(G, "r02_launches.csv")):
        if r["Metric Name"] == "gpu__time_duration.sum":
            name = r["Kernel Name"].split("(")[0].replace("void ", "")
            tot[name] += us_of(r)
            cnt[name] += 1
    T = sum(tot.values())
    out = ["# ncu launch list of ONE eagerly launched cfg-2 step (profiles/r02_launches_one_step.csv), summed per kernel",
           "# (launches are serialised and cache-cold under ncu; the graph replays the same kernels in %.2f ms)" % bench["ms_per_step"],
           "# kernel, launches, us, share"]
    out += [f"{k}, {cnt[k]}, {v:.1f}, {v / T:.3f}" for k, v in tot.most_common()]
    fp = sum(v for k, v in tot.items() if "igemm_fprop" in k)
    out.append(f"# igemm_fprop_kernel (all modes): {fp:.0f} us of {T:.0f} us = {fp / T:.3f} of the serialised step "
               f"(bench.py roofline.share_of_step {bench['roofline']['share_of_step']:.3f}: event pass, same schedule)")
    open(os.path.join(P, "r02_launches_summary.txt"), "w").write("\n".join(out) + "\n")
    # ---- per-launch metrics of the tensor-core kernels
    by, order = {}, []
    for r in rows_of(os.path.join(G, "r02_igemm_step_metrics.csv")):
        k = (int(r["ID"]), r["Kernel Name"])
        if k not in by:
            order.append(k)
        by.setdefault(k, {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    pl = bench["roofline"]["per_launch"]
    assert len(pl) == len(order), (len(pl), len(order))
    peak = bench["roofline"]["peak"]
    lines = ["# ncu per-launch metrics of the 65 tensor-core launches of ONE eagerly launched cfg-2 step (B = 256, bf16, final round-2 build)",
             "# command: scripts/r2_final_profile.sh (ncu --profile-from-start off --clock-control none -k regex:igemm --metrics ... python scripts/profile_one_step.py)",
             "# launches are serialised and cache-cold under ncu; GFLOP = algorithmic (bench.py per_launch, same order); tensor% = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
             "idx,kernel,kind,gflop,us,tflops,tensor_pipe_pct,dram_mb,l2_mb,sm_throughput_pct,warp_inst_M,grid,regs,smem_kb"]
    agg = {}
    for n, (k, p) in enumerate(zip(order, pl)):
        m = by[k]
        val = lambda key: m[key][0]
        us = val("gpu__time_duration.sum") / (1000 if m["gpu__time_duration.sum"][1] in ("ns", "nsecond") else 1)
        scale = lambda key: {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m[key][1], 1)
        dram = val("dram__bytes_read.sum") * scale("dram__bytes_read.sum") + val("dram__bytes_write.sum") * scale("dram__bytes_write.sum")
        l2 = val("lts__t_bytes.sum") * scale("lts__t_bytes.sum")
        smem = val("launch__shared_mem_per_block_dynamic") * scale("launch__shared_mem_per_block_dynamic") / 1024
        tp = val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
        name = k[1].split("(")[0].replace("void ", "")
        assert (p["kind"] == "wgrad") == ("wgrad" in name), (n, p["kind"], name)
        lines.append(f'{n},"{name}",{p["kind"]},{p["gflop"]},{us:.1f},{p["gflop"] / us * 1e3:.0f},{tp:.1f},{dram / 1e6:.1f},'
                     f'{l2 / 1e6:.0f},{val("sm__throughput.avg.pct_of_peak_sustained_elapsed"):.0f},'
                     f'{val("smsp__inst_executed.sum") / 1e6:.1f},{int(val("launch__grid_size"))},'
                     f'{int(val("launch__registers_per_thread"))},{smem:.0f}')
        t = agg.setdefault(p["kind"], dict(n=0, us=0.0, gf=0.0, dram=0.0, tp=0.0))
        t["n"] += 1; t["us"] += us; t["gf"] += p["gflop"]; t["dram"] += dram; t["tp"] += tp * us
    for kind, t in agg.items():
        tf = t["gf"] / t["us"] * 1e3
        lines.append(f"# {kind}: {t['n']} launches, {t['us']:.0f} us, {t['gf']:.0f} GFLOP -> {tf:.0f} TFLOP/s = {tf / peak:.3f} of the "
                     f"sustained bf16 peak ({peak}); time-weighted tensor pipe active {t['tp'] / t['us']:.1f} %; DRAM "
                     f"{t['dram'] / t['n'] / 1e6:.1f} MB / launch")
    open(os.path.join(P, "r02_ncu_igemm_step.csv"), "w").write("\n".join(lines) + "\n")
    f = agg["fprop"]
    json.dump({"dram_bytes_per_launch": f["dram"] / f["n"], "launches": f["n"],
               "tensor_pipe_active_pct_time_weighted": f["tp"] / f["us"], "us_per_launch_under_ncu": f["us"] / f["n"],
               "source": "profiles/r02_ncu_igemm_step.csv (final round-2 build): ncu --profile-from-start off --clock-control none "
                         "-k regex:igemm over the 45 fprop-type launches of one eagerly launched cfg-2 step "
                         "(scripts/profile_one_step.py), dram__bytes_read.sum + dram__bytes_write.sum averaged per launch; "
                         "outputs that fit stay in the 126 MB L2 within a kernel's window"},
              open(os.path.join(P, "fprop_traffic.json"), "w"), indent=1)
    print("\n".join(lines[-2:]))
    print(out[-1])


if __name__ == "__main__":
    main()
