"""Turns the two ncu captures of DESIGN.md §7 into the tracked summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/X_launches.csv            -> markdown table of the LAST wave's launches
  python tools/summarize_ncu.py full gpurun_out/X.ncu-rep [traffic.json]     -> markdown tables of every captured launch
                                                                                (+ DRAM traffic per extend launch as JSON)
  python tools/summarize_ncu.py source gpurun_out/X.ncu-rep                  -> per-source-line / per-phase instruction shares
  python tools/summarize_ncu.py counters raw.csv device_only.json out.json   -> per-RAY hardware counters of the extend kernel (warp
                                                                                instructions, active threads, DRAM bytes): what bench.py's
                                                                                roofline scales by its own live launch times
"""
import csv
import io
import json
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def short(name):
    for k in ("k_trace_spec", "k_trace_packet", "k_shade", "k_raygen", "k_accumulate", "k_tally", "k_commit_bounce"):
        if k in name:
            if k in ("k_trace_spec", "k_trace_packet"):      # first template argument: ANY (shadow rays); the packet kernel counts as a traversal launch of its class
                any_ = "(bool)1, " in name or "<1, " in name
                return "k_trace_spec (connect)" if any_ else "k_trace_spec (extend)"
            return k
    return name.split("(")[0][-40:]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = [(short(r[k]), float(r[v].replace(",", "")) / 1e3) for r in rows]
    last = max(i for i, (n, _) in enumerate(seq) if n == "k_raygen")
    wave = seq[last:]
    end = next((i for i, (n, _) in enumerate(wave) if n == "k_tally"), len(wave) - 1)
    wave = wave[:end + 1]
    tot = sum(t for _, t in wave)
    print("| kernel | duration us |\n|---|---|")
    for n, t in wave:
        print(f"| {n} | {t:.1f} |")
    share = {}
    for n, t in wave:
        share[n] = share.get(n, 0) + t
    print(f"\nSum {tot:.0f} us per wave. Shares: " + ", ".join(f"{n} {t / tot:.3f}" for n, t in share.items()))


def raw_rows(rep):
    # `rep` is an .ncu-rep, or the CSV that `ncu -i X.ncu-rep --page raw --csv` printed (reports of a whole wave exceed the
    # 64 MiB that gpurun brings back, so the export runs on the GPU box)
    out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]], dict(zip(hdr, rows[1]))


def full(rep, traffic_out=None):
    rows, units = raw_rows(rep)
    print("| kernel | ms | DRAM read GB | DRAM write MB | active threads / warp inst | issue active % | L1 hit % | L2 hit % | L1 throughput % | achieved occupancy % |\n|---|---|---|---|---|---|---|---|---|---|")
    rd = wr = 0.0
    n_ext = 0

    def f(d, k, scale=1.0):
        try:
            return float(d[k].replace(",", "")) * scale
        except Exception:
            return float("nan")
    for d in rows:
        name = short(d["Kernel Name"])
        tu = units["gpu__time_duration.sum"]
        ms = f(d, "gpu__time_duration.sum", {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}.get(tu, 1.0))
        ru, wu = units["dram__bytes_read.sum"], units["dram__bytes_write.sum"]
        sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        r_b, w_b = f(d, "dram__bytes_read.sum", sc.get(ru, 1.0)), f(d, "dram__bytes_write.sum", sc.get(wu, 1.0))
        if name == "k_trace_spec (extend)":
            rd += r_b; wr += w_b; n_ext += 1
        print(f"| {name} | {ms:.3f} | {r_b / 1e9:.3f} | {w_b / 1e6:.1f} | {f(d, 'smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | "
              f"{f(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {f(d, 'l1tex__t_sector_hit_rate.pct'):.1f} | "
              f"{f(d, 'lts__t_sector_hit_rate.pct'):.1f} | {f(d, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{f(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} |")
    print("\n(units as exported by ncu: time " + units["gpu__time_duration.sum"] + ", dram " + units["dram__bytes_read.sum"] + ")")
    print("\n## Detail: first extend, first shade, first connect\n")
    firsts, heads = [], []
    for want in ("k_trace_spec (extend)", "k_shade", "k_trace_spec (connect)"):
        hit = [d for d in rows if short(d["Kernel Name"]) == want]
        if hit:
            firsts.append(hit[0]); heads.append(want + " #0")
            if want == "k_trace_spec (extend)" and len(hit) > 1:
                firsts.append(hit[1]); heads.append(want + " #1")
    print("| metric | unit | " + " | ".join(heads) + " |\n|---|---|" + "---|" * len(heads))
    for m in RAW:
        if m in firsts[0]:
            print(f"| `{m}` | {units.get(m, '')} | " + " | ".join(d[m] for d in firsts) + " |")
    if traffic_out and n_ext:
        json.dump({"kernel": "k_trace_spec<false,false> (extend, merged)", "launches": n_ext, "dram_bytes_read_sum": rd, "dram_bytes_write_sum": wr,
                   "traffic_bytes_per_launch": (rd + wr) / n_ext, "source": f"ncu --set full --clock-control none, {rep} ({n_ext} extend launches of one wave)"},
                  open(traffic_out, "w"), indent=1)


def counters(rep, device_only_json, out_path):
    """Sums over the extend launches of ONE wave (ncu capture of `bench.py --device-only --steps S`), divided by the extend rays of that wave
    (the counters the same command prints): per-ray figures that do not depend on the wave size."""
    rows, units = raw_rows(rep)
    dev = json.load(open(device_only_json))
    sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tsc = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}

    def f(d, k, scale=1.0):
        return float(d[k].replace(",", "")) * scale
    ext = [d for d in rows if short(d["Kernel Name"]) == "k_trace_spec (extend)"]
    per_bounce = [r for r in dev["extend_rays_per_bounce"] if r]
    waves = dev["steps"] and max(1, len(ext) // max(1, len(per_bounce)))
    rays = sum(per_bounce)                     # all extend rays the command traced in its timed region ...
    # ... the capture may hold the warm-up waves too: keep the LAST len(per_bounce) extend launches (the timed wave)
    ext = ext[-len(per_bounce):]
    inst = sum(f(d, "smsp__inst_executed.sum") for d in ext)
    thr = sum(f(d, "smsp__inst_executed.sum") * f(d, "smsp__thread_inst_executed_per_inst_executed.ratio") for d in ext)
    dram = sum(f(d, "dram__bytes_read.sum", sc.get(units["dram__bytes_read.sum"], 1.0)) + f(d, "dram__bytes_write.sum", sc.get(units["dram__bytes_write.sum"], 1.0)) for d in ext)
    ms = sum(f(d, "gpu__time_duration.sum", tsc.get(units["gpu__time_duration.sum"], 1.0)) for d in ext)
    out = {"kernel": "k_trace_spec<ANY=false> (extend: every instantiation the wave launches)", "launches": len(ext), "rays": rays,
           "warp_inst_per_ray": inst / rays, "active_threads_per_warp": thr / inst, "dram_bytes_per_ray": dram / rays,
           "warp_inst_sum": inst, "dram_bytes_sum": dram, "ms_sum_under_ncu": ms,
           "per_launch": [{"bounce": i + 1, "rays": per_bounce[i], "ms": f(d, "gpu__time_duration.sum", tsc.get(units["gpu__time_duration.sum"], 1.0)),
                           "warp_inst": f(d, "smsp__inst_executed.sum"), "active_threads_per_warp": f(d, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                           "issue_active_pct": f(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "dram_bytes": f(d, "dram__bytes_read.sum", sc.get(units["dram__bytes_read.sum"], 1.0)) + f(d, "dram__bytes_write.sum", sc.get(units["dram__bytes_write.sum"], 1.0))}
                          for i, d in enumerate(ext)],
           "source": f"ncu --set full --clock-control none of `bench.py --device-only --steps {dev['steps']}` ({rep}): the {len(ext)} extend launches of the timed wave; rays from the same command's counters"}
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps({k: v for k, v in out.items() if k != "per_launch"}, indent=1))


def source(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur = fn = hdr = None
    per_fn = {}
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            fn = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < 9 or r[2] != "-":
            continue
        try:
            line, samples, inst, thr = int(r[0]), int(r[4]), int(r[7]), int(r[8])
        except ValueError:
            continue
        per_fn.setdefault(short(fn), {})[(cur, line)] = (samples, inst, thr, r[1].strip()[:110])
    for name, agg in per_fn.items():
        ts, ti, tt = (sum(v[k] for v in agg.values()) for k in range(3))
        print(f"\n### {name}: {ts} stall samples, {ti} warp instructions, {tt / max(ti, 1):.1f} active threads per instruction\n")
        print("| file:line | stall samples % | warp instructions % | threads / instruction | source |\n|---|---|---|---|---|")
        for (f_, l_), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:24]:
            print(f"| {f_}:{l_} | {100 * v[0] / ts:.1f} | {100 * v[1] / ti:.1f} | {v[2] / max(v[1], 1):.1f} | `{v[3]}` |")


if __name__ == "__main__":
    {"launches": launches, "full": full, "source": source, "counters": counters}[sys.argv[1]](*sys.argv[2:])
