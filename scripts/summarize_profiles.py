"""Turn the ncu reports / launch lists under gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r02
"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

KEYS = ['gpu__time_duration.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__shared_mem_per_block_dynamic', 'launch__cluster_size',
        'launch__grid_size', 'launch__block_size', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return [(dict(zip(rows[0], r)), dict(zip(rows[0], rows[1]))) for r in rows[2:]]


def table(f, v, u):
    f.write("| metric | value | unit |\n|---|---|---|\n")
    for k in KEYS:
        if k in v:
            f.write(f"| {k} | {v[k]} | {u[k]} |\n")
    f.write("\n")


def to_bytes(x, unit):
    return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    traffic_path = os.path.join(OUT, "aug_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    for key, rep, cmd in (("B4096_s224", f"aug_{tag}_final.ncu-rep", "python scripts/prof_aug.py 4096 224 0"),
                          ("B1024_s224", f"aug_{tag}_b1024.ncu-rep", "python scripts/prof_aug.py 1024 224 0")):
        rep = os.path.join(G, rep)
        if not os.path.exists(rep):
            continue
        with open(os.path.join(OUT, f"{tag}_aug_strip_{key}_ncu.md"), "w") as f:
            f.write(f"# K1 aug_strip_kernel (the default variant), build of {tag}: `ncu --set full --clock-control none`\n\n"
                    f"Command: `{cmd}` ({key[1:].split('_')[0]} slices 512x512 u16 -> views 224x224 bf16; blur / solarize off as in bench.py).\n"
                    "Times under ncu are cold-cache/serialised; bench.py's CUDA-event time is the number of record.\n\n")
            for v, u in raw(rep):
                f.write(f"## {v.get('Kernel Name', '')[:90]}\n\n")
                table(f, v, u)
                rd = to_bytes(v['dram__bytes_read.sum'], u['dram__bytes_read.sum'])
                wr = to_bytes(v['dram__bytes_write.sum'], u['dram__bytes_write.sum'])
                traffic[key] = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "kernel": "aug_strip_kernel",
                                "source": os.path.basename(rep), "command": cmd}
    with open(traffic_path, "w") as f:
        json.dump(traffic, f, indent=1)
    rep = os.path.join(G, f"ntxent_{tag}_final.ncu-rep")
    if os.path.exists(rep):
        with open(os.path.join(OUT, f"{tag}_ntxent_ncu.md"), "w") as f:
            f.write(f"# K2/K3 ntxent_tile_kernel, build of {tag}: `ncu --set full --clock-control none`\n\n"
                    "Command: `python scripts/prof_loss.py 4096 128 3` (2N = 8192 rows, D = 128, one rank; forward = <0>, backward = <1>).\n\n")
            for v, u in raw(rep):
                f.write(f"## {v.get('Kernel Name', '')[:90]}\n\n")
                table(f, v, u)
    rep = os.path.join(G, f"ntxent_{tag}_wide.ncu-rep")
    if os.path.exists(rep):
        with open(os.path.join(OUT, f"{tag}_ntxent_wide_ncu.md"), "w") as f:
            f.write(f"# NT-Xent with wide embeddings, build of {tag}: `ncu --set full --clock-control none`\n\n"
                    "Command: `python scripts/prof_loss.py 8192 2048 1` (2N = 16384 rows, D = 2048, one rank -- the per-GPU shape of "
                    "cfg4 on two GPUs is 16384 x 32768; forward = ntxent_tile_kernel<0, 0>, backward = ntxent_tile_kernel<1, 1> "
                    "writing W + wu_gemm_kernel computing dU = W . U_all).\n\n")
            for v, u in raw(rep):
                f.write(f"## {v.get('Kernel Name', '')[:90]}\n\n")
                table(f, v, u)
    lst = os.path.join(G, f"launches_{tag}.csv")
    if os.path.exists(lst):
        rows = [r for r in csv.reader(open(lst)) if len(r) > 10 and r[0].isdigit()]
        agg = collections.OrderedDict()
        for r in rows:
            name = r[4].split("(")[0].replace("void ", "")[:60]
            agg.setdefault(name, []).append(float(r[-1]))
        tot = sum(sum(v) for v in agg.values())
        with open(os.path.join(OUT, f"{tag}_bench_launch_list.md"), "w") as f:
            f.write(f"# Launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-weak` ({tag}, cfg3 at N = 1: 4096 slices)\n\n"
                    "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` (cold-cache, serialised: compare shares).\n\n"
                    "| kernel | launches | avg us | share of GPU time |\n|---|---|---|---|\n")
            for k, v in agg.items():
                f.write(f"| {k} | {len(v)} | {sum(v) / len(v) / 1e3:.2f} | {100 * sum(v) / tot:.1f} % |\n")


if __name__ == "__main__":
    main()
