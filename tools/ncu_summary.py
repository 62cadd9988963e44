"""Summarise ncu artefacts into profiles/ (run here, no GPU):
  python tools/ncu_summary.py launches <csv> <out.md> "<command line>"
  python tools/ncu_summary.py full <rep.ncu-rep> <out.md>"""
import collections, csv, subprocess, sys

KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__occupancy_limit_registers", "occ limit regs"),
        ("launch__occupancy_limit_shared_mem", "occ limit smem"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor inst"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("smsp__inst_executed.sum", "warp instructions")]
STALLS = ["barrier", "short_scoreboard", "long_scoreboard", "mio_throttle", "math_pipe_throttle", "not_selected", "wait",
          "dispatch_stall", "lg_throttle", "branch_resolving", "no_instruction", "sleeping", "membar", "tex_throttle"]


def launches(path, out, cmd):
    rows = list(csv.reader(open(path)))
    i0 = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[i0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[i0 + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]].split("(")[0]
        v = float(r[ix["Metric Value"]])
        unit = r[ix["Metric Unit"]]
        v = v / 1e3 if unit == "ns" else (v * 1e3 if unit == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list\n\nCommand: `{cmd}`\n(per-launch times are cold-cache and serialised: compare SHARES)\n\n")
        f.write(f"total captured: {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {100 * a[1] / tot:.1f}% |\n")


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of `{rep.split('/')[-1]}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')}\n\n| metric | value |\n|---|---|\n")
            for k, label in KEYS:
                if k in d and d[k] != "":
                    f.write(f"| {label} (`{k}`) | {d[k]} {u.get(k, '')} |\n")
            st = []
            for s in STALLS:
                k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
                if k in d and d[k]:
                    st.append((float(d[k]), s))
            f.write("| warp stalls per issue (top) | " + ", ".join(f"{s} {v:.2f}" for v, s in sorted(st, reverse=True)[:6]) + " |\n\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        full(sys.argv[2], sys.argv[3])
